// Halo exchange between neighbouring strips of a domain-decomposed tissue: ONE kernel per
// exchange point pushes this rank's boundary values straight into the neighbours' windows
// (peer stores over NVLink), publishes an epoch flag with a system-scope release and, in the
// fused mode, waits for the neighbours' flags of the same epoch (bounded spin).
//
//   X1 (after k_mem):    ghost-cell cc_mid[I] + Vmem of the cells whose gap-junction partner lives
//                        on the neighbour (cells.nn_i), and the membrane->env fluxes of membranes
//                        whose env square the neighbour owns (cells.map_mem2ecm)
//   X2 (after k_envacc): cc_env rows (radius-2 transport stencil, sim.py:2209-2254) and raw env
//                        voltage rows (9-tap Gaussian + gradient, ion_current.py:101-109)
#include "kparams.cuh"
#include "xchg.cuh"

#define XCHG_PUSH 1
#define XCHG_WAIT 2
__global__ void __launch_bounds__(256)
k_xchg(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ XPlan X,
       const int which, const int buf, const int mode, const int flux_ell)
{
    const int NI = P.n_ions;
    const int C = P.n_cells, E = P.ny * P.nx, nx = P.nx;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int gsz = gridDim.x * blockDim.x;
    __shared__ int s_last;

    if (mode & XCHG_PUSH) {
        for (int k = 0; k < X.n_nbr; ++k) {
            const XNbr& nb = X.nb[k];
            if (which == 0) {
                const double* __restrict__ cm = A.cc_mid[buf];
                const double* __restrict__ vc = A.vm_cell[buf];
                double* __restrict__ dcm = nb.cc_mid[buf];
                double* __restrict__ dvc = nb.vm_cell[buf];
                const int nc = nb.n_send_cells;
                for (int t = gtid; t < nc * (NI + 1); t += gsz) {
                    const int i = t / nc, j = t - i * nc;
                    const int c = __ldg(nb.send_cells + j);
                    if (i < NI) dcm[i * nb.Cn + nb.recv_cell0 + j] = cm[i * C + c];
                    else dvc[nb.recv_cell0 + j] = vc[c];
                }
                const int nf = nb.n_send_flux;
                for (int t = gtid; t < nf * NI; t += gsz) {
                    const int j = t / NI, i = t - j * NI;
                    // the receiver's remote slots are [slot][ion] whichever kernel produced the fluxes here
                    double v;
                    if (flux_ell) v = A.flux_ell[(size_t)__ldg(nb.send_flux_ell + j) + i * 32];
                    else v = A.flux_slots[__ldg(nb.send_flux + j) * NI + i];
                    nb.flux[(nb.recv_slot0 + j) * NI + i] = v;
                }
            } else {
                const int ncc = nb.cc_rows * nx;
                for (int t = gtid; t < ncc * NI; t += gsz) {
                    const int i = t / ncc, j = t - i * ncc;
                    nb.cc_env[buf][i * nb.En + nb.cc_dst_row0 * nx + j] = A.cc_env[buf][i * E + nb.cc_src_row0 * nx + j];
                }
                const int nv = nb.v_rows * nx;
                for (int t = gtid; t < nv; t += gsz) nb.v_raw[nb.v_dst_row0 * nx + t] = A.v_raw[nb.v_src_row0 * nx + t];
            }
        }
        // publish: the CTA's stores are ordered before thread 0's system-scope fence by the barrier
        // (fences are cumulative), and the last CTA to arrive raises the flags
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned int prev = atomicAdd(X.done_ctr + which, 1u);
            s_last = (prev == gridDim.x - 1) ? 1 : 0;
        }
        __syncthreads();
        if (!s_last) return;
        if (threadIdx.x == 0) {
            X.done_ctr[which] = 0;
            const unsigned long long e = X.epoch[which] + 1ull;
            X.epoch[which] = e;
            __threadfence_system();
            for (int k = 0; k < X.n_nbr; ++k)
                st_release_sys(X.nb[k].flags + which * 2 + (1 - X.nb[k].side), e);
        }
    } else if (blockIdx.x != 0) return;

    if ((mode & XCHG_WAIT) && threadIdx.x == 0) {
        const unsigned long long e = X.epoch[which];
        const unsigned long long t0 = globaltimer_ns();
        for (int k = 0; k < X.n_nbr; ++k) {
            const unsigned long long* f = X.my_flags + which * 2 + X.nb[k].side;
            while (ld_acquire_sys(f) < e) {
                if (globaltimer_ns() - t0 > X.timeout_ns) { atomicOr(A.status, ST_XCHG_TIMEOUT); break; }
            }
        }
    }
}

void launch_xchg(const KParams& P, const KArrays& A, const XPlan& X, int which, int buf, int mode, int flux_ell, cudaStream_t st)
{
    // enough CTAs to cover the largest block copy; the wait-only form needs one
    long long work = 1;
    if (mode & XCHG_PUSH) {
        for (int k = 0; k < X.n_nbr; ++k) {
            const XNbr& nb = X.nb[k];
            long long w = which == 0 ? (long long)nb.n_send_cells * (P.n_ions + 1) + (long long)nb.n_send_flux * P.n_ions
                                     : (long long)nb.cc_rows * P.nx * P.n_ions + (long long)nb.v_rows * P.nx;
            if (w > work) work = w;
        }
    }
    int grid = (int)((work + 2047) / 2048);     // 8 stores per thread: the blocks are a few hundred KB
    if (grid < 1) grid = 1;
    if (grid > 48) grid = 48;
    if (!(mode & XCHG_PUSH)) grid = 1;
    k_xchg<<<grid, 256, 0, st>>>(P, A, X, which, buf, mode, flux_ell);
}
