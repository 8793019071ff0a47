// General network / gene regulatory network (the part implemented so far):
//   MasterOfNetworks.run_loop                 betse/science/chemistry/networks.py:2805-2982
//   Molecule.transport -> stb.molecule_mover  networks.py:5670-5700, sim_toolbox.py:909-1153 (gap-junction branch)
//   write_growth_and_decay / write_reactions  networks.py:1187-1572 (the strings; compiled by betse_b200/ratelaw.py)
//
// k_net       one thread per cell: every growth/decay and reaction rate from the OLD concentrations
//             (the reference evals all strings first, networks.py:2826-2853), delta = reaction_matrix . rates,
//             c += delta*dt (networks.py:2856-2914); substances that do not pass gap junctions are checked for
//             negative values here (sim_toolbox.py:1124-1150 raises, networks.py:2945 clamps).
// k_net_gj    substances that pass gap junctions: GHK flux between the two cells of every membrane from the
//             updated concentrations, zero at the cluster boundary (sim_toolbox.py:976-1006), summed per cell;
// k_net_gj_apply  c += dt*delta*time_dilation_factor, negative check.
#include "kparams.cuh"
#include "network.cuh"

#define FLOAT_NONCE 1.0e-25
#define ST_NEG_NET 16u

__global__ void __launch_bounds__(128)
k_net(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int cur)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_cells_owned) return;
    const int C = P.n_cells, M = P.n_mems_owned;
    double r[NET_MAX_RATES];
    for (int j = 0; j < N.n_rates; ++j) {
        double v = rl_eval(N, j, c, -1, A, C, M, cur, 0.0);
        if (j < N.K && N.gmask && !N.gmask[(size_t)j * C + c]) v = 0.0;     // mat[trgs] = rts[trgs], networks.py:2844-2846
        r[j] = v;
        N.rates[(size_t)j * C + c] = v;
    }
    unsigned int flags = 0;
    for (int k = 0; k < N.K; ++k) {
        double d = 0.0;
        for (int j = 0; j < N.n_rates; ++j) d += __ldg(N.stoich + k * N.n_rates + j) * r[j];   // np.dot(reaction_matrix, all_rates)
        double cn = N.c[(size_t)k * C + c] + d * P.dt;                      // networks.py:2914
        if (__ldg(N.Dgj + k) < 0.0 && cn < 0.0) { flags |= ST_NEG_NET; cn = 0.0; }
        N.c[(size_t)k * C + c] = cn;
    }
    if (flags) atomicOr(A.status, flags);
}

// gap-junction transport of substance k: per-cell sum of -f_gj*mem_sa (one warp per tile, the packing of k_mem)
__global__ void __launch_bounds__(BT_TPB)
k_net_gj(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int k, const int cur,
         const int nonces)
{
    __shared__ double s_all[(BT_TPB / 32) * 32];
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (BT_TPB / 32) + (threadIdx.x >> 5);
    if (tile >= P.n_tiles) return;
    double* s_f = s_all + (threadIdx.x >> 5) * 32;
    const int4 td = __ldg(reinterpret_cast<const int4*>(A.tile_desc) + tile);
    const int c0 = td.x, nc = td.y, m0 = td.z, nm = td.w;
    const int C = P.n_cells;
    const double* __restrict__ cc = N.c + (size_t)k * C;
    double fsa = 0.0;
    if (lane < nm) {
        const int m = m0 + lane;
        const int c = __ldg(A.mem_to_cells + m);
        const int nnp = __ldg(A.nn_cell_flag + m);
        const int cn = nnp & 0x7fffffff;
        if (nnp >= 0) {                                                    // fgj_X[cells.bflags_mems] = 0
            const double gjb = A.gj_block ? __ldg(A.gj_block + m) : P.gj_block;
            const double D = (__ldg(N.Dgj + k) * gjb) * A.gjopen[m];          // Dgj*sim.gj_block*sim.gjopen
            double va = A.vm_cell[cur][cn], vb = A.vm_cell[cur][c];
            if (P.polar) { va = A.vm_pol[cur][__ldg(A.nn_i + m)]; vb = A.vm_pol[cur][m]; }
            else if (P.has_phi) { va -= __ldg(A.phi_b_old + __ldg(A.map_mem2ecm + __ldg(A.nn_i + m))); vb -= __ldg(A.phi_b_old + __ldg(A.map_mem2ecm + m)); }
            double vBA = va - vb;                                          // sim.vgj (sim.py:2166) ...
            for (int q = 0; q < nonces; ++q) vBA += FLOAT_NONCE;           // ... after the in-place `vBA += 1e-25` of every earlier electroflux call
            const double zc = __ldg(N.z + k) + FLOAT_NONCE;
            const double alpha = ((zc * vBA) * P.F) / P.RT_p;              // p.T (sim_toolbox.py:986)
            const double ex = exp(-alpha), deno = -expm1(-alpha);
            const double cA = cc[c], cB = cc[cn];                          // cX_mems[mem_i], cX_mems[nn_i]
            const double f = -((D * alpha) / P.gj_len) * ((cB - cA * ex) / deno);
            fsa = -f * __ldg(A.mem_sa + m);
        }
    }
    s_f[lane] = fsa;
    __syncwarp();
    if (lane < nc) {
        const int c = c0 + lane;
        const int jb = __ldg(A.cell_mem_ptr + c) - m0, je = __ldg(A.cell_mem_ptr + c + 1) - m0;
        double S = 0.0;
        for (int j = jb; j < je; ++j) S += s_f[j];
        N.gj_delta[c] = S / __ldg(A.cell_vol + c);                         // delta_cco, sim_toolbox.py:994
    }
}

__global__ void __launch_bounds__(256)
k_net_gj_apply(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int k)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_cells_owned) return;
    double cn = N.c[(size_t)k * P.n_cells + c] + (P.dt * N.gj_delta[c]) * __ldg(N.tdf + k);   // sim_toolbox.py:999
    if (cn < 0.0) { atomicOr(A.status, ST_NEG_NET); cn = 0.0; }
    N.c[(size_t)k * P.n_cells + c] = cn;
}

void launch_net(const KParams& P, const KArrays& A, const KNet& N, const double* h_Dgj, int n_ions, int cur, cudaStream_t st)
{
    if (N.K <= 0) return;
    k_net<<<(P.n_cells_owned + 127) / 128, 128, 0, st>>>(P, A, N, cur);
    const int grid = (P.n_tiles + (BT_TPB / 32) - 1) / (BT_TPB / 32);
    int nonces = n_ions;
    for (int k = 0; k < N.K; ++k) {
        if (h_Dgj[k] < 0.0) continue;
        ++nonces;
        k_net_gj<<<grid, BT_TPB, 0, st>>>(P, A, N, k, cur, nonces);
        k_net_gj_apply<<<(P.n_cells_owned + 255) / 256, 256, 0, st>>>(P, A, N, k);
    }
}
