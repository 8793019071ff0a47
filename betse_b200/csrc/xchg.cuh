// Device-visible plan of the halo exchange between neighbouring strips (include/betse_b200.h,
// "Multi-GPU").  Pointers of a neighbour are addresses inside ITS window, valid in this process
// (CUDA-IPC mapping over NVLink, or the raw pointer when both ranks share a process/device).
#pragma once
#include <stdint.h>

struct XNbr {
    double *cc_mid[2], *vm_cell[2], *flux, *cc_env[2], *v_raw;
    unsigned long long *flags;      // the neighbour's flags, [2 exchange points][2 sides]
    int Cn, En;                     // the neighbour's leading dimensions (cells, env points)
    int side;                       // 0: the neighbour is the strip below, 1: above
    int n_send_cells, recv_cell0;
    int n_send_flux, recv_slot0;
    const int *send_cells, *send_flux;
    const int *send_flux_ell;       // position of the same membranes' fluxes in flux_ell (k_cell), or null
    int cc_rows, cc_src_row0, cc_dst_row0;
    int v_rows, v_src_row0, v_dst_row0;
    int push_lo, push_hi;           // local rows [lo, hi) of X2 that go to this neighbour (union of the cc_env and voltage rows)
};

struct XPlan {
    int n_nbr;
    XNbr nb[2];
    unsigned long long *my_flags;   // [2][2] written by the neighbours
    unsigned long long *epoch;      // [2]    exchanges completed on this rank, per exchange point
    unsigned int *done_ctr;         // [2]    CTAs that finished pushing
    unsigned long long timeout_ns;
    // pushes fused into the producing kernels (k_cell: X1, k_envacc_ell: X2)
    int side_k[2];                  // index into nb[] of the neighbour on side 0 / 1, or -1
    int n_bblocks;                  // k_cell blocks that hold ghost-send cells or remote-flux membranes
    int blk_n0, blk_n1;             // they lie within the first blk_n0 / last blk_n1 blocks, which k_cell runs first
    int env_n0, env_n1, n_push_ctas;   // k_envacc_ell CTAs (256 squares) at the lower / upper end of the owned rows that push
};

#ifdef __CUDACC__
#define ST_XCHG_TIMEOUT 8u
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// The wait half of an exchange point, inside the kernel that consumes the neighbours' values (KParams.xwait): ONE thread
// of the CTA spins (bounded) until every neighbour's flag of exchange point `which` has reached this rank's own epoch —
// the push kernel of the same point ran earlier on this stream and moved the epoch — then the CTA's barrier releases the
// other threads.  Only CTAs that read neighbour-written rows call it; the others start at once, so the wait overlaps the
// interior of the kernel instead of sitting between two kernels.
__device__ __forceinline__ void xchg_wait_cta(const KParams& P, const KArrays& A, const int which)
{
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const unsigned long long e = *reinterpret_cast<const volatile unsigned long long*>(A.xepoch + which);
        const unsigned long long t0 = globaltimer_ns();
        for (int side = 0; side < 2; ++side) {
            if (!((P.xsides >> side) & 1)) continue;
            const unsigned long long* f = A.xflags + which * 2 + side;
            while (ld_acquire_sys(f) < e) {
                if (globaltimer_ns() - t0 > P.x_timeout_ns) { atomicOr(A.status, ST_XCHG_TIMEOUT); break; }
            }
        }
    }
    __syncthreads();
}
// The publish half, called by ONE thread of every CTA / warp that pushed values of exchange point `which` once the stores
// of its whole group are ordered before it (__syncthreads / __syncwarp): the last of `n_groups` to arrive moves the epoch
// and raises this rank's flag in every neighbour's window.
__device__ __forceinline__ void xchg_publish(const XPlan& X, const int which, const int n_groups)
{
    __threadfence_system();
    const unsigned int prev = atomicAdd(X.done_ctr + which, 1u);
    if (prev != (unsigned int)n_groups - 1u) return;
    __threadfence();
    X.done_ctr[which] = 0;
    const unsigned long long e = X.epoch[which] + 1ull;
    X.epoch[which] = e;
    __threadfence_system();
    for (int k = 0; k < X.n_nbr; ++k) st_release_sys(X.nb[k].flags + which * 2 + (1 - X.nb[k].side), e);
}
#endif
