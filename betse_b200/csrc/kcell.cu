// k_cell — the membranes->cells kernel of the tissue step with ONE LANE PER CELL (same arithmetic as
// kernels.cu:k_mem through the shared functions of kmath.cuh, bit-identical results).
//
// Why (profiles/r01i_*: the lane-per-membrane kernels issue ~28 warp instructions per membrane, two thirds of them
// staging, segmented sums and pipeline bookkeeping, and every membrane lane recomputes what belongs to its CELL):
//   * Vmem is a per-cell quantity in the default mode (sim.py:2029), so the membrane-side GHK table, the pumps'
//     equilibrium constant and every pump factor that depends on the cell's concentrations are formed ONCE per cell;
//   * a lane walks its cell's membranes k = 0..nm-1, so the membranes->cell sums (update_Co, sim_toolbox.py:1177)
//     are plain register accumulations in membrane order — no shared memory staging of fluxes, no shuffles;
//   * per-membrane constants live in a sliced-ELL "cell pack" (SELL-32: block b = cells 32b..32b+31, row k of the
//     block holds membrane k of each of its cells; a row is DmS[I][32], mem_sa[32], partner[32], env square[32]), so
//     lane = cell reads them coalesced, and because neighbouring cells have neighbouring partners and env squares,
//     the gathers of one row touch two or three lines instead of 32.
//
// Two builds:
//   * k_cell_pipe (default): the rows of the pack are ONE stream; every warp owns a contiguous range of blocks, i.e. a
//     contiguous range of rows, and runs a three-stage software pipeline over it with cp.async (LDGSTS) into per-warp
//     shared-memory rings — the row's index pair four rows ahead, everything addressed through it (env concentrations
//     at the membrane's square, the partner cell's concentrations and Vmem, the transported Ca) together with the
//     row's DmS / area / gap-junction state two rows ahead, the next block's per-cell state with its first row.  Every
//     lane copies only what it consumes itself, so cp.async.wait_group is the only synchronisation.  (The register
//     build below spent 46 % of its stall samples on the first use of a gathered value and on the dependent loads of a
//     block's prologue, with two warps per scheduler to hide them: profiles/r02a_*.)  Block boundaries travel with the
//     data: bits 29/30 of a row's env-square word mark the last / first row of a block, bit 31 a lane without membrane.
//   * k_cell (register build; meshes with a one-membrane block, ion profiles whose rings do not fit): persistent warps
//     draw blocks in order, the gathers of membrane k+1 load into a second register buffer while membrane k is computed.
//
// Reference lines as in k_mem: sim.py:1193-1283, 2086-2111, 2162-2206; sim_toolbox.py:18-182, 1155-1207;
// channels/gap_junction.py:53-77; ion_current.py:19; sim.py:2027-2029.
#include <stdlib.h>
#include <algorithm>
#include <stdint.h>
#include "kmath.cuh"
#include "xchg.cuh"

#define KC_WARPS 4
#define KC_ROWB(NI) (((NI) + 2) * 256)          // bytes of one row of the cell pack
#define KC_INV   0x80000000u                     // env-square word: the lane has no membrane in this row
#define KC_FIRST 0x40000000u                     //                  first row of its block
#define KC_LAST  0x20000000u                     //                  last row of its block
#define KC_QMASK 0x1fffffffu
#define KC_DEPS_SMEM (116 * 1024)

template <int NI>
struct MemIn { double co[NI], cnb[NI], vnb, cao; int nnp; };

__device__ __forceinline__ uint32_t kc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kc_cp8(uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void kc_cp4(uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void kc_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void kc_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// hint: pull [p, p + bytes) into L2 (one bulk prefetch, no registers held)
__device__ __forceinline__ void l2_prefetch(const void* p, unsigned bytes)
{
    if (bytes == 0) return;
    const unsigned long long a = (unsigned long long)p;
    const unsigned long long a0 = a & ~15ull;
    const unsigned n = (unsigned)(((a + bytes + 15ull) & ~15ull) - a0);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(n) : "memory");
}

// ---- the flux math of one membrane (lane = cell), shared by both builds.  Inputs: the cell-side terms, the membrane's
//      gathered values; adds to the cell's sums, stores the membrane -> env fluxes and the new gap-junction state.
template <int NI>
struct CellSide {
    double cin[NI], cinAm[NI], Bm[NI], Sm[NI], Sg[NI];
    double vm_own, keq;
    NaKCell nkc;
    CaCell cac;
    bool ca_on;
};

template <int NI>
__device__ __forceinline__ void cell_prologue(const KParams& P, CellSide<NI>& S, const double vm_own, const double cCa_fresh, unsigned int& flags)
{
    constexpr int iNa = StdProf<NI>::iNa, iK = StdProf<NI>::iK, iCa = StdProf<NI>::iCa;
    S.vm_own = vm_own;
    MemSide ms;
    mem_side(vm_own, P, ms);
    S.keq = ms.keq;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        double Am;
        ghk_pick(ms.t, StdProf<NI>::z(i), Am, S.Bm[i]);
        S.cinAm[i] = __dmul_rn(S.cin[i], Am);
        S.Sm[i] = 0.0; S.Sg[i] = 0.0;
    }
    nak_cell(S.cin[iNa], S.cin[iK], P, S.nkc);
    S.cac.g1 = S.cac.g2 = 0.0;
    S.ca_on = iCa >= 0 && P.alpha_Ca > 0.0;
    if (S.ca_on) {
        double cCai = cCa_fresh;                       // the fresh cell value (update_intra, sim.py:2310)
        if (cCai != cCai) flags |= ST_NAN_CONC;
        if (cCai < 0.0) cCai = 0.0;
        ca_cell(cCai, ms.keq, P, S.cac);
    }
}

template <int NI>
__device__ __forceinline__ double membrane_fluxes(const KParams& P, CellSide<NI>& S, const double* __restrict__ co, const double* __restrict__ cnb,
                                                  const double vnb, const double cao, const int nnp, const double* __restrict__ DmS,
                                                  const double sa, double g, double* __restrict__ fl, unsigned int& flags,
                                                  double* __restrict__ frem = nullptr)
{
    constexpr int iNa = StdProf<NI>::iNa, iK = StdProf<NI>::iK, iCa = StdProf<NI>::iCa;
    // gap-junction side: vgj and its GHK table with p.T (sim.py:2166, 2197), gating sub-step g' = g*gc1 + gc2
    const double vgj0 = vnb - S.vm_own;
    const double ag1 = ((vgj0 + FLOAT_NONCE) * P.F) * P.inv_RT_p;
    GhkAB tg;
    ghk_table(ag1, tg);
    double gc1, gc2;
    gj_gate_map(vgj0, P, P.gj_block, gc1, gc2);
    const double sa_g = (nnp < 0) ? 0.0 : sa;     // no gap-junction flux at boundary membranes (sim.py:2199-2201)
    double fNa = 0.0, fK = 0.0;
    if (P.alpha_NaK > 0.0) {
        fNa = nak_flux(S.nkc, S.keq, co[iNa], co[iK], P.NaK_block, P);
        fK = -(2.0 / 3.0) * fNa;
        fNa = P.rho_pump * fNa;
        fK = P.rho_pump * fK;
    }
    double fCa = 0.0;
    if (S.ca_on) {
        double cCao = cao;
        if (cCao != cCao) flags |= ST_NAN_CONC;
        if (cCao < 0.0) cCao = 0.0;
        fCa = ca_flux(S.cac, cCao, P);
        fCa = P.rho_pump * fCa;
        fCa = P.rho_pump * fCa;                 // applied twice in the reference (sim.py:2141, 2155)
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        double Ag, Bg;
        ghk_pick(tg, StdProf<NI>::z(i), Ag, Bg);
        double fsa = ghk_mem_flux(DmS[i], S.cinAm[i], co[i], S.Bm[i]);
        if (i == iNa) fsa = fma(fNa, sa, fsa);
        if (i == iK) fsa = fma(fK, sa, fsa);
        if (i == iCa) fsa = fma(fCa, sa, fsa);
        g = fma(g, gc1, gc2);                                      // once per ion (sim.py:1272 -> 2180-2183)
        const double fg = ghk_gj_flux(P.Dgj_len[i], __dmul_rn(g, sa_g), cnb[i], Ag, S.cin[i], Bg);
        S.Sm[i] = __dadd_rn(S.Sm[i], fsa);
        S.Sg[i] = __dadd_rn(S.Sg[i], fg);
        fl[i * 32] = fsa;
        if (frem) frem[i] = fsa;          // the membrane's env square belongs to a neighbouring strip: its slot there, [slot][ion]
    }
    return g;
}

// update_Co + update_all_concs, charge and Vmem of the cell (sim_toolbox.py:1177-1181; sim.py:2105-2111;
// ion_current.py:19; sim.py:2027-2029)
template <int NI>
__device__ __forceinline__ void cell_epilogue(const KParams& P, const KArrays& A, const CellSide<NI>& S, const double* cc, const double vol,
                                              const double dvt, const int c, const int nxt, unsigned int& flags,
                                              const XPlan* X = nullptr, const int2 gs = make_int2(-1, -1))
{
    const int C = P.n_cells;
    const double rvol = fast_rcp(vol);
    double rho = 0.0;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        double cm_new, cn_new;
        cell_conc_update(cc[i], S.Sm[i], S.Sg[i], rvol, P.dt, cm_new, cn_new);
        if (cn_new != cn_new) flags |= ST_NAN_CONC;
        if (cn_new < 0.0) { cn_new = 0.0; flags |= ST_NEG; }         // no_negs, sim.py:2111
        A.cc_cells[(size_t)i * C + c] = cn_new;
        A.cc_mid[nxt][(size_t)i * C + c] = cm_new;                   // the stale cc_at_mem (quirk list)
        if (X) {
            // this cell is a ghost on a neighbouring strip: exchange point X1, stored straight into its window
            if (gs.x >= 0) { const XNbr& nb = X->nb[X->side_k[0]]; nb.cc_mid[nxt][(size_t)i * nb.Cn + gs.x] = cm_new; }
            if (gs.y >= 0) { const XNbr& nb = X->nb[X->side_k[1]]; nb.cc_mid[nxt][(size_t)i * nb.Cn + gs.y] = cm_new; }
        }
        rho = fma(P.zF[i], cn_new, rho);
    }
    if (A.extra_rho_cells) rho += ldg(A.extra_rho_cells + c);
    A.rho_cells[c] = rho;
    const double vmn = P.inv_cm * (rho * dvt);
    if (vmn != vmn) flags |= ST_NAN_VM;
    A.vm_cell[nxt][c] = vmn;
    if (X) {
        if (gs.x >= 0) X->nb[X->side_k[0]].vm_cell[nxt][gs.x] = vmn;
        if (gs.y >= 0) X->nb[X->side_k[1]].vm_cell[nxt][gs.y] = vmn;
    }
}

// ---------------------------------------------------------------------------- register build
// one block of 32 cells: lane = cell
template <int NI>
__device__ __forceinline__ void cell_task(const KParams& P, const KArrays& A, const int cur, const int task, const int lane, unsigned int& flags,
                                          const XPlan* X = nullptr)
{
    constexpr int iCa = StdProf<NI>::iCa;
    constexpr int ROWB = KC_ROWB(NI);
    const int nxt = cur ^ 1;
    const int C = P.n_cells, E = P.ny * P.nx;
    const int2 h0 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + task);          // {first row, first membrane}
    const int2 h1 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + task + 1);
    const int row0 = h0.x, Kb = h1.x - h0.x;
    const int c = task * 32 + lane;
    const bool valid = c < P.n_cells_owned;
    int m_beg = 0, nm = 0;
    if (valid) { m_beg = ldgi(A.cell_mem_ptr + c); nm = ldgi(A.cell_mem_ptr + c + 1) - m_beg; }
    const char* __restrict__ rows = A.cpack + (size_t)row0 * ROWB;
    const double* __restrict__ gjs = A.gjopen + m_beg;
    const double* __restrict__ cmid = A.cc_mid[cur];
    const double* __restrict__ vmc = A.vm_cell[cur];
    const double* __restrict__ cenv = A.cc_env[cur];
    const double* __restrict__ cenvCa = A.cc_env[nxt] + (size_t)(iCa >= 0 ? iCa : 0) * E;   // Ca after transport (sim.py:1282 after 2254)

    // strips with fused pushes: does this block hold cells that are ghosts on a neighbour / membranes of its env squares?
    int2 bx = make_int2(-1, -1), gs = make_int2(-1, -1);
    if (X) bx = __ldg(reinterpret_cast<const int2*>(A.blk_x) + task);
    // the block whose streams this task pulls into L2 (issued after the first membrane, below)
    const int up = task + P.pf_dist;
    int2 u0 = make_int2(0, 0), u1 = make_int2(0, 0);
    if (P.pf_dist > 0 && up < P.n_blocks) {
        u0 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + up);
        u1 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + up + 1);
    }

    // the index pair of membrane k is loaded two membranes ahead (ia / ib), so that the gathers through them do not
    // wait for a second round trip
    int2 ia = make_int2(0, 0), ib = make_int2(0, 0);
    auto idx_load = [&](int2& x, const int k) {
        if (k < nm) {
            const char* r = rows + (size_t)k * ROWB;
            x.x = __ldcs(reinterpret_cast<const int*>(r + (NI + 1) * 256) + lane);
            x.y = __ldcs(reinterpret_cast<const int*>(r + (NI + 1) * 256 + 128) + lane);
        }
    };
    auto gather = [&](MemIn<NI>& x, const int2& ix, const int k) {
        if (k < nm) {
            const unsigned q = (unsigned)ix.y & KC_QMASK;
            const unsigned cn = (unsigned)(ix.x & 0x7fffffff);
#pragma unroll
            for (int i = 0; i < NI; ++i) x.co[i] = (cenv + (size_t)i * E)[q];
#pragma unroll
            for (int i = 0; i < NI; ++i) x.cnb[i] = (cmid + (size_t)i * C)[cn];
            x.vnb = vmc[cn];
            x.cao = (iCa >= 0) ? cenvCa[q] : 0.0;
            x.nnp = ix.x;
        }
    };

    CellSide<NI> S;
    double cc[NI], vm_own = 0.0, vol = 1.0, dvt = 0.0;
#pragma unroll
    for (int i = 0; i < NI; ++i) { cc[i] = 0.0; S.cin[i] = 0.0; }
    MemIn<NI> a, b;
    idx_load(ia, 0);
    idx_load(ib, 1);
    gather(a, ia, 0);
    if (valid) {
        vm_own = vmc[c];
#pragma unroll
        for (int i = 0; i < NI; ++i) S.cin[i] = (cmid + (size_t)i * C)[c];
#pragma unroll
        for (int i = 0; i < NI; ++i) cc[i] = A.cc_cells[(size_t)i * C + c];
        vol = ldg(A.cell_vol + c);
        dvt = ldg(A.diviterm + c);
    }
    cell_prologue<NI>(P, S, vm_own, cc[iCa >= 0 ? iCa : 0], flags);
    if (bx.x >= 0) gs = __ldg(reinterpret_cast<const int2*>(A.ghost_tab) + bx.x + lane);

    auto compute = [&](const MemIn<NI>& x, const int k) {
        if (k < nm) {
            const double* __restrict__ r = reinterpret_cast<const double*>(rows + (size_t)k * ROWB) + lane;
            double DmS[NI];
#pragma unroll
            for (int i = 0; i < NI; ++i) DmS[i] = __ldcs(r + i * 32);
            const double sa = __ldcs(r + NI * 32);
            double* __restrict__ fl = A.flux_ell + ((size_t)(row0 + k) * NI) * 32 + lane;
            double* frem = nullptr;
            if (bx.y >= 0) {
                const int rs = ldgi(A.rslot_tab + bx.y + k * 32 + lane);
                if (rs >= 0) frem = X->nb[X->side_k[rs >> 30]].flux + (size_t)(rs & 0x3fffffff) * NI;
            }
            A.gjopen[m_beg + k] = membrane_fluxes<NI>(P, S, x.co, x.cnb, x.vnb, x.cao, x.nnp, DmS, sa, gjs[k], fl, flags, frem);
        }
    };

    // ---- the cell's membranes, two register buffers: membrane k is computed while the gathers of k+1 load
#pragma unroll 1
    for (int k = 0; k < Kb; k += 2) {
        gather(b, ib, k + 1);
        idx_load(ia, k + 2);
        compute(a, k);
        if (k == 0 && u1.x > u0.x) {
            // streams of block `up`, one bulk prefetch per array and lane: its rows, gjopen, the cells' own state
            const int cu = up * 32;
            const int ncu = min(32, P.n_cells_owned - cu);
            if (lane == 0) l2_prefetch(A.cpack + (size_t)u0.x * ROWB, (unsigned)(u1.x - u0.x) * ROWB);
            else if (lane == 1) l2_prefetch(A.gjopen + u0.y, (unsigned)(u1.y - u0.y) * 8u);
            else if (lane < NI + 2) l2_prefetch(A.cc_cells + (size_t)(lane - 2) * C + cu, ncu * 8u);
            else if (lane < 2 * NI + 2) l2_prefetch(cmid + (size_t)(lane - NI - 2) * C + cu, ncu * 8u);
            else if (lane == 2 * NI + 2) l2_prefetch(vmc + cu, ncu * 8u);
            else if (lane == 2 * NI + 3) l2_prefetch(A.cell_vol + cu, ncu * 8u);
            else if (lane == 2 * NI + 4) l2_prefetch(A.diviterm + cu, ncu * 8u);
            else if (lane == 2 * NI + 5) l2_prefetch(A.cell_mem_ptr + cu, (ncu + 1) * 4u);
        }
        if (k + 1 < Kb) {
            gather(a, ia, k + 2);
            idx_load(ib, k + 3);
            compute(b, k + 1);
        }
    }
    if (valid) cell_epilogue<NI>(P, A, S, cc, vol, dvt, c, nxt, flags, (bx.x >= 0) ? X : nullptr, gs);
    if (X && (bx.x >= 0 || bx.y >= 0)) {
        // this block pushed: the last of the boundary blocks raises this rank's X1 flag on the neighbours
        __syncwarp();
        if (lane == 0) xchg_publish(*X, 0, X->n_bblocks);
    }
}

// Persistent: every warp draws tickets (ticket t = block t of the cell pack, so the blocks finish as a wavefront through
// the tissue); kc_persist = 0: one block per warp.  After a block, its group's completion counter moves (release): the
// env accumulation kernel running NEXT TO this one (k_envacc_ell, `deps`) consumes a block's fluxes out of L2 as soon as
// every block that feeds its env squares has finished.
template <int NI>
__device__ __forceinline__ void k_cell_body(const KParams& P, const KArrays& A, const int cur, const XPlan* X)
{
    const int lane = threadIdx.x & 31;
    unsigned int flags = 0;
    for (;;) {
        int t = 0;
        if (P.kc_persist) {
            if (lane == 0) t = atomicAdd(A.ticket, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
        } else t = (int)(blockIdx.x * KC_WARPS + (threadIdx.x >> 5));
        if (t >= P.n_blocks) break;
        // strips: the blocks at the two ends of the strip first (cells are numbered row by row, the strip's edge cells sit
        // in its first and last blocks), so that the neighbours have their values long before they need them
        if (X) {
            const int n0 = X->blk_n0, n1 = X->blk_n1;
            if (t >= n0) t = (t < n0 + n1) ? P.n_blocks - n1 + (t - n0) : n0 + (t - n0 - n1);
        }
        cell_task<NI>(P, A, cur, t, lane, flags, X);
        if (A.cell_done) {
            // a RELEASE on the counter, not __threadfence(): a gpu-scope acq_rel fence makes ptxas invalidate the SM's
            // whole L1 (CCTL.IVALL), which the other warps' gathers live on
            __syncwarp();
            if (lane == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(A.cell_done + t / KC_GRP) : "memory");
        }
        if (!P.kc_persist) break;
    }
    if (flags) atomicOr(A.status, flags);
}

// (more resident warps at fewer registers were measured and lost: 255 registers / 8 warps per SM 0.258 ms, 200 / 10 warps
// 0.278 ms, 184 / 11 warps 0.305 ms, 168 / 12 warps 0.31 ms — the spills cost more than the occupancy buys;
// profiles/r02j_sweep_kcell_regs.txt)
// REGS = 0: all 255 registers (two CTAs fill the register file); REGS = 1: capped at 208, which leaves 12288 registers per
// SM — one 256-thread CTA of the env kernels (k_ion, k_envacc_ell: <= 48 registers) runs next to the two k_cell CTAs
template <int NI, int MINB>
__global__ void __launch_bounds__(KC_WARPS * 32, MINB)
k_cell(const __grid_constant__ KParams P, const KArrays A, const int cur) { k_cell_body<NI>(P, A, cur, nullptr); }

// decomposed tissue: the same kernel with exchange point X1 inside it (xchg.cuh)
template <int NI>
__global__ void __launch_bounds__(KC_WARPS * 32, 2)
k_cell_x(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ XPlan X, const int cur) { k_cell_body<NI>(P, A, cur, &X); }

template <int NI>
__global__ void __maxnreg__(208)
k_cell_share(const __grid_constant__ KParams P, const KArrays A, const int cur) { k_cell_body<NI>(P, A, cur, nullptr); }

// ---------------------------------------------------------------------------- pipelined build
template <int NI>
struct KcPipe {
    static constexpr int IDX_SLOTS = 8, BULK_SLOTS = 3, CELL_SLOTS = 2, MB_SLOTS = 4;
    static constexpr int BULK_D = 3 * NI + 4;                 // co[NI], cnb[NI], DmS[NI], vnb, cao, sa, gj: doubles per lane
    static constexpr int CELL_D = 2 * NI + 3;                 // cin[NI], cc[NI], vm, vol, diviterm
    static constexpr int O_IDX = 0;
    static constexpr int O_MB = O_IDX + IDX_SLOTS * 256;      // first membrane of the lane's cell, per block in flight
    static constexpr int O_BULK = O_MB + MB_SLOTS * 128;
    static constexpr int O_CELL = O_BULK + BULK_SLOTS * BULK_D * 256;
    static constexpr int WARP_B = O_CELL + CELL_SLOTS * CELL_D * 256;
    static constexpr int CTA_B = KC_WARPS * WARP_B;
};

template <int NI>
__global__ void __launch_bounds__(KC_WARPS * 32, 2)
k_cell_pipe(const __grid_constant__ KParams P, const KArrays A, const int cur)
{
    using L = KcPipe<NI>;
    extern __shared__ __align__(128) char kc_sm[];
    constexpr int iCa = StdProf<NI>::iCa;
    constexpr int ROWB = KC_ROWB(NI);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nxt = cur ^ 1;
    const int C = P.n_cells, E = P.ny * P.nx;
    // this warp's blocks [b0, b1) = rows [r0, r1) of the pack
    const long long gw = (long long)blockIdx.x * KC_WARPS + warp, W = (long long)gridDim.x * KC_WARPS;
    const int b0 = (int)(gw * P.n_blocks / W), b1 = (int)((gw + 1) * P.n_blocks / W);
    if (b0 >= b1) return;
    const int r0 = ldgi(A.blk_row0 + 2 * b0), r1 = ldgi(A.blk_row0 + 2 * b1);
    char* const base = kc_sm + (size_t)warp * L::WARP_B;
    const uint32_t sb = kc_smem_u32(base);
    const double* __restrict__ cmid = A.cc_mid[cur];
    const double* __restrict__ vmc = A.vm_cell[cur];
    const double* __restrict__ cenv = A.cc_env[cur];
    const double* __restrict__ cenvCa = A.cc_env[nxt] + (size_t)(iCa >= 0 ? iCa : 0) * E;   // Ca after transport (sim.py:1282 after 2254)

    // ---- stage 1: index pair of row r (and, with the first row of a block, the first membrane of the lane's cell)
    int bp1 = b0 - 1, next1 = r0;                    // block of the stage-1 cursor, first row of the block after it
    auto stage1 = [&](const int r) {
        if (r < r1) {
            if (r == next1) {
                ++bp1;
                next1 = ldgi(A.blk_row0 + 2 * (bp1 + 1));
                const int c = bp1 * 32 + lane;
                if (c < P.n_cells_owned) kc_cp4(sb + L::O_MB + (bp1 & (L::MB_SLOTS - 1)) * 128 + lane * 4, A.cell_mem_ptr + c);
            }
            const char* row = A.cpack + (size_t)r * ROWB + (NI + 1) * 256;
            const uint32_t d = sb + L::O_IDX + (r & (L::IDX_SLOTS - 1)) * 256 + lane * 4;
            kc_cp4(d, reinterpret_cast<const int*>(row) + lane);
            kc_cp4(d + 128, reinterpret_cast<const int*>(row + 128) + lane);
        }
        kc_commit();
    };
    // ---- stage 2: everything of row r that is addressed through its index pair, the row's constants, the gap-junction
    //      state; with the first row of a block the per-cell state of the block
    int bp2 = b0 - 1, k2 = 0, mb2 = 0;
    auto stage2 = [&](const int r) {
        if (r < r1) {
            const int* ix = reinterpret_cast<const int*>(base + L::O_IDX + (r & (L::IDX_SLOTS - 1)) * 256);
            const int nnp = ix[lane];
            const unsigned ew = (unsigned)ix[32 + lane];
            if (ew & KC_FIRST) {
                ++bp2; k2 = 0;
                const int c = bp2 * 32 + lane;
                mb2 = reinterpret_cast<const int*>(base + L::O_MB + (bp2 & (L::MB_SLOTS - 1)) * 128)[lane];
                if (c < P.n_cells_owned) {
                    const uint32_t d = sb + L::O_CELL + (bp2 & 1) * (L::CELL_D * 256) + lane * 8;
#pragma unroll
                    for (int i = 0; i < NI; ++i) kc_cp8(d + i * 256, cmid + (size_t)i * C + c);
#pragma unroll
                    for (int i = 0; i < NI; ++i) kc_cp8(d + (NI + i) * 256, A.cc_cells + (size_t)i * C + c);
                    kc_cp8(d + (2 * NI) * 256, vmc + c);
                    kc_cp8(d + (2 * NI + 1) * 256, A.cell_vol + c);
                    kc_cp8(d + (2 * NI + 2) * 256, A.diviterm + c);
                }
            }
            if (!(ew & KC_INV)) {
                const unsigned q = ew & KC_QMASK;
                const unsigned cn = (unsigned)nnp & 0x7fffffffu;
                const uint32_t d = sb + L::O_BULK + (r % L::BULK_SLOTS) * (L::BULK_D * 256) + lane * 8;
                const double* row = reinterpret_cast<const double*>(A.cpack + (size_t)r * ROWB) + lane;
#pragma unroll
                for (int i = 0; i < NI; ++i) kc_cp8(d + i * 256, cenv + (size_t)i * E + q);
#pragma unroll
                for (int i = 0; i < NI; ++i) kc_cp8(d + (NI + i) * 256, cmid + (size_t)i * C + cn);
#pragma unroll
                for (int i = 0; i < NI; ++i) kc_cp8(d + (2 * NI + i) * 256, row + i * 32);
                kc_cp8(d + (3 * NI) * 256, vmc + cn);
                if (iCa >= 0) kc_cp8(d + (3 * NI + 1) * 256, cenvCa + q);
                kc_cp8(d + (3 * NI + 2) * 256, row + NI * 32);
                kc_cp8(d + (3 * NI + 3) * 256, A.gjopen + mb2 + k2);
            }
            ++k2;
        }
        kc_commit();
    };

    // ---- fill the pipeline: index pairs of rows r0..r0+3, then rows r0 and r0+1 (commit groups arranged like the steady
    //      state's: stage 1, stage 2, stage 1, stage 2)
    stage1(r0); stage1(r0 + 1); stage1(r0 + 2); stage1(r0 + 3);
    kc_wait<0>();
    kc_commit();
    stage2(r0);
    kc_commit();
    stage2(r0 + 1);

    unsigned int flags = 0;
    CellSide<NI> S;
    double vol = 1.0, dvt = 0.0;
    int bc = b0 - 1, kc = 0, mbc = 0;
    bool valid = false;
#pragma unroll 1
    for (int r = r0; r < r1; ++r) {
        stage1(r + 4);
        kc_wait<4>();                       // index pair of row r+2 has landed
        stage2(r + 2);
        kc_wait<4>();                       // row r has landed
        const int* ix = reinterpret_cast<const int*>(base + L::O_IDX + (r & (L::IDX_SLOTS - 1)) * 256);
        const int nnp = ix[lane];
        const unsigned ew = (unsigned)ix[32 + lane];
        if (ew & KC_FIRST) {
            ++bc; kc = 0;
            const int c = bc * 32 + lane;
            valid = c < P.n_cells_owned;
            mbc = reinterpret_cast<const int*>(base + L::O_MB + (bc & (L::MB_SLOTS - 1)) * 128)[lane];
            const double* cs = reinterpret_cast<const double*>(base + L::O_CELL + (bc & 1) * (L::CELL_D * 256)) + lane;
            double vm_own = 0.0, cCa = 0.0;
#pragma unroll
            for (int i = 0; i < NI; ++i) S.cin[i] = valid ? cs[i * 32] : 0.0;
            if (valid) {
                vm_own = cs[(2 * NI) * 32];
                vol = cs[(2 * NI + 1) * 32];
                dvt = cs[(2 * NI + 2) * 32];
                cCa = cs[(NI + (iCa >= 0 ? iCa : 0)) * 32];
            }
            cell_prologue<NI>(P, S, vm_own, cCa, flags);
        }
        if (!(ew & KC_INV)) {
            const double* bs = reinterpret_cast<const double*>(base + L::O_BULK + (r % L::BULK_SLOTS) * (L::BULK_D * 256)) + lane;
            double co[NI], cnb[NI], DmS[NI];
#pragma unroll
            for (int i = 0; i < NI; ++i) { co[i] = bs[i * 32]; cnb[i] = bs[(NI + i) * 32]; DmS[i] = bs[(2 * NI + i) * 32]; }
            const double vnb = bs[(3 * NI) * 32];
            const double cao = (iCa >= 0) ? bs[(3 * NI + 1) * 32] : 0.0;
            const double sa = bs[(3 * NI + 2) * 32];
            const double g = bs[(3 * NI + 3) * 32];
            double* __restrict__ fl = A.flux_ell + ((size_t)r * NI) * 32 + lane;
            A.gjopen[mbc + kc] = membrane_fluxes<NI>(P, S, co, cnb, vnb, cao, nnp, DmS, sa, g, fl, flags);
        }
        ++kc;
        if ((ew & KC_LAST) && valid) {
            const int c = bc * 32 + lane;
            const double* cs = reinterpret_cast<const double*>(base + L::O_CELL + (bc & 1) * (L::CELL_D * 256)) + lane;
            double cc[NI];
#pragma unroll
            for (int i = 0; i < NI; ++i) cc[i] = cs[(NI + i) * 32];
            cell_epilogue<NI>(P, A, S, cc, vol, dvt, c, nxt, flags);
        }
    }
    kc_wait<0>();
    if (flags) atomicOr(A.status, flags);
}

static int kc_env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// ---------------------------------------------------------------------------- cell pack
// constant part: membrane areas and index rows in SELL-32 order, and the flux position of every membrane
__global__ void k_pack_cell_const(const __grid_constant__ KParams P, const KArrays A, int* __restrict__ mem_ell)
{
    const int task = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (task >= P.n_blocks) return;
    const int row0 = A.blk_row0[2 * task], Kb = A.blk_row0[2 * task + 2] - row0;
    const int c = task * 32 + lane;
    int m_beg = 0, nm = 0;
    if (c < P.n_cells_owned) { m_beg = A.cell_mem_ptr[c]; nm = A.cell_mem_ptr[c + 1] - m_beg; }
    const int ni = P.n_ions;
    const size_t rowb = (size_t)(ni + 2) * 256;
    char* pack = const_cast<char*>(A.cpack);
    for (int k = 0; k < Kb; ++k) {
        char* r = pack + (size_t)(row0 + k) * rowb;
        double sa = 0.0;
        int nnp = (int)0x80000000;
        unsigned esq = KC_INV;
        if (k < nm) {
            const int m = m_beg + k;
            sa = A.mem_sa[m]; nnp = A.nn_cell_flag[m]; esq = (unsigned)A.map_mem2ecm[m];
            mem_ell[m] = (int)(((size_t)(row0 + k) * ni) * 32 + lane);
        }
        if (k == 0) esq |= KC_FIRST;            // block boundaries travel with the rows (k_cell_pipe)
        if (k == Kb - 1) esq |= KC_LAST;
        reinterpret_cast<double*>(r + (size_t)ni * 256)[lane] = sa;
        reinterpret_cast<int*>(r + (size_t)(ni + 1) * 256)[lane] = nnp;
        reinterpret_cast<int*>(r + (size_t)(ni + 1) * 256 + 128)[lane] = (int)esq;
    }
}

// DmS rows = (Dm * -(rho_channel/tm)) * mem_sa from the canonical [ion][membrane] array: after an upload of Dm_cells
// and after betse_set_schedule (rho_channel)
__global__ void k_pack_cell_dm(const __grid_constant__ KParams P, const KArrays A)
{
    const int task = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (task >= P.n_blocks) return;
    const int row0 = A.blk_row0[2 * task], Kb = A.blk_row0[2 * task + 2] - row0;
    const int c = task * 32 + lane;
    int m_beg = 0, nm = 0;
    if (c < P.n_cells_owned) { m_beg = A.cell_mem_ptr[c]; nm = A.cell_mem_ptr[c + 1] - m_beg; }
    const int ni = P.n_ions;
    const size_t rowb = (size_t)(ni + 2) * 256;
    char* pack = const_cast<char*>(A.cpack);
    const double Dtm = -(P.inv_tm * P.rho_channel);
    for (int k = 0; k < Kb; ++k) {
        double* r = reinterpret_cast<double*>(pack + (size_t)(row0 + k) * rowb) + lane;
        for (int i = 0; i < ni; ++i) {
            double v = 0.0;
            if (k < nm) {
                const int m = m_beg + k;
                v = __dmul_rn(__dmul_rn(A.Dm[(size_t)i * P.n_mems_owned + m], Dtm), A.mem_sa[m]);
            }
            r[i * 32] = v;
        }
    }
}

// flux position of every env-square slot: >= 0 local membrane (stride 32 between ions, flux_ell), < 0 remote slot
// s >= n_mems_owned written by a neighbouring rank into the window's [slot][ion] array: -(s*ni) - 1
__global__ void k_slot_off(const int* __restrict__ slot_idx, const int* __restrict__ mem_ell, int* __restrict__ slot_off,
                           const int n, const int Mo, const int ni)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int s = slot_idx[j];
    slot_off[j] = (s < Mo) ? mem_ell[s] : -(s * ni) - 1;
}

__global__ void k_gather_int(int* __restrict__ dst, const int* __restrict__ src, const int* __restrict__ idx, const int n)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dst[j] = src[idx[j]];
}

void launch_gather_int(int* dst, const int* src, const int* idx, int n, cudaStream_t st)
{
    if (n > 0) k_gather_int<<<(n + 255) / 256, 256, 0, st>>>(dst, src, idx, n);
}

void launch_pack_cell_const(const KParams& P, const KArrays& A, int* mem_ell, cudaStream_t st)
{
    if (!A.cpack || P.n_blocks <= 0) return;
    k_pack_cell_const<<<(P.n_blocks + 7) / 8, 256, 0, st>>>(P, A, mem_ell);
}

void launch_pack_cell_dm(const KParams& P, const KArrays& A, cudaStream_t st)
{
    if (!A.cpack || P.n_blocks <= 0) return;
    k_pack_cell_dm<<<(P.n_blocks + 7) / 8, 256, 0, st>>>(P, A);
}

void launch_slot_off(const int* slot_idx, const int* mem_ell, int* slot_off, int n, int Mo, int ni, cudaStream_t st)
{
    if (n <= 0) return;
    k_slot_off<<<(n + 255) / 256, 256, 0, st>>>(slot_idx, mem_ell, slot_off, n, Mo, ni);
}

size_t cell_pack_row_bytes(int ni) { return (size_t)(ni + 2) * 256; }

// ---------------------------------------------------------------------------- membrane -> env exchange (ELL fluxes)
// update_Co env branch + div_env (sim_toolbox.py:1189-1234), env charge, raw env voltage (ion_current.py:75-97):
// kernels.cu:k_envacc with the fluxes read where k_cell left them.  Same summation order (global membrane index).
// DEPS: the kernel runs NEXT TO k_cell (second stream): CTA b first waits until every group of cell blocks that feeds its
// squares has finished (env_dep[b] = {first, last} group; CTAs are dispatched in order and the cell blocks finish in order,
// so a CTA hardly waits and the fluxes it reads were written microseconds ago — they come out of L2, not DRAM).
// XP (decomposed tissue, xchg.cuh): exchange point X2 inside the kernel — the CTAs at the two ends of the owned rows run
// first (blockIdx is permuted), store their rows of cc_env and of the raw env voltage straight into the neighbours' halo
// rows as well, and the last of them raises this rank's X2 flag there.
template <int NI, bool DEPS, bool XP>
__device__ __forceinline__ void envacc_ell_body(const KParams& P, const KArrays& A, const int nxt, const XPlan* X)
{
    int b = blockIdx.x;
    if (XP) {
        // [lower end | upper end | interior]
        const int g = gridDim.x, n0 = X->env_n0, n1 = X->env_n1;
        if (b >= n0) b = (b < n0 + n1) ? g - n1 + (b - n0) : n0 + (b - n0 - n1);
    }
    const int k = P.ya0 * P.nx + b * blockDim.x + threadIdx.x;
    const int E = P.nx * P.ny;
    const bool in = k < P.ya1 * P.nx;
    // everything that does not depend on the fluxes first
    int s0 = 0, s1 = 0;
    double acc[NI], cv[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) { acc[i] = 0.0; cv[i] = 0.0; }
    if (in) {
        s0 = ldgi(A.slot_ptr + k); s1 = ldgi(A.slot_ptr + k + 1);
#pragma unroll
        for (int i = 0; i < NI; ++i) cv[i] = A.cc_env[nxt][(size_t)i * E + k];
    }
    if (XP) {
        // every CTA that pushes waits for the neighbours' X1 first: it tells that they are past the kernels that still read
        // (field of the last step) or write (their own redundant transport of the halo rows) what this CTA is about to
        // store into their windows; the CTAs whose squares take remote flux slots are among them
        if (blockIdx.x < X->env_n0 + X->env_n1) xchg_wait_cta(P, A, 0);
    } else if (!DEPS && P.xwait) {
        // decomposed tissue: squares near the strip edges take fluxes the neighbours pushed (exchange point X1)
        const int k0 = P.ya0 * P.nx + b * blockDim.x;
        if (k0 / P.nx < P.xw_lo || (k0 + (int)blockDim.x - 1) / P.nx >= P.xw_hi) xchg_wait_cta(P, A, 0);
    }
    if (DEPS) {
        if (threadIdx.x == 0) {
            const int2 dep = __ldg(reinterpret_cast<const int2*>(A.env_dep) + blockIdx.x);
            for (int g = dep.x; g <= dep.y; ++g) {
                const int want = min(KC_GRP, P.n_blocks - g * KC_GRP);
                int got;
                do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(got) : "l"(A.cell_done + g) : "memory");
                    if (got < want) __nanosleep(200);
                } while (got < want);
            }
        }
        __syncthreads();
    }
    if (in) {
        for (int j = s0; j < s1; ++j) {
            const int off = ldgi(A.slot_off + j);
            if (off >= 0) {
                // DEPS: L2 loads (ld.cg) — the producers' stores are in L2, this SM's L1 may hold an older line
#pragma unroll
                for (int i = 0; i < NI; ++i) acc[i] += DEPS ? __ldcg(A.flux_ell + (size_t)off + i * 32) : A.flux_ell[(size_t)off + i * 32];
            } else {
                const double* __restrict__ f = A.flux_slots + (size_t)(-(off + 1));
#pragma unroll
                for (int i = 0; i < NI; ++i) acc[i] += f[i];
            }
        }
        const int y = k / P.nx, x = k - y * P.nx;
        double rho = 0.0;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const double delta_env = (-acc[i]) / P.env_vol_div;                     // sim_toolbox.py:1229
            const double c = cv[i] + delta_env * P.dt;
            A.cc_env[nxt][(size_t)i * E + k] = c;
            if (XP) {
                for (int q = 0; q < X->n_nbr; ++q) {
                    const XNbr& nb = X->nb[q];
                    if (y >= nb.cc_src_row0 && y < nb.cc_src_row0 + nb.cc_rows)
                        nb.cc_env[nxt][(size_t)i * nb.En + (size_t)(nb.cc_dst_row0 + (y - nb.cc_src_row0)) * P.nx + x] = c;
                }
            }
            rho = fma(P.zF[i], c, rho);
        }
        if (A.extra_rho_env) rho += ldg(A.extra_rho_env + k);
        A.rho_env[k] = rho;
        const double vr = (s1 > s0) ? ((rho * P.env_vol_div) / P.memsa_mean) / P.ko_eo_er : 0.0;   // ion_current.py:93-97
        A.v_raw[k] = vr;
        if (XP) {
            for (int q = 0; q < X->n_nbr; ++q) {
                const XNbr& nb = X->nb[q];
                if (y >= nb.v_src_row0 && y < nb.v_src_row0 + nb.v_rows)
                    nb.v_raw[(size_t)(nb.v_dst_row0 + (y - nb.v_src_row0)) * P.nx + x] = vr;
            }
        }
    }
    if (XP) {
        // a CTA of the two end groups pushed (or could have): the last of them publishes
        if (blockIdx.x < X->env_n0 + X->env_n1) {
            __syncthreads();
            if (threadIdx.x == 0) xchg_publish(*X, 1, X->n_push_ctas);
        }
    }
}

template <int NI, bool DEPS>
__global__ void __launch_bounds__(256, DEPS ? 5 : 4)
k_envacc_ell(const __grid_constant__ KParams P, const KArrays A, const int nxt) { envacc_ell_body<NI, DEPS, false>(P, A, nxt, nullptr); }

template <int NI>
__global__ void __launch_bounds__(256, 4)
k_envacc_ell_x(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ XPlan X, const int nxt) { envacc_ell_body<NI, false, true>(P, A, nxt, &X); }

// decomposed tissue, exchange point X2 inside the kernel
void launch_envacc_ell_x(int ni, const KParams& P, const KArrays& A, const XPlan& X, int nxt, cudaStream_t st)
{
    const int n = (P.ya1 - P.ya0) * P.nx;
    if (n <= 0) return;
    const int g = (n + 255) / 256;
    switch (ni) {
        case 4: k_envacc_ell_x<4><<<g, 256, 0, st>>>(P, A, X, nxt); break;
        case 5: k_envacc_ell_x<5><<<g, 256, 0, st>>>(P, A, X, nxt); break;
        case 6: k_envacc_ell_x<6><<<g, 256, 0, st>>>(P, A, X, nxt); break;
        default: k_envacc_ell_x<7><<<g, 256, 0, st>>>(P, A, X, nxt); break;
    }
}

// number of 256-square CTAs of k_envacc_ell at the lower / upper end of the accumulation rows that cover rows [lo, hi)
void envacc_end_ctas(const KParams& P, int lo, int hi, int* n_lower, int* n_upper)
{
    const int n = (P.ya1 - P.ya0) * P.nx, g = (n + 255) / 256;
    const int mid = (P.ya0 + P.ya1) / 2;
    *n_lower = *n_upper = 0;
    if (hi <= lo || g <= 0) return;
    const long long k_lo = (long long)(lo - P.ya0) * P.nx, k_hi = (long long)(hi - P.ya0) * P.nx;     // squares [k_lo, k_hi)
    const int b_lo = (int)std::max(0LL, k_lo / 256), b_hi = (int)std::min((long long)g - 1, (k_hi - 1) / 256);
    if (lo < mid) *n_lower = b_hi + 1;            // from CTA 0 up to the last one touching the range
    else *n_upper = g - b_lo;                     // from the first one touching the range to the last CTA
}

// deps: run next to k_cell (the caller launches it on a second stream and has built env_dep for CTAs of 256 squares)
void launch_envacc_ell(int ni, const KParams& P, const KArrays& A, int nxt, int deps, cudaStream_t st)
{
    const int n = (P.ya1 - P.ya0) * P.nx;
    if (n <= 0) return;
    const int g = (n + 255) / 256;
    if (deps) {
        // KC_DEPS_SMEM of (unused) dynamic shared memory: at most ONE of these CTAs is resident per SM, so that its waiting
        // CTAs can never keep k_cell — which they wait for — off the SMs
        switch (ni) {
            case 4: k_envacc_ell<4, true><<<g, 256, KC_DEPS_SMEM, st>>>(P, A, nxt); break;
            case 5: k_envacc_ell<5, true><<<g, 256, KC_DEPS_SMEM, st>>>(P, A, nxt); break;
            case 6: k_envacc_ell<6, true><<<g, 256, KC_DEPS_SMEM, st>>>(P, A, nxt); break;
            default: k_envacc_ell<7, true><<<g, 256, KC_DEPS_SMEM, st>>>(P, A, nxt); break;
        }
        return;
    }
    switch (ni) {
        case 4: k_envacc_ell<4, false><<<g, 256, 0, st>>>(P, A, nxt); break;
        case 5: k_envacc_ell<5, false><<<g, 256, 0, st>>>(P, A, nxt); break;
        case 6: k_envacc_ell<6, false><<<g, 256, 0, st>>>(P, A, nxt); break;
        default: k_envacc_ell<7, false><<<g, 256, 0, st>>>(P, A, nxt); break;
    }
}

// ---------------------------------------------------------------------------- launch
bool kcell_enabled()
{
    static int v = -1;
    if (v < 0) v = kc_env_int("BETSE_KCELL", 1) ? 1 : 0;
    return v == 1;
}

static int g_kc_sms = 148;

void kcell_set_sms(int n) { if (n > 0) g_kc_sms = n; }

// the pipelined build needs two CTAs (8 warps) per SM in 227 KB of shared memory (each CTA also costs 1 KB of system
// use) and blocks of at least two rows (its per-cell stage is double buffered: the fetch runs two rows ahead)
template <int NI>
static bool pipe_fits(int kb_min)
{
    static int v = -1;
    if (v < 0) v = kc_env_int("BETSE_KCELL_PIPE", 0) ? 1 : 0;    // measured slower than the register build (profiles/r02e_*): opt-in
    return v == 1 && kb_min >= 2 && 2 * ((size_t)KcPipe<NI>::CTA_B + 1024) <= (size_t)227 * 1024;
}

bool kcell_pipe_fits(int ni, int kb_min)
{
    switch (ni) {
        case 4: return pipe_fits<4>(kb_min);
        case 5: return pipe_fits<5>(kb_min);
        case 6: return pipe_fits<6>(kb_min);
        case 7: return pipe_fits<7>(kb_min);
        default: return false;
    }
}

template <int NI>
static cudaError_t prep_cell_t(int kb_min)
{
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_envacc_ell<NI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, KC_DEPS_SMEM))) return e;
    if (!pipe_fits<NI>(kb_min)) return cudaSuccess;
    if ((e = cudaFuncSetAttribute(k_cell_pipe<NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, KcPipe<NI>::CTA_B))) return e;
    return cudaFuncSetAttribute(k_cell_pipe<NI>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}

// not capturable: once per context
cudaError_t prepare_cell(int ni, int kb_min)
{
    switch (ni) {
        case 4: return prep_cell_t<4>(kb_min);
        case 5: return prep_cell_t<5>(kb_min);
        case 6: return prep_cell_t<6>(kb_min);
        case 7: return prep_cell_t<7>(kb_min);
        default: return cudaSuccess;
    }
}

template <int NI>
static void launch_cell_t(const KParams& P, const KArrays& A, int cur, cudaStream_t st)
{
    static int minb = -1;
    if (minb < 0) minb = kc_env_int("BETSE_KCELL_MINB", 2);      // register build: resident CTAs (of 4 warps) per SM, 2 = 255 registers, 3 = 168
    const int need = (P.n_blocks + KC_WARPS - 1) / KC_WARPS;
    if (pipe_fits<NI>(P.kb_min)) {
        const int grid = need < g_kc_sms * 2 ? need : g_kc_sms * 2;
        k_cell_pipe<NI><<<grid, KC_WARPS * 32, KcPipe<NI>::CTA_B, st>>>(P, A, cur);
        return;
    }
    static int share = -1;
    if (share < 0) share = kc_env_int("BETSE_KCELL_SHARE", 0);   // 208-register build: the env kernels run next to it (measured: no gain, profiles/r02f_*)
    const int mb = minb <= 2 ? 2 : 3;
    const int grid = (!P.kc_persist || need < g_kc_sms * mb) ? need : g_kc_sms * mb;
    if (mb == 2 && share) k_cell_share<NI><<<grid, KC_WARPS * 32, 0, st>>>(P, A, cur);
    else if (mb == 2) k_cell<NI, 2><<<grid, KC_WARPS * 32, 0, st>>>(P, A, cur);
    else k_cell<NI, 3><<<grid, KC_WARPS * 32, 0, st>>>(P, A, cur);
}

// decomposed tissue, exchange point X1 inside the kernel (register build, persistent)
void launch_cell_x(int ni, const KParams& P, const KArrays& A, const XPlan& X, int cur, cudaStream_t st)
{
    const int need = (P.n_blocks + KC_WARPS - 1) / KC_WARPS;
    const int grid = (!P.kc_persist || need < g_kc_sms * 2) ? need : g_kc_sms * 2;
    switch (ni) {
        case 4: k_cell_x<4><<<grid, KC_WARPS * 32, 0, st>>>(P, A, X, cur); break;
        case 5: k_cell_x<5><<<grid, KC_WARPS * 32, 0, st>>>(P, A, X, cur); break;
        case 6: k_cell_x<6><<<grid, KC_WARPS * 32, 0, st>>>(P, A, X, cur); break;
        default: k_cell_x<7><<<grid, KC_WARPS * 32, 0, st>>>(P, A, X, cur); break;
    }
}

// register build: the caller zeroes A.ticket on the same stream before every launch
void launch_cell(int ni, const KParams& P, const KArrays& A, int cur, cudaStream_t st)
{
    switch (ni) {
        case 4: launch_cell_t<4>(P, A, cur, st); break;
        case 5: launch_cell_t<5>(P, A, cur, st); break;
        case 6: launch_cell_t<6>(P, A, cur, st); break;
        default: launch_cell_t<7>(P, A, cur, st); break;
    }
}
