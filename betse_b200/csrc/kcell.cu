// k_cell — the membranes->cells kernel of the tissue step with ONE LANE PER CELL (same arithmetic as
// kernels.cu:k_mem through the shared functions of kmath.cuh, bit-identical results).
//
// Why (profiles/r01i_*: the lane-per-membrane kernels issue ~28 warp instructions per membrane, two thirds of them
// staging, segmented sums and pipeline bookkeeping, and every membrane lane recomputes what belongs to its CELL):
//   * Vmem is a per-cell quantity in the default mode (sim.py:2029), so the membrane-side GHK table, the pumps'
//     equilibrium constant and every pump factor that depends on the cell's concentrations are formed ONCE per cell;
//   * a lane walks its cell's membranes k = 0..nm-1, so the membranes->cell sums (update_Co, sim_toolbox.py:1177)
//     are plain register accumulations in membrane order — no shared memory staging of fluxes, no shuffles;
//   * per-membrane constants live in a sliced-ELL "cell pack" (SELL-32: a block is 32 cells, row k of the block holds
//     membrane k of each of its cells; a row is DmS[I][32], mem_sa[32], partner[32], env square[32]), so lane = cell
//     reads them coalesced, and because neighbouring cells have neighbouring partners and env squares, the gathers of
//     one row touch two or three lines instead of 32;
//   * persistent warps draw blocks in order; the gathers of membrane k+1 load into a second register buffer while
//     membrane k is computed, and a block pulls the streams of the block pf_dist tickets ahead into L2 with bulk prefetches.
//
// Three kernels over the same block routine (cell_task):
//   * k_cell        undivided tissue, blocks = 32 consecutive cells; every membrane's membrane->env fluxes go to flux_ell
//                   and k_envacc_ell sums them per env square afterwards;
//   * k_cell_x      strip of a decomposed tissue: the same, with exchange point X1 inside it (xchg.cuh);
//   * k_cell_patch  undivided tissue, the membrane->env exchange done ON CHIP: a CTA owns a PATCH of 128 cells that are
//                   neighbours in space (KArrays.pcell: the blocks are composed from a spatial sort of the cells, not
//                   from their numbering); its warps leave the fluxes of their membranes in shared memory, and after one
//                   barrier the CTA's threads sum them — in the order of the global membrane index, exactly like
//                   k_envacc_ell — for every env square ALL of whose membranes belong to the patch, into sq_sum [I][E].
//                   Only membranes of squares that several patches share write their fluxes to flux_ell as well (bit
//                   KC_BORDER of the row's env-square word).  k_envacc_ell then takes an owned square's sums with one
//                   load per ion instead of walking its slots (bit 31 of slot_ptrp marks them).  The env accumulation
//                   stays a kernel of its own because the transport of the non-Ca ions (k_ion) runs NEXT TO the membrane
//                   kernel and must have finished before its result is advanced.  The 576 MB round trip of the exchange
//                   slots (1/3 of the step's DRAM traffic) shrinks to the border share + 64 MB of sums.
//
// Reference lines as in k_mem: sim.py:1193-1283, 2086-2111, 2162-2206; sim_toolbox.py:18-182, 1155-1234;
// channels/gap_junction.py:53-77; ion_current.py:19, 75-97; sim.py:2027-2029.
#include <stdlib.h>
#include <algorithm>
#include <stdint.h>
#include "kmath.cuh"
#include "xchg.cuh"

#define KC_WARPS 4
#define KC_ROWB(NI) (((NI) + 2) * 256)          // bytes of one row of the cell pack
#define KC_INV    0x80000000u                    // env-square word: the lane has no membrane in this row
#define KC_BORDER 0x10000000u                    //                  the membrane's env square is shared between patches
#define KC_QMASK  0x0fffffffu

template <int NI>
struct MemIn { double co[NI], cnb[NI], vnb, cao; int nnp, bs; };     // bs: compact border slot of the membrane (patch build), or -1

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
        if (frem) frem[i] = fsa;          // the env square belongs to a neighbouring strip / is shared between patches: [slot][ion]
    }
    return g;
}

// update_Co + update_all_concs, charge and Vmem of the cell (sim_toolbox.py:1177-1181; sim.py:2105-2111;
// ion_current.py:19; sim.py:2027-2029)
template <int NI, bool DEFER = false>
__device__ __forceinline__ void cell_epilogue(const KParams& P, const KArrays& A, const CellSide<NI>& S, const double* cc, const double vol,
                                              const double dvt, const int c, const int nxt, unsigned int& flags,
                                              const XPlan* X = nullptr, const int2 gs = make_int2(-1, -1))
{
    const int C = P.n_cells;
    if (DEFER) {
        // channels / networks act between the ion loop and update_all_concs (sim.py:1290-1357): the sums go to k_cell_update
#pragma unroll
        for (int i = 0; i < NI; ++i) { A.dsum_m[(size_t)i * C + c] = S.Sm[i]; A.dsum_g[(size_t)i * C + c] = S.Sg[i]; }
        return;
    }
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

// ---------------------------------------------------------------------------- one block of 32 cells: lane = cell
// fsm: the block's part of the CTA's flux buffer in shared memory (k_cell_patch), [row][ion][32]; null: fluxes to flux_ell
template <int NI, bool PATCH = false, bool DEFER = false>
__device__ __forceinline__ void cell_task(const KParams& P, const KArrays& A, const int cur, const int task, const int lane, unsigned int& flags,
                                          const XPlan* X = nullptr, double* __restrict__ fsm = nullptr, const int pf_up = -1, const int c_next = -1)
{
    constexpr int iCa = StdProf<NI>::iCa;
    constexpr int ROWB = KC_ROWB(NI);
    const int nxt = cur ^ 1;
    const int C = P.n_cells, E = P.ny * P.nx;
    const int2 h0 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + task);          // {first row, first membrane}
    const int2 h1 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + task + 1);
    const int row0 = h0.x, Kb = h1.x - h0.x;
    const int c = PATCH ? ldgi(A.pcell + task * 32 + lane) : task * 32 + lane;        // patches: blocks are spatial neighbours
    const bool valid = (!PATCH || c >= 0) && c < P.n_cells_owned;
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
    if (!PATCH && X) bx = __ldg(reinterpret_cast<const int2*>(A.blk_x) + task);
    // the block whose streams this task pulls into L2 (issued after the first membrane, below)
    const int up = PATCH ? pf_up : task + P.pf_dist;
    int2 u0 = make_int2(0, 0), u1 = make_int2(0, 0);
    if (P.pf_dist > 0 && up < P.n_blocks) {
        u0 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + up);
        u1 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + up + 1);
    }

    // patch build: first membrane of the cell this lane takes in the CTA's next patch (for the prefetch of its gap junctions)
    int mb_next = 0;
    if (PATCH && c_next >= 0) mb_next = ldgi(A.cell_mem_ptr + c_next);
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
            // patch build: the env square is shared between patches — the membrane's compact border slot, fetched with the gathers
            if (PATCH) x.bs = ((unsigned)ix.y & KC_BORDER) ? ldgi(A.bslot + (size_t)(row0 + k) * 32 + lane) : -1;
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
    if (!PATCH && bx.x >= 0) gs = __ldg(reinterpret_cast<const int2*>(A.ghost_tab) + bx.x + lane);

    auto compute = [&](const MemIn<NI>& x, const int k) {
        if (k < nm) {
            const double* __restrict__ r = reinterpret_cast<const double*>(rows + (size_t)k * ROWB) + lane;
            double DmS[NI];
#pragma unroll
            for (int i = 0; i < NI; ++i) DmS[i] = __ldcs(r + i * 32);
            const double sa = __ldcs(r + NI * 32);
            double* __restrict__ fg = A.flux_ell + ((size_t)(row0 + k) * NI) * 32 + lane;
            double* __restrict__ fl = PATCH ? fsm + (k * NI) * 32 + lane : fg;
            double* frem = nullptr;
            // patch build, shared env square: the fluxes also go to the membrane's compact border slot (8 doubles, [slot][ion]
            // like a remote slot), from where k_envacc_patch sums them
            if (PATCH && x.bs >= 0) frem = A.flux_slots + (size_t)x.bs * 8;
            if (!PATCH && bx.y >= 0) {
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
            // streams of block `up`, one bulk prefetch per array and lane: its rows; with blocks of consecutive cells also
            // gjopen and the cells' own state
            const int cu = up * 32;
            const int ncu = min(32, P.n_cells_owned - cu);
            if (lane == 0) l2_prefetch(A.cpack + (size_t)u0.x * ROWB, (unsigned)(u1.x - u0.x) * ROWB);
            else if (PATCH) { /* the cells of a patch are not consecutive: every lane pulls its next cell's state, below */ }
            else if (lane == 1) l2_prefetch(A.gjopen + u0.y, (unsigned)(u1.y - u0.y) * 8u);
            else if (lane < NI + 2) l2_prefetch(A.cc_cells + (size_t)(lane - 2) * C + cu, ncu * 8u);
            else if (lane < 2 * NI + 2) l2_prefetch(cmid + (size_t)(lane - NI - 2) * C + cu, ncu * 8u);
            else if (lane == 2 * NI + 2) l2_prefetch(vmc + cu, ncu * 8u);
            else if (lane == 2 * NI + 3) l2_prefetch(A.cell_vol + cu, ncu * 8u);
            else if (lane == 2 * NI + 4) l2_prefetch(A.diviterm + cu, ncu * 8u);
            else if (lane == 2 * NI + 5) l2_prefetch(A.cell_mem_ptr + cu, (ncu + 1) * 4u);
        }
        if (PATCH && k == 0 && c_next >= 0) {
            // patch build: the state of the cell this lane takes in the CTA's next patch, one line each
#pragma unroll
            for (int i = 0; i < NI; ++i) prefetch_l2(A.cc_cells + (size_t)i * C + c_next);
#pragma unroll
            for (int i = 0; i < NI; ++i) prefetch_l2(cmid + (size_t)i * C + c_next);
            prefetch_l2(vmc + c_next);
            prefetch_l2(A.cell_vol + c_next);
            prefetch_l2(A.diviterm + c_next);
            prefetch_l2(A.gjopen + mb_next);
            prefetch_l2(A.gjopen + mb_next + 4);
        }
        if (k + 1 < Kb) {
            gather(a, ia, k + 2);
            idx_load(ib, k + 3);
            compute(b, k + 1);
        }
    }
    if (valid) cell_epilogue<NI, DEFER>(P, A, S, cc, vol, dvt, c, nxt, flags, (!PATCH && bx.x >= 0) ? X : nullptr, gs);
    if (!PATCH && X && (bx.x >= 0 || bx.y >= 0)) {
        // this block pushed: the last of the boundary blocks raises this rank's X1 flag on the neighbours
        __syncwarp();
        if (lane == 0) xchg_publish(*X, 0, X->n_bblocks);
    }
}

// Persistent: every warp draws tickets (ticket t = block t of the cell pack, so the blocks finish as a wavefront through
// the tissue); kc_persist = 0: one block per warp.
template <int NI, bool DEFER = false>
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
        cell_task<NI, false, DEFER>(P, A, cur, t, lane, flags, X);
        // a compiler-only fence that keeps the block index live to the end of the iteration: without it ptxas moves the next
        // ticket's atomic and header loads up into this block's epilogue and the kernel loses 6 % (0.258 -> 0.275 ms at 1 M
        // cells, same SASS instruction mix; A/B on one box, profiles/r02l_sweep_patch.txt)
        __syncwarp();
        asm volatile("" ::"r"(t) : "memory");
        if (!P.kc_persist) break;
    }
    if (flags) atomicOr(A.status, flags);
}

// (measured and lost, profiles/r02e_*, r02f_*, r02j_*: a cp.async ring pipeline over the rows of the pack — 0.378 ms against 0.269;
// more resident warps at fewer registers — 255 registers / 8 warps per SM 0.258 ms, 200 / 10 warps 0.278, 184 / 11 warps
// 0.305, 168 / 12 warps 0.31: the spills cost more than the occupancy buys; a 208-register build that leaves room for the
// env kernels next to it, with the env accumulation consuming the fluxes out of L2 behind a completion counter — 0.49 ms
// per step against 0.42: one 256-thread CTA per SM is latency-bound and becomes the critical path)
template <int NI, int MINB>
__global__ void __launch_bounds__(KC_WARPS * 32, MINB)
k_cell(const __grid_constant__ KParams P, const KArrays A, const int cur) { k_cell_body<NI>(P, A, cur, nullptr); }

// deferred-update mode (channels / networks): the membranes -> cell sums go to dsum_m / dsum_g for k_cell_update
template <int NI>
__global__ void __launch_bounds__(KC_WARPS * 32, 2)
k_cell_defer(const __grid_constant__ KParams P, const KArrays A, const int cur) { k_cell_body<NI, true>(P, A, cur, nullptr); }

// decomposed tissue: the same kernel with exchange point X1 inside it (xchg.cuh)
template <int NI>
__global__ void __launch_bounds__(KC_WARPS * 32, 2)
k_cell_x(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ XPlan X, const int cur) { k_cell_body<NI>(P, A, cur, &X); }

// ---------------------------------------------------------------------------- membrane -> env exchange of one env square
// update_Co env branch + div_env (sim_toolbox.py:1189-1234), env charge, raw env voltage (ion_current.py:75-97) from the
// square's summed fluxes: shared by k_envacc_ell, k_envacc_list and the patch build, so that all three give the same bits
template <int NI>
__device__ __forceinline__ double env_square_finish(const KParams& P, const KArrays& A, const int nxt, const int k, const double* cv,
                                                    const double* acc, const bool has_mems, double* c_out)
{
    const int E = P.nx * P.ny;
    double rho = 0.0;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const double delta_env = (-acc[i]) / P.env_vol_div;                     // sim_toolbox.py:1229
        const double c = cv[i] + delta_env * P.dt;
        A.cc_env[nxt][(size_t)i * E + k] = c;
        c_out[i] = c;
        rho = fma(P.zF[i], c, rho);
    }
    if (A.extra_rho_env) rho += ldg(A.extra_rho_env + k);
    A.rho_env[k] = rho;
    const double vr = has_mems ? ((rho * P.env_vol_div) / P.memsa_mean) / P.ko_eo_er : 0.0;   // ion_current.py:93-97
    A.v_raw[k] = vr;
    return vr;
}

// ---------------------------------------------------------------------------- patch build
// the env squares a patch owns: its table (A.ptab at A.ptab_ptr[patch], copied to shared memory while the blocks run) =
// [n, q[n], (first slot | count << 16)[n], slots as u16...]; a slot = position of a membrane's fluxes in the CTA's buffer
// ((warp*kb_max + row)*NI*32 + lane), in global membrane order
template <int NI>
__device__ __forceinline__ void patch_env(const KParams& P, const KArrays& A, const int* __restrict__ tab, const double* __restrict__ fsm)
{
    const int E = P.nx * P.ny;
    const int n_owned = tab[0];
    const int* __restrict__ qs = tab + 1;
    const int* __restrict__ sc = qs + n_owned;
    const unsigned short* __restrict__ slots = reinterpret_cast<const unsigned short*>(sc + n_owned);
    for (int t = threadIdx.x; t < n_owned; t += blockDim.x) {
        const int q = qs[t];
        const unsigned w = (unsigned)sc[t];
        const int s0 = (int)(w & 0xffffu), cnt = (int)(w >> 16);
        double acc[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) acc[i] = 0.0;
        for (int j = 0; j < cnt; ++j) {
            const int pos = slots[s0 + j];
#pragma unroll
            for (int i = 0; i < NI; ++i) acc[i] += fsm[pos + i * 32];
        }
#pragma unroll
        for (int i = 0; i < NI; ++i) A.sq_sum[(size_t)i * E + q] = acc[i];
    }
}

// Static schedule: CTA b takes patches b, b + grid, ... — the next patch is known, so its table is fetched and the state of
// its cells pulled into L2 while the current one is computed.  Shared memory: two flux buffers and two table buffers; ONE
// barrier per patch (the fluxes of all four blocks and the table are in place; the other buffers are free again because
// every thread came through this barrier after its env phase of the patch before).
template <int NI>
__global__ void __launch_bounds__(KC_WARPS * 32, 2)
k_cell_patch(const __grid_constant__ KParams P, const KArrays A, const int cur)
{
    extern __shared__ __align__(16) double kc_fsm[];          // [2][KC_WARPS][kb_max][NI][32] fluxes | [2][ptab_max] table ints
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_warp = P.kb_max * NI * 32, per_buf = KC_WARPS * per_warp;
    int* const tab_sm = reinterpret_cast<int*>(kc_fsm + 2 * per_buf);
    unsigned int flags = 0;
    int buf = 0;
    int patch = blockIdx.x;
    int off = 0, end = 0;
    if (patch < P.n_patches) { off = ldgi(A.ptab_ptr + patch); end = ldgi(A.ptab_ptr + patch + 1); }
    for (; patch < P.n_patches; patch += gridDim.x, buf ^= 1) {
        const int next = patch + gridDim.x;
        double* fsm = kc_fsm + buf * per_buf;
        int* tab = tab_sm + buf * P.ptab_max;
        for (int j = threadIdx.x; j < end - off; j += blockDim.x) {
            const uint32_t d = (uint32_t)__cvta_generic_to_shared(tab + j);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(A.ptab + off + j) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        int c_next = -1;
        if (next < P.n_patches) {
            c_next = ldgi(A.pcell + (size_t)next * (KC_WARPS * 32) + threadIdx.x);
            off = ldgi(A.ptab_ptr + next); end = ldgi(A.ptab_ptr + next + 1);
        }
        cell_task<NI, true>(P, A, cur, patch * KC_WARPS + warp, lane, flags, nullptr, fsm + warp * per_warp,
                      next < P.n_patches ? next * KC_WARPS + warp : P.n_blocks, c_next);
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        patch_env<NI>(P, A, tab, fsm);
    }
    if (flags) atomicOr(A.status, flags);
}

static int kc_env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// ---------------------------------------------------------------------------- cell pack
// constant part: membrane areas and index rows in SELL-32 order, and the flux position of every membrane.
// mem_bidx (patch build, else null): compact border slot of the membranes whose env square several patches share, else -1
__global__ void k_pack_cell_const(const __grid_constant__ KParams P, const KArrays A, int* mem_ell, const int* __restrict__ mem_bidx)
{
    const int task = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (task >= P.n_blocks) return;
    const int row0 = A.blk_row0[2 * task], Kb = A.blk_row0[2 * task + 2] - row0;
    const int c = A.pcell ? A.pcell[task * 32 + lane] : task * 32 + lane;
    int m_beg = 0, nm = 0;
    if (c >= 0 && c < P.n_cells_owned) { m_beg = A.cell_mem_ptr[c]; nm = A.cell_mem_ptr[c + 1] - m_beg; }
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
            int bi = -1;
            if (mem_bidx) bi = mem_bidx[m];
            if (bi >= 0) { esq |= KC_BORDER; const_cast<int*>(A.bslot)[(size_t)(row0 + k) * 32 + lane] = bi; }
            // where k_envacc_* finds the membrane's fluxes: in flux_ell (stride 32 between ions), or in its compact border slot
            mem_ell[m] = (bi >= 0) ? -(bi * 8) - 1 : (int)(((size_t)(row0 + k) * ni) * 32 + lane);
        }
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
    const int c = A.pcell ? A.pcell[task * 32 + lane] : task * 32 + lane;
    int m_beg = 0, nm = 0;
    if (c >= 0 && c < P.n_cells_owned) { m_beg = A.cell_mem_ptr[c]; nm = A.cell_mem_ptr[c + 1] - m_beg; }
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

void launch_pack_cell_const(const KParams& P, const KArrays& A, int* mem_ell, const int* mem_bidx, cudaStream_t st)
{
    if (!A.cpack || P.n_blocks <= 0) return;
    k_pack_cell_const<<<(P.n_blocks + 7) / 8, 256, 0, st>>>(P, A, mem_ell, mem_bidx);
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
// XP (decomposed tissue, xchg.cuh): exchange point X2 inside the kernel — the CTAs at the two ends of the owned rows run
// first (blockIdx is permuted), store their rows of cc_env and of the raw env voltage straight into the neighbours' halo
// rows as well, and the last of them raises this rank's X2 flag there.
// PATCH (after k_cell_patch): CTAs [0, n_dense) take the squares a patch owns (bit 31 of slot_ptrp[k]; their fluxes are
// already summed in sq_sum: one load per ion), the CTAs behind them the squares of A.out_sq — those several patches share
// and those without membranes — as a dense list, so that no warp mixes the one-load path with the slot walk.
template <int NI, bool XP, bool PATCH>
__device__ __forceinline__ void envacc_ell_body(const KParams& P, const KArrays& A, const int nxt, const XPlan* X)
{
    int b = blockIdx.x;
    if (XP) {
        // [lower end | upper end | interior]
        const int g = gridDim.x, n0 = X->env_n0, n1 = X->env_n1;
        if (b >= n0) b = (b < n0 + n1) ? g - n1 + (b - n0) : n0 + (b - n0 - n1);
    }
    int k = P.ya0 * P.nx + b * blockDim.x + threadIdx.x;
    bool in = k < P.ya1 * P.nx;
    const int E = P.nx * P.ny;
    bool listed = false;
    if (PATCH) {
        const int n_dense = ((P.ya1 - P.ya0) * P.nx + (int)blockDim.x - 1) / (int)blockDim.x;
        if (b >= n_dense) {
            const int t = (b - n_dense) * blockDim.x + threadIdx.x;
            listed = true;
            in = t < P.n_out_sq;
            k = in ? ldgi(A.out_sq + t) : 0;
        }
    }
    // everything that does not depend on the fluxes first
    int s0 = 0, s1 = 0;
    double acc[NI], cv[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) { acc[i] = 0.0; cv[i] = 0.0; }
    bool summed = false;
    if (in) {
        if (PATCH && !listed) {
            summed = ((unsigned)ldgi(A.slot_ptrp + k) >> 31) != 0;
            in = summed;                                  // the dense CTAs leave the other squares to the list
            s0 = 0; s1 = 1;
        } else { s0 = ldgi(A.slot_ptr + k); s1 = ldgi(A.slot_ptr + k + 1); }
    }
    if (in) {
#pragma unroll
        for (int i = 0; i < NI; ++i) cv[i] = A.cc_env[nxt][(size_t)i * E + k];
    }
    if (XP) {
        // every CTA that pushes waits for the neighbours' X1 first: it tells that they are past the kernels that still read
        // (field of the last step) or write (their own redundant transport of the halo rows) what this CTA is about to
        // store into their windows; the CTAs whose squares take remote flux slots are among them
        if (blockIdx.x < X->env_n0 + X->env_n1) xchg_wait_cta(P, A, 0);
    } else if (!PATCH && P.xwait) {
        // decomposed tissue: squares near the strip edges take fluxes the neighbours pushed (exchange point X1)
        const int k0 = P.ya0 * P.nx + b * blockDim.x;
        if (k0 / P.nx < P.xw_lo || (k0 + (int)blockDim.x - 1) / P.nx >= P.xw_hi) xchg_wait_cta(P, A, 0);
    }
    if (in) {
        if (PATCH && summed) {
#pragma unroll
            for (int i = 0; i < NI; ++i) acc[i] = A.sq_sum[(size_t)i * E + k];
        } else
        // the square's slots four at a time: all positions first, then all fluxes, then the sums in slot order — small
        // tissues and thin strips are bound by the dependent round trips of this walk, not by bandwidth
        for (int j0 = s0; j0 < s1; j0 += 4) {
            int off[4];
            double v[4][NI];
#pragma unroll
            for (int u = 0; u < 4; ++u) off[u] = (j0 + u < s1) ? ldgi(A.slot_off + j0 + u) : 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                // local membrane: flux_ell, stride 32 between ions; remote slot: the window's [slot][ion] array
                const double* __restrict__ f = (off[u] >= 0) ? A.flux_ell + (size_t)off[u] : A.flux_slots + (size_t)(-(off[u] + 1));
                const int st = (off[u] >= 0) ? 32 : 1;
#pragma unroll
                for (int i = 0; i < NI; ++i) v[u][i] = (j0 + u < s1) ? f[i * st] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (j0 + u < s1) {
#pragma unroll
                    for (int i = 0; i < NI; ++i) acc[i] += v[u][i];
                }
        }
        double c[NI];
        const double vr = env_square_finish<NI>(P, A, nxt, k, cv, acc, s1 > s0, c);
        if (XP) {
            const int y = k / P.nx, x = k - y * P.nx;
            for (int q = 0; q < X->n_nbr; ++q) {
                const XNbr& nb = X->nb[q];
                if (y >= nb.cc_src_row0 && y < nb.cc_src_row0 + nb.cc_rows) {
#pragma unroll
                    for (int i = 0; i < NI; ++i)
                        nb.cc_env[nxt][(size_t)i * nb.En + (size_t)(nb.cc_dst_row0 + (y - nb.cc_src_row0)) * P.nx + x] = c[i];
                }
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

template <int NI>
__global__ void __launch_bounds__(256, 4)
k_envacc_ell(const __grid_constant__ KParams P, const KArrays A, const int nxt) { envacc_ell_body<NI, false, false>(P, A, nxt, nullptr); }

template <int NI>
__global__ void __launch_bounds__(256, 4)
k_envacc_patch(const __grid_constant__ KParams P, const KArrays A, const int nxt) { envacc_ell_body<NI, false, true>(P, A, nxt, nullptr); }

template <int NI>
__global__ void __launch_bounds__(256, 4)
k_envacc_ell_x(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ XPlan X, const int nxt) { envacc_ell_body<NI, true, false>(P, A, nxt, &X); }

#define KC_DISPATCH_NI(ni, CALL)                  \
    switch (ni) {                                 \
        case 4: { constexpr int NI = 4; CALL; } break; \
        case 5: { constexpr int NI = 5; CALL; } break; \
        case 6: { constexpr int NI = 6; CALL; } break; \
        default: { constexpr int NI = 7; CALL; } break; \
    }

// decomposed tissue, exchange point X2 inside the kernel
void launch_envacc_ell_x(int ni, const KParams& P, const KArrays& A, const XPlan& X, int nxt, cudaStream_t st)
{
    const int n = (P.ya1 - P.ya0) * P.nx;
    if (n <= 0) return;
    const int g = (n + 255) / 256;
    KC_DISPATCH_NI(ni, (k_envacc_ell_x<NI><<<g, 256, 0, st>>>(P, A, X, nxt)));
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

void launch_envacc_ell(int ni, const KParams& P, const KArrays& A, int nxt, cudaStream_t st)
{
    const int n = (P.ya1 - P.ya0) * P.nx;
    if (n <= 0) return;
    const int g = (n + 255) / 256;
    if (P.n_patches > 0) { const int gp = g + (P.n_out_sq + 255) / 256; KC_DISPATCH_NI(ni, (k_envacc_patch<NI><<<gp, 256, 0, st>>>(P, A, nxt))); }
    else { KC_DISPATCH_NI(ni, (k_envacc_ell<NI><<<g, 256, 0, st>>>(P, A, nxt))); }
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

// dynamic shared memory of the patch build: two flux buffers of [KC_WARPS][kb_max][ni][32] doubles; two CTAs per SM
size_t kcell_patch_smem(int ni, int kb_max, int ptab_max) { return (size_t)2 * KC_WARPS * kb_max * ni * 32 * sizeof(double) + (size_t)2 * ptab_max * sizeof(int); }

bool kcell_patch_fits(int ni, int kb_max)
{
    static int v = -1;
    // measured (profiles/r02l_*): the on-chip exchange halves the DRAM traffic of the env accumulation (415 -> 278 MB) and cuts
    // the membrane kernel's writes by 40 %, but the kernel is latency-bound, not bandwidth-bound: with the cells of a warp
    // in two rows instead of one and a barrier per patch it takes 0.387 ms against 0.259 — opt-in with BETSE_PATCH=1
    if (v < 0) v = kc_env_int("BETSE_PATCH", 0) ? 1 : 0;
    // (a patch's table: 1 + 2 squares + slots / 2 ints, at most ~128 cells x kb_max membranes: 8 KB is ample)
    return v == 1 && ni >= 4 && ni <= 7 && kb_max > 0 && 2 * (kcell_patch_smem(ni, kb_max, 2048) + 1024) <= (size_t)227 * 1024 &&
           (size_t)KC_WARPS * kb_max * ni * 32 <= 65535;
}

// not capturable: once per context
cudaError_t prepare_cell(int ni, int kb_max, int ptab_max)
{
    if (ptab_max <= 0 || !kcell_patch_fits(ni, kb_max)) return cudaSuccess;
    const int smem = (int)kcell_patch_smem(ni, kb_max, ptab_max);
    cudaError_t e = cudaSuccess;
    KC_DISPATCH_NI(ni, (e = cudaFuncSetAttribute(k_cell_patch<NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)));
    if (e) return e;
    // two CTAs' worth of shared memory and no more: the gathers live on what is left for L1
    const int pct = (int)((2 * ((size_t)smem + 1024) * 100 + (size_t)228 * 1024 - 1) / ((size_t)228 * 1024));
    KC_DISPATCH_NI(ni, (e = cudaFuncSetAttribute(k_cell_patch<NI>, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct)));
    return e;
}

// decomposed tissue, exchange point X1 inside the kernel (persistent)
void launch_cell_x(int ni, const KParams& P, const KArrays& A, const XPlan& X, int cur, cudaStream_t st)
{
    const int need = (P.n_blocks + KC_WARPS - 1) / KC_WARPS;
    const int grid = (!P.kc_persist || need < g_kc_sms * 2) ? need : g_kc_sms * 2;
    KC_DISPATCH_NI(ni, (k_cell_x<NI><<<grid, KC_WARPS * 32, 0, st>>>(P, A, X, cur)));
}

// the caller zeroes A.ticket on the same stream before every launch
void launch_cell(int ni, const KParams& P, const KArrays& A, int cur, cudaStream_t st)
{
    if (P.n_patches > 0) {
        const int grid = P.n_patches < g_kc_sms * 2 ? P.n_patches : g_kc_sms * 2;
        const size_t smem = kcell_patch_smem(ni, P.kb_max, P.ptab_max);
        KC_DISPATCH_NI(ni, (k_cell_patch<NI><<<grid, KC_WARPS * 32, smem, st>>>(P, A, cur)));
        return;
    }
    static int minb = -1;
    if (minb < 0) minb = kc_env_int("BETSE_KCELL_MINB", 2);      // resident CTAs (of 4 warps) per SM, 2 = 255 registers, 3 = 168
    const int need = (P.n_blocks + KC_WARPS - 1) / KC_WARPS;
    const int grid = (!P.kc_persist || need < g_kc_sms * minb) ? need : g_kc_sms * (minb <= 2 ? 2 : 3);
    if (P.defer) { KC_DISPATCH_NI(ni, (k_cell_defer<NI><<<grid, KC_WARPS * 32, 0, st>>>(P, A, cur))); }
    else if (minb <= 2) { KC_DISPATCH_NI(ni, (k_cell<NI, 2><<<grid, KC_WARPS * 32, 0, st>>>(P, A, cur))); }
    else { KC_DISPATCH_NI(ni, (k_cell<NI, 3><<<grid, KC_WARPS * 32, 0, st>>>(P, A, cur))); }
}
