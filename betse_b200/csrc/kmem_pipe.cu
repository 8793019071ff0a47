// k_mem_pipe — the membranes->cells kernel of the tissue step as a persistent, software-pipelined
// warp kernel (same arithmetic as kernels.cu:k_mem; shared helpers in kmath.cuh).
//
// Why: k_mem is latency- and LSU-bound, not bandwidth-bound (ncu summaries profiles/r01d, r01e:
// 38 % issue utilisation with long-scoreboard stalls; a cp.async-only pipeline moved the stall to
// the LSU — 50 LDGSTS per tile at ~8 cycles each; one TMA bulk copy per row costs ~12 issue slots
// per copy because UBLKCP takes uniform operands and per-lane copies are serialised).  A tile's
// inputs arrive in three dependent round trips (tile descriptor -> membrane index arrays ->
// gathers through those indices).  Here every warp walks a strided sequence of tiles and keeps
// the NEXT tiles' inputs in flight while it does the arithmetic of the current one:
//
//   K = the tile's constant block ("tile pack", fixed size, built by the host + k_pack_dm):
//       DmS[I][32] = (Dm*(-rho_channel/tm))*mem_sa, mem_sa[32], cell_vol[8], diviterm[8],
//       mem_to_cells[32], nn_cell_flag[32], map_mem2ecm[32], cell_mem_ptr[9(+3)] -> ONE TMA bulk
//       copy (cp.async.bulk by lane 0, completion on an mbarrier), two tiles ahead; every offset
//       inside the block is a compile-time constant;
//   G = state and gathers, one tile ahead: Vmem, cc_cells, cc_mid of the tile's cells by cp.async (LDGSTS); the
//       per-membrane gathers through the landed indices of K — env concentrations at the membrane's env square,
//       partner-cell concentrations and Vmem, transported Ca, gjopen — by cp.async into shared memory
//       (BETSE_KMEM_GREG=0) or, default, by plain loads into REGISTERS issued after the flux phase of the current tile
//       (168 registers, no spills at 3 CTAs x 4 warps; 113 of 268 shared-memory wavefronts per tile less)
//
//   iteration t:  wait (mbarrier of t-1, cp.async group)   -> K(t+1), G(t) have landed
//                 K(t), G(t) -> registers
//                 issue K(t+2) into K(t)'s buffer, G(t+1)
//                 arithmetic of tile t (staging of f*sa aliases the consumed gather rows)
//
// Reference lines as in k_mem: sim.py:1193-1283, 2086-2111, 2162-2206; sim_toolbox.py:18-182,
// 1155-1207; channels/gap_junction.py:53-77; ion_current.py:19; sim.py:2027-2029.
#include <stdlib.h>
#include <stdint.h>
#include "kmath.cuh"

#define KP_MAXC BT_TILE_MAXC                         // cells per tile (host packing, capi.cu)
#define KP_SST(NI) ((NI) | 1)                        // staging stride (doubles) of f*sa [membrane][ion]: odd => the membrane-lane stores,
                                                     // the (cell, ion)-lane sums and the slot copy all stay (nearly) bank-conflict free
// tile pack block, in doubles: DmS[NI][32], sa[32], vol[MAXC], dvt[MAXC] | ints: m2c[32], nnp[32], e[32], ptr[MAXC+4]
#define KP_KD(NI) (((NI) + 1) * 32 + 2 * KP_MAXC)    // doubles before the int section
#define KP_K(NI) (KP_KD(NI) + (96 + KP_MAXC + 4) / 2)
// per-warp shared memory, in doubles
#define KP_G(NI) ((2 * (NI) + 3) * 32)               // co[NI][32], cnb[NI][32], vnb[32], cao[32], gj[32]; then staging 2*32*KP_SST
#define KP_C(NI) (KP_MAXC + 2 * (NI) * KP_MAXC)      // vmo[c], cc[c][NI] (then the updated values), cmid[c][NI]
#define KP_AUX (KP_MAXC + 8)                         // cell_mem_ptr[MAXC+1] (6 doubles) + cell_vol[MAXC]: tiles with > 32/NI cells
#define KP_WARP(NI) (2 + 2 * KP_K(NI) + 2 * KP_G(NI) + 2 * KP_C(NI) + KP_AUX)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp8(uint32_t s, const void* g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(g));
}
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KP_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra KP_DONE;\n"
        "bra KP_WAIT;\n"
        "KP_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, unsigned bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// the constant block of tile `tile` -> Kdst; always completes one phase of `bar`
template <int NI>
__device__ __forceinline__ void issue_K(const KArrays& A, const int tile, const bool valid, const uint32_t Kdst,
                                        const uint32_t bar, const int lane)
{
    if (lane == 0) {
        if (valid) {
            mbar_expect_tx(bar, KP_K(NI) * 8);
            bulk_g2s(Kdst, A.tile_pack + (size_t)tile * (KP_K(NI) * 8), KP_K(NI) * 8, bar);
        } else mbar_arrive(bar);
    }
}

// state + gathers of tile td through the indices in its (landed) constant block K
template <int NI, bool GREG = false>
__device__ __forceinline__ void issue_G(const KArrays& A, const int4 td, const double* K, double* G, double* Cs,
                                        const int lane, const int o0, const int C, const int E, const int cur)
{
    const int c0 = td.x, nc = td.y, m0 = td.z, nm = td.w;
    if (!GREG && lane < nm) {
        const int* Ki = reinterpret_cast<const int*>(K + KP_KD(NI));
        const int cn = Ki[32 + lane] & 0x7fffffff;
        const int e = Ki[64 + lane];
        const double* __restrict__ cenv = A.cc_env[cur] + e;
        const double* __restrict__ cmid = A.cc_mid[cur] + cn;
        const uint32_t g = smem_u32(G + lane);
#pragma unroll
        for (int i = 0; i < NI; ++i) cp8(g + i * 256, cenv + (size_t)i * E);
#pragma unroll
        for (int i = 0; i < NI; ++i) cp8(g + (NI + i) * 256, cmid + (size_t)i * C);
        cp8(g + (2 * NI) * 256, A.vm_cell[cur] + cn);
        if (StdProf<NI>::iCa >= 0) cp8(g + (2 * NI + 1) * 256, A.cc_env[cur ^ 1] + (size_t)StdProf<NI>::iCa * E + e);
        cp8(g + (2 * NI + 2) * 256, A.gjopen + m0 + lane);
    }
    const uint32_t cs = smem_u32(Cs);
    if (lane < nc) cp8(cs + lane * 8, A.vm_cell[cur] + c0 + lane);
    if (lane < nc * NI) {                              // first round of (cell, ion) pairs: o0 = ion*C + cell-in-tile
        cp8(cs + (KP_MAXC + lane) * 8, A.cc_cells + o0 + c0);
        cp8(cs + (KP_MAXC + NI * KP_MAXC + lane) * 8, A.cc_mid[cur] + o0 + c0);
    }
    for (int q = lane + 32; q < nc * NI; q += 32) {
        const int lc = q / NI, i = q - lc * NI;
        const size_t o = (size_t)i * C + c0 + lc;
        cp8(cs + (KP_MAXC + q) * 8, A.cc_cells + o);
        cp8(cs + (KP_MAXC + NI * KP_MAXC + q) * 8, A.cc_mid[cur] + o);
    }
}

// GREG variant: the per-membrane gathers of the NEXT tile go straight into registers (plain loads issued after the flux
// phase of the current tile, when most registers are free again) instead of through cp.async + shared memory: fewer
// shared-memory wavefronts, the pipe the kernel is bound by
template <int NI>
struct GReg { double co[NI], cnb[NI], vm_nb, cao, g; };

template <int NI>
__device__ __forceinline__ void load_G(const KArrays& A, const int4 td, const double* K, GReg<NI>& R,
                                       const int lane, const int C, const int E, const int cur)
{
    const int m0 = td.z, nm = td.w;
    if (lane < nm) {
        const int* Ki = reinterpret_cast<const int*>(K + KP_KD(NI));
        const int cn = Ki[32 + lane] & 0x7fffffff;
        const int e = Ki[64 + lane];
        // uniform row bases + a 32-bit lane index: one address instruction per load
        const double* __restrict__ cenv = A.cc_env[cur];
        const double* __restrict__ cmid = A.cc_mid[cur];
        const unsigned ue = (unsigned)e, ucn = (unsigned)cn;
#pragma unroll
        for (int i = 0; i < NI; ++i) R.co[i] = (cenv + (size_t)i * E)[ue];
#pragma unroll
        for (int i = 0; i < NI; ++i) R.cnb[i] = (cmid + (size_t)i * C)[ucn];
        R.vm_nb = A.vm_cell[cur][cn];
        R.cao = (StdProf<NI>::iCa >= 0) ? A.cc_env[cur ^ 1][(size_t)StdProf<NI>::iCa * E + e] : 0.0;
        R.g = A.gjopen[m0 + lane];
    }
}

template <int NI, int WPC, int MINB, bool GREG>
__global__ void __launch_bounds__(WPC * 32, MINB)
k_mem_pipe(const __grid_constant__ KParams P, const KArrays A, const int cur)
{
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int W = gridDim.x * WPC;                    // tile stride: consecutive warps take consecutive tiles
    int tile = blockIdx.x * WPC + wib;
    const int nt = P.n_tiles;
    double* base = sm + (size_t)wib * KP_WARP(NI);
    // rings (buffer b = 0/1) by plain arithmetic: an indexed pointer array would live in local memory
    double* const K0 = base + 2;
    double* const G0 = K0 + 2 * KP_K(NI);
    double* const C0 = G0 + 2 * KP_G(NI);
    double* const s_aux = C0 + 2 * KP_C(NI);
#define KB(b) (K0 + (b) * KP_K(NI))
#define GB(b) (G0 + (b) * KP_G(NI))
#define CB(b) (C0 + (b) * KP_C(NI))
    const uint32_t bar0 = smem_u32(base);             // two mbarriers (one per iteration parity)

    constexpr int iNa = StdProf<NI>::iNa, iK = StdProf<NI>::iK, iCa = StdProf<NI>::iCa;
    const int nxt = cur ^ 1;
    const int C = P.n_cells, E = P.ny * P.nx;
    const int4* __restrict__ TD = reinterpret_cast<const int4*>(A.tile_desc);
    const int4 zero4 = make_int4(0, 0, 0, 0);
    unsigned int flags = 0;

    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    // tile-independent lane roles of the later phases: lane = (cell, ion) pair; slot-copy element p = lane + 32k
    const int q0c = lane / NI, q0i = lane - q0c * NI;
    const int o0 = q0i * C + q0c;
    int cpy[NI];                                       // staging index of element p = lane + 32 k of the [membrane][ion] slot block
#pragma unroll
    for (int k = 0; k < NI; ++k) { const int p = lane + 32 * k; const int m = p / NI; cpy[k] = m * KP_SST(NI) + (p - m * NI); }

    // ---- prologue = pseudo-iterations -2 (barrier 0: K(t)) and -1 (barrier 1: K(t+1)); then G(t)
    int4 td0 = (tile < nt) ? __ldg(TD + tile) : zero4;
    int4 td1 = (tile + W < nt) ? __ldg(TD + tile + W) : zero4;
    if (td0.w == 0) return;
    issue_K<NI>(A, tile, true, smem_u32(KB(0)), bar0, lane);
    issue_K<NI>(A, tile + W, tile + W < nt, smem_u32(KB(1)), bar0 + 8, lane);
    mbar_wait(bar0, 0);
    issue_G<NI, GREG>(A, td0, KB(0), GB(0), CB(0), lane, o0, C, E, cur);
    GReg<NI> R;
    if (GREG) load_G<NI>(A, td0, KB(0), R, lane, C, E, cur);

    int it = 0;
    for (; tile < nt; tile += W, ++it) {
        const int pb = it & 1;                         // buffer parity of the current tile
        // the constant block issued during iteration it-1 (barrier (it+1)&1, its use number (it+1)>>1) and the cp.asyncs
        mbar_wait(bar0 + 8 * (pb ^ 1), ((it + 1) >> 1) & 1);
        cp_wait_all();
        __syncwarp();
        const int c0 = td0.x, nc = td0.y, m0 = td0.z, nm = td0.w;
        const double* K = KB(pb);
        double* G = GB(pb);
        double* Cs = CB(pb);
        double* c_cc = Cs + KP_MAXC;                   // [cell][ion]; overwritten in place by the updated values
        const double* c_cmi = c_cc + NI * KP_MAXC;
        const int* Ki = reinterpret_cast<const int*>(K + KP_KD(NI));

        // ---- this tile's constants, state and gathers -> registers (the K block is overwritten below)
        int lc = 0, nnp = 0, jb0 = 0, je0 = 0;
        double sa = 0.0, g = 0.0, vm_nb = 0.0, cCao = 0.0, vol0 = 1.0, dvt = 0.0;
        double DmS[NI], co[NI], cnb[NI];
        const bool act = lane < nm;
        if (lane < nc * NI) { jb0 = Ki[96 + q0c] - m0; je0 = Ki[96 + q0c + 1] - m0; vol0 = K[(NI + 1) * 32 + q0c]; }
        if (lane < nc) dvt = K[(NI + 1) * 32 + KP_MAXC + lane];
        if (act) {
            lc = Ki[lane] - c0;
            nnp = Ki[32 + lane];
#pragma unroll
            for (int i = 0; i < NI; ++i) DmS[i] = K[i * 32 + lane];
            sa = K[NI * 32 + lane];
            if (GREG) {
#pragma unroll
                for (int i = 0; i < NI; ++i) { co[i] = R.co[i]; cnb[i] = R.cnb[i]; }
                vm_nb = R.vm_nb; cCao = R.cao; g = R.g;
            } else {
#pragma unroll
                for (int i = 0; i < NI; ++i) co[i] = G[i * 32 + lane];
#pragma unroll
                for (int i = 0; i < NI; ++i) cnb[i] = G[(NI + i) * 32 + lane];
                vm_nb = G[(2 * NI) * 32 + lane];
                if (iCa >= 0) cCao = G[(2 * NI + 1) * 32 + lane];
                g = G[(2 * NI + 2) * 32 + lane];
            }
        }
        if (nc * NI > 32) {                            // later rounds of the (cell, ion) phase: keep ptr/vol of all cells
            if (lane <= nc) reinterpret_cast<int*>(s_aux)[lane] = Ki[96 + lane];
            if (lane < nc) s_aux[6 + lane] = K[(NI + 1) * 32 + lane];
        }
        __syncwarp();
        // ---- keep the pipeline full: K(t+2) into the buffer just drained, G(t+1)
        const int4 td2 = (tile + 2 * W < nt) ? __ldg(TD + tile + 2 * W) : zero4;
        issue_K<NI>(A, tile + 2 * W, tile + 2 * W < nt, smem_u32(K), bar0 + 8 * pb, lane);
        issue_G<NI, GREG>(A, td1, KB(pb ^ 1), GB(pb ^ 1), CB(pb ^ 1), lane, o0, C, E, cur);

        double* s_m = G;                               // [32][KP_SST] f_mem*sa   (the consumed gather rows)
        double* s_g = G + 32 * KP_SST(NI);             // [32][KP_SST] f_gj*sa

        // ---- lanes = membranes
        if (act) {
            const int m = m0 + lane;
            double cin[NI];
#pragma unroll
            for (int i = 0; i < NI; ++i) cin[i] = c_cmi[lc * NI + i];
            const double vm_own = Cs[lc];
            double cCai = 0.0;
            if (iCa >= 0) cCai = c_cc[lc * NI + iCa];
            // membrane side (per-cell quantities, recomputed by every membrane lane here) and gap-junction side
            MemSide ms;
            mem_side(vm_own, P, ms);
            const GhkAB& tm = ms.t;
            const double keq = ms.keq;
            const double vgj0 = vm_nb - vm_own;
            const double ag1 = ((vgj0 + FLOAT_NONCE) * P.F) * P.inv_RT_p;
            GhkAB tg;
            ghk_table(ag1, tg);
            double gc1, gc2;                           // gating sub-step g' = g*gc1 + gc2 (gap_junction.py:56-72)
            gj_gate_map(vgj0, P, P.gj_block, gc1, gc2);
            const double sa_g = (nnp < 0) ? 0.0 : sa;   // no gap-junction flux at boundary membranes (sim.py:2199-2201)

            double fNa = 0.0, fK = 0.0;
            if (P.alpha_NaK > 0.0) {
                NaKCell nc_;
                nak_cell(cin[iNa], cin[iK], P, nc_);
                fNa = nak_flux(nc_, keq, co[iNa], co[iK], P.NaK_block, P);
                fK = -(2.0 / 3.0) * fNa;
                fNa = P.rho_pump * fNa;
                fK = P.rho_pump * fK;
            }
            double fCa = 0.0;
            if (iCa >= 0 && P.alpha_Ca > 0.0) {
                if (cCai != cCai || cCao != cCao) flags |= ST_NAN_CONC;
                if (cCai < 0.0) cCai = 0.0;
                if (cCao < 0.0) cCao = 0.0;
                CaCell cac;
                ca_cell(cCai, keq, P, cac);
                fCa = ca_flux(cac, cCao, P);
                fCa = P.rho_pump * fCa;
                fCa = P.rho_pump * fCa;                 // applied twice in the reference (sim.py:2141, 2155)
            }

#pragma unroll
            for (int i = 0; i < NI; ++i) {
                double Am, Bm, Ag, Bg;
                ghk_pick(tm, StdProf<NI>::z(i), Am, Bm);
                ghk_pick(tg, StdProf<NI>::z(i), Ag, Bg);
                double fsa = ghk_mem_flux(DmS[i], __dmul_rn(cin[i], Am), co[i], Bm);
                if (i == iNa) fsa = fma(fNa, sa, fsa);
                if (i == iK) fsa = fma(fK, sa, fsa);
                if (i == iCa) fsa = fma(fCa, sa, fsa);
                g = fma(g, gc1, gc2);                                      // once per ion (sim.py:1272 -> 2180-2183)
                s_m[lane * KP_SST(NI) + i] = fsa;
                s_g[lane * KP_SST(NI) + i] = ghk_gj_flux(P.Dgj_len[i], __dmul_rn(g, sa_g), cnb[i], Ag, cin[i], Bg);
            }
            A.gjopen[m] = g;
        }
        __syncwarp();
        // the next tile's per-membrane gathers, into registers (its constant block K(t+1) landed before this iteration)
        if (GREG && td1.w > 0) load_G<NI>(A, td1, KB(pb ^ 1), R, lane, C, E, cur);

        // ---- the warp's slice of the membrane->env exchange slots ([membrane][ion], contiguous)
        {
            double* __restrict__ dst = A.flux_slots + (size_t)m0 * NI + lane;
            const int np = nm * NI;
            double v[NI];
#pragma unroll
            for (int k = 0; k < NI; ++k) v[k] = s_m[cpy[k]];
#pragma unroll
            for (int k = 0; k < NI; ++k) if (lane + 32 * k < np) dst[32 * k] = v[k];
        }

        // ---- lanes = (cell, ion) pairs: membranes -> cells (update_Co + update_all_concs)
        for (int q = lane; q < nc * NI; q += 32) {
            int i = q0i, jb = jb0, je = je0;
            size_t oc = (size_t)(o0 + c0);
            double vol = vol0;
            if (q != lane) {
                const int qc = q / NI;
                i = q - qc * NI;
                jb = reinterpret_cast<const int*>(s_aux)[qc] - m0; je = reinterpret_cast<const int*>(s_aux)[qc + 1] - m0;
                vol = s_aux[6 + qc];
                oc = (size_t)i * C + c0 + qc;
            }
            const double* pm = s_m + jb * KP_SST(NI) + i;
            const double* pg = s_g + jb * KP_SST(NI) + i;
            const int n = je - jb;
            double Sm = 0.0, Sg = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { if (k < n) { Sm += pm[k * KP_SST(NI)]; Sg += pg[k * KP_SST(NI)]; } }
            for (int k = 8; k < n; ++k) { Sm += pm[k * KP_SST(NI)]; Sg += pg[k * KP_SST(NI)]; }
            const double rvol = fast_rcp(vol);
            double cm_new, cn_new;
            cell_conc_update(c_cc[q], Sm, Sg, rvol, P.dt, cm_new, cn_new);
            if (cn_new != cn_new) flags |= ST_NAN_CONC;
            if (cn_new < 0.0) { cn_new = 0.0; flags |= ST_NEG; }         // no_negs, sim.py:2111
            A.cc_cells[oc] = cn_new;
            A.cc_mid[nxt][oc] = cm_new;                                  // the stale cc_at_mem (quirk list)
            c_cc[q] = cn_new;
        }
        __syncwarp();

        // ---- lanes = cells: charge and Vmem (ion_current.py:19; sim.py:2027-2029)
        if (lane < nc) {
            const int c = c0 + lane;
            double rho = 0.0;
#pragma unroll
            for (int i = 0; i < NI; ++i) rho = fma(P.zF[i], c_cc[lane * NI + i], rho);
            if (A.extra_rho_cells) rho += ldg(A.extra_rho_cells + c);
            A.rho_cells[c] = rho;
            const double vmn = P.inv_cm * (rho * dvt);
            if (vmn != vmn) flags |= ST_NAN_VM;
            A.vm_cell[nxt][c] = vmn;
        }
        td0 = td1; td1 = td2;
    }
    cp_wait_all();
    if (flags) atomicOr(A.status, flags);
#undef KB
#undef GB
#undef CB
}

// ---------------------------------------------------------------------------- tile pack
// (Re)build the DmS rows of every tile's constant block from the canonical [ion][membrane] array:
// DmS = (Dm * -(rho_channel/tm)) * mem_sa, the operand order of kernels.cu:k_mem.  Runs after an
// upload of Dm_cells (init state, scheduled interventions) and after betse_set_schedule.
__global__ void k_pack_dm(const __grid_constant__ KParams P, const KArrays A)
{
    const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (tile >= P.n_tiles) return;
    const int4 td = __ldg(reinterpret_cast<const int4*>(A.tile_desc) + tile);
    const int ni = P.n_ions;
    double* K = reinterpret_cast<double*>(const_cast<char*>(A.tile_pack)) + (size_t)tile * ((ni + 1) * 32 + 2 * KP_MAXC + (96 + KP_MAXC + 4) / 2);
    if (lane < td.w) {
        const double Dtm = -(P.inv_tm * P.rho_channel);
        const double sa = A.mem_sa[td.z + lane];
        for (int i = 0; i < ni; ++i) K[i * 32 + lane] = (A.Dm[(size_t)i * P.n_mems_owned + td.z + lane] * Dtm) * sa;
    }
}

void launch_pack_dm(const KParams& P, const KArrays& A, cudaStream_t st)
{
    if (!A.tile_pack || P.n_tiles <= 0) return;
    k_pack_dm<<<(P.n_tiles + 7) / 8, 256, 0, st>>>(P, A);
}

// The constant part of every block (membrane areas, cell volumes, index rows), built on the device from the
// arrays betse_create uploaded anyway.  One warp per tile; unused slots stay zero (the pack is memset first).
__global__ void k_pack_const(const __grid_constant__ KParams P, const KArrays A)
{
    const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (tile >= P.n_tiles) return;
    const int4 td = __ldg(reinterpret_cast<const int4*>(A.tile_desc) + tile);
    const int c0 = td.x, nc = td.y, m0 = td.z, nm = td.w, ni = P.n_ions;
    const int kd = (ni + 1) * 32 + 2 * KP_MAXC;
    double* K = reinterpret_cast<double*>(const_cast<char*>(A.tile_pack)) + (size_t)tile * (kd + (96 + KP_MAXC + 4) / 2);
    int* Ki = reinterpret_cast<int*>(K + kd);
    if (lane < nm) {
        K[ni * 32 + lane] = A.mem_sa[m0 + lane];
        Ki[lane] = A.mem_to_cells[m0 + lane];
        Ki[32 + lane] = A.nn_cell_flag[m0 + lane];
        Ki[64 + lane] = A.map_mem2ecm[m0 + lane];
    }
    if (lane < nc) {
        K[(ni + 1) * 32 + lane] = A.cell_vol[c0 + lane];
        K[(ni + 1) * 32 + KP_MAXC + lane] = A.diviterm[c0 + lane];
    }
    if (lane <= nc) Ki[96 + lane] = A.cell_mem_ptr[c0 + lane];
}

void launch_pack_const(const KParams& P, const KArrays& A, cudaStream_t st)
{
    if (!A.tile_pack || P.n_tiles <= 0) return;
    k_pack_const<<<(P.n_tiles + 7) / 8, 256, 0, st>>>(P, A);
}

// byte size of one block (capi.cu:create_impl allocates n_tiles of them)
unsigned tile_pack_size(int ni) { return (unsigned)(((ni + 1) * 32 + 2 * KP_MAXC + (96 + KP_MAXC + 4) / 2) * 8); }

// ---------------------------------------------------------------------------- launch
// BETSE_KMEM_PIPE=0 falls back to the one-tile-per-warp kernel; BETSE_KMEM_WARPS = resident warps
// per SM (multiples of 4: CTAs of 4 warps).
static int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

bool kmem_pipe_enabled()
{
    static int v = -1;
    if (v < 0) v = env_int("BETSE_KMEM_PIPE", 1) ? 1 : 0;
    return v == 1;
}

template <int NI, int WPC, int MINB>
static cudaError_t prep_pipe()
{
    const int smem = (int)(WPC * KP_WARP(NI) * sizeof(double));
    cudaError_t e = cudaFuncSetAttribute(k_mem_pipe<NI, WPC, MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e) return e;
    if ((e = cudaFuncSetAttribute(k_mem_pipe<NI, WPC, MINB, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100))) return e;
    if ((e = cudaFuncSetAttribute(k_mem_pipe<NI, WPC, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))) return e;
    return cudaFuncSetAttribute(k_mem_pipe<NI, WPC, MINB, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}

template <int NI>
static cudaError_t prep_pipe_ni()
{
    cudaError_t e;
    if ((e = prep_pipe<NI, 4, 2>())) return e;
    if ((e = prep_pipe<NI, 4, 3>())) return e;
    if ((e = prep_pipe<NI, 5, 3>())) return e;
    if ((e = prep_pipe<NI, 4, 4>())) return e;
    return cudaSuccess;
}

cudaError_t prepare_mem_pipe(int ni)
{
    switch (ni) {
        case 4: return prep_pipe_ni<4>();
        case 5: return prep_pipe_ni<5>();
        case 6: return prep_pipe_ni<6>();
        case 7: return prep_pipe_ni<7>();
        default: return cudaSuccess;
    }
}

template <int NI, int WPC, int MINB>
static void launch_pipe_cfg(const KParams& P, const KArrays& A, int n_sms, int cur, cudaStream_t st)
{
    const size_t smem = WPC * KP_WARP(NI) * sizeof(double);
    int grid = n_sms * MINB;
    const int need = (P.n_tiles + WPC - 1) / WPC;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    static int greg = -1;
    if (greg < 0) greg = env_int("BETSE_KMEM_GREG", 1) ? 1 : 0;     // measured: k_mem 0.344 -> 0.335 ms at 1M cells (profiles/r01_sweeps.txt)
    if (greg) k_mem_pipe<NI, WPC, MINB, true><<<grid, WPC * 32, smem, st>>>(P, A, cur);
    else k_mem_pipe<NI, WPC, MINB, false><<<grid, WPC * 32, smem, st>>>(P, A, cur);
}

// resident warps per SM: 16 = 4 CTAs x 4 warps (128 registers), 15 = 3 x 5, 12 = 3 x 4, 8 = 2 x 4; a CTA count that
// does not fit in shared memory (7-ion profile) falls back to the next smaller one
template <int NI>
static void launch_pipe_t(const KParams& P, const KArrays& A, int n_sms, int cur, cudaStream_t st)
{
    static int warps = -1;
    if (warps < 0) warps = env_int("BETSE_KMEM_WARPS", 12);
    const size_t per_warp = KP_WARP(NI) * sizeof(double);
    const size_t cap = 227 * 1024;
    if (warps >= 16 && (4 * per_warp + 1024) * 4 <= cap + 4096) launch_pipe_cfg<NI, 4, 4>(P, A, n_sms, cur, st);
    else if (warps >= 15 && (5 * per_warp + 1024) * 3 <= cap + 3072) launch_pipe_cfg<NI, 5, 3>(P, A, n_sms, cur, st);
    else if (warps >= 12 && (4 * per_warp + 1024) * 3 <= cap + 3072) launch_pipe_cfg<NI, 4, 3>(P, A, n_sms, cur, st);
    else launch_pipe_cfg<NI, 4, 2>(P, A, n_sms, cur, st);
}

void launch_mem_pipe(int ni, const KParams& P, const KArrays& A, int n_sms, int cur, cudaStream_t st)
{
    switch (ni) {
        case 4: launch_pipe_t<4>(P, A, n_sms, cur, st); break;
        case 5: launch_pipe_t<5>(P, A, n_sms, cur, st); break;
        case 6: launch_pipe_t<6>(P, A, n_sms, cur, st); break;
        default: launch_pipe_t<7>(P, A, n_sms, cur, st); break;
    }
}
