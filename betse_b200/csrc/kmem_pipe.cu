// k_mem_pipe — the membranes->cells kernel of the tissue step as a persistent, software-pipelined
// warp kernel (same arithmetic as kernels.cu:k_mem; shared helpers in kmath.cuh).
//
// Why: k_mem is latency- and LSU-bound, not bandwidth-bound (ncu, profiles/r01d, r01e: 38 % issue
// utilisation, long-scoreboard stalls; a first cp.async-only pipeline moved the stall to the LSU:
// 50 LDGSTS per tile at ~8 cycles each).  A tile's inputs arrive in three dependent round trips
// (tile descriptor -> membrane index arrays -> gathers through those indices).  Here every warp
// walks a strided sequence of tiles and keeps the NEXT tiles' inputs in flight while it does the
// arithmetic of the current one:
//
//   S = per-membrane streams of a tile (Dm[I], mem_sa, gjopen, mem_to_cells, nn_cell_flag,
//       map_mem2ecm): contiguous rows -> ONE TMA bulk copy per row (cp.async.bulk, issued by one
//       lane per row, completion on an mbarrier), two tiles ahead;
//   C = per-cell block (cell_vol, diviterm, Vmem, cc_cells[I], cc_mid[I], cell_mem_ptr): rows
//       again, one tile ahead, same mbarrier;
//   G = gathers through the landed indices of S (env concentrations at the membrane's env square,
//       partner-cell concentrations and Vmem, transported Ca): cp.async (LDGSTS), one tile ahead.
//
//   iteration t:  wait (mbarrier of t-1, cp.async group)   -> S(t+1), C(t), G(t) have landed
//                 S(t), G(t) -> registers
//                 issue S(t+2) into S(t)'s buffer, C(t+1), G(t+1)
//                 arithmetic of tile t (staging of f*sa aliases the consumed G(t) buffer)
//
// Bulk copies need 16-byte aligned addresses and sizes: a row is fetched as the aligned superset of
// [start, start+n) and read back at the offset (start*esize mod 16)/esize.  Array bases are 16-byte
// aligned (cudaMalloc / 256-byte carved window) and allocations carry 16 bytes of slack (capi.cu).
//
// Reference lines as in k_mem: sim.py:1193-1283, 2086-2111, 2162-2206; sim_toolbox.py:18-182,
// 1155-1207; channels/gap_junction.py:53-77; ion_current.py:19; sim.py:2027-2029.
#include <stdlib.h>
#include <stdint.h>
#include "kmath.cuh"

#define KP_MAXC 10                                   // cells per tile (host packing, capi.cu)
#define KP_ROW8 34                                   // doubles per membrane row (32 + alignment slack; 272 B)
#define KP_ROW4 36                                   // ints per membrane index row (144 B)
#define KP_CROW8 12                                  // doubles per cell row (10 + slack; 96 B)
#define KP_CROW4 16                                  // ints of the cell_mem_ptr row (11 + slack; 64 B)
#define KP_SST 33                                    // staging stride (doubles): conflict-free f*sa [ion][membrane]
// per-warp shared memory, in doubles
#define KP_S(NI) (((NI) + 2) * KP_ROW8 + 3 * KP_ROW4 / 2)
#define KP_G(NI) ((2 * (NI) + 2) * 32)               // co[NI][32], cnb[NI][32], vnb[32], cao[32]; then staging 2*NI*33
#define KP_C(NI) ((3 + 2 * (NI)) * KP_CROW8 + KP_CROW4 / 2)
#define KP_WARP(NI) (2 + 2 * KP_S(NI) + 2 * KP_G(NI) + 2 * KP_C(NI) + (NI) * KP_MAXC)

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

// One lane = one row of the S or C block (set up once per kernel).
struct KRow {
    const char* src;   // row base in global memory (null: this lane has no row)
    int sh;            // log2(element size)
    int kind;          // 0: membrane row (S), 1: cell row (C), 2: cell_mem_ptr row (C, nc+1 elements)
    int dst;           // byte offset inside the S / C buffer
};

template <int NI>
__device__ __forceinline__ KRow make_row(const KArrays& A, const int lane, const int C, const int Mo, const int cur)
{
    KRow r;
    r.src = nullptr; r.sh = 3; r.kind = 0; r.dst = 0;
    if (lane < NI) { r.src = (const char*)(A.Dm + (size_t)lane * Mo); r.dst = lane * (KP_ROW8 * 8); }
    else if (lane == NI) { r.src = (const char*)A.mem_sa; r.dst = NI * (KP_ROW8 * 8); }
    else if (lane == NI + 1) { r.src = (const char*)A.gjopen; r.dst = (NI + 1) * (KP_ROW8 * 8); }
    else if (lane < NI + 5) {
        const int j = lane - (NI + 2);
        r.src = (const char*)(j == 0 ? A.mem_to_cells : j == 1 ? A.nn_cell_flag : A.map_mem2ecm);
        r.sh = 2; r.dst = (NI + 2) * (KP_ROW8 * 8) + j * (KP_ROW4 * 4);
    } else if (lane < 3 * NI + 9) {
        const int j = lane - (NI + 5);
        r.kind = 1; r.dst = j * (KP_CROW8 * 8);
        if (j == 0) r.src = (const char*)A.cell_vol;
        else if (j == 1) r.src = (const char*)A.diviterm;
        else if (j == 2) r.src = (const char*)A.vm_cell[cur];
        else if (j < 3 + NI) r.src = (const char*)(A.cc_cells + (size_t)(j - 3) * C);
        else if (j < 3 + 2 * NI) r.src = (const char*)(A.cc_mid[cur] + (size_t)(j - 3 - NI) * C);
        else { r.src = (const char*)A.cell_mem_ptr; r.sh = 2; r.kind = 2; }
    }
    return r;
}

// rows of S(tdS) -> Sdst and of C(tdC) -> Cdst, completion on `bar`
__device__ __forceinline__ void issue_rows(const KRow& r, const int4 tdS, const int4 tdC, const uint32_t Sdst,
                                           const uint32_t Cdst, const uint32_t bar, const int lane)
{
    int start, n;
    uint32_t dstb;
    if (r.kind == 0) { start = tdS.z; n = tdS.w; dstb = Sdst; }
    else { start = tdC.x; n = tdC.y ? tdC.y + (r.kind == 2 ? 1 : 0) : 0; dstb = Cdst; }
    if (r.src == nullptr) n = 0;
    const uintptr_t a = (uintptr_t)r.src + ((size_t)start << r.sh);
    const unsigned head = (unsigned)(a & 15);
    const unsigned bytes = n ? ((head + ((unsigned)n << r.sh) + 15u) & ~15u) : 0u;
    const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
    if (total) {
        if (lane == 0) mbar_expect_tx(bar, total);
        if (bytes) bulk_g2s(dstb + r.dst, (const void*)(a - head), bytes, bar);
    }
}

// gathers of tile td through the indices in its (landed) S buffer
template <int NI>
__device__ __forceinline__ void issue_G(const KArrays& A, const int4 td, const double* S, double* G, const int lane,
                                        const int C, const int E, const int cur)
{
    if (lane < td.w) {
        const int* Si = reinterpret_cast<const int*>(S + (NI + 2) * KP_ROW8) + (td.z & 3) + lane;
        const int cn = Si[KP_ROW4] & 0x7fffffff;
        const int e = Si[2 * KP_ROW4];
        const double* __restrict__ cenv = A.cc_env[cur] + e;
        const double* __restrict__ cmid = A.cc_mid[cur] + cn;
        const uint32_t g = smem_u32(G + lane);
#pragma unroll
        for (int i = 0; i < NI; ++i) cp8(g + i * 256, cenv + (size_t)i * E);
#pragma unroll
        for (int i = 0; i < NI; ++i) cp8(g + (NI + i) * 256, cmid + (size_t)i * C);
        cp8(g + (2 * NI) * 256, A.vm_cell[cur] + cn);
        if (StdProf<NI>::iCa >= 0) cp8(g + (2 * NI + 1) * 256, A.cc_env[cur ^ 1] + (size_t)StdProf<NI>::iCa * E + e);
    }
}

template <int NI, int WPC, int MINB>
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
    double* const S0 = base + 2;
    double* const G0 = S0 + 2 * KP_S(NI);
    double* const C0 = G0 + 2 * KP_G(NI);
    double* const s_cc = C0 + 2 * KP_C(NI);
#define SB(b) (S0 + (b) * KP_S(NI))
#define GB(b) (G0 + (b) * KP_G(NI))
#define CB(b) (C0 + (b) * KP_C(NI))
    const uint32_t bar0 = smem_u32(base);             // two mbarriers (one per iteration parity)

    constexpr int iNa = StdProf<NI>::iNa, iK = StdProf<NI>::iK, iCa = StdProf<NI>::iCa;
    const int nxt = cur ^ 1;
    const int C = P.n_cells, E = P.ny * P.nx, Mo = P.n_mems_owned;
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
    const KRow row = make_row<NI>(A, lane, C, Mo, cur);
    // tile-independent lane roles of the later phases: lane = (cell, ion) pair; slot-copy element p = lane + 32k
    const int q0c = lane / NI, q0i = lane - q0c * NI;
    int cpy[NI];                                       // staging index of element p = lane + 32 k of the [membrane][ion] slot block
#pragma unroll
    for (int k = 0; k < NI; ++k) { const int p = lane + 32 * k; const int m = p / NI; cpy[k] = (p - m * NI) * KP_SST + m; }

    // ---- prologue = pseudo-iterations -2 (barrier 0: S(t)) and -1 (barrier 1: S(t+1), C(t));
    //      then G(t) through the landed indices of S(t)
    int4 td0 = (tile < nt) ? __ldg(TD + tile) : zero4;
    int4 td1 = (tile + W < nt) ? __ldg(TD + tile + W) : zero4;
    int4 td2 = (tile + 2 * W < nt) ? __ldg(TD + tile + 2 * W) : zero4;
    if (td0.w == 0) return;
    issue_rows(row, td0, zero4, smem_u32(SB(0)), smem_u32(CB(0)), bar0, lane);
    issue_rows(row, td1, td0, smem_u32(SB(1)), smem_u32(CB(0)), bar0 + 8, lane);
    mbar_wait(bar0, 0);
    issue_G<NI>(A, td0, SB(0), GB(0), lane, C, E, cur);

    int it = 0;
    for (; tile < nt; tile += W, ++it) {
        const int pb = it & 1;                         // buffer parity of the current tile
        // everything issued during iteration it-1 (barrier (it+1)&1, its use number ((it+1)>>1)) and the gathers
        mbar_wait(bar0 + 8 * (pb ^ 1), ((it + 1) >> 1) & 1);
        cp_wait_all();
        __syncwarp();
        const int c0 = td0.x, nc = td0.y, m0 = td0.z, nm = td0.w;
        double* S = SB(pb);
        double* G = GB(pb);
        const double* Cc = CB(pb);
        // offsets of the aligned supersets (see the header)
        const int oS0 = m0 & 1, oS1 = (m0 + Mo) & 1, oC0 = c0 & 1, oC1 = (c0 + C) & 1;
        const int* c_ptr = reinterpret_cast<const int*>(Cc + (3 + 2 * NI) * KP_CROW8) + (c0 & 3);
        const double* c_vol = Cc + oC0;
        const double* c_dvt = Cc + KP_CROW8 + oC0;
        const double* c_vmo = Cc + 2 * KP_CROW8 + oC0;
        const double* c_cc = Cc + 3 * KP_CROW8;        // row i at i*KP_CROW8 + (i odd ? oC1 : oC0)
        const double* c_cmi = Cc + (3 + NI) * KP_CROW8;

        // ---- this tile's streams and gathers -> registers
        int lc = 0, nnp = 0;
        double sa = 0.0, g = 0.0, vm_nb = 0.0, cCao = 0.0;
        double Dm[NI], co[NI], cnb[NI];
        const bool act = lane < nm;
        if (act) {
            const int* Si = reinterpret_cast<const int*>(S + (NI + 2) * KP_ROW8) + (m0 & 3) + lane;
            lc = Si[0] - c0;
            nnp = Si[KP_ROW4];
#pragma unroll
            for (int i = 0; i < NI; ++i) Dm[i] = S[i * KP_ROW8 + ((i & 1) ? oS1 : oS0) + lane];
            sa = S[NI * KP_ROW8 + oS0 + lane];
            g = S[(NI + 1) * KP_ROW8 + oS0 + lane];
#pragma unroll
            for (int i = 0; i < NI; ++i) co[i] = G[i * 32 + lane];
#pragma unroll
            for (int i = 0; i < NI; ++i) cnb[i] = G[(NI + i) * 32 + lane];
            vm_nb = G[(2 * NI) * 32 + lane];
            if (iCa >= 0) cCao = G[(2 * NI + 1) * 32 + lane];
        }
        __syncwarp();
        // ---- keep the pipeline full: S(t+2) into the buffer just drained, C(t+1), G(t+1)
        const int4 td3 = (tile + 3 * W < nt) ? __ldg(TD + tile + 3 * W) : zero4;
        issue_rows(row, td2, td1, smem_u32(S), smem_u32(CB(pb ^ 1)), bar0 + 8 * pb, lane);
        issue_G<NI>(A, td1, SB(pb ^ 1), GB(pb ^ 1), lane, C, E, cur);

        double* s_m = G;                               // [NI][33] f_mem*sa   (the consumed gather buffer)
        double* s_g = G + NI * KP_SST;                 // [NI][33] f_gj*sa

        // ---- lanes = membranes
        if (act) {
            const int m = m0 + lane;
            const bool bnd = nnp < 0;
            double cin[NI];
#pragma unroll
            for (int i = 0; i < NI; ++i) cin[i] = c_cmi[i * KP_CROW8 + ((i & 1) ? oC1 : oC0) + lc];
            const double vm_own = c_vmo[lc];
            double cCai = 0.0;
            if (iCa >= 0) cCai = c_cc[iCa * KP_CROW8 + ((iCa & 1) ? oC1 : oC0) + lc];
            // membrane side: electroflux adds 1e-25 to vBA (sim_toolbox.py:54); alpha for z = +1
            const double a1 = ((vm_own + FLOAT_NONCE) * P.F) * P.inv_RT_sim;
            GhkAB tm;
            const double keq = P.K0 * fast_rcp(ghk_table(a1, tm));   // K0/e1 = exp(-dG/RT + F vm/RT): pump Keq
            // gap junction: vgj and its GHK table with p.T (sim.py:2166, 2197)
            const double vgj0 = vm_nb - vm_own;
            const double ag1 = ((vgj0 + FLOAT_NONCE) * P.F) * P.inv_RT_p;
            GhkAB tg;
            ghk_table(ag1, tg);
            double gc1, gc2;                           // gating sub-step g' = g*gc1 + gc2 (gap_junction.py:56-72)
            gj_gate_map(vgj0, P, P.gj_block, gc1, gc2);

            // ---- Na/K-ATPase (sim_toolbox.py:71-122)
            double fNa = 0.0, fK = 0.0;
            if (P.alpha_NaK > 0.0) {
                const double cNao = co[iNa], cNai = cin[iNa], cKo = co[iK], cKi = cin[iK];
                const double a = cNao * 1e-3, b = cKi * 1e-3;
                const double Qn = (P.QnNK0 * (a * a * a)) * (b * b);
                const double a2 = cNai * 1e-3, b2 = cKo * 1e-3;
                double Qd = (P.QdNK0 * (a2 * a2 * a2)) * (b2 * b2);
                if (Qd == 0.0) Qd = 1.0e-15;
                const double QdK = Qd * keq;
                const double u = cNai * P.inv_KmNK_Na, w = cKo * P.inv_KmNK_K, t = P.tNK;
                const double u3 = u * u * u, w2 = w * w;
                const double num = ((u3 * w2) * t) * (QdK - Qn);
                const double den = (((1.0 + u3) * (1.0 + w2)) * (1.0 + t)) * QdK;
                fNa = ((-3.0 * P.NaK_block) * P.alpha_NaK) * fast_div(num, den);
                fK = -(2.0 / 3.0) * fNa;
                fNa = P.rho_pump * fNa;
                fK = P.rho_pump * fK;
            }
            // ---- Ca-ATPase (sim.py:2126-2155, sim_toolbox.py:124-182)
            double fCa = 0.0;
            if (iCa >= 0 && P.alpha_Ca > 0.0) {
                if (cCai != cCai || cCao != cCao) flags |= ST_NAN_CONC;
                if (cCai < 0.0) cCai = 0.0;
                if (cCao < 0.0) cCao = 0.0;
                const double Qn = P.QnCa0 * cCao;
                double Qd = P.cATP * cCai;
                if (Qd == 0.0) Qd = 1.0e-16;
                const double QdK = Qd * ((keq * keq) * P.inv_K0);
                const double x = cCai * P.inv_KmCa_Ca, t = P.tCa;
                const double num = (x * t) * (QdK - Qn);
                const double den = ((1.0 + x) * (1.0 + t)) * QdK;
                fCa = -P.alpha_Ca * fast_div(num, den);
                fCa = P.rho_pump * fCa;
                fCa = P.rho_pump * fCa;                 // applied twice in the reference (sim.py:2141, 2155)
            }

            const double Dtm = -(P.inv_tm * P.rho_channel);
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                double Am, Bm, Ag, Bg;
                ghk_pick(tm, StdProf<NI>::z(i), Am, Bm);
                ghk_pick(tg, StdProf<NI>::z(i), Ag, Bg);
                double f = (Dm[i] * Dtm) * (cin[i] * Am - co[i] * Bm);     // sim_toolbox.py:58-65
                if (i == iNa) f += fNa;
                if (i == iK) f += fK;
                if (i == iCa) f += fCa;
                g = fma(g, gc1, gc2);                                      // once per ion (sim.py:1272 -> 2180-2183)
                double fg = -((P.Dgj_surf[i] * g) * P.inv_gjl) * (cnb[i] * Ag - cin[i] * Bg);   // sim.py:2191-2197
                if (bnd) fg = 0.0;
                s_m[i * KP_SST + lane] = f * sa;
                s_g[i * KP_SST + lane] = fg * sa;
            }
            A.gjopen[m] = g;
        }
        __syncwarp();

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
            int qc = q0c, i = q0i;
            if (q != lane) { qc = q / NI; i = q - qc * NI; }
            const int c = c0 + qc;
            const int jb = c_ptr[qc] - m0, je = c_ptr[qc + 1] - m0;
            const double* pm = s_m + i * KP_SST + jb;
            const double* pg = s_g + i * KP_SST + jb;
            const int n = je - jb;
            double vm_[8], vg_[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { vm_[k] = (k < n) ? pm[k] : 0.0; vg_[k] = (k < n) ? pg[k] : 0.0; }
            double Sm = 0.0, Sg = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { if (k < n) { Sm += vm_[k]; Sg += vg_[k]; } }
            for (int k = 8; k < n; ++k) { Sm += pm[k]; Sg += pg[k]; }
            const int oc = (i & 1) ? oC1 : oC0;
            const double rvol = fast_rcp(c_vol[qc]);
            const double cm_new = c_cc[i * KP_CROW8 + oc + qc] + (Sm * rvol) * P.dt;   // sim_toolbox.py:1177-1181
            double cn_new = cm_new + P.dt * ((-Sg) * rvol);                            // sim.py:2105-2108
            if (cn_new != cn_new) flags |= ST_NAN_CONC;
            if (cn_new < 0.0) { cn_new = 0.0; flags |= ST_NEG; }                       // no_negs, sim.py:2111
            A.cc_cells[(size_t)i * C + c] = cn_new;
            A.cc_mid[nxt][(size_t)i * C + c] = cm_new;                                 // the stale cc_at_mem (quirk list)
            s_cc[q] = cn_new;
        }
        __syncwarp();

        // ---- lanes = cells: charge and Vmem (ion_current.py:19; sim.py:2027-2029)
        if (lane < nc) {
            const int c = c0 + lane;
            double rho = 0.0;
#pragma unroll
            for (int i = 0; i < NI; ++i) rho = fma(P.zF[i], s_cc[lane * NI + i], rho);
            if (A.extra_rho_cells) rho += ldg(A.extra_rho_cells + c);
            A.rho_cells[c] = rho;
            const double vmn = P.inv_cm * (rho * c_dvt[lane]);
            if (vmn != vmn) flags |= ST_NAN_VM;
            A.vm_cell[nxt][c] = vmn;
        }
        td0 = td1; td1 = td2; td2 = td3;
    }
    cp_wait_all();
    if (flags) atomicOr(A.status, flags);
#undef SB
#undef GB
#undef CB
}

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

template <int NI, int MINB>
static cudaError_t prep_pipe()
{
    const int smem = (int)(4 * KP_WARP(NI) * sizeof(double));
    cudaError_t e = cudaFuncSetAttribute(k_mem_pipe<NI, 4, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e) return e;
    return cudaFuncSetAttribute(k_mem_pipe<NI, 4, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}

template <int NI>
static cudaError_t prep_pipe_ni()
{
    cudaError_t e;
    if ((e = prep_pipe<NI, 2>())) return e;
    if ((e = prep_pipe<NI, 3>())) return e;
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

template <int NI>
static void launch_pipe_t(const KParams& P, const KArrays& A, int n_sms, int cur, cudaStream_t st)
{
    const size_t smem = 4 * KP_WARP(NI) * sizeof(double);
    static int ctas_per_sm = -1;
    if (ctas_per_sm < 0) { ctas_per_sm = env_int("BETSE_KMEM_WARPS", 12) / 4; if (ctas_per_sm < 1) ctas_per_sm = 1; }
    int cps = ctas_per_sm;
    while (cps > 1 && (smem + 1024) * cps > 228 * 1024) --cps;   // what fits next to each other on one SM
    int grid = n_sms * cps;
    const int need = (P.n_tiles + 3) / 4;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (cps >= 3) k_mem_pipe<NI, 4, 3><<<grid, 128, smem, st>>>(P, A, cur);
    else k_mem_pipe<NI, 4, 2><<<grid, 128, smem, st>>>(P, A, cur);
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
