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
//     the gathers of one row touch two or three lines instead of 32;
//   * persistent warps draw tickets (blocks in order).  STAGED build: the whole block of the NEXT ticket — its rows are
//     contiguous — arrives by ONE TMA bulk copy (cp.async.bulk + mbarrier) in shared memory while the current block is
//     computed (12 KB per warp in flight all the time: the first build, which kept one membrane's loads in registers,
//     stalled at 59 % of the copy bandwidth for lack of bytes in flight, profiles/r02a_*); the gathers of membrane k+1
//     load into a second register buffer while membrane k is computed.  Register build (ragged meshes whose widest
//     block does not fit the stage): rows are loaded one membrane ahead, blocks further ahead are pulled into L2 by
//     bulk prefetches.
//
// Reference lines as in k_mem: sim.py:1193-1283, 2086-2111, 2162-2206; sim_toolbox.py:18-182, 1155-1207;
// channels/gap_junction.py:53-77; ion_current.py:19; sim.py:2027-2029.
#include <stdlib.h>
#include <stdint.h>
#include "kmath.cuh"

#define KC_WARPS 4
#define KC_ROWB(NI) (((NI) + 2) * 256)          // bytes of one row of the cell pack

template <int NI>
struct MemIn { double co[NI], cnb[NI], vnb, cao; int nnp; };

__device__ __forceinline__ uint32_t kc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kc_mbar_init(uint32_t bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void kc_mbar_expect_tx(uint32_t bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void kc_mbar_wait(uint32_t bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KC_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra KC_DONE;\n"
        "bra KC_WAIT;\n"
        "KC_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void kc_bulk_g2s(uint32_t dst, const void* src, unsigned bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// hint: pull [p, p + bytes) into L2 (one bulk prefetch, no registers held)
__device__ __forceinline__ void l2_prefetch(const void* p, unsigned bytes)
{
    if (bytes == 0) return;
    const unsigned long long a = (unsigned long long)p;
    const unsigned long long a0 = a & ~15ull;
    const unsigned n = (unsigned)(((a + bytes + 15ull) & ~15ull) - a0);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(n) : "memory");
}

// ---- one block of 32 cells: lane = cell.  STAGED: `stage` holds the block's rows and its gap-junction states.
template <int NI, bool STAGED, bool FUSE>
__device__ __forceinline__ void cell_task(const KParams& P, const KArrays& A, const int cur, const int task, const int lane,
                                          const char* __restrict__ stage, unsigned int& flags)
{
    constexpr int iNa = StdProf<NI>::iNa, iK = StdProf<NI>::iK, iCa = StdProf<NI>::iCa;
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
    // row k of the block: shared memory (staged) or the cell pack itself
    const char* __restrict__ rows = STAGED ? stage : (A.cpack + (size_t)row0 * ROWB);
    const double* __restrict__ gjs = STAGED ? reinterpret_cast<const double*>(stage + (size_t)P.kb_max * ROWB) + (m_beg - (h0.y & ~1))
                                            : (A.gjopen + m_beg);
    const double* __restrict__ cmid = A.cc_mid[cur];
    const double* __restrict__ vmc = A.vm_cell[cur];
    const double* __restrict__ cenv = A.cc_env[cur];
    const double* __restrict__ cenvCa = A.cc_env[nxt] + (size_t)(iCa >= 0 ? iCa : 0) * E;   // Ca after transport (sim.py:1282 after 2254)

    // the block whose streams this task pulls into L2 (register build; issued after the first membrane, below)
    const int up = task + P.pf_dist;
    int2 u0 = make_int2(0, 0), u1 = make_int2(0, 0);
    if (!STAGED && P.pf_dist > 0 && up < P.n_blocks) {
        u0 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + up);
        u1 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + up + 1);
    }

    // register build: the index pair of membrane k is loaded two membranes ahead (ia / ib), so that the gathers through
    // them do not wait for a second round trip; staged build: they are in shared memory
    int2 ia = make_int2(0, 0), ib = make_int2(0, 0);
    auto idx_load = [&](int2& x, const int k) {
        if (!STAGED && k < nm) {
            const char* r = rows + (size_t)k * ROWB;
            x.x = __ldcs(reinterpret_cast<const int*>(r + (NI + 1) * 256) + lane);
            x.y = __ldcs(reinterpret_cast<const int*>(r + (NI + 1) * 256 + 128) + lane);
        }
    };
    auto gather = [&](MemIn<NI>& x, const int2& ix, const int k) {
        if (k < nm) {
            const char* r = rows + (size_t)k * ROWB;
            const int nnp = STAGED ? reinterpret_cast<const int*>(r + (NI + 1) * 256)[lane] : ix.x;
            const unsigned q = (unsigned)(STAGED ? reinterpret_cast<const int*>(r + (NI + 1) * 256 + 128)[lane] : ix.y);
            const unsigned cn = (unsigned)(nnp & 0x7fffffff);
#pragma unroll
            for (int i = 0; i < NI; ++i) x.co[i] = (cenv + (size_t)i * E)[q];
#pragma unroll
            for (int i = 0; i < NI; ++i) x.cnb[i] = (cmid + (size_t)i * C)[cn];
            x.vnb = vmc[cn];
            x.cao = (iCa >= 0) ? cenvCa[q] : 0.0;
            x.nnp = nnp;
        }
    };

    // ---- this cell
    double cc[NI], cin[NI], vm_own = 0.0, vol = 1.0, dvt = 0.0;
#pragma unroll
    for (int i = 0; i < NI; ++i) { cc[i] = 0.0; cin[i] = 0.0; }
    MemIn<NI> a, b;
    idx_load(ia, 0);
    idx_load(ib, 1);
    gather(a, ia, 0);
    if (valid) {
        vm_own = vmc[c];
#pragma unroll
        for (int i = 0; i < NI; ++i) cin[i] = (cmid + (size_t)i * C)[c];
#pragma unroll
        for (int i = 0; i < NI; ++i) cc[i] = A.cc_cells[(size_t)i * C + c];
        vol = ldg(A.cell_vol + c);
        dvt = ldg(A.diviterm + c);
    }

    // ---- per-cell part of the flux math (kmath.cuh)
    MemSide ms;
    mem_side(vm_own, P, ms);
    double cinAm[NI], Bm[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        double Am;
        ghk_pick(ms.t, StdProf<NI>::z(i), Am, Bm[i]);
        cinAm[i] = __dmul_rn(cin[i], Am);
    }
    NaKCell nkc;
    nak_cell(cin[iNa], cin[iK], P, nkc);
    CaCell cac;
    cac.g1 = cac.g2 = 0.0;
    const bool ca_on = iCa >= 0 && P.alpha_Ca > 0.0;
    if (ca_on) {
        double cCai = cc[iCa >= 0 ? iCa : 0];          // the fresh cell value (update_intra, sim.py:2310)
        if (cCai != cCai) flags |= ST_NAN_CONC;
        if (cCai < 0.0) cCai = 0.0;
        ca_cell(cCai, ms.keq, P, cac);
    }
    double Sm[NI], Sg[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) { Sm[i] = 0.0; Sg[i] = 0.0; }

    auto compute = [&](const MemIn<NI>& x, const int k) {
        if (k < nm) {
            const double* __restrict__ r = reinterpret_cast<const double*>(rows + (size_t)k * ROWB) + lane;
            double DmS[NI];
#pragma unroll
            for (int i = 0; i < NI; ++i) DmS[i] = STAGED ? r[i * 32] : __ldcs(r + i * 32);
            const double sa = STAGED ? r[NI * 32] : __ldcs(r + NI * 32);
            double g = gjs[k];
            // gap-junction side: vgj and its GHK table with p.T (sim.py:2166, 2197), gating sub-step g' = g*gc1 + gc2
            const double vgj0 = x.vnb - vm_own;
            const double ag1 = ((vgj0 + FLOAT_NONCE) * P.F) * P.inv_RT_p;
            GhkAB tg;
            ghk_table(ag1, tg);
            double gc1, gc2;
            gj_gate_map(vgj0, P, P.gj_block, gc1, gc2);
            const double sa_g = (x.nnp < 0) ? 0.0 : sa;     // no gap-junction flux at boundary membranes (sim.py:2199-2201)
            double fNa = 0.0, fK = 0.0;
            if (P.alpha_NaK > 0.0) {
                fNa = nak_flux(nkc, ms.keq, x.co[iNa], x.co[iK], P.NaK_block, P);
                fK = -(2.0 / 3.0) * fNa;
                fNa = P.rho_pump * fNa;
                fK = P.rho_pump * fK;
            }
            double fCa = 0.0;
            if (ca_on) {
                double cCao = x.cao;
                if (cCao != cCao) flags |= ST_NAN_CONC;
                if (cCao < 0.0) cCao = 0.0;
                fCa = ca_flux(cac, cCao, P);
                fCa = P.rho_pump * fCa;
                fCa = P.rho_pump * fCa;                 // applied twice in the reference (sim.py:2141, 2155)
            }
            double* __restrict__ fl = A.flux_ell + ((size_t)(row0 + k) * NI) * 32 + lane;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                double Ag, Bg;
                ghk_pick(tg, StdProf<NI>::z(i), Ag, Bg);
                double fsa = ghk_mem_flux(DmS[i], cinAm[i], x.co[i], Bm[i]);
                if (i == iNa) fsa = fma(fNa, sa, fsa);
                if (i == iK) fsa = fma(fK, sa, fsa);
                if (i == iCa) fsa = fma(fCa, sa, fsa);
                g = fma(g, gc1, gc2);                                      // once per ion (sim.py:1272 -> 2180-2183)
                const double fg = ghk_gj_flux(P.Dgj_len[i], __dmul_rn(g, sa_g), x.cnb[i], Ag, cin[i], Bg);
                Sm[i] = __dadd_rn(Sm[i], fsa);
                Sg[i] = __dadd_rn(Sg[i], fg);
                fl[i * 32] = fsa;
            }
            A.gjopen[m_beg + k] = g;
        }
    };

    // ---- the cell's membranes, two register buffers: membrane k is computed while the gathers of k+1 load
#pragma unroll 1
    for (int k = 0; k < Kb; k += 2) {
        gather(b, ib, k + 1);
        idx_load(ia, k + 2);
        compute(a, k);
        if (!STAGED && k == 0 && u1.x > u0.x) {
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

    // ---- update_Co + update_all_concs, charge and Vmem of the cell (sim_toolbox.py:1177-1181; sim.py:2105-2111;
    //      ion_current.py:19; sim.py:2027-2029)
    if (valid) {
        const double rvol = fast_rcp(vol);
        double rho = 0.0;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            double cm_new, cn_new;
            cell_conc_update(cc[i], Sm[i], Sg[i], rvol, P.dt, cm_new, cn_new);
            if (cn_new != cn_new) flags |= ST_NAN_CONC;
            if (cn_new < 0.0) { cn_new = 0.0; flags |= ST_NEG; }         // no_negs, sim.py:2111
            A.cc_cells[(size_t)i * C + c] = cn_new;
            A.cc_mid[nxt][(size_t)i * C + c] = cm_new;                   // the stale cc_at_mem (quirk list)
            rho = fma(P.zF[i], cn_new, rho);
        }
        if (A.extra_rho_cells) rho += ldg(A.extra_rho_cells + c);
        A.rho_cells[c] = rho;
        const double vmn = P.inv_cm * (rho * dvt);
        if (vmn != vmn) flags |= ST_NAN_VM;
        A.vm_cell[nxt][c] = vmn;
    }
    if (FUSE) {
        // publish: the fluxes of this block are visible before its group's counter moves.  A RELEASE on the counter,
        // not __threadfence(): a gpu-scope acq_rel fence makes ptxas invalidate the SM's whole L1 (CCTL.IVALL), which the
        // other warps' gathers live on
        __syncwarp();
        if (lane == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(A.cell_done + task / KC_GRP) : "memory");
    }
}

// ---- fused env task: membrane -> env exchange of KC_ENV_CHUNK squares (the body of k_envacc_ell), run inside the
//      membrane kernel once every cell block that feeds these squares has published its fluxes
template <int NI>
__device__ __forceinline__ void env_task(const KParams& P, const KArrays& A, const int nxt, const int v, const int lane)
{
    const int2 dep = __ldg(reinterpret_cast<const int2*>(A.env_dep) + v);       // groups [x, y] of cell tasks that feed the chunk
    const int E = P.nx * P.ny;
    const int base = v * KC_ENV_CHUNK;
    constexpr int J = KC_ENV_CHUNK / 32;
    // everything that does not depend on the fluxes first: slot ranges and the transported concentrations
    int s0[J], s1[J];
    double cenv[J][NI];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int k = base + j * 32 + lane;
        s0[j] = s1[j] = 0;
        if (k < E) {
            s0[j] = ldgi(A.slot_ptr + k); s1[j] = ldgi(A.slot_ptr + k + 1);
#pragma unroll
            for (int i = 0; i < NI; ++i) cenv[j][i] = A.cc_env[nxt][(size_t)i * E + k];
        }
    }
    if (lane == 0) {
        for (int g = dep.x; g <= dep.y; ++g) {
            const int want = min(KC_GRP, P.n_blocks - g * KC_GRP);
            const volatile int* ctr = A.cell_done + g;
            while (*ctr < want) __nanosleep(100);
        }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int k = base + j * 32 + lane;
        if (k >= E) continue;
        double acc[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) acc[i] = 0.0;
        for (int jj = s0[j]; jj < s1[j]; ++jj) {
            const int off = ldgi(A.slot_off + jj);
            // L2 loads (ld.cg): the producers' stores are in L2, this SM's L1 may hold an older line
            if (off >= 0) {
#pragma unroll
                for (int i = 0; i < NI; ++i) acc[i] += __ldcg(A.flux_ell + (size_t)off + i * 32);
            } else {
                const double* __restrict__ f = A.flux_slots + (size_t)(-(off + 1));
#pragma unroll
                for (int i = 0; i < NI; ++i) acc[i] += __ldcg(f + i);
            }
        }
        double rho = 0.0;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const double delta_env = (-acc[i]) / P.env_vol_div;                     // sim_toolbox.py:1229
            const double c = cenv[j][i] + delta_env * P.dt;
            A.cc_env[nxt][(size_t)i * E + k] = c;
            rho = fma(P.zF[i], c, rho);
        }
        if (A.extra_rho_env) rho += ldg(A.extra_rho_env + k);
        A.rho_env[k] = rho;
        A.v_raw[k] = (s1[j] > s0[j]) ? ((rho * P.env_vol_div) / P.memsa_mean) / P.ko_eo_er : 0.0;   // ion_current.py:93-97
    }
}

// Persistent: every warp draws tickets; ticket t is block t of the cell pack, or (FUSE) entry t of the host-built
// schedule, which interleaves the env tasks a fixed lag behind the cell blocks that feed them.  A waiting env task only
// waits for cell tasks with SMALLER tickets; each of those is running, or is the next ticket of a warp whose current
// ticket is smaller still — by induction over the ticket order somebody always makes progress.
template <int NI, int MINB, bool STAGED, bool FUSE>
__global__ void __launch_bounds__(KC_WARPS * 32, MINB)
k_cell(const __grid_constant__ KParams P, const KArrays A, const int cur)
{
    extern __shared__ __align__(128) char kc_sm[];
    constexpr int ROWB = KC_ROWB(NI);
    const int lane = threadIdx.x & 31;
    unsigned int flags = 0;
    const int n_tickets = FUSE ? P.n_sched : P.n_blocks;
    auto claim = [&]() {
        int t = 0;
        if (lane == 0) t = atomicAdd(A.ticket, 1);
        return __shfl_sync(0xffffffffu, t, 0);
    };
    if (!STAGED) {
        for (;;) {
            const int t = P.kc_persist ? claim() : (int)(blockIdx.x * KC_WARPS + (threadIdx.x >> 5));
            if (t >= n_tickets) break;
            if (FUSE) {
                const int code = ldgi(A.sched + t);
                if (code < 0) env_task<NI>(P, A, cur ^ 1, code & 0x7fffffff, lane);
                else cell_task<NI, false, FUSE>(P, A, cur, code, lane, nullptr, flags);
            } else cell_task<NI, false, false>(P, A, cur, t, lane, nullptr, flags);
            if (!P.kc_persist) break;
        }
    } else {
        // per warp: two mbarriers, two stages of [kb_max rows | gap-junction states of the block]
        const size_t sst = (((size_t)P.kb_max * ROWB + ((size_t)P.kb_max * 32 + 2) * 8) + 127) & ~(size_t)127;
        char* base = kc_sm + (size_t)(threadIdx.x >> 5) * (128 + 2 * sst);
        const uint32_t bar = kc_smem_u32(base);
        if (lane == 0) {
            kc_mbar_init(bar, 1);
            kc_mbar_init(bar + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        // one TMA bulk copy brings a block's rows, a second its gap-junction states (from an even membrane index: 16-byte
        // alignment); ticket -> block through the schedule when fused (env tasks need no stage)
        auto task_of = [&](const int t) { return FUSE ? ldgi(A.sched + t) : t; };
        auto issue = [&](const int code, const int s) {
            if (code < 0) return;
            const int2 h0 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + code);
            const int2 h1 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + code + 1);
            if (lane == 0) {
                const unsigned rb = (unsigned)(h1.x - h0.x) * ROWB;
                const int ma = h0.y & ~1;
                const unsigned gb = (unsigned)(((h1.y - ma) * 8 + 15) & ~15);
                char* st = base + 128 + (size_t)s * sst;
                kc_mbar_expect_tx(bar + 8 * s, rb + gb);
                kc_bulk_g2s(kc_smem_u32(st), A.cpack + (size_t)h0.x * ROWB, rb, bar + 8 * s);
                kc_bulk_g2s(kc_smem_u32(st + (size_t)P.kb_max * ROWB), A.gjopen + ma, gb, bar + 8 * s);
            }
        };
        // cell blocks alternate between the two stages; at most two are in flight (the current and the next ticket)
        int t_cur = claim();
        if (t_cur < n_tickets) {
            int c_cur = task_of(t_cur);
            unsigned ph = 0;                               // phase parity of the two barriers, bit s
            int fill = 0, s_cur = 0;
            if (c_cur >= 0) { issue(c_cur, fill); s_cur = fill; fill ^= 1; }
            for (;;) {
                const int t_next = claim();
                const int c_next = t_next < n_tickets ? task_of(t_next) : -1;
                int s_next = 0;
                // the next block goes into the other stage (its last reader finished: __syncwarp below)
                if (t_next < n_tickets && c_next >= 0) { issue(c_next, fill); s_next = fill; fill ^= 1; }
                if (c_cur < 0) env_task<NI>(P, A, cur ^ 1, c_cur & 0x7fffffff, lane);
                else {
                    kc_mbar_wait(bar + 8 * s_cur, (ph >> s_cur) & 1u);
                    ph ^= 1u << s_cur;
                    cell_task<NI, true, FUSE>(P, A, cur, c_cur, lane, base + 128 + (size_t)s_cur * sst, flags);
                    __syncwarp();
                }
                if (t_next >= n_tickets) break;
                c_cur = c_next;
                s_cur = s_next;
            }
        }
    }
    if (flags) atomicOr(A.status, flags);
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
        int nnp = (int)0x80000000, esq = 0;
        if (k < nm) {
            const int m = m_beg + k;
            sa = A.mem_sa[m]; nnp = A.nn_cell_flag[m]; esq = A.map_mem2ecm[m];
            mem_ell[m] = (int)(((size_t)(row0 + k) * ni) * 32 + lane);
        }
        reinterpret_cast<double*>(r + (size_t)ni * 256)[lane] = sa;
        reinterpret_cast<int*>(r + (size_t)(ni + 1) * 256)[lane] = nnp;
        reinterpret_cast<int*>(r + (size_t)(ni + 1) * 256 + 128)[lane] = esq;
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
template <int NI>
__global__ void __launch_bounds__(256)
k_envacc_ell(const __grid_constant__ KParams P, const KArrays A, const int nxt)
{
    const int k = P.ya0 * P.nx + blockIdx.x * blockDim.x + threadIdx.x;
    const int E = P.nx * P.ny;
    if (k >= P.ya1 * P.nx) return;
    const int s0 = ldgi(A.slot_ptr + k), s1 = ldgi(A.slot_ptr + k + 1);
    double acc[NI], cv[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) { acc[i] = 0.0; cv[i] = A.cc_env[nxt][(size_t)i * E + k]; }
    for (int j = s0; j < s1; ++j) {
        const int off = ldgi(A.slot_off + j);
        if (off >= 0) {
#pragma unroll
            for (int i = 0; i < NI; ++i) acc[i] += A.flux_ell[(size_t)off + i * 32];
        } else {
            const double* __restrict__ f = A.flux_slots + (size_t)(-(off + 1));
#pragma unroll
            for (int i = 0; i < NI; ++i) acc[i] += f[i];
        }
    }
    double rho = 0.0;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const double delta_env = (-acc[i]) / P.env_vol_div;                     // sim_toolbox.py:1229
        const double c = cv[i] + delta_env * P.dt;
        A.cc_env[nxt][(size_t)i * E + k] = c;
        rho = fma(P.zF[i], c, rho);
    }
    if (A.extra_rho_env) rho += ldg(A.extra_rho_env + k);
    A.rho_env[k] = rho;
    A.v_raw[k] = (s1 > s0) ? ((rho * P.env_vol_div) / P.memsa_mean) / P.ko_eo_er : 0.0;   // ion_current.py:93-97
}

void launch_envacc_ell(int ni, const KParams& P, const KArrays& A, int nxt, cudaStream_t st)
{
    const int n = (P.ya1 - P.ya0) * P.nx;
    if (n <= 0) return;
    const int g = (n + 255) / 256;
    switch (ni) {
        case 4: k_envacc_ell<4><<<g, 256, 0, st>>>(P, A, nxt); break;
        case 5: k_envacc_ell<5><<<g, 256, 0, st>>>(P, A, nxt); break;
        case 6: k_envacc_ell<6><<<g, 256, 0, st>>>(P, A, nxt); break;
        default: k_envacc_ell<7><<<g, 256, 0, st>>>(P, A, nxt); break;
    }
}

// ---------------------------------------------------------------------------- launch
static int kc_env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

bool kcell_enabled()
{
    static int v = -1;
    if (v < 0) v = kc_env_int("BETSE_KCELL", 1) ? 1 : 0;
    return v == 1;
}

static int g_kc_sms = 148;

void kcell_set_sms(int n) { if (n > 0) g_kc_sms = n; }

// dynamic shared memory of the staged build: per warp two barriers (128 bytes) + two stages
static size_t kc_stage_smem(int ni, int kb_max)
{
    const size_t sst = (((size_t)kb_max * cell_pack_row_bytes(ni) + ((size_t)kb_max * 32 + 2) * 8) + 127) & ~(size_t)127;
    return KC_WARPS * (128 + 2 * sst);
}

// the staged build needs two CTAs (8 warps) per SM in 227 KB of shared memory, each CTA also costs 1 KB of system use
bool kcell_staged_fits(int ni, int kb_max)
{
    static int v = -1;
    if (v < 0) v = kc_env_int("BETSE_KCELL_STAGED", 1) ? 1 : 0;
    return v == 1 && kb_max > 0 && 2 * (kc_stage_smem(ni, kb_max) + 1024) <= (size_t)227 * 1024;
}

template <int NI>
static cudaError_t prep_cell_t(int kb_max)
{
    if (!kcell_staged_fits(NI, kb_max)) return cudaSuccess;
    const int smem = (int)kc_stage_smem(NI, kb_max);
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_cell<NI, 2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))) return e;
    if ((e = cudaFuncSetAttribute(k_cell<NI, 2, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100))) return e;
    if ((e = cudaFuncSetAttribute(k_cell<NI, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))) return e;
    return cudaFuncSetAttribute(k_cell<NI, 2, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}

// not capturable: once per context
cudaError_t prepare_cell(int ni, int kb_max)
{
    switch (ni) {
        case 4: return prep_cell_t<4>(kb_max);
        case 5: return prep_cell_t<5>(kb_max);
        case 6: return prep_cell_t<6>(kb_max);
        case 7: return prep_cell_t<7>(kb_max);
        default: return cudaSuccess;
    }
}

template <int NI, bool FUSE>
static void launch_cell_f(const KParams& P, const KArrays& A, int cur, cudaStream_t st)
{
    static int minb = -1;
    if (minb < 0) minb = kc_env_int("BETSE_KCELL_MINB", 2);      // register build: resident CTAs (of 4 warps) per SM, 2 = 255 registers, 3 = 168
    const int need = ((FUSE ? P.n_sched : P.n_blocks) + KC_WARPS - 1) / KC_WARPS;
    if (kcell_staged_fits(NI, P.kb_max)) {
        const int grid = need < g_kc_sms * 2 ? need : g_kc_sms * 2;
        k_cell<NI, 2, true, FUSE><<<grid, KC_WARPS * 32, kc_stage_smem(NI, P.kb_max), st>>>(P, A, cur);
        return;
    }
    const int mb = minb <= 2 ? 2 : 3;
    const int grid = (!P.kc_persist || need < g_kc_sms * mb) ? need : g_kc_sms * mb;
    if (mb == 2) k_cell<NI, 2, false, FUSE><<<grid, KC_WARPS * 32, 0, st>>>(P, A, cur);
    else k_cell<NI, 3, false, FUSE><<<grid, KC_WARPS * 32, 0, st>>>(P, A, cur);
}

template <int NI>
static void launch_cell_t(const KParams& P, const KArrays& A, int cur, int fuse, cudaStream_t st)
{
    if (fuse) launch_cell_f<NI, true>(P, A, cur, st);
    else launch_cell_f<NI, false>(P, A, cur, st);
}

// the caller zeroes A.ticket (and, fused, A.cell_done) on the same stream before every launch
void launch_cell(int ni, const KParams& P, const KArrays& A, int cur, int fuse, cudaStream_t st)
{
    switch (ni) {
        case 4: launch_cell_t<4>(P, A, cur, fuse, st); break;
        case 5: launch_cell_t<5>(P, A, cur, fuse, st); break;
        case 6: launch_cell_t<6>(P, A, cur, fuse, st); break;
        default: launch_cell_t<7>(P, A, cur, fuse, st); break;
    }
}
