// k_cell — the membranes->cells kernel of the tissue step with ONE LANE PER CELL (same arithmetic as
// kernels.cu:k_mem through the shared functions of kmath.cuh, bit-identical results).
//
// Why (profiles/r01i_*: the lane-per-membrane kernels issue ~28 warp instructions per membrane, two thirds of them
// staging, segmented sums and pipeline bookkeeping, and every membrane lane recomputes what belongs to its CELL):
//   * Vmem is a per-cell quantity in the default mode (sim.py:2029), so the membrane-side GHK table, the pumps'
//     equilibrium constant and every pump factor that depends on the cell's concentrations are formed ONCE per cell;
//   * a lane walks its cell's membranes k = 0..nm-1, so the membranes->cell sums (update_Co, sim_toolbox.py:1177)
//     are plain register accumulations in membrane order — no shared memory, no shuffles, no barriers;
//   * per-membrane constants live in a sliced-ELL "cell pack" (SELL-32: block b = cells 32b..32b+31, row k of the
//     block holds membrane k of each of its cells), so lane = cell reads them coalesced, and because neighbouring
//     cells have neighbouring partners and env squares, the gathers of one row touch two or three lines instead of 32;
//   * the loads of membrane k+1 (and the indices of k+2) are in flight while membrane k is computed: two register
//     buffers, no cp.async / mbarrier machinery.
//
// Reference lines as in k_mem: sim.py:1193-1283, 2086-2111, 2162-2206; sim_toolbox.py:18-182, 1155-1207;
// channels/gap_junction.py:53-77; ion_current.py:19; sim.py:2027-2029.
#include <stdlib.h>
#include <stdint.h>
#include "kmath.cuh"

#define KC_WARPS 4

template <int NI>
struct MemIn { double DmS[NI], co[NI], cnb[NI], vnb, cao, g, sa; int nnp; };
struct MemIdx { int nnp, esq; };

// hint: pull [p, p + bytes) into L2 (one bulk prefetch, no registers held): the streams of a task some hundred tasks
// ahead, so that the loads that later need them pay L2, not DRAM, latency
__device__ __forceinline__ void l2_prefetch(const void* p, unsigned bytes)
{
    if (bytes == 0) return;
    const unsigned long long a = (unsigned long long)p;
    const unsigned long long a0 = a & ~15ull;
    const unsigned n = (unsigned)(((a + bytes + 15ull) & ~15ull) - a0);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(n) : "memory");
}

// ---- one block of 32 cells: lane = cell
template <int NI, bool FUSE>
__device__ __forceinline__ void cell_task(const KParams& P, const KArrays& A, const int cur, const int task, const int lane,
                                          unsigned int& flags)
{
    constexpr int iNa = StdProf<NI>::iNa, iK = StdProf<NI>::iK, iCa = StdProf<NI>::iCa;
    const int nxt = cur ^ 1;
    const int C = P.n_cells, E = P.ny * P.nx;
    const size_t R32 = (size_t)P.ell_R32;
    const int2 h0 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + task);          // {first row, first membrane}
    const int2 h1 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + task + 1);
    const int row0 = h0.x, Kb = h1.x - h0.x;
    // header of the block whose streams this task pulls into L2 (issued after the first membrane, below)
    const int up = task + P.pf_dist;
    int2 u0 = make_int2(0, 0), u1 = make_int2(0, 0);
    if (P.pf_dist > 0 && up < P.n_blocks) {
        u0 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + up);
        u1 = __ldg(reinterpret_cast<const int2*>(A.blk_row0) + up + 1);
    }
    const int c = task * 32 + lane;
    const bool valid = c < P.n_cells_owned;
    int m_beg = 0, nm = 0;
    if (valid) { m_beg = ldgi(A.cell_mem_ptr + c); nm = ldgi(A.cell_mem_ptr + c + 1) - m_beg; }
    const unsigned e0 = (unsigned)row0 * 32u + (unsigned)lane;      // ELL element of membrane k = 0
    const double* __restrict__ cmid = A.cc_mid[cur];
    const double* __restrict__ vmc = A.vm_cell[cur];
    const double* __restrict__ cenv = A.cc_env[cur];
    const double* __restrict__ cenvCa = A.cc_env[nxt] + (size_t)(iCa >= 0 ? iCa : 0) * E;   // Ca after transport (sim.py:1282 after 2254)

    auto idx_load = [&](MemIdx& x, const int k) {
        if (k < nm) {
            x.nnp = __ldcs(A.ell_nnp + e0 + 32u * k);
            x.esq = __ldcs(A.ell_esq + e0 + 32u * k);
        }
    };
    auto gather = [&](MemIn<NI>& x, const MemIdx& ix, const int k) {
        if (k < nm) {
            const unsigned e = e0 + 32u * k;
#pragma unroll
            for (int i = 0; i < NI; ++i) x.DmS[i] = __ldcs(A.ell_DmS + i * R32 + e);
            x.sa = __ldcs(A.ell_sa + e);
            x.g = A.gjopen[m_beg + k];
            const unsigned cn = (unsigned)(ix.nnp & 0x7fffffff), q = (unsigned)ix.esq;
#pragma unroll
            for (int i = 0; i < NI; ++i) x.co[i] = (cenv + (size_t)i * E)[q];
#pragma unroll
            for (int i = 0; i < NI; ++i) x.cnb[i] = (cmid + (size_t)i * C)[cn];
            x.vnb = vmc[cn];
            x.cao = (iCa >= 0) ? cenvCa[q] : 0.0;
            x.nnp = ix.nnp;
        }
    };

    MemIdx ia, ib;
    ia.nnp = ia.esq = ib.nnp = ib.esq = 0;
    idx_load(ia, 0);
    idx_load(ib, 1);
    // ---- this cell
    double cc[NI], cin[NI], vm_own = 0.0, vol = 1.0, dvt = 0.0;
#pragma unroll
    for (int i = 0; i < NI; ++i) { cc[i] = 0.0; cin[i] = 0.0; }
    if (valid) {
        vm_own = vmc[c];
#pragma unroll
        for (int i = 0; i < NI; ++i) cin[i] = (cmid + (size_t)i * C)[c];
#pragma unroll
        for (int i = 0; i < NI; ++i) cc[i] = A.cc_cells[(size_t)i * C + c];
        vol = ldg(A.cell_vol + c);
        dvt = ldg(A.diviterm + c);
    }
    MemIn<NI> a, b;
    gather(a, ia, 0);

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
            // gap-junction side: vgj and its GHK table with p.T (sim.py:2166, 2197), gating sub-step g' = g*gc1 + gc2
            const double vgj0 = x.vnb - vm_own;
            const double ag1 = ((vgj0 + FLOAT_NONCE) * P.F) * P.inv_RT_p;
            GhkAB tg;
            ghk_table(ag1, tg);
            double gc1, gc2;
            gj_gate_map(vgj0, P, P.gj_block, gc1, gc2);
            const double sa_g = (x.nnp < 0) ? 0.0 : x.sa;   // no gap-junction flux at boundary membranes (sim.py:2199-2201)
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
            double g = x.g;
            double* __restrict__ fl = A.flux_ell + ((size_t)(row0 + k) * NI) * 32 + lane;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                double Ag, Bg;
                ghk_pick(tg, StdProf<NI>::z(i), Ag, Bg);
                double fsa = ghk_mem_flux(x.DmS[i], cinAm[i], x.co[i], Bm[i]);
                if (i == iNa) fsa = fma(fNa, x.sa, fsa);
                if (i == iK) fsa = fma(fK, x.sa, fsa);
                if (i == iCa) fsa = fma(fCa, x.sa, fsa);
                g = fma(g, gc1, gc2);                                      // once per ion (sim.py:1272 -> 2180-2183)
                const double fg = ghk_gj_flux(P.Dgj_len[i], __dmul_rn(g, sa_g), x.cnb[i], Ag, cin[i], Bg);
                Sm[i] = __dadd_rn(Sm[i], fsa);
                Sg[i] = __dadd_rn(Sg[i], fg);
                fl[i * 32] = fsa;
            }
            A.gjopen[m_beg + k] = g;
        }
    };

    // ---- the cell's membranes, two register buffers: membrane k is computed while k+1 (and the indices of k+2) load
#pragma unroll 1
    for (int k = 0; k < Kb; k += 2) {
        gather(b, ib, k + 1);
        idx_load(ia, k + 2);
        compute(a, k);
        if (k == 0 && u1.x > u0.x) {
            // streams of block `up`, one bulk prefetch per array and lane: cell-pack rows, gjopen, the cells' own state
            const unsigned rb = (unsigned)(u1.x - u0.x) * 256u;                 // bytes of the block's rows (doubles)
            const size_t r0 = (size_t)u0.x * 32;
            const int cu = up * 32;
            const int ncu = min(32, P.n_cells_owned - cu);
            if (lane < NI) l2_prefetch(A.ell_DmS + lane * R32 + r0, rb);
            else if (lane == NI) l2_prefetch(A.ell_sa + r0, rb);
            else if (lane == NI + 1) l2_prefetch(A.ell_nnp + r0, rb / 2);
            else if (lane == NI + 2) l2_prefetch(A.ell_esq + r0, rb / 2);
            else if (lane == NI + 3) l2_prefetch(A.gjopen + u0.y, (unsigned)(u1.y - u0.y) * 8u);
            else if (lane < 2 * NI + 4) l2_prefetch(A.cc_cells + (size_t)(lane - NI - 4) * C + cu, ncu * 8u);
            else if (lane < 3 * NI + 4) l2_prefetch(cmid + (size_t)(lane - 2 * NI - 4) * C + cu, ncu * 8u);
            else if (lane == 3 * NI + 4) l2_prefetch(vmc + cu, ncu * 8u);
            else if (lane == 3 * NI + 5) l2_prefetch(A.cell_vol + cu, ncu * 8u);
            else if (lane == 3 * NI + 6) l2_prefetch(A.diviterm + cu, ncu * 8u);
            else if (lane == 3 * NI + 7) l2_prefetch(A.cell_mem_ptr + cu, (ncu + 1) * 4u);
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
        // publish: the fluxes of this block are visible before its group's counter moves
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(A.cell_done + task / KC_GRP, 1);
    }
}

// ---- fused env task: membrane -> env exchange of KC_ENV_CHUNK squares (the body of k_envacc_ell), run inside the
//      membrane kernel once every cell block that feeds these squares has published its fluxes: the fluxes are then
//      read from L2, they need not wait for the whole tissue and a kernel boundary
template <int NI>
__device__ __forceinline__ void env_task(const KParams& P, const KArrays& A, const int nxt, const int v, const int lane)
{
    const int2 dep = __ldg(reinterpret_cast<const int2*>(A.env_dep) + v);       // groups [x, y] of cell tasks that feed the chunk
    if (lane == 0) {
        for (int g = dep.x; g <= dep.y; ++g) {
            const int want = min(KC_GRP, P.n_blocks - g * KC_GRP);
            const volatile int* ctr = A.cell_done + g;
            while (*ctr < want) __nanosleep(64);
        }
        __threadfence();
    }
    __syncwarp();
    const int E = P.nx * P.ny;
    const int base = v * KC_ENV_CHUNK;
#pragma unroll 1
    for (int j = 0; j < KC_ENV_CHUNK / 32; ++j) {
        const int k = base + j * 32 + lane;
        if (k >= E) break;
        const int s0 = ldgi(A.slot_ptr + k), s1 = ldgi(A.slot_ptr + k + 1);
        double acc[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) acc[i] = 0.0;
        for (int jj = s0; jj < s1; ++jj) {
            const int off = ldgi(A.slot_off + jj);
            // L2 loads: the producer's stores went to L2, this SM's L1 may hold nothing newer
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
            double c = A.cc_env[nxt][(size_t)i * E + k];
            const double delta_env = (-acc[i]) / P.env_vol_div;                     // sim_toolbox.py:1229
            c = c + delta_env * P.dt;
            A.cc_env[nxt][(size_t)i * E + k] = c;
            rho = fma(P.zF[i], c, rho);
        }
        if (A.extra_rho_env) rho += ldg(A.extra_rho_env + k);
        A.rho_env[k] = rho;
        A.v_raw[k] = (s1 > s0) ? ((rho * P.env_vol_div) / P.memsa_mean) / P.ko_eo_er : 0.0;   // ion_current.py:93-97
    }
}

// Persistent: every warp draws tickets; ticket t is block t of the cell pack, or (FUSE) entry t of the host-built
// schedule, which interleaves the env tasks a fixed lag behind the cell blocks that feed them.  A waiting env task only
// ever waits for cell tasks with SMALLER tickets, i.e. tasks that running warps already hold: no deadlock.
template <int NI, int MINB, bool FUSE>
__global__ void __launch_bounds__(KC_WARPS * 32, MINB)
k_cell(const __grid_constant__ KParams P, const KArrays A, const int cur)
{
    const int lane = threadIdx.x & 31;
    unsigned int flags = 0;
    const int n_tickets = FUSE ? P.n_sched : P.n_blocks;
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(A.ticket, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tickets) break;
        if (FUSE) {
            const int code = ldgi(A.sched + t);
            if (code < 0) env_task<NI>(P, A, cur ^ 1, code & 0x7fffffff, lane);
            else cell_task<NI, true>(P, A, cur, code, lane, flags);
        } else cell_task<NI, false>(P, A, cur, t, lane, flags);
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
    for (int k = 0; k < Kb; ++k) {
        const size_t e = (size_t)(row0 + k) * 32 + lane;
        double sa = 0.0;
        int nnp = (int)0x80000000, esq = 0;
        if (k < nm) {
            const int m = m_beg + k;
            sa = A.mem_sa[m]; nnp = A.nn_cell_flag[m]; esq = A.map_mem2ecm[m];
            mem_ell[m] = (int)(((size_t)(row0 + k) * ni) * 32 + lane);
        }
        const_cast<double*>(A.ell_sa)[e] = sa;
        const_cast<int*>(A.ell_nnp)[e] = nnp;
        const_cast<int*>(A.ell_esq)[e] = esq;
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
    const size_t R32 = (size_t)P.ell_R32;
    const double Dtm = -(P.inv_tm * P.rho_channel);
    for (int k = 0; k < Kb; ++k) {
        const size_t e = (size_t)(row0 + k) * 32 + lane;
        for (int i = 0; i < ni; ++i) {
            double v = 0.0;
            if (k < nm) {
                const int m = m_beg + k;
                v = __dmul_rn(__dmul_rn(A.Dm[(size_t)i * P.n_mems_owned + m], Dtm), A.mem_sa[m]);
            }
            const_cast<double*>(A.ell_DmS)[i * R32 + e] = v;
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
    if (!A.ell_sa || P.n_blocks <= 0) return;
    k_pack_cell_const<<<(P.n_blocks + 7) / 8, 256, 0, st>>>(P, A, mem_ell);
}

void launch_pack_cell_dm(const KParams& P, const KArrays& A, cudaStream_t st)
{
    if (!A.ell_DmS || P.n_blocks <= 0) return;
    k_pack_cell_dm<<<(P.n_blocks + 7) / 8, 256, 0, st>>>(P, A);
}

void launch_slot_off(const int* slot_idx, const int* mem_ell, int* slot_off, int n, int Mo, int ni, cudaStream_t st)
{
    if (n <= 0) return;
    k_slot_off<<<(n + 255) / 256, 256, 0, st>>>(slot_idx, mem_ell, slot_off, n, Mo, ni);
}

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
    double acc[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) acc[i] = 0.0;
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
        double c = A.cc_env[nxt][(size_t)i * E + k];
        const double delta_env = (-acc[i]) / P.env_vol_div;                     // sim_toolbox.py:1229
        c = c + delta_env * P.dt;
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

template <int NI, bool FUSE>
static void launch_cell_f(const KParams& P, const KArrays& A, int cur, cudaStream_t st)
{
    static int minb = -1;
    if (minb < 0) minb = kc_env_int("BETSE_KCELL_MINB", 2);      // resident CTAs (of 4 warps) per SM: 2 = 255 registers, 3 = 168, 4 = 128
    const int need = ((FUSE ? P.n_sched : P.n_blocks) + KC_WARPS - 1) / KC_WARPS;
    const int mb = minb <= 2 ? 2 : (minb == 3 ? 3 : 4);
    const int grid = need < g_kc_sms * mb ? need : g_kc_sms * mb;
    if (mb == 2) k_cell<NI, 2, FUSE><<<grid, KC_WARPS * 32, 0, st>>>(P, A, cur);
    else if (mb == 3) k_cell<NI, 3, FUSE><<<grid, KC_WARPS * 32, 0, st>>>(P, A, cur);
    else k_cell<NI, 4, FUSE><<<grid, KC_WARPS * 32, 0, st>>>(P, A, cur);
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
