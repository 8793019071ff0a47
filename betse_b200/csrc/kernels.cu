// Hand-written fp64 sm_100a kernels of the BETSE tissue step (one timestep of
// Simulator._run_sim_core_loop, betse/science/sim.py:1169-1365).
//
//   k_ion     env-grid electrodiffusion of every ion          (Simulator.update_ecm, sim.py:2209-2254)
//   k_mem     membranes -> cells: Na/K-ATPase, Ca-ATPase, GHK membrane flux, gap-junction gating
//             and GHK flux, membrane->cell segmented sums, concentration + charge + Vmem update
//             (sim.py:1193-1283, 2086-2111, 2162-2206; sim_toolbox.py:18-182, 1155-1207;
//              channels/gap_junction.py:53-77; ion_current.py:19; sim.py:2027-2029)
//   k_envacc  membrane -> env-grid exchange, env charge, raw env voltage
//             (sim_toolbox.py:1189-1234; ion_current.py:75-97)
//   k_field   9x9 separable Gaussian of the env voltage + its gradient (ion_current.py:101-109)
//   k_envmix  no-ECM bath mixing (sim_toolbox.py:1198-1205)
//   k_diag    sampled-step diagnostics (ion_current.py:22-47; sim.py:2016-2045, 2083)
//
// Layouts are SoA: [ion][cell], [ion][membrane], [ion][env point]; the membranes of a cell
// are contiguous, so a CTA owns a contiguous run of cells and of membranes.
#include <stdlib.h>
#include "kparams.cuh"

#include "kmath.cuh"
#include "xchg.cuh"

static int g_n_sms = 148;

// Shared-memory plan of k_mem, per warp (doubles): staged flux*sa [32][NI] for the membrane and
// the gap-junction flux (AoS: the warp's slice of flux_slots is a straight copy), and a scratch
// of 256 doubles (generic build: the GHK A/B tables of the two sides, [8][32]; then the updated
// concentrations of the tile's cells).
#define KM_WARP_DOUBLES(NI) (64 * (NI) + 256)
#define KM_SMEM_DOUBLES(NI) (8 * KM_WARP_DOUBLES(NI))

// One WARP per tile of whole cells with <= 32 membranes (host-built packing, tile_desc):
// lanes are membranes while the fluxes are formed, then (cell, ion) pairs for the membrane->cell
// sums (fixed summation order => bit-reproducible), then cells for charge and Vmem.  Warps never
// meet at a block barrier, so a warp stalled on a gather does not hold up its neighbours.
template <int NI, bool PHI, int MINB, int PROF>
__global__ void __launch_bounds__(BT_TPB, MINB)
k_mem(const __grid_constant__ KParams P, const KArrays A, const int cur, const int diag_in)
{
    constexpr bool S = (PROF != 0);
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (BT_TPB / 32) + (threadIdx.x >> 5);
    if (tile >= P.n_tiles) return;
    double* s_m = sm + (threadIdx.x >> 5) * KM_WARP_DOUBLES(NI);   // [32][NI] f_mem*sa
    double* s_g = s_m + 32 * NI;                                    // [32][NI] f_gj*sa
    double* s_t = s_g + 32 * NI;                                    // 256 doubles of scratch

    const bool is_ecm = S ? true : (P.is_ecm != 0);
    const bool vsens = S ? true : (P.v_sensitive_gj != 0);
    const bool cl_open = S ? true : (P.cluster_open != 0);
    const bool fast_ecm = S ? false : (P.fast_update_ecm != 0);
    const bool diag = S ? false : (diag_in != 0);
    const int iNa = S ? StdProf<NI>::iNa : P.iNa, iK = S ? StdProf<NI>::iK : P.iK, iCa = S ? StdProf<NI>::iCa : P.iCa;

    const int nxt = cur ^ 1;
    const int4 td = __ldg(reinterpret_cast<const int4*>(A.tile_desc) + tile);   // {c0, nc, m0, nm}
    // the tile some rounds ahead: its streamed inputs are pulled into L2 at the end of this warp's work
    int4 tf = make_int4(0, 0, 0, 0);
    if (P.pf_tiles > 0 && tile + P.pf_tiles < P.n_tiles) tf = __ldg(reinterpret_cast<const int4*>(A.tile_desc) + tile + P.pf_tiles);
    const int c0 = td.x, nc = td.y, m0 = td.z, nm = td.w;
    const int C = P.n_cells;
    const int E = P.ny * P.nx;
    const int Mo = P.n_mems_owned;
    const double* __restrict__ cmid = A.cc_mid[cur];
    const double* __restrict__ vmc = A.vm_cell[cur];
    const double* __restrict__ cenv = A.cc_env[cur];
    unsigned int flags = 0;

    // inputs of the later phases, requested now so that they are in flight during the flux math:
    // lane q = (cell, ion) pair (first round), lane = cell
    const int q0lc = lane / NI, q0i = lane - q0lc * NI;
    const bool q0ok = lane < nc * NI;
    int q0jb = 0, q0je = 0;
    double q0vol = 1.0, q0cc = 0.0, dvt = 0.0;
    if (q0ok) {
        q0jb = ldgi(A.cell_mem_ptr + c0 + q0lc) - m0; q0je = ldgi(A.cell_mem_ptr + c0 + q0lc + 1) - m0;
        q0vol = ldg(A.cell_vol + c0 + q0lc);
        q0cc = A.cc_cells[q0i * C + c0 + q0lc];
    }
    if (lane < nc) dvt = ldg(A.diviterm + c0 + lane);

    // ---- lanes = membranes
    if (lane < nm) {
        const int m = m0 + lane;
        const int c = ldgi(A.mem_to_cells + m);
        const int nnp = ldgi(A.nn_cell_flag + m);
        const int cn = nnp & 0x7fffffff;
        const bool bnd = nnp < 0;
        const int e = ldgi(A.map_mem2ecm + m);
        const double sa = ldg(A.mem_sa + m);
        double g = A.gjopen[m];
        double Dm[NI], co[NI], cnb[NI], cin[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) Dm[i] = ldg(A.Dm + i * Mo + m);
#pragma unroll
        for (int i = 0; i < NI; ++i) cin[i] = cmid[i * C + c];
#pragma unroll
        for (int i = 0; i < NI; ++i) co[i] = is_ecm ? cenv[i * E + e] : A.cenv_u[cur * 8 + i];
#pragma unroll
        for (int i = 0; i < NI; ++i) cnb[i] = cmid[i * C + cn];
        double vm_own = vmc[c];
        double vm_nb = vmc[cn];
        double cCai = 0.0, cCao = 0.0;
        if (iCa >= 0) {                               // Ca-ATPase inputs: fresh cell value, env after transport
            cCai = A.cc_cells[iCa * C + c];
            if (is_ecm) cCao = A.cc_env[nxt][iCa * E + e];
        }
        if (PHI) {
            if (P.polar) {                          // Vmem is membrane state (sim.py:2059-2061); vgj = vm[nn_i] - vm (sim.py:2166)
                vm_own = A.vm_pol[cur][m];
                vm_nb = A.vm_pol[cur][ldgi(A.nn_i + m)];
            } else {
                vm_own -= ldg(A.phi_b_old + e);     // sim.vm of the step's start: update_V of the previous step (sim.py:2029)
                vm_nb -= ldg(A.phi_b_old + ldgi(A.map_mem2ecm + ldgi(A.nn_i + m)));
            }
        }
        // membrane side: electroflux adds 1e-25 to vBA (sim_toolbox.py:54); alpha for z = +1; pump Keq = K0/e1
        MemSide ms;
        const double a1 = mem_side(vm_own, P, ms);
        const GhkAB& tm = ms.t;
        const double keq = ms.keq;
        // gap junction: vgj and its GHK table with p.T (sim.py:2166, 2197)
        const double vgj0 = vm_nb - vm_own;
        const double ag1 = ((vgj0 + FLOAT_NONCE) * P.F) * P.inv_RT_p;
        GhkAB tg;
        ghk_table(ag1, tg);
        if (!S) {
            s_t[0 * 32 + lane] = tm.A1; s_t[1 * 32 + lane] = tm.B1; s_t[2 * 32 + lane] = tm.A2; s_t[3 * 32 + lane] = tm.B2;
            s_t[4 * 32 + lane] = tg.A1; s_t[5 * 32 + lane] = tg.B1; s_t[6 * 32 + lane] = tg.A2; s_t[7 * 32 + lane] = tg.B2;
        }
        // gating (gap_junction.py:56-72): one implicit-Euler sub-step is the affine map g' = g*gc1 + gc2
        const double gjb = (!S && A.gj_block) ? ldg(A.gj_block + m) : P.gj_block;
        double gc1 = 0.0, gc2;
        if (vsens) gj_gate_map(vgj0, P, gjb, gc1, gc2);
        else gc2 = gjb * ldg(A.gj_w + m);             // static gap junctions, sim.py:2186
        const bool closed_bnd = (!cl_open) && bnd;

        // ---- Na/K-ATPase (sim_toolbox.py:71-122), kmath.cuh: nak_cell / nak_flux
        double fNa = 0.0, fK = 0.0;
        if (P.alpha_NaK > 0.0) {
            double cNai = 0.0, cKi = 0.0, cNao = 0.0, cKo = 0.0;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                if (i == iNa) { cNao = co[i]; cNai = cin[i]; }
                if (i == iK) { cKo = co[i]; cKi = cin[i]; }
            }
            NaKCell nc_;
            nak_cell(cNai, cKi, P, nc_);
            const double blk = (!S && A.NaK_block) ? ldg(A.NaK_block + m) : P.NaK_block;
            fNa = nak_flux(nc_, keq, cNao, cKo, blk, P);
            fK = -(2.0 / 3.0) * fNa;
            if (diag) A.rate_NaK[m] = -fNa;
            fNa = P.rho_pump * fNa;
            fK = P.rho_pump * fK;
            if (closed_bnd) { fNa = 0.0; fK = 0.0; }
        } else if (diag) A.rate_NaK[m] = 0.0;

        // ---- Ca-ATPase (sim.py:2126-2155, sim_toolbox.py:124-182), kmath.cuh: ca_cell / ca_flux
        double fCa = 0.0;
        if (iCa >= 0 && P.alpha_Ca > 0.0) {
            if (!is_ecm) {
#pragma unroll
                for (int i = 0; i < NI; ++i) if (i == iCa) cCao = co[i];
            }
            if (cCai != cCai || cCao != cCao) flags |= ST_NAN_CONC;
            if (cCai < 0.0) cCai = 0.0;
            if (cCao < 0.0) cCao = 0.0;
            CaCell cac;
            ca_cell(cCai, keq, P, cac);
            fCa = ca_flux(cac, cCao, P);
            fCa = P.rho_pump * fCa;
            if (closed_bnd) fCa = 0.0;
            fCa = P.rho_pump * fCa;                         // applied twice in the reference (sim.py:2141, 2155)
        }

        const double Dtm = -(P.inv_tm * P.rho_channel);
        const double sa_g = bnd ? 0.0 : sa;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            double Am, Bm, Ag, Bg;
            if (S) {
                ghk_pick(tm, StdProf<NI>::z(i), Am, Bm);
                ghk_pick(tg, StdProf<NI>::z(i), Ag, Bg);
            } else if (P.zi[i] != 0) {     // table slot of (A, B) for this valence: host-built ia/ib
                Am = s_t[P.ia[i] * 32 + lane]; Bm = s_t[P.ib[i] * 32 + lane];
                Ag = s_t[(4 + P.ia[i]) * 32 + lane]; Bg = s_t[(4 + P.ib[i]) * 32 + lane];
            } else {
                const double2 xm = ghk_generic(P.z[i], a1), xg = ghk_generic(P.z[i], ag1);
                Am = xm.x; Bm = xm.y; Ag = xg.x; Bg = xg.y;
            }
            // electroflux (sim_toolbox.py:58-65), cA = env, cB = cell; the flux is formed already times the
            // membrane area ((Dm*Dtm)*sa is what the cell pack of k_cell stores: same operand order)
            const double cinAm = __dmul_rn(cin[i], Am);
            const double DmS = __dmul_rn(__dmul_rn(Dm[i], Dtm), sa);
            double fsa = closed_bnd ? 0.0 : ghk_mem_flux(DmS, cinAm, co[i], Bm);
            if (i == iNa) fsa = fma(fNa, sa, fsa);
            if (i == iK) fsa = fma(fK, sa, fsa);
            if (i == iCa) fsa = fma(fCa, sa, fsa);
            // gap junction: gating advances once per ion (sim.py:1272 -> 2180-2183)
            g = fma(g, gc1, gc2);
            // cA = this cell, cB = partner cell (sim.py:2191-2197); zero at boundary membranes through sa_g
            s_m[lane * NI + i] = fsa;
            s_g[lane * NI + i] = ghk_gj_flux(P.Dgj_len[i], __dmul_rn(g, sa_g), cnb[i], Ag, cin[i], Bg);
            if ((is_ecm && fast_ecm) || diag) {     // per-area fluxes (fast ECM update, sampled-step diagnostics)
                const double pm_ = fma(-co[i], Bm, cinAm);
                double f = closed_bnd ? 0.0 : ((Dm[i] * Dtm) * pm_);
                if (i == iNa) f += fNa;
                if (i == iK) f += fK;
                if (i == iCa) f += fCa;
                if (is_ecm && fast_ecm) A.flux_slots[m * NI + i] = f;
                if (diag) { A.fl_mem[i * Mo + m] = f; A.fl_gj[i * Mo + m] = bnd ? 0.0 : ghk_gj_flux(P.Dgj_len[i], g, cnb[i], Ag, cin[i], Bg); }
            }
        }
        A.gjopen[m] = g;
    }
    __syncwarp();

    // ---- the warp's slice of the membrane->env exchange slots: a straight, coalesced copy
    if (is_ecm && !fast_ecm) {
        double* __restrict__ dst = A.flux_slots + m0 * NI;
        for (int p = lane; p < nm * NI; p += 32) dst[p] = s_m[p];
    }

    // ---- lanes = (cell, ion) pairs: membranes -> cells (update_Co + update_all_concs)
    double* s_cc = s_t;      // [nc][NI] updated concentrations (nc*NI <= 10*8 = 80)
    double* s_sm = s_t + 96; // [nc][NI] per-cell sum of f_mem*sa (no-ECM bath)
    for (int q = lane; q < nc * NI; q += 32) {
        const int lc = q / NI, i = q - lc * NI;
        const int c = c0 + lc;
        int jb = q0jb, je = q0je;
        double vol = q0vol, cc = q0cc;
        if (q != lane) {
            jb = ldgi(A.cell_mem_ptr + c) - m0; je = ldgi(A.cell_mem_ptr + c + 1) - m0;
            vol = ldg(A.cell_vol + c);
            cc = A.cc_cells[i * C + c];
        }
        double Sm = 0.0, Sg = 0.0;
        for (int j = jb; j < je; ++j) { Sm += s_m[j * NI + i]; Sg += s_g[j * NI + i]; }
        if (!S && P.defer) {              // channels act before update_all_concs: k_cell_update applies these
            A.dsum_m[i * C + c] = Sm; A.dsum_g[i * C + c] = Sg;
            if (!is_ecm) s_sm[q] = Sm;    // the bath still takes this step's membrane fluxes (k_envmix)
            continue;
        }
        const double rvol = fast_rcp(vol);
        double cm_new, cn_new;
        cell_conc_update(cc, Sm, Sg, rvol, P.dt, cm_new, cn_new);
        if (cn_new != cn_new) flags |= ST_NAN_CONC;
        if (cn_new < 0.0) { cn_new = 0.0; flags |= ST_NEG; }      // no_negs, sim.py:2111
        A.cc_cells[i * C + c] = cn_new;
        A.cc_mid[nxt][i * C + c] = cm_new;                        // the stale cc_at_mem (quirk list)
        s_cc[q] = cn_new;
        if (!is_ecm) s_sm[q] = Sm;
    }
    __syncwarp();

    // ---- lanes = cells: charge and Vmem (ion_current.py:19; sim.py:2027-2029)
    if (lane < nc && (S || !P.defer)) {
        const int c = c0 + lane;
        double rho = 0.0;
#pragma unroll
        for (int i = 0; i < NI; ++i) rho = fma(P.zF[i], s_cc[lane * NI + i], rho);
        if (A.extra_rho_cells) rho += ldg(A.extra_rho_cells + c);
        A.rho_cells[c] = rho;
        const double vmn = P.inv_cm * (rho * dvt);
        if (vmn != vmn) flags |= ST_NAN_VM;
        A.vm_cell[nxt][c] = vmn;
    }
    if (!is_ecm && lane < NI) {      // per-tile partial of sum_m f*sa for the bath (no-ECM)
        double s = 0.0;
        for (int lc = 0; lc < nc; ++lc) s += s_sm[lc * NI + lane];
        A.cenv_part[tile * 8 + lane] = s;
    }
    if (flags) atomicOr(A.status, flags);

    // ---- L2 prefetch of a later tile's streamed membrane arrays (<= 32 membranes: <= 3 lines of
    // doubles, <= 2 of int32 per array; lane = line)
    if (tf.w > 0 && lane < 3) {
        const int mf = tf.z;
        const size_t o8 = ((size_t)mf * 8 & ~(size_t)127) + (size_t)lane * 128;
#pragma unroll
        for (int i = 0; i < NI; ++i) prefetch_l2((const char*)(A.Dm + i * Mo) + o8);
        prefetch_l2((const char*)A.mem_sa + o8);
        prefetch_l2((const char*)A.gjopen + o8);
        if (lane < 2) {
            const size_t o4 = ((size_t)mf * 4 & ~(size_t)127) + (size_t)lane * 128;
            prefetch_l2((const char*)A.mem_to_cells + o4);
            prefetch_l2((const char*)A.nn_cell_flag + o4);
            prefetch_l2((const char*)A.map_mem2ecm + o4);
        }
    }
}

// no-ECM: cX_env = mean(cX_env + (-flux*(mem_sa/vol_env))*dt)  (sim_toolbox.py:1200-1205)
__global__ void k_envmix(const __grid_constant__ KParams P, const KArrays A, const int cur)
{
    __shared__ double red[256];
    const int i = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < P.n_tiles; b += blockDim.x) s += A.cenv_part[(size_t)b * 8 + i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double c = A.cenv_u[cur * 8 + i];
        const double mean_delta = ((-red[0] / P.vol_env) * P.dt) / (double)P.n_mems_owned;
        A.cenv_u[(cur ^ 1) * 8 + i] = c + mean_delta;
    }
}

// ---------------------------------------------------------------------------- env grid
// Env-grid electrodiffusion of every ion (Simulator.update_ecm, sim.py:2209-2254), tiled through
// shared memory: a CTA owns IT_Y x IT_X outputs; per ion it stages the concentration tile with a
// 2-point halo (Dirichlet edge fill applied on load), forms the Nernst-Planck flux once per point
// of tile + 1-point halo (central gradient, one-sided at the world edge; last step's E field,
// staged once for all ions), then takes fd.divergence with fd.diff's edge convention and steps
// forward Euler.  Rows are local rows of a strip [y0, y0+ny) of the world; only rows
// [yi0, yi1) are produced.
#define IT_X 32
#define IT_Y 16
// blockIdx.z selects the ion: ions [ion0 + z] (a CTA per ion and tile keeps small strips busy and lets the Ca row,
// which the Ca-ATPase reads after transport, run ahead of the others).  INTERIOR tiles — tile + 2-point halo strictly
// inside the world and off its Dirichlet ring, all output rows wanted — take a build without any bounds / edge logic
// (round 1 spent ~300 thread-instructions per point-ion, most of them on those tests: 1.7 TB/s).
template <bool INTERIOR>
__device__ __forceinline__ void ion_tile(const KParams& P, const KArrays& A, const int cur, const int diag, const int i,
                                         double (*sC)[IT_X + 4], double (*sFx)[IT_X + 3], double (*sFy)[IT_X + 3])
{
    const int nx = P.nx, ny = P.ny;
    const int E = nx * ny;
    const int tx0 = blockIdx.x * IT_X, ty0 = P.yi0 + blockIdx.y * IT_Y;
    const int tid = threadIdx.x;
    const int gny = P.ny_global;
    const double inv_d = P.inv_delta, inv_2d = P.inv_2delta;
    const double* __restrict__ c = A.cc_env[cur] + (size_t)i * E;
    const double* __restrict__ D = A.Denv + (size_t)i * E;
    const double cb = P.cbound[i];
    const double zq = P.z[i] * P.q;

    for (int t = tid; t < (IT_Y + 4) * (IT_X + 4); t += 256) {
        const int ly = t / (IT_X + 4), lx = t - ly * (IT_X + 4);
        const int y = ty0 + ly - 2, x = tx0 + lx - 2;
        double v = 0.0;
        if (INTERIOR) v = c[y * nx + x];
        else if (x >= 0 && x < nx && y >= 0 && y < ny) {
            const int g = y + P.y0;
            const bool edge = (g == 0 || g == gny - 1 || x == 0 || x == nx - 1);
            v = edge ? cb : c[y * nx + x];          // Dirichlet fill, sim.py:2211-2217
        }
        sC[ly][lx] = v;
    }
    // interior tiles: the diffusion constant and the field at this thread's flux points are requested BEFORE the barrier, so
    // that one memory round trip serves both phases (the kernel is latency-bound: profiles/r02k_*: 11 long-scoreboard stalls
    // per issue)
    constexpr int NFP = ((IT_Y + 2) * (IT_X + 2) + 255) / 256;
    double pD[NFP], pEx[NFP], pEy[NFP];
    if (INTERIOR) {
#pragma unroll
        for (int j = 0; j < NFP; ++j) {
            const int t = tid + j * 256;
            pD[j] = pEx[j] = pEy[j] = 0.0;
            if (t < (IT_Y + 2) * (IT_X + 2)) {
                const int ly = t / (IT_X + 2), lx = t - ly * (IT_X + 2);
                const int k = (ty0 + ly - 1) * nx + (tx0 + lx - 1);
                pD[j] = D[k]; pEx[j] = A.E_x[k]; pEy[j] = A.E_y[k];
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NFP; ++j) {
        const int t = tid + j * 256;
        if (t >= (IT_Y + 2) * (IT_X + 2)) break;
        const int ly = t / (IT_X + 2), lx = t - ly * (IT_X + 2);
        const int y = ty0 + ly - 1, x = tx0 + lx - 1;
        double fx = 0.0, fy = 0.0;
        if (INTERIOR || (x >= 0 && x < nx && y >= 0 && y < ny)) {
            const int g = y + P.y0;
            const int k = y * nx + x;
            const double cc = sC[ly + 1][lx + 1];
            double gcx, gcy;                                   // fd.gradient, finitediff.py:1236-1266
            if (!INTERIOR && x == 0) gcx = (sC[ly + 1][lx + 2] - cc) * inv_d;
            else if (!INTERIOR && x == nx - 1) gcx = (cc - sC[ly + 1][lx]) * inv_d;
            else gcx = -(sC[ly + 1][lx] - sC[ly + 1][lx + 2]) * inv_2d;
            if (!INTERIOR && g == 0) gcy = (sC[ly + 2][lx + 1] - cc) * inv_d;
            else if (!INTERIOR && g == gny - 1) gcy = (cc - sC[ly][lx + 1]) * inv_d;
            else gcy = -(sC[ly][lx + 1] - sC[ly + 2][lx + 1]) * inv_2d;
            const double Dk = INTERIOR ? pD[j] : D[k];
            const double Exk = INTERIOR ? pEx[j] : A.E_x[k], Eyk = INTERIOR ? pEy[j] : A.E_y[k];
            const double al = (Dk * zq) * P.inv_kbT_sim;       // nernst_planck_flux, sim_toolbox.py:409-411
            fx = -Dk * gcx - (al * (-Exk)) * cc;
            fy = -Dk * gcy - (al * (-Eyk)) * cc;
            if (diag && ly >= 1 && ly <= IT_Y && lx >= 1 && lx <= IT_X && y < P.yi1) {
                A.fl_env_x[(size_t)i * E + k] = fx;
                A.fl_env_y[(size_t)i * E + k] = fy;
            }
        }
        sFx[ly][lx] = fx; sFy[ly][lx] = fy;
    }
    __syncthreads();
    // fd.divergence(-fx, -fy) with fd.diff's edge rows (finitediff.py:1268-1311)
#pragma unroll
    for (int h = 0; h < IT_Y / 8; ++h) {
        const int oy = (tid >> 5) + 8 * h, ox = tid & 31;
        const int y = ty0 + oy, x = tx0 + ox;
        if (INTERIOR || (x < nx && y < P.yi1)) {
            const int g = y + P.y0;
            const int ly = oy + 1, lx = ox + 1;
            double dx, dy;
            if (!INTERIOR && x == 0) dx = ((-sFx[ly][lx]) - (-sFx[ly][lx + 1])) * inv_d;
            else if (!INTERIOR && x == nx - 1) dx = ((-sFx[ly][lx - 1]) - (-sFx[ly][lx])) * inv_d;
            else dx = -((-sFx[ly][lx - 1]) - (-sFx[ly][lx + 1])) * inv_2d;
            if (!INTERIOR && g == 0) dy = -((-sFy[ly + 1][lx]) - (-sFy[ly][lx])) * inv_d;
            else if (!INTERIOR && g == gny - 1) dy = -((-sFy[ly][lx]) - (-sFy[ly - 1][lx])) * inv_d;
            else dy = -((-sFy[ly - 1][lx]) - (-sFy[ly + 1][lx])) * inv_2d;
            A.cc_env[cur ^ 1][(size_t)i * E + y * nx + x] = sC[oy + 2][ox + 2] + (dx + dy) * P.dt;
        }
    }
}

__global__ void __launch_bounds__(256)
k_ion(const __grid_constant__ KParams P, const KArrays A, const int cur, const int diag, const int ion0)
{
    __shared__ double sC[IT_Y + 4][IT_X + 4];
    __shared__ double sFx[IT_Y + 2][IT_X + 3], sFy[IT_Y + 2][IT_X + 3];
    const int tx0 = blockIdx.x * IT_X, ty0 = P.yi0 + blockIdx.y * IT_Y;
    const int g0 = ty0 + P.y0;
    // tile + halo 2 inside the local rows and strictly inside the world's Dirichlet ring; every output row wanted
    const bool interior = tx0 - 2 >= 1 && tx0 + IT_X + 1 <= P.nx - 2 && g0 - 2 >= 1 && g0 + IT_Y + 1 <= P.ny_global - 2 &&
                          ty0 - 2 >= 0 && ty0 + IT_Y + 1 <= P.ny - 1 && ty0 + IT_Y <= P.yi1;
    if (interior) ion_tile<true>(P, A, cur, diag, ion0 + blockIdx.z, sC, sFx, sFy);
    else ion_tile<false>(P, A, cur, diag, ion0 + blockIdx.z, sC, sFx, sFy);
}

// fd.integrator (finitediff.py:1479-1512) applied to the transported field (sim.py:2249-2252).
__global__ void k_ion_smooth(const __grid_constant__ KParams P, const KArrays A, const int nxt, const int n_ions)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int nx = P.nx, ny = P.ny;
    if (x >= nx || y >= ny) return;
    const int E = nx * ny, k = y * nx + x;
    const int gy = y + P.y0;
    const bool edge = (gy == 0 || gy == P.ny_global - 1 || x == 0 || x == nx - 1);
    const double sh = P.sharpness, sides = (1.0 - sh) / 4.0;
    for (int i = 0; i < n_ions; ++i) {
        const double* __restrict__ c = A.cc_env[nxt] + (size_t)i * E;
        double v = c[k];
        if (!edge) {
            // F = sharp*P; F[0:-1] += s*nP; F[1:] += s*sP; F[:,0:-1] += s*eP; F[:,1:] += s*wP
            double f = sh * v;
            if (y + 1 < ny) f += sides * c[k + nx];
            if (y > 0) f += sides * c[k - nx];
            f += sides * c[k + 1];
            f += sides * c[k - 1];
            v = f;
        }
        A.scratch_env[i * E + k] = v;
    }
}

// membrane -> env exchange (update_Co env branch + div_env), env charge, raw env voltage.
template <int NI>
__global__ void __launch_bounds__(256)
k_envacc(const __grid_constant__ KParams P, const KArrays A, const int nxt, const int apply_flux)
{
    const int k = P.ya0 * P.nx + blockIdx.x * blockDim.x + threadIdx.x;
    const int E = P.nx * P.ny;
    if (k >= P.ya1 * P.nx) return;
    const int s0 = ldgi(A.slot_ptr + k), s1 = ldgi(A.slot_ptr + k + 1);
    double acc[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) acc[i] = 0.0;
    if (!apply_flux) {
        // charge / raw voltage only (the update_V call that precedes the loop, sim.py:1041)
    } else if (P.fast_update_ecm) {
        if (s1 > s0) {                       // flux_env[map_mem2ecm] = flux: the last writer wins
            const int s = ldgi(A.slot_idx + s1 - 1);
#pragma unroll
            for (int i = 0; i < NI; ++i) acc[i] = A.flux_slots[s * NI + i];
        }
    } else {
        for (int j = s0; j < s1; ++j) {
            const int s = ldgi(A.slot_idx + j);
#pragma unroll
            for (int i = 0; i < NI; ++i) acc[i] += A.flux_slots[s * NI + i];
        }
    }
    double rho = 0.0;
    const double msa = P.fast_update_ecm ? ldg(A.memsa_env + k) : 0.0;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        double c = A.cc_env[nxt][i * E + k];
        double delta_env;
        if (P.fast_update_ecm) delta_env = ((-acc[i]) * msa) / P.ecm_vol;        // sim_toolbox.py:1222-1225
        else delta_env = (-acc[i]) / P.env_vol_div;                             // sim_toolbox.py:1229
        if (apply_flux) {
            c = c + delta_env * P.dt;
            A.cc_env[nxt][i * E + k] = c;
        }
        rho = fma(P.zF[i], c, rho);
    }
    if (A.extra_rho_env) rho += ldg(A.extra_rho_env + k);
    A.rho_env[k] = rho;
    // ion_current.py:93-97
    A.v_raw[k] = (s1 > s0) ? ((rho * P.env_vol_div) / P.memsa_mean) / P.ko_eo_er : 0.0;
}

// v_env = gaussian_filter(v_raw, sigma=1, mode='constant') + Phi_b ; E = -grad(screen*v_env)
// Shared-memory tile: 32x32 outputs, halo 5 (4 Gaussian + 1 gradient); zero fill outside the
// world reproduces scipy's mode='constant', cval=0 in both separable passes.
#define FH 5
// FT: 32 for large grids; 16 when 32x32 tiles would leave most SMs without a CTA (small tissues, thin strips of a
// decomposed tissue: the kernel is then a chain of four barrier-separated phases, and its length is what counts)
template <int FT>
__global__ void __launch_bounds__(256)
k_field_t(const __grid_constant__ KParams P, const KArrays A)
{
    __shared__ double sA[FT + 2 * FH][FT + 2 * FH + 1];   // raw
    __shared__ double sB[FT + 2][FT + 2 * FH + 1];        // after the axis-0 (y) pass
    __shared__ double sC[FT + 2][FT + 2 + 1];             // screen * v_env
    const int nx = P.nx, ny = P.ny;
    const int x0 = blockIdx.x * FT, y0 = P.yf0 + blockIdx.y * FT;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int gny = P.ny_global;
    // decomposed tissue: tiles that reach into the halo rows read what the neighbours pushed (exchange point X2)
    if (P.xwait && (y0 - FH < P.y_own0 || y0 + FT + FH > P.y_own1)) xchg_wait_cta(P, A, 1);
    for (int t = tid; t < (FT + 2 * FH) * (FT + 2 * FH); t += 256) {
        const int ly = t / (FT + 2 * FH), lx = t % (FT + 2 * FH);
        const int y = y0 + ly - FH, x = x0 + lx - FH;
        const int gy = y + P.y0;
        double v = 0.0;
        if (x >= 0 && x < nx && gy >= 0 && gy < gny && y >= 0 && y < ny) v = A.v_raw[y * nx + x];
        sA[ly][lx] = v;
    }
    __syncthreads();
    const double w0 = P.gw[0], w1 = P.gw[1], w2 = P.gw[2], w3 = P.gw[3], w4 = P.gw[4];
    // axis 0 (rows): scipy correlate1d symmetric form  c*w0 + sum_k (a[-k]+a[+k])*w_k, k = 4..1
    for (int t = tid; t < (FT + 2) * (FT + 2 * FH); t += 256) {
        const int ly = t / (FT + 2 * FH), lx = t % (FT + 2 * FH);   // ly: output row y0-1+ly
        const int r = ly + FH - 1;
        double s = sA[r][lx] * w0;
        s += (sA[r - 4][lx] + sA[r + 4][lx]) * w4;
        s += (sA[r - 3][lx] + sA[r + 3][lx]) * w3;
        s += (sA[r - 2][lx] + sA[r + 2][lx]) * w2;
        s += (sA[r - 1][lx] + sA[r + 1][lx]) * w1;
        // rows outside the world do not exist: their intermediate is the zero padding
        const int gy = y0 - 1 + ly + P.y0;
        sB[ly][lx] = (gy >= 0 && gy < gny) ? s : 0.0;
    }
    __syncthreads();
    for (int t = tid; t < (FT + 2) * (FT + 2); t += 256) {
        const int ly = t / (FT + 2), lx = t % (FT + 2);             // output col x0-1+lx
        const int cidx = lx + FH - 1;
        double s = sB[ly][cidx] * w0;
        s += (sB[ly][cidx - 4] + sB[ly][cidx + 4]) * w4;
        s += (sB[ly][cidx - 3] + sB[ly][cidx + 3]) * w3;
        s += (sB[ly][cidx - 2] + sB[ly][cidx + 2]) * w2;
        s += (sB[ly][cidx - 1] + sB[ly][cidx + 1]) * w1;
        const int y = y0 - 1 + ly, x = x0 - 1 + lx;
        double v = 0.0;
        if (x >= 0 && x < nx && y >= 0 && y < ny) {
            v = s;
            if (P.has_phi) v += A.phi_b[y * nx + x];
            if (ly >= 1 && ly <= FT && lx >= 1 && lx <= FT && y < P.yf1) A.v_env[y * nx + x] = v;
        }
        sC[ly][lx] = P.screen * v;
    }
    __syncthreads();
    const double d = P.delta;
    for (int t = tid; t < FT * FT; t += 256) {
        const int ly = t / FT + 1, lx = t % FT + 1;
        const int y = y0 + ly - 1, x = x0 + lx - 1;
        if (x >= nx || y >= P.yf1) continue;
        const int gy = y + P.y0;
        double gx, gyv;
        if (x == 0) gx = (sC[ly][lx + 1] - sC[ly][lx]) / d;
        else if (x == nx - 1) gx = (sC[ly][lx] - sC[ly][lx - 1]) / d;
        else gx = -(sC[ly][lx - 1] - sC[ly][lx + 1]) / (2.0 * d);
        if (gy == 0) gyv = (sC[ly + 1][lx] - sC[ly][lx]) / d;
        else if (gy == gny - 1) gyv = (sC[ly][lx] - sC[ly - 1][lx]) / d;
        else gyv = -(sC[ly - 1][lx] - sC[ly + 1][lx]) / (2.0 * d);
        A.E_x[y * nx + x] = -gx;
        A.E_y[y * nx + x] = -gyv;
    }
}

// ---------------------------------------------------------------------------- diagnostics
// Sampled-step quantities of get_current / update_V that the loop itself never reads
// (cell_polarizability == 0): Jmem, Jgj, smoothed Jn, I_mem, J_cell, Jc, sigma_cell, E_cell, Emc,
// per-membrane vm, vm_ave, dvm.  Same CTA packing as k_mem.
template <int NI>
__global__ void __launch_bounds__(BT_TPB)
k_diag(const __grid_constant__ KParams P, const KArrays A, const int newb)
{
    __shared__ double s_a[BT_TPB], s_b[BT_TPB];
    __shared__ double s_c[BT_MAX_CTA_CELLS], s_d[BT_MAX_CTA_CELLS];
    const bool polar = P.polar != 0;
    const int tid = threadIdx.x;
    const int c0 = ldgi(A.cta_cell_start + blockIdx.x), c1 = ldgi(A.cta_cell_start + blockIdx.x + 1);
    const int m0 = ldgi(A.cell_mem_ptr + c0), m1 = ldgi(A.cell_mem_ptr + c1);
    const int nm = m1 - m0, nc = c1 - c0;
    const int C = P.n_cells, Mo = P.n_mems_owned;
    const int m = m0 + tid;
    const bool act = tid < nm;
    int c = 0; double sa = 0, nxv = 0, nyv = 0, Jn0 = 0, vm = 0, vmo = 0;
    if (act) {
        c = ldgi(A.mem_to_cells + m);
        sa = ldg(A.mem_sa + m); nxv = ldg(A.mem_nx + m); nyv = ldg(A.mem_ny + m);
        double dm = 0.0, dg = 0.0;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            dm = fma(P.zF[i], A.fl_mem[(size_t)i * Mo + m], dm);
            dg = fma(P.zF[i], A.fl_gj[(size_t)i * Mo + m], dg);
        }
        // sim.extra_J_mem: uploaded, or (channels + substances_affect_charge) this step's channel currents
        const double Jmem = -dm + ((P.chan_charge && A.chanJ) ? A.chanJ[m] : (A.extra_J_mem ? ldg(A.extra_J_mem + m) : 0.0));
        A.Jmem[m] = Jmem; A.Jgj[m] = dg;
        Jn0 = Jmem + dg;
        double phi = 0.0, phi_o = 0.0;
        if (P.has_phi) { phi = ldg(A.phi_b + ldgi(A.map_mem2ecm + m)); phi_o = ldg(A.phi_b_old + ldgi(A.map_mem2ecm + m)); }
        if (polar) {
            vmo = A.vm_pol[newb ^ 1][m];               // Vmem the step started with; the new one needs Jn (below)
        } else {
            vm = A.vm_cell[newb][c] - phi;
            vmo = A.vm_cell[newb ^ 1][c] - phi_o;
            A.vm_mem[m] = vm;
            A.dvm[m] = (vm - vmo) / P.dt;
        }
        // gap-junction field of this step (update_gj, sim.py:2166-2172): from the Vmem the step started with
        {
            const int nnp = ldgi(A.nn_cell_flag + m);
            double vnb = A.vm_cell[newb ^ 1][nnp & 0x7fffffff];
            if (polar) vnb = A.vm_pol[newb ^ 1][ldgi(A.nn_i + m)];
            else if (P.has_phi) vnb -= ldg(A.phi_b_old + ldgi(A.map_mem2ecm + ldgi(A.nn_i + m)));
            const double Egj = -(vnb - vmo) / P.gj_len;
            A.E_gj_x[m] = Egj * nxv; A.E_gj_y[m] = Egj * nyv;
        }
    }
    s_a[tid] = act ? Jn0 * sa : 0.0;
    s_b[tid] = act ? vm : 0.0;
    __syncthreads();
    if (tid < nc) {
        const int cc = c0 + tid;
        const int jb = ldgi(A.cell_mem_ptr + cc) - m0, je = ldgi(A.cell_mem_ptr + cc + 1) - m0;
        double s = 0.0, sv = 0.0;
        for (int j = jb; j < je; ++j) { s += s_a[j]; sv += s_b[j]; }
        s_c[tid] = s / ldg(A.cell_sa + cc);
        if (!polar) A.vm_ave[cc] = sv / ldg(A.num_mems + cc);
        else {                                         // sigma_cell (sim.py:2018-2019) is needed by the Vmem update
            double sg = 0.0;
#pragma unroll
            for (int i = 0; i < NI; ++i) sg += (((P.sig_k[i] * A.cc_cells[i * C + cc]) * P.D_free[i]) * 0.1) / P.R_T_p;
            s_d[tid] = sg / (double)NI;
        }
    }
    __syncthreads();
    double Jn = 0.0, Rr = 1.0;
    if (act) {
        const double nmem = ldg(A.num_mems + c);
        const double wm = (P.smooth_cells * nmem - 1.0) / (P.smooth_cells * nmem);   // sim.py:782-786
        const double wo = 1.0 / (P.smooth_cells * nmem);
        Jn = wm * Jn0 + s_c[c - c0] * wo;
        A.Jn[m] = Jn;
        A.I_mem[m] = -Jn * sa;
        if (polar) {
            // implicit-Euler Vmem towards vm_o = rho_surf/cm through the cytosol conductance (sim.py:2057-2061)
            Rr = ldg(A.R_rads + m);
            const double sg = s_d[c - c0];
            const double vm_o = A.vm_cell[newb][c];
            const double dtcm = P.dt / P.cm;
            vm = ((vmo - dtcm * Jn) + ((P.dt * sg) * vm_o) / (P.cm * Rr)) / (1.0 + (P.dt * sg) / (P.cm * Rr));
            if (vm != vm) atomicOr(A.status, (unsigned)ST_NAN_VM);
            A.vm_pol[newb][m] = vm;
            A.vm_mem[m] = vm;
            A.dvm[m] = (vm - vmo) / P.dt;
        }
    }
    __syncthreads();
    if (polar) {
        // vm_ave, then the cell field from the Vmem spread over each cell (sim.py:2064-2076)
        s_b[tid] = act ? vm : 0.0;
        __syncthreads();
        if (tid < nc) {
            const int cc = c0 + tid;
            const int jb = ldgi(A.cell_mem_ptr + cc) - m0, je = ldgi(A.cell_mem_ptr + cc + 1) - m0;
            double sv = 0.0;
            for (int j = jb; j < je; ++j) sv += s_b[j];
            const double va = sv / ldg(A.num_mems + cc);
            A.vm_ave[cc] = va;
            s_c[tid] = va;
        }
        __syncthreads();
        double gE = 0.0;
        if (act) gE = (vm - s_c[c - c0]) / (Rr * (P.true_cell_size / P.cell_radius));
        __syncthreads();
        s_a[tid] = act ? ((-gE) * nxv) * sa : 0.0;
        s_b[tid] = act ? ((-gE) * nyv) * sa : 0.0;
        __syncthreads();
        if (tid < nc) {
            const int cc = c0 + tid;
            const int jb = ldgi(A.cell_mem_ptr + cc) - m0, je = ldgi(A.cell_mem_ptr + cc + 1) - m0;
            double sx = 0.0, sy = 0.0;
            for (int j = jb; j < je; ++j) { sx += s_a[j]; sy += s_b[j]; }
            const double csa = ldg(A.cell_sa + cc);
            A.E_cell_x[cc] = sx / csa; A.E_cell_y[cc] = sy / csa;
        }
        __syncthreads();
    }
    s_a[tid] = act ? (Jn * nxv) * sa : 0.0;
    s_b[tid] = act ? (Jn * nyv) * sa : 0.0;
    __syncthreads();
    if (tid < nc) {
        const int cc = c0 + tid;
        const int jb = ldgi(A.cell_mem_ptr + cc) - m0, je = ldgi(A.cell_mem_ptr + cc + 1) - m0;
        double sx = 0.0, sy = 0.0;
        for (int j = jb; j < je; ++j) { sx += s_a[j]; sy += s_b[j]; }
        const double csa = ldg(A.cell_sa + cc);
        const double Jx = sx / csa, Jy = sy / csa;
        double sg = 0.0;
#pragma unroll
        for (int i = 0; i < NI; ++i) sg += (((P.sig_k[i] * A.cc_cells[i * C + cc]) * P.D_free[i]) * 0.1) / P.R_T_p;
        sg = sg / (double)NI;
        A.J_cell_x[cc] = Jx; A.J_cell_y[cc] = Jy; A.sigma_cell[cc] = sg;
        if (!polar) { A.E_cell_x[cc] = Jx / sg; A.E_cell_y[cc] = Jy / sg; }
        s_c[tid] = Jx; s_d[tid] = Jy;
    }
    __syncthreads();
    if (act) {
        const double Jx = s_c[c - c0], Jy = s_d[c - c0];
        A.Jc[m] = Jx * nxv + Jy * nyv;
        const double sg = A.sigma_cell[c];
        if (polar) A.Emc[m] = A.E_cell_x[c] * nxv + A.E_cell_y[c] * nyv;      // sim.py:2079-2080
        else A.Emc[m] = (Jx / sg) * nxv + (Jy / sg) * nyv;
    }
}

// per-membrane Vmem for download: vm = vm_cell[cell] - Phi_b[map_mem2ecm]
__global__ void k_expand_vm(const __grid_constant__ KParams P, const KArrays A, const int cur)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P.n_mems_owned || P.polar) return;      // polar: vm_mem already holds the membrane state
    double phi = 0.0;
    if (P.has_phi) phi = A.phi_b[A.map_mem2ecm[m]];
    A.vm_mem[m] = A.vm_cell[cur][A.mem_to_cells[m]] - phi;
}

// vm_ave = np.dot(M_sum_mems, vm)/num_mems (sim.py:2038) for download
__global__ void k_vm_ave(const __grid_constant__ KParams P, const KArrays A)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_cells_owned) return;
    double s = 0.0;
    for (int m = A.cell_mem_ptr[c]; m < A.cell_mem_ptr[c + 1]; ++m) s += A.vm_mem[m];
    A.vm_ave[c] = s / A.num_mems[c];
}

// ---------------------------------------------------------------------------- launchers
// tuning/testing switches: BETSE_KMEM_GENERIC=1 forces the run-time-configured build
static bool kmem_generic()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("BETSE_KMEM_GENERIC"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

// register budget of k_mem: measured best at 2 CTAs/SM (128 registers, no spills: 0.46 ms at 1 M cells)
// against 3 (80 registers, 0.52 ms) and 4 (64 registers, 0.77 ms); BETSE_KMEM_MINB=3|4 selects the others
static int kmem_minb()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("BETSE_KMEM_MINB"); v = (e && e[0] == '3') ? 3 : (e && e[0] == '4') ? 4 : 2; }
    return v;
}

// kmem_pipe.cu: the persistent, software-pipelined build of the specialised kernel
bool kmem_pipe_enabled();
cudaError_t prepare_mem_pipe(int ni);
void launch_mem_pipe(int ni, const KParams& P, const KArrays& A, int n_sms, int cur, cudaStream_t st);

// kcell.cu: the lane-per-cell build of the specialised kernel (fluxes in flux_ell)
bool kcell_enabled();
void kcell_set_sms(int n);
void launch_cell(int ni, const KParams& P, const KArrays& A, int cur, cudaStream_t st);

// the specialised builds apply to the shipped ion profiles with the default feature switches
template <int NI>
static bool std_profile(const KParams& P, const KArrays& A, int diag, bool allow_defer = false)
{
    bool ok = NI <= 7 && P.is_ecm && P.v_sensitive_gj && P.cluster_open && !P.fast_update_ecm && !diag &&
              !A.gj_block && !A.NaK_block && P.iNa == StdProf<NI>::iNa && P.iK == StdProf<NI>::iK &&
              P.iCa == StdProf<NI>::iCa && !kmem_generic() && (allow_defer || !P.defer) && !P.has_phi && !P.polar;
    for (int i = 0; i < NI && ok; ++i) ok = (P.zi[i] == StdProf<NI>::z(i)) && P.zi[i] != 0;
    return ok;
}

// 0: k_mem (run-time configured), 1: k_mem_pipe, 2: k_cell
int mem_kernel_kind(int ni, const KParams& P, const KArrays& A, int diag)
{
    // k_cell also runs the deferred-update mode of channels / networks (it leaves the membranes -> cell sums in dsum_m / dsum_g
    // for k_cell_update); k_mem_pipe does not
    const bool kc = kcell_enabled() && A.cpack && !(P.defer && (!A.dsum_m || P.n_patches > 0 || P.defer_slots));
    bool sp;
    switch (ni) {
        case 4: sp = std_profile<4>(P, A, diag, kc); break;
        case 5: sp = std_profile<5>(P, A, diag, kc); break;
        case 6: sp = std_profile<6>(P, A, diag, kc); break;
        case 7: sp = std_profile<7>(P, A, diag, kc); break;
        default: sp = false;
    }
    if (!sp) return 0;
    if (kc) return 2;
    if (P.defer) return 0;
    return (kmem_pipe_enabled() && A.tile_pack) ? 1 : 0;
}

// returns 1 when the membrane -> env fluxes went to flux_ell (k_cell)
template <int NI>
static int launch_mem_t(const KParams& P, const KArrays& A, int n_ctas, int cur, int diag, cudaStream_t st)
{
    const size_t smem = (size_t)KM_SMEM_DOUBLES(NI) * sizeof(double);
    const int grid = (P.n_tiles + (BT_TPB / 32) - 1) / (BT_TPB / 32);
    const bool minb2 = kmem_minb() == 2;
    const int kind = mem_kernel_kind(NI, P, A, diag);
    const bool std_prof = NI <= 7 && std_profile<(NI <= 7 ? NI : 7)>(P, A, diag);
    if (P.has_phi || P.polar) k_mem<NI, true, 2, 0><<<grid, BT_TPB, smem, st>>>(P, A, cur, diag);
    else if (kind == 2) { launch_cell(NI, P, A, cur, st); return 1; }
    else if (kind == 1) launch_mem_pipe(NI, P, A, g_n_sms, cur, st);
    else if (std_prof && minb2) k_mem<NI, false, 2, 1><<<grid, BT_TPB, smem, st>>>(P, A, cur, diag);
    else if (std_prof && kmem_minb() == 4) k_mem<NI, false, 4, 1><<<grid, BT_TPB, smem, st>>>(P, A, cur, diag);
    else if (std_prof) k_mem<NI, false, 3, 1><<<grid, BT_TPB, smem, st>>>(P, A, cur, diag);
    else k_mem<NI, false, 2, 0><<<grid, BT_TPB, smem, st>>>(P, A, cur, diag);
    return 0;
}

template <typename K>
static cudaError_t prep_one(K kern, int smem)
{
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e) return e;
    return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}

template <int NI>
static cudaError_t prepare_mem_t()
{
    const int smem = (int)(KM_SMEM_DOUBLES(NI) * sizeof(double));
    cudaError_t e;
    if ((e = prep_one(k_mem<NI, true, 2, 0>, smem))) return e;
    if ((e = prep_one(k_mem<NI, false, 2, 0>, smem))) return e;
    if ((e = prep_one(k_mem<NI, false, 2, 1>, smem))) return e;
    if ((e = prep_one(k_mem<NI, false, 4, 1>, smem))) return e;
    return prep_one(k_mem<NI, false, 3, 1>, smem);
}

// not capturable: called once per context before the first launch
cudaError_t prepare_kernels(int ni)
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) { g_n_sms = sms; kcell_set_sms(sms); }
    cudaError_t ep = prepare_mem_pipe(ni);
    if (ep) return ep;
    switch (ni) {
        case 4: return prepare_mem_t<4>();
        case 5: return prepare_mem_t<5>();
        case 6: return prepare_mem_t<6>();
        case 7: return prepare_mem_t<7>();
        default: return prepare_mem_t<8>();
    }
}

int launch_mem(int ni, const KParams& P, const KArrays& A, int n_ctas, int cur, int diag, cudaStream_t st)
{
    switch (ni) {
        case 4: return launch_mem_t<4>(P, A, n_ctas, cur, diag, st);
        case 5: return launch_mem_t<5>(P, A, n_ctas, cur, diag, st);
        case 6: return launch_mem_t<6>(P, A, n_ctas, cur, diag, st);
        case 7: return launch_mem_t<7>(P, A, n_ctas, cur, diag, st);
        default: return launch_mem_t<8>(P, A, n_ctas, cur, diag, st);
    }
}

// ions [ion0, ion0 + n)
void launch_ion(const KParams& P, const KArrays& A, int nx, int cur, int diag, int ion0, int n, cudaStream_t st)
{
    const int rows = P.yi1 - P.yi0;
    if (rows <= 0 || n <= 0) return;
    dim3 b(256), g((nx + IT_X - 1) / IT_X, (rows + IT_Y - 1) / IT_Y, n);
    k_ion<<<g, b, 0, st>>>(P, A, cur, diag, ion0);
}

void launch_ion_smooth(int ni, const KParams& P, const KArrays& A, int ny, int nx, int nxt, cudaStream_t st)
{
    dim3 b(32, 8), g((nx + 31) / 32, (ny + 7) / 8);
    k_ion_smooth<<<g, b, 0, st>>>(P, A, nxt, ni);
}

void launch_envacc(int ni, const KParams& P, const KArrays& A, int E, int nxt, int apply, cudaStream_t st)
{
    const int n = (P.ya1 - P.ya0) * P.nx;
    if (n <= 0) return;
    const int g = (n + 255) / 256;
    switch (ni) {
        case 4: k_envacc<4><<<g, 256, 0, st>>>(P, A, nxt, apply); break;
        case 5: k_envacc<5><<<g, 256, 0, st>>>(P, A, nxt, apply); break;
        case 6: k_envacc<6><<<g, 256, 0, st>>>(P, A, nxt, apply); break;
        case 7: k_envacc<7><<<g, 256, 0, st>>>(P, A, nxt, apply); break;
        default: k_envacc<8><<<g, 256, 0, st>>>(P, A, nxt, apply); break;
    }
}

// rho_cells and Vmem from the current concentrations (ion_current.py:19; sim.py:2027-2029)
__global__ void k_cell_charge(const __grid_constant__ KParams P, const KArrays A, const int cur)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_cells_owned) return;
    double rho = 0.0;
    for (int i = 0; i < P.n_ions; ++i) rho = fma(P.zF[i], A.cc_cells[(size_t)i * P.n_cells + c], rho);
    if (A.extra_rho_cells) rho += A.extra_rho_cells[c];
    A.rho_cells[c] = rho;
    A.vm_cell[cur][c] = P.inv_cm * (rho * A.diviterm[c]);
}

void launch_cell_charge(const KParams& P, const KArrays& A, int C, int cur, cudaStream_t st)
{
    k_cell_charge<<<(C + 255) / 256, 256, 0, st>>>(P, A, cur);
}

void launch_field(const KParams& P, const KArrays& A, int ny, int nx, cudaStream_t st)
{
    const int rows = P.yf1 - P.yf0;
    if (rows <= 0) return;
    static const int small_ok = [] { const char* e = getenv("BETSE_FIELD16"); return (e && e[0] == '0') ? 0 : 1; }();
    const int n32 = ((nx + 31) / 32) * ((rows + 31) / 32);
    dim3 b(32, 8);
    if (small_ok && n32 < 2 * 148) k_field_t<16><<<dim3((nx + 15) / 16, (rows + 15) / 16), b, 0, st>>>(P, A);
    else k_field_t<32><<<dim3((nx + 31) / 32, (rows + 31) / 32), b, 0, st>>>(P, A);
}

void launch_envmix(int ni, const KParams& P, const KArrays& A, int cur, cudaStream_t st)
{
    k_envmix<<<ni, 256, 0, st>>>(P, A, cur);
}

void launch_diag(int ni, const KParams& P, const KArrays& A, int n_ctas, int newb, cudaStream_t st)
{
    switch (ni) {
        case 4: k_diag<4><<<n_ctas, BT_TPB, 0, st>>>(P, A, newb); break;
        case 5: k_diag<5><<<n_ctas, BT_TPB, 0, st>>>(P, A, newb); break;
        case 6: k_diag<6><<<n_ctas, BT_TPB, 0, st>>>(P, A, newb); break;
        case 7: k_diag<7><<<n_ctas, BT_TPB, 0, st>>>(P, A, newb); break;
        default: k_diag<8><<<n_ctas, BT_TPB, 0, st>>>(P, A, newb); break;
    }
}

void launch_expand_vm(const KParams& P, const KArrays& A, int M, int C, int cur, cudaStream_t st)
{
    k_expand_vm<<<(M + 255) / 256, 256, 0, st>>>(P, A, cur);
    k_vm_ave<<<(C + 255) / 256, 256, 0, st>>>(P, A);
}
