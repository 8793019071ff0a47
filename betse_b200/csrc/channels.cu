// Voltage-gated channels of the general network:
//   MasterOfNetworks.run_loop_channels        betse/science/chemistry/networks.py:3115-3213
//   VgNaABC/VgKABC/VgCaABC/VgClABC.run        betse/science/channels/vg_na.py:75-112 (and siblings)
//   ChannelsABC.update_mh                     betse/science/channels/channelsabc.py:40-60
// The ~35 Hodgkin-Huxley models are DATA (betse_b200/channels.py: a table of gating terms held to
// the reference classes by tests/test_channels_table.py); k_chan evaluates any of them.
//
// Semantics that matter: a channel's GHK flux is applied IMMEDIATELY (update_Co inside the loop
// over channels), so the next channel sees the updated cell and env concentrations, and all of it
// happens between the flux computation of the ion loop and update_all_concs (sim.py:1290-1357).
// Hence, with channels, k_mem only DEFERS its membrane->cell sums (KParams.defer), the channels
// run one after the other (k_chan: gates, flux, cell update; k_chan_env: env update), and
// k_cell_update then applies the deferred sums exactly as k_mem's tail would have.
#include "kparams.cuh"
#include "channels.cuh"
#include "network.cuh"

#define FLOAT_NONCE 1.0e-25
#define ST_NAN_VM 1u
#define ST_NAN_CONC 2u
#define ST_NEG 4u

// same reciprocal as k_mem's tail (kernels.cu: fast_rcp) so that the deferred update is bit-identical to the fused one
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// One copy of the term evaluator in the kernel (not inlined): eight inlined copies of ten exp / division variants made
// k_chan_cell 9 k instructions long and its warps stalled on instruction fetch (profiles/r02s_ncu_chan_summary.csv).
__device__ __noinline__ double gate_term_eval(const int type, const double p0, const double p1, const double p2, const double p3, const double U)
{
    switch (type) {
        case 0: return p0;
        case 1: return p0 + p1 / (1.0 + exp((U - p2) / p3));
        case 2: return p0 * U + p1;
        case 3: return p0 + p1 * exp((p2 - U) / p3);
        case 4: { const double x = (U + p2) / p3; return p0 + p1 * exp(-(x * x)); }
        case 5: { const double x = U - p1; return p0 * x / (1.0 - exp(-x / p2)); }
        case 7: return (U >= p2) ? 1.0 : p0 + p1 * exp(-U / p3);   // vg_ca.py Cav3p1: tau overwritten with 1 above the cut
        case 8: return p0 / cosh((U - p1) / p2);                      // vg_morrislecar.py: 1/cosh time constants, Kir_ML
        case 9: return p0 + p1 * tanh((U - p2) / p3);             // vg_morrislecar.py: 0.5*(1 + tanh(...))
        default: { const double x = -U - p1; return p0 * x / (1.0 - exp(-x / p2)); }
    }
}

__device__ __forceinline__ double gate_term(const KTerm& t, double U)
{
    return gate_term_eval(t.type, t.p[0], t.p[1], t.p[2], t.p[3], U);
}

__device__ __forceinline__ double gate_quantity(const KChan& ch, int q, double U)
{
    const double a = gate_term(ch.a[q], U);
    if (ch.kind[q] == 0) return a;
    const double b = gate_term(ch.b[q], U);
    return ch.kind[q] == 1 ? a / (a + b) : 1.0 / (a + b);
}

__device__ __forceinline__ double ipow(double x, int n)
{
    double r = 1.0;
    for (int k = 0; k < n; ++k) r *= x;
    return r;
}

// ChannelsABC.update_mh (channelsabc.py:40-60): implicit step of both gates at the membrane voltage vm.  The roundings are
// spelled out (fma / single operations) because k_chan and k_chan_cell must agree bit for bit.
__device__ __forceinline__ void gate_advance(const KChan& ch, const double vm, double& m, double& h)
{
    const double U = fma(vm, 1000.0, ch.shift);                    // V = vm[targets]*1000 + v_corr (vg_na.py:91)
    const double mInf = gate_quantity(ch, 0, U), mTau = gate_quantity(ch, 1, U);
    const double hInf = gate_quantity(ch, 2, U), hTau = gate_quantity(ch, 3, U);
    const double dt = ch.dt_tu;                                    // p.dt*time_unit (channelsabc.py:56)
    m = fma(mTau, m, __dmul_rn(dt, mInf)) / __dadd_rn(mTau, dt);
    h = fma(hTau, h, __dmul_rn(dt, hInf)) / __dadd_rn(hTau, dt);
}

// stb.electroflux with the cell-side factors formed by the caller (sim_toolbox.py:18-69): coef = -(DChan*alpha/tm)
__device__ __forceinline__ double chan_flux(const double coef, const double cB, const double cA, const double ex, const double deno, const double rho)
{
    return __dmul_rn(__dmul_rn(coef, fma(-cA, ex, cB) / deno), rho);
}

// One warp per tile of whole cells (the packing of k_mem): lanes = membranes for gates and flux,
// then lanes = cells for the immediate concentration update of the conducted ion.
// `slots`: where the channel leaves f*sa of every membrane for the env side of its update_Co.
__device__ __forceinline__ void chan_apply(const KParams& P, const KArrays& A, const KChan& ch, const KNet& N, const int cur,
                                           double* __restrict__ slots, double* s_f, const int lane, const int tile, const int4 td)
{
    const int c0 = td.x, nc = td.y, m0 = td.z, nm = td.w;
    const int C = P.n_cells, E = P.ny * P.nx;
    const int ion = ch.ion;
    double* __restrict__ ccell = A.cc_cells + ion * C;
    const double* __restrict__ cenv = A.cc_env[cur ^ 1] + ion * E;   // after this step's transport (+ earlier channels)

    double fsa = 0.0;
    if (lane < nm) {
        const int m = m0 + lane;
        const int c = __ldg(A.mem_to_cells + m);
        const int e = __ldg(A.map_mem2ecm + m);
        double vm = A.vm_cell[cur][c];
        if (P.polar) vm = A.vm_pol[cur][m];
        else if (P.has_phi) vm -= __ldg(A.phi_b_old + e);
        double Pm = 0.0;
        if (ch.frozen) Pm = ch.P[m];
        else if (!ch.mask || ch.mask[m]) {
            double mm = ch.m[m], hh = ch.h[m];
            gate_advance(ch, vm, mm, hh);
            ch.m[m] = mm; ch.h[m] = hh;
            Pm = ipow(mm, ch.mpow) * ipow(hh, ch.hpow);            // vg_na.py:104
        }
        if (!ch.frozen) ch.P[m] = Pm;
        // moddy = eval(chan.alpha_eval_string) in the membrane zone (networks.py:3147; compiled by ratelaw.py)
        const double moddy = (ch.mod_prog >= 0) ? rl_eval(N, ch.mod_prog, c, m, A, C, P.n_mems_owned, cur, vm) : 1.0;
        const double DChan = ((Pm * ch.rel_perm) * ch.maxDm) * moddy;   // networks.py:3164
        if (ch.D) ch.D[m] = DChan;
        // stb.electroflux(cenv[map_mem2ecm], cmem, DChan, tm, z, vm, sim.T, rho=rho_channel), sim_toolbox.py:18-69
        const double alpha = ((P.z[ion] + FLOAT_NONCE) * (vm + FLOAT_NONCE) * P.F) / P.RT_sim;
        const double ex = exp(-alpha), deno = -expm1(-alpha);
        const double cB = ccell[c], cA = P.is_ecm ? cenv[e] : A.cenv_u[cur * 8 + ion];
        const double f = chan_flux(-((DChan * alpha) / P.tm), cB, cA, ex, deno, P.rho_channel);
        fsa = __dmul_rn(f, __ldg(A.mem_sa + m));
        if (P.is_ecm) slots[m] = fsa;
        if (ch.flux) ch.flux[m] = f;
        if (A.chanJ) A.chanJ[m] += (-f * P.F) * P.z[ion];         // extra_J_mem += -f_ED*p.F*zzz, networks.py:3199
    }
    s_f[lane] = fsa;
    __syncwarp();
    if (!P.is_ecm && lane == 0) {                                 // no-ECM: the tile's share of sum_m f*sa (k_chan_mix)
        double S = 0.0;
        for (int j = 0; j < nm; ++j) S += s_f[j];
        A.chan_part[tile] = S;
    }
    if (lane < nc) {                                              // update_Co cell branch, sim_toolbox.py:1177-1181
        const int c = c0 + lane;
        const int jb = __ldg(A.cell_mem_ptr + c) - m0, je = __ldg(A.cell_mem_ptr + c + 1) - m0;
        double S = 0.0;
        for (int j = jb; j < je; ++j) S += s_f[j];
        ccell[c] = fma(S / __ldg(A.cell_vol + c), P.dt, ccell[c]);
    }
}

__global__ void __launch_bounds__(BT_TPB)
k_chan(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KChan ch, const __grid_constant__ KNet N,
       const int cur)
{
    __shared__ double s_all[(BT_TPB / 32) * 32];
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (BT_TPB / 32) + (threadIdx.x >> 5);
    if (tile >= P.n_tiles) return;
    const int4 td = __ldg(reinterpret_cast<const int4*>(A.tile_desc) + tile);
    chan_apply(P, A, ch, N, cur, A.chan_slots, s_all + (threadIdx.x >> 5) * 32, lane, tile, td);
}

// The deferred tail of k_mem (update_all_concs, sim.py:2086-2111; charge and Vmem, ion_current.py:19,
// sim.py:2027-2029), applied to the concentrations the channels left behind: one cell.  Shared by k_cell_update and by the
// last channel pass of a step (k_chan_cell with `fuse_update`: nothing else runs between the two, and a cell's update
// reads and writes that cell only).
__device__ __forceinline__ void cell_update_one(const KParams& P, const KArrays& A, const int c, const int cur, unsigned int& flags)
{
    const int C = P.n_cells, nxt = cur ^ 1;
    const double rvol = fast_rcp(__ldg(A.cell_vol + c));
    double rho = 0.0;
    // every input first (the stores below would otherwise fence the next ion's loads: one memory round trip per ion)
    double cc[BT_MAX_IONS], Sm[BT_MAX_IONS], Sg[BT_MAX_IONS];
#pragma unroll
    for (int i = 0; i < BT_MAX_IONS; ++i) {
        cc[i] = Sm[i] = Sg[i] = 0.0;
        if (i < P.n_ions) { cc[i] = A.cc_cells[i * C + c]; Sm[i] = A.dsum_m[i * C + c]; Sg[i] = A.dsum_g[i * C + c]; }
    }
    const double xr = A.extra_rho_cells ? __ldg(A.extra_rho_cells + c) : 0.0;
    const double dvt = __ldg(A.diviterm + c);
#pragma unroll
    for (int i = 0; i < BT_MAX_IONS; ++i) if (i < P.n_ions) {
        const double cm_new = cc[i] + (Sm[i] * rvol) * P.dt;
        double cn_new = cm_new + P.dt * ((-Sg[i]) * rvol);
        if (cn_new != cn_new) flags |= ST_NAN_CONC;
        if (cn_new < 0.0) { cn_new = 0.0; flags |= ST_NEG; }
        A.cc_cells[i * C + c] = cn_new;
        A.cc_mid[nxt][i * C + c] = cm_new;
        rho = fma(P.zF[i], cn_new, rho);
    }
    if (A.extra_rho_cells) rho += xr;
    A.rho_cells[c] = rho;
    const double vmn = P.inv_cm * (rho * dvt);
    if (vmn != vmn) flags |= ST_NAN_VM;
    A.vm_cell[nxt][c] = vmn;
}

__global__ void __launch_bounds__(256)
k_cell_update(const __grid_constant__ KParams P, const KArrays A, const int cur)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_cells_owned) return;
    unsigned int flags = 0;
    cell_update_one(P, A, c, cur, flags);
    if (flags) atomicOr(A.status, flags);
}

// ---------------------------------------------------------------------------- channels with ONE LANE PER CELL
// In the default mode Vmem is a per-cell quantity (sim.py:2029), so the gates of a channel that sits on every membrane —
// m, h, the open probability, DChan, and the exp / expm1 pair of its GHK flux — have ONE value per cell: k_chan evaluates
// them once per membrane (six times per cell, ~10 exp and ~10 divisions each; the kernel is bound by fp64 issue, not by
// memory: profiles/r02r_launches_c3.csv).  k_chan_cell walks the cell pack of k_cell (kcell.cu: SELL-32, lane = cell)
// instead: gates once per cell, then the cell's membranes in order — env concentration gathered through the row's env
// square, flux, f*sa into the pass's ELL exchange array, the membranes->cell sum as a register accumulation in membrane
// order (the order of k_chan's sum).  Same expressions as chan_apply, evaluated once instead of six times.
// A PASS = consecutive channels of the reference's application order that conduct DIFFERENT ions: none of them reads
// what another's update_Co moves (each reads its own ion and Vmem, which no channel moves), so they run back to back
// inside one kernel and one env kernel closes the update_Co of every ion of the pass.
// Gate state lives per cell (KChan.mc/hc/Pc/Dc) while this path is in use; betse_channel_state expands it (k_chan_expand).
// Eligibility (capi.cu:chan_cell_eligible): extracellular spaces, cell pack of consecutive cells, no polarizability, no
// boundary potential, no target mask, no network modulation, initial gates uniform within every cell.
#define KC_QMASK 0x0fffffffu                     // kcell.cu: env-square word of a pack row
#define KCH_TPB 64                               // threads per CTA of k_chan_cell
#define KCH_KREG 6                               // membranes of a cell whose row constants stay in registers
struct KChanPack { int n; KChan ch[KCH_PACK]; };

__global__ void __launch_bounds__(KCH_TPB, 11)              // <= 93 registers: 100 k cells are resident in one wave
k_chan_cell(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KChanPack pk, double* __restrict__ ell,
            const int cur, const int diag, const int fuse_update)
{
    const int lane = threadIdx.x & 31;
    const int task = blockIdx.x * (KCH_TPB / 32) + (threadIdx.x >> 5);
    if (task >= P.n_blocks) return;
    const int c = task * 32 + lane;
    if (c >= P.n_cells_owned) return;
    const int ni = P.n_ions;
    const size_t rowb = (size_t)(ni + 2) * 256;
    const int row0 = __ldg(A.blk_row0 + 2 * task);
    const int m_beg = __ldg(A.cell_mem_ptr + c), nm = __ldg(A.cell_mem_ptr + c + 1) - m_beg;
    const int C = P.n_cells, E = P.ny * P.nx;
    const double vm = A.vm_cell[cur][c];
    const double vol = __ldg(A.cell_vol + c);
    const char* __restrict__ rows = A.cpack + (size_t)row0 * rowb;
    // the rows' areas and env squares of the first KCH_KREG membranes stay in registers for the whole pass, and every
    // channel requests its env concentrations BEFORE its gate arithmetic: with ~20 warps per SM nothing else hides the
    // dependent gather (index -> concentration) of each membrane
    double sa[KCH_KREG];
    unsigned q[KCH_KREG];
#pragma unroll
    for (int k = 0; k < KCH_KREG; ++k) {
        sa[k] = 0.0; q[k] = 0;
        if (k < nm) {
            const char* r = rows + (size_t)k * rowb;
            sa[k] = __ldg(reinterpret_cast<const double*>(r + (size_t)ni * 256) + lane);
            q[k] = (unsigned)__ldg(reinterpret_cast<const int*>(r + (size_t)(ni + 1) * 256 + 128) + lane) & KC_QMASK;
        }
    }
    for (int j = 0; j < pk.n; ++j) {
        const KChan& ch = pk.ch[j];
        const int ion = ch.ion;
        double* ccell = A.cc_cells + (size_t)ion * C;             // (read again by the fused cell update below: not restrict)
        const double* __restrict__ cenv = A.cc_env[cur ^ 1] + (size_t)ion * E;
        double cAk[KCH_KREG];
#pragma unroll
        for (int k = 0; k < KCH_KREG; ++k) cAk[k] = (k < nm) ? cenv[q[k]] : 0.0;
        const double cB = ccell[c];
        double Pm;
        if (ch.frozen) Pm = ch.Pc[c];
        else {
            double mm = ch.mc[c], hh = ch.hc[c];
            gate_advance(ch, vm, mm, hh);
            ch.mc[c] = mm; ch.hc[c] = hh;
            Pm = ipow(mm, ch.mpow) * ipow(hh, ch.hpow);
            ch.Pc[c] = Pm;
        }
        const double DChan = ((Pm * ch.rel_perm) * ch.maxDm) * 1.0;
        ch.Dc[c] = DChan;
        const double alpha = ((P.z[ion] + FLOAT_NONCE) * (vm + FLOAT_NONCE) * P.F) / P.RT_sim;
        const double ex = exp(-alpha), deno = -expm1(-alpha);
        const double coef = -((DChan * alpha) / P.tm);
        const double zF = P.z[ion];
        double S = 0.0;
#pragma unroll
        for (int k = 0; k < KCH_KREG; ++k) if (k < nm) {
            const double f = chan_flux(coef, cB, cAk[k], ex, deno, P.rho_channel);
            const double fsa = __dmul_rn(f, sa[k]);
            ell[((size_t)(row0 + k) * ni + j) * 32 + lane] = fsa;
            ch.fell[(size_t)(row0 + k) * 32 + lane] = f;
            if (diag && A.chanJ) A.chanJ[m_beg + k] += (-f * P.F) * zF;
            S = __dadd_rn(S, fsa);
        }
        for (int k = KCH_KREG; k < nm; ++k) {
            const char* r = rows + (size_t)k * rowb;
            const double sak = __ldg(reinterpret_cast<const double*>(r + (size_t)ni * 256) + lane);
            const unsigned qk = (unsigned)__ldg(reinterpret_cast<const int*>(r + (size_t)(ni + 1) * 256 + 128) + lane) & KC_QMASK;
            const double f = chan_flux(coef, cB, cenv[qk], ex, deno, P.rho_channel);
            const double fsa = __dmul_rn(f, sak);
            ell[((size_t)(row0 + k) * ni + j) * 32 + lane] = fsa;
            ch.fell[(size_t)(row0 + k) * 32 + lane] = f;
            if (diag && A.chanJ) A.chanJ[m_beg + k] += (-f * P.F) * zF;
            S = __dadd_rn(S, fsa);
        }
        ccell[c] = fma(S / vol, P.dt, cB);
    }
    if (fuse_update) {
        // the step's last channel pass: update_all_concs, charge and Vmem of this cell right away (k_cell_update's arithmetic)
        unsigned int flags = 0;
        cell_update_one(P, A, c, cur, flags);
        if (flags) atomicOr(A.status, flags);
    }
}

// update_Co env branch of every channel of a pass (k_chan_env per conducted ion), fluxes read where k_chan_cell left them
struct KPassIons { int n; int ion[KCH_PACK]; };

__global__ void __launch_bounds__(256)
k_chan_env_cell(const __grid_constant__ KParams P, const KArrays A, const KPassIons pi, const double* __restrict__ ell, const int nxt)
{
    const int k = P.ya0 * P.nx + blockIdx.x * blockDim.x + threadIdx.x;
    const int E = P.nx * P.ny;
    if (k >= P.ya1 * P.nx) return;
    const int s0 = __ldg(A.slot_ptr + k), s1 = __ldg(A.slot_ptr + k + 1);
    if (s1 == s0) return;
    double acc[KCH_PACK], cv[KCH_PACK];
#pragma unroll
    for (int q = 0; q < KCH_PACK; ++q) { acc[q] = 0.0; cv[q] = (q < pi.n) ? A.cc_env[nxt][(size_t)pi.ion[q] * E + k] : 0.0; }
    // the square's slots eight at a time: all positions first, then all fluxes, then the sums in slot order — one memory
    // round trip per stage instead of two per slot
    for (int j0 = s0; j0 < s1; j0 += 8) {
        int off[8];
        double v[8][KCH_PACK];
#pragma unroll
        for (int u = 0; u < 8; ++u) off[u] = (j0 + u < s1) ? __ldg(A.slot_off + j0 + u) : -1;
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int q = 0; q < KCH_PACK; ++q) v[u][q] = (off[u] >= 0 && q < pi.n) ? ell[(size_t)off[u] + q * 32] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int q = 0; q < KCH_PACK; ++q) if (j0 + u < s1) acc[q] += v[u][q];
    }
#pragma unroll
    for (int q = 0; q < KCH_PACK; ++q) if (q < pi.n) A.cc_env[nxt][(size_t)pi.ion[q] * E + k] = cv[q] + ((-acc[q]) / P.env_vol_div) * P.dt;
}

// ---------------------------------------------------------------------------- channels under the FAST solver
// MasterOfNetworks.run_fast_loop_channels (networks.py:3217-3280): the gates advance at the membrane's potential (the cell
// average of the equivalent-circuit solver), the open probability becomes a conductance
//   G = stb.get_conductivity(DChan, z, cbar, tm, p) * geo_conv = DChan * (q z^2 F cbar / (tm kb T)) * geo_conv
// (sim_toolbox.py:1363-1372; the bracket per ion comes from the host as `coef`), and the current against the FIXED
// reversal potential of Simulator.fast_sim_init (sim.py:1393-1452) joins extra_J_mem: J = G (vm - E_rev).  No
// concentration moves.  One thread per membrane, the channels of a pack back to back; `first` = this launch starts the sum.
struct KFastChan { double coef[BT_MAX_IONS], rev_E[BT_MAX_IONS]; };

__global__ void __launch_bounds__(256)
k_fast_chan(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KChanPack pk, const __grid_constant__ KFastChan fc,
            const double* __restrict__ vm_ave, double* __restrict__ extra_J, const int first)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P.n_mems_owned) return;
    const double vm = vm_ave[__ldg(A.mem_to_cells + m)];
    double acc = first ? 0.0 : extra_J[m];
    for (int j = 0; j < pk.n; ++j) {
        const KChan& ch = pk.ch[j];
        double Pm = 0.0;
        if (ch.frozen) Pm = ch.P[m];
        else if (!ch.mask || ch.mask[m]) {
            double mm = ch.m[m], hh = ch.h[m];
            gate_advance(ch, vm, mm, hh);
            ch.m[m] = mm; ch.h[m] = hh;
            Pm = ipow(mm, ch.mpow) * ipow(hh, ch.hpow);
        }
        if (!ch.frozen) ch.P[m] = Pm;
        const double DChan = ((Pm * ch.rel_perm) * ch.maxDm) * 1.0;
        if (ch.D) ch.D[m] = DChan;
        const double J = (DChan * fc.coef[ch.ion]) * (vm - fc.rev_E[ch.ion]);
        acc += J;
        if (ch.flux) ch.flux[m] = J / (P.z[ch.ion] * P.F);                 // chan_flux = J_ED/(zzz*p.F), networks.py:3280
    }
    extra_J[m] = acc;
}

// all channels of the ctx, KCH_PACK per launch (coef / rev_E: [n_ions])
void launch_fast_chan(const KParams& P, const KArrays& A, const KChan* chs, int n, const double* coef, const double* rev_E,
                      const double* vm_ave, double* extra_J, cudaStream_t st)
{
    KFastChan fc;
    for (int i = 0; i < BT_MAX_IONS; ++i) { fc.coef[i] = i < P.n_ions ? coef[i] : 0.0; fc.rev_E[i] = i < P.n_ions ? rev_E[i] : 0.0; }
    for (int k0 = 0; k0 < n; k0 += KCH_PACK) {
        KChanPack pk;
        pk.n = n - k0 < KCH_PACK ? n - k0 : KCH_PACK;
        for (int j = 0; j < KCH_PACK; ++j) pk.ch[j] = chs[k0 + (j < pk.n ? j : 0)];
        k_fast_chan<<<(P.n_mems_owned + 255) / 256, 256, 0, st>>>(P, A, pk, fc, vm_ave, extra_J, k0 == 0 ? 1 : 0);
    }
}

// per-cell gate state -> the per-membrane arrays of the C ABI (betse_channel_state; leaving the per-cell path)
__global__ void __launch_bounds__(256)
k_chan_expand(const __grid_constant__ KChan ch, const int* __restrict__ mem_to_cells, const int* __restrict__ mem_ell, const int Mo, const int ni)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= Mo) return;
    const int c = __ldg(mem_to_cells + m);
    ch.m[m] = ch.mc[c]; ch.h[m] = ch.hc[c]; ch.P[m] = ch.Pc[c]; ch.D[m] = ch.Dc[c];
    const int pos = __ldg(mem_ell + m);
    ch.flux[m] = ch.fell[(size_t)(pos / (ni * 32)) * 32 + (pos & 31)];
}

// update_Co env branch for one channel: cc_env[ion] += div_env(-f)*dt (sim_toolbox.py:1189-1195, 1209-1234)
__global__ void __launch_bounds__(256)
k_chan_env(const __grid_constant__ KParams P, const KArrays A, const int ion, const int nxt)
{
    const int k = P.ya0 * P.nx + blockIdx.x * blockDim.x + threadIdx.x;
    const int E = P.nx * P.ny;
    if (k >= P.ya1 * P.nx) return;
    const int s0 = __ldg(A.slot_ptr + k), s1 = __ldg(A.slot_ptr + k + 1);
    if (s1 == s0) return;
    double acc = 0.0;
    for (int j = s0; j < s1; ++j) acc += A.chan_slots[__ldg(A.slot_idx + j)];
    double* c = A.cc_env[nxt] + ion * E + k;
    *c = *c + ((-acc) / P.env_vol_div) * P.dt;
}

// update_Co without extracellular spaces (sim_toolbox.py:1198-1205): cX_env = mean(cX_env + (-f*mem_sa/vol_env)*dt) —
// the bath is one number per ion; the channel's flux moves it before the next channel reads it.
__global__ void __launch_bounds__(256)
k_chan_mix(const __grid_constant__ KParams P, const KArrays A, const int ion, const int cur)
{
    __shared__ double red[256];
    double s = 0.0;
    for (int b = threadIdx.x; b < P.n_tiles; b += blockDim.x) s += A.chan_part[b];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double c = A.cenv_u[cur * 8 + ion];
        A.cenv_u[cur * 8 + ion] = c + ((-red[0] / P.vol_env) * P.dt) / (double)P.n_mems_owned;
    }
}

// channels [0, n) of `chs` as one pass on the per-cell path (n <= min(KCH_PACK, n_ions), distinct ions); ell: [rows][I][32]
void launch_chan_cell(const KParams& P, const KArrays& A, const KChan* chs, int n, double* ell, int cur, int diag, int fuse_update, cudaStream_t st)
{
    KChanPack pk;
    KPassIons pi;
    pk.n = pi.n = n;
    for (int j = 0; j < n; ++j) { pk.ch[j] = chs[j]; pi.ion[j] = chs[j].ion; }
    for (int j = n; j < KCH_PACK; ++j) { pk.ch[j] = chs[0]; pi.ion[j] = 0; }
    k_chan_cell<<<(P.n_blocks + KCH_TPB / 32 - 1) / (KCH_TPB / 32), KCH_TPB, 0, st>>>(P, A, pk, ell, cur, diag, fuse_update);
    const int ne = (P.ya1 - P.ya0) * P.nx;
    if (ne > 0) k_chan_env_cell<<<(ne + 255) / 256, 256, 0, st>>>(P, A, pi, ell, cur ^ 1);
}

void launch_chan_expand(const KChan& ch, const int* mem_to_cells, const int* mem_ell, int Mo, int ni, cudaStream_t st)
{
    if (Mo > 0) k_chan_expand<<<(Mo + 255) / 256, 256, 0, st>>>(ch, mem_to_cells, mem_ell, Mo, ni);
}

void launch_chan(const KParams& P, const KArrays& A, const KChan& ch, const KNet& N, int cur, cudaStream_t st)
{
    const int grid = (P.n_tiles + (BT_TPB / 32) - 1) / (BT_TPB / 32);
    k_chan<<<grid, BT_TPB, 0, st>>>(P, A, ch, N, cur);
    if (P.is_ecm) {
        const int n = (P.ya1 - P.ya0) * P.nx;
        if (n > 0) k_chan_env<<<(n + 255) / 256, 256, 0, st>>>(P, A, ch.ion, cur ^ 1);
    } else k_chan_mix<<<1, 256, 0, st>>>(P, A, ch.ion, cur);
}

// update_Co of one ion for a membrane flux handed in by the host: the dynamic-noise random walk on the protein
// concentration (sim.py:1322-1339; the draw itself is np.random.random(mdl) on the host, the reference's own stream)
__global__ void __launch_bounds__(BT_TPB)
k_flux_apply(const __grid_constant__ KParams P, const KArrays A, const int ion, const double* __restrict__ flux)
{
    __shared__ double s_all[(BT_TPB / 32) * 32];
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (BT_TPB / 32) + (threadIdx.x >> 5);
    if (tile >= P.n_tiles) return;
    double* s_f = s_all + (threadIdx.x >> 5) * 32;
    const int4 td = __ldg(reinterpret_cast<const int4*>(A.tile_desc) + tile);
    const int c0 = td.x, nc = td.y, m0 = td.z, nm = td.w;
    double* __restrict__ ccell = A.cc_cells + (size_t)ion * P.n_cells;
    double fsa = 0.0;
    if (lane < nm) {
        fsa = flux[m0 + lane] * __ldg(A.mem_sa + m0 + lane);
        if (P.is_ecm) A.chan_slots[m0 + lane] = fsa;
    }
    s_f[lane] = fsa;
    __syncwarp();
    if (!P.is_ecm && lane == 0) {
        double S = 0.0;
        for (int j = 0; j < nm; ++j) S += s_f[j];
        A.chan_part[tile] = S;
    }
    if (lane < nc) {
        const int c = c0 + lane;
        const int jb = __ldg(A.cell_mem_ptr + c) - m0, je = __ldg(A.cell_mem_ptr + c + 1) - m0;
        double S = 0.0;
        for (int j = jb; j < je; ++j) S += s_f[j];
        ccell[c] = ccell[c] + (S / __ldg(A.cell_vol + c)) * P.dt;
    }
}

void launch_flux_apply(const KParams& P, const KArrays& A, int ion, const double* flux, cudaStream_t st)
{
    const int grid = (P.n_tiles + (BT_TPB / 32) - 1) / (BT_TPB / 32);
    k_flux_apply<<<grid, BT_TPB, 0, st>>>(P, A, ion, flux);
}

// env side of an immediate update_Co for fluxes left in chan_slots / chan_part (ligand-gated channels)
void launch_chan_env(const KParams& P, const KArrays& A, int ion, int cur, cudaStream_t st)
{
    if (P.is_ecm) {
        const int n = (P.ya1 - P.ya0) * P.nx;
        if (n > 0) k_chan_env<<<(n + 255) / 256, 256, 0, st>>>(P, A, ion, cur ^ 1);
    } else k_chan_mix<<<1, 256, 0, st>>>(P, A, ion, cur);
}

void launch_cell_update(const KParams& P, const KArrays& A, int cur, cudaStream_t st)
{
    k_cell_update<<<(P.n_cells_owned + 255) / 256, 256, 0, st>>>(P, A, cur);
}
