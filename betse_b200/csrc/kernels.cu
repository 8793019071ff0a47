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
#include "kparams.cuh"

#define FLOAT_NONCE 1.0e-25   // sim_toolbox.py:52

#define ST_NAN_VM 1u
#define ST_NAN_CONC 2u
#define ST_NEG 4u

__device__ __forceinline__ double ldg(const double* p) { return __ldg(p); }
__device__ __forceinline__ int ldgi(const int* p) { return __ldg(p); }

// exp(-alpha_i) and 1/(-expm1(-alpha_i)) for the valence classes, derived from ONE exp/expm1
// pair evaluated at -alpha_1 (alpha_i = z_i*alpha_1 exactly for z = +-1, +-2).
struct GhkBase {
    double e1, inv_e1, em1, rden_p1, inv_e1p1;
    __device__ __forceinline__ void init(double a1) {
        const double x = -a1;
        e1 = exp(x);
        em1 = expm1(x);
        inv_e1 = 1.0 / e1;
        rden_p1 = 1.0 / (-em1);
        inv_e1p1 = 1.0 / (e1 + 1.0);
    }
    // ex = exp(-z a1); rden = 1/(-expm1(-z a1))
    __device__ __forceinline__ void get(int zi, double z, double a1, double& alpha, double& ex, double& rden) const {
        if (zi == 1) { alpha = a1; ex = e1; rden = rden_p1; }
        else if (zi == -1) { alpha = -a1; ex = inv_e1; rden = -e1 * rden_p1; }
        else if (zi == 2) { alpha = 2.0 * a1; ex = e1 * e1; rden = rden_p1 * inv_e1p1; }
        else if (zi == -2) { alpha = -2.0 * a1; ex = inv_e1 * inv_e1; rden = -(e1 * e1) * rden_p1 * inv_e1p1; }
        else {  // generic valence (incl. 0: the reference adds 1e-25 to z, sim_toolbox.py:56)
            alpha = (z + FLOAT_NONCE) * a1;
            ex = exp(-alpha);
            rden = 1.0 / (-expm1(-alpha));
        }
    }
};

template <int NI>
__global__ void __launch_bounds__(BT_TPB, 2)
k_mem(const KParams* __restrict__ Pp, const KArrays A, const int cur, const int diag)
{
    extern __shared__ double sm[];
    double* s_mem = sm;                       // [NI][TPB]   f_mem*sa per membrane
    double* s_gj = sm + NI * BT_TPB;          // [NI][TPB]   f_gj*sa per membrane
    double* s_cc = sm + 2 * NI * BT_TPB;      // [NI][MAX_CTA_CELLS] updated cell concentrations
    double* s_sum = s_cc + NI * BT_MAX_CTA_CELLS;  // [NI][MAX_CTA_CELLS] per-cell sum of f_mem*sa (no-ECM bath)

    const KParams& P = *Pp;
    const int tid = threadIdx.x;
    const int nxt = cur ^ 1;
    const int c0 = ldgi(A.cta_cell_start + blockIdx.x);
    const int c1 = ldgi(A.cta_cell_start + blockIdx.x + 1);
    const int m0 = ldgi(A.cell_mem_ptr + c0);
    const int m1 = ldgi(A.cell_mem_ptr + c1);
    const int nm = m1 - m0, nc = c1 - c0;
    const int C = P.n_cells;
    const int E = P.ny * P.nx;
    const int Mo = P.n_mems_owned;
    const double* __restrict__ cmid = A.cc_mid[cur];
    const double* __restrict__ vmc = A.vm_cell[cur];
    unsigned int flags = 0;

    if (tid < nm) {
        const int m = m0 + tid;
        const int c = ldgi(A.mem_to_cells + m);
        const int nnp = ldgi(A.nn_cell_flag + m);
        const int cn = nnp & 0x7fffffff;
        const bool bnd = nnp < 0;
        const int e = ldgi(A.map_mem2ecm + m);
        const double sa = ldg(A.mem_sa + m);
        double g = A.gjopen[m];
        double vm_own = vmc[c];
        double vm_nb = vmc[cn];
        if (P.has_phi) {
            vm_own -= ldg(A.phi_b + e);
            vm_nb -= ldg(A.phi_b + ldgi(A.map_mem2ecm + ldgi(A.nn_i + m)));
        }
        // ---- shared transcendental bases
        const double v = vm_own + FLOAT_NONCE;          // electroflux: vBA += 1e-25
        const double a1 = (v * P.F) / P.RT_sim;         // alpha for z = +1
        GhkBase gm; gm.init(a1);
        const double vgj0 = vm_nb - vm_own;             // sim.py:2166
        const double vg = vgj0 + FLOAT_NONCE;
        const double ag1 = (vg * P.F) / P.RT_p;         // GJ flux uses p.T (sim.py:2197)
        GhkBase gg; gg.init(ag1);
        double gnum = 0.0, rgden = 1.0;
        if (P.v_sensitive_gj) {                         // gap_junction.py:56-72
            const double V1 = 1.0e3 * fabs(vgj0);
            const double al = 0.0013 * exp(-0.077 * (V1 - P.gj_vthresh));
            double be = 0.0013 * exp(0.14 * (V1 - P.gj_vthresh));
            be = be / (1.0 + 50.0 * be);
            const double dtm = P.dt * 1.0e3;
            gnum = dtm * (al + be * P.gj_min);
            rgden = 1.0 / (1.0 + al * dtm + be * dtm);
        }
        const double gjb = A.gj_block ? ldg(A.gj_block + m) : P.gj_block;
        const double gjw = P.v_sensitive_gj ? 1.0 : ldg(A.gj_w + m);
        const double inv_tm = 1.0 / P.tm;
        const double inv_gjl = 1.0 / P.gj_len;
        const bool closed_bnd = (!P.cluster_open) && bnd;

        // ---- Na/K-ATPase (sim_toolbox.py:71-122); Keq = exp(-dG/RT + F vm/RT) = K0/e1
        double fNa = 0.0, fK = 0.0;
        const double K0 = P.K0;                         // exp(-deltaGATP/(R*T_sim))
        if (P.alpha_NaK > 0.0) {
            const double cNai = cmid[P.iNa * C + c], cKi = cmid[P.iK * C + c];
            double cNao, cKo;
            if (P.is_ecm) { cNao = A.cc_env[cur][P.iNa * E + e]; cKo = A.cc_env[cur][P.iK * E + e]; }
            else { cNao = A.cenv_u[cur * 8 + P.iNa]; cKo = A.cenv_u[cur * 8 + P.iK]; }
            const double a = cNao * 1e-3, b = cKi * 1e-3;
            const double Qn = (((P.cADP * 1e-3) * (P.cPi * 1e-3)) * (a * a * a)) * (b * b);
            const double a2 = cNai * 1e-3, b2 = cKo * 1e-3;
            double Qd = ((P.cATP * 1e-3) * (a2 * a2 * a2)) * (b2 * b2);
            if (Qd == 0.0) Qd = 1.0e-15;
            const double Q = Qn / Qd;
            const double Keq = K0 * gm.inv_e1;
            const double u = cNai / P.KmNK_Na, w = cKo / P.KmNK_K, t = P.cATP / P.KmNK_ATP;
            const double u3 = u * u * u, w2 = w * w;
            const double fwd = ((u3 * w2) * t) / (((1.0 + u3) * (1.0 + w2)) * (1.0 + t));
            const double blk = A.NaK_block ? ldg(A.NaK_block + m) : P.NaK_block;
            fNa = (((-3.0 * blk) * P.alpha_NaK) * fwd) * (1.0 - (Q / Keq));
            fK = -(2.0 / 3.0) * fNa;
            if (diag) A.rate_NaK[m] = -fNa;
            fNa = P.rho_pump * fNa;
            fK = P.rho_pump * fK;
            if (closed_bnd) { fNa = 0.0; fK = 0.0; }
        } else if (diag) A.rate_NaK[m] = 0.0;

        // ---- Ca-ATPase (sim.py:2126-2155, sim_toolbox.py:124-182): fresh cell value, env after transport
        double fCa = 0.0;
        if (P.iCa >= 0 && P.alpha_Ca > 0.0) {
            double cCai = A.cc_cells[P.iCa * C + c];
            double cCao = P.is_ecm ? A.cc_env[nxt][P.iCa * E + e] : A.cenv_u[cur * 8 + P.iCa];
            if (cCai != cCai || cCao != cCao) flags |= ST_NAN_CONC;
            if (cCai < 0.0) cCai = 0.0;
            if (cCao < 0.0) cCao = 0.0;
            const double Qn = (P.cADP * P.cPi) * cCao;
            double Qd = P.cATP * cCai;
            if (Qd == 0.0) Qd = 1.0e-16;
            const double Q = Qn / Qd;
            const double Keq = (K0 * gm.inv_e1) * gm.inv_e1;
            const double numo = (cCai / P.KmCa_Ca) * (P.cATP / P.KmCa_ATP);
            const double deno = (1.0 + (cCai / P.KmCa_Ca)) * (1.0 + (P.cATP / P.KmCa_ATP));
            fCa = (-P.alpha_Ca * (numo / deno)) * (1.0 - (Q / Keq));
            fCa = P.rho_pump * fCa;
            if (closed_bnd) fCa = 0.0;
            fCa = P.rho_pump * fCa;
        }

#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const double cin = cmid[i * C + c];
            const double cout = P.is_ecm ? A.cc_env[cur][i * E + e] : A.cenv_u[cur * 8 + i];
            const double Dm = ldg(A.Dm + (size_t)i * Mo + m);
            double alpha, ex, rden;
            gm.get(P.zi[i], P.z[i], a1, alpha, ex, rden);
            // electroflux (sim_toolbox.py:58-65): -((Dc*alpha)/d)*((cB - cA*exp(-alpha))/deno)*rho
            double f = -((Dm * alpha) * inv_tm) * ((cin - cout * ex) * rden) * P.rho_channel;
            if (closed_bnd) f = 0.0;
            if (i == P.iNa) f += fNa;
            if (i == P.iK) f += fK;
            if (i == P.iCa) f += fCa;
            // gap junction: gating advances once per ion (sim.py:1272 -> 2180-2183)
            if (P.v_sensitive_gj) g = gjb * ((g + gnum) * rgden);
            else g = gjb * gjw;
            const double cnb = cmid[i * C + cn];
            gg.get(P.zi[i], P.z[i], ag1, alpha, ex, rden);
            double fg = -(((P.Dgj_surf[i] * g) * alpha) * inv_gjl) * ((cnb - cin * ex) * rden);
            if (bnd) fg = 0.0;
            const double pm = f * sa;
            s_mem[i * BT_TPB + tid] = pm;
            s_gj[i * BT_TPB + tid] = fg * sa;
            if (P.is_ecm) A.flux_slots[(size_t)m * NI + i] = P.fast_update_ecm ? f : pm;
            if (diag) { A.fl_mem[(size_t)i * Mo + m] = f; A.fl_gj[(size_t)i * Mo + m] = fg; }
        }
        A.gjopen[m] = g;
    }
    __syncthreads();

    // ---- membranes -> cells (update_Co + update_all_concs), one thread per (ion, cell)
    for (int p = tid; p < NI * nc; p += BT_TPB) {
        const int i = p / nc, lc = p - i * nc;
        const int c = c0 + lc;
        const int jb = ldgi(A.cell_mem_ptr + c) - m0, je = ldgi(A.cell_mem_ptr + c + 1) - m0;
        double Sm = 0.0, Sg = 0.0;
        for (int j = jb; j < je; ++j) { Sm += s_mem[i * BT_TPB + j]; Sg += s_gj[i * BT_TPB + j]; }
        const double vol = ldg(A.cell_vol + c);
        const double cc = A.cc_cells[i * C + c];
        const double cm_new = cc + (Sm / vol) * P.dt;             // sim_toolbox.py:1177-1181
        double cn_new = cm_new + P.dt * ((-Sg) / vol);            // sim.py:2105-2108
        if (cn_new != cn_new) flags |= ST_NAN_CONC;
        if (cn_new < 0.0) { cn_new = 0.0; flags |= ST_NEG; }      // no_negs, sim.py:2111
        A.cc_cells[i * C + c] = cn_new;
        A.cc_mid[nxt][i * C + c] = cm_new;                        // the stale cc_at_mem (quirk list)
        s_cc[i * BT_MAX_CTA_CELLS + lc] = cn_new;
        s_sum[i * BT_MAX_CTA_CELLS + lc] = Sm;
    }
    __syncthreads();

    // ---- charge and Vmem (ion_current.py:19; sim.py:2027-2029)
    if (tid < nc) {
        const int c = c0 + tid;
        double rho = 0.0;
#pragma unroll
        for (int i = 0; i < NI; ++i) rho = fma(P.zF[i], s_cc[i * BT_MAX_CTA_CELLS + tid], rho);
        if (A.extra_rho_cells) rho += ldg(A.extra_rho_cells + c);
        A.rho_cells[c] = rho;
        const double vmn = P.inv_cm * (rho * ldg(A.diviterm + c));
        if (vmn != vmn) flags |= ST_NAN_VM;
        A.vm_cell[nxt][c] = vmn;
    }
    if (!P.is_ecm && tid < NI) {      // per-CTA partial of sum_m f*sa for the bath (no-ECM)
        double s = 0.0;
        for (int lc = 0; lc < nc; ++lc) s += s_sum[tid * BT_MAX_CTA_CELLS + lc];
        A.cenv_part[(size_t)blockIdx.x * 8 + tid] = s;
    }
    if (flags) atomicOr(A.status, flags);
}

// no-ECM: cX_env = mean(cX_env + (-flux*(mem_sa/vol_env))*dt)  (sim_toolbox.py:1200-1205)
__global__ void k_envmix(const KParams* __restrict__ Pp, const KArrays A, const int cur)
{
    const KParams& P = *Pp;
    __shared__ double red[256];
    const int i = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < P.n_ctas; b += blockDim.x) s += A.cenv_part[(size_t)b * 8 + i];
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
// One thread per env point, all ions: Dirichlet edge fill, central gradient (one-sided at the
// world edge), Nernst-Planck flux with last step's E field, divergence with finitediff.diff's
// edge convention, forward Euler.  Rows are local rows of a strip [y0, y0+ny) of the world.
template <int NI>
__global__ void __launch_bounds__(256)
k_ion(const KParams* __restrict__ Pp, const KArrays A, const int cur, const int diag)
{
    const KParams& P = *Pp;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int nx = P.nx, ny = P.ny;
    if (x >= nx || y >= ny) return;
    const int E = nx * ny;
    const int gy = y + P.y0, gny = P.ny_global;
    const bool top = (gy == gny - 1), bot = (gy == 0), lef = (x == 0), rig = (x == nx - 1);
    // neighbours needed: (y, x+-1), (y+-1, x), (y, x+-2), (y+-2, x); clamp inside the strip –
    // points whose stencil leaves the strip are halo points, recomputed by their owner.
    const int k = y * nx + x;
    const double d = P.delta, inv_d = 1.0 / d, inv_2d = 1.0 / (2.0 * d);
    const double inv_kbT = 1.0 / P.kbT_sim;
    const double* __restrict__ Ex = A.E_x;
    const double* __restrict__ Ey = A.E_y;

    auto edge = [&](int yy, int xx) -> bool {
        const int g = yy + P.y0;
        return g == 0 || g == gny - 1 || xx == 0 || xx == nx - 1;
    };
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const double* __restrict__ c = A.cc_env[cur] + (size_t)i * E;
        const double* __restrict__ D = A.Denv + (size_t)i * E;
        const double cb = P.cbound[i];
        const double z = P.z[i];
        auto C = [&](int yy, int xx) -> double {           // concentration with the Dirichlet fill
            yy = min(max(yy, 0), ny - 1); xx = min(max(xx, 0), nx - 1);
            return edge(yy, xx) ? cb : c[yy * nx + xx];
        };
        auto gcx = [&](int yy, int xx) -> double {         // fd.gradient, finitediff.py:1236-1266
            if (xx == 0) return (C(yy, 1) - C(yy, 0)) * inv_d;
            if (xx == nx - 1) return (C(yy, nx - 1) - C(yy, nx - 2)) * inv_d;
            return -(C(yy, xx - 1) - C(yy, xx + 1)) * inv_2d;
        };
        auto gcy = [&](int yy, int xx) -> double {
            const int g = yy + P.y0;
            if (g == 0) return (C(yy + 1, xx) - C(yy, xx)) * inv_d;
            if (g == gny - 1) return (C(yy, xx) - C(yy - 1, xx)) * inv_d;
            return -(C(yy - 1, xx) - C(yy + 1, xx)) * inv_2d;
        };
        auto FX = [&](int yy, int xx) -> double {          // nernst_planck_flux, sim_toolbox.py:409-411
            yy = min(max(yy, 0), ny - 1); xx = min(max(xx, 0), nx - 1);
            const int kk = yy * nx + xx;
            const double Dk = D[kk];
            const double al = ((Dk * z) * P.q) * inv_kbT;
            return -Dk * gcx(yy, xx) - (al * (-Ex[kk])) * C(yy, xx);
        };
        auto FY = [&](int yy, int xx) -> double {
            yy = min(max(yy, 0), ny - 1); xx = min(max(xx, 0), nx - 1);
            const int kk = yy * nx + xx;
            const double Dk = D[kk];
            const double al = ((Dk * z) * P.q) * inv_kbT;
            return -Dk * gcy(yy, xx) - (al * (-Ey[kk])) * C(yy, xx);
        };
        // fd.divergence(-fx, -fy) with fd.diff's edge rows (finitediff.py:1268-1311)
        double dx, dy;
        if (lef) dx = ((-FX(y, 0)) - (-FX(y, 1))) * inv_d;
        else if (rig) dx = ((-FX(y, nx - 2)) - (-FX(y, nx - 1))) * inv_d;
        else dx = -((-FX(y, x - 1)) - (-FX(y, x + 1))) * inv_2d;
        if (bot) dy = -((-FY(y + 1, x)) - (-FY(y, x))) * inv_d;
        else if (top) dy = -((-FY(y, x)) - (-FY(y - 1, x))) * inv_d;
        else dy = -((-FY(y - 1, x)) - (-FY(y + 1, x))) * inv_2d;
        const double c0 = edge(y, x) ? cb : c[k];
        A.cc_env[cur ^ 1][(size_t)i * E + k] = c0 + (dx + dy) * P.dt;
        if (diag) {
            A.fl_env_x[(size_t)i * E + k] = FX(y, x);
            A.fl_env_y[(size_t)i * E + k] = FY(y, x);
        }
    }
}

// fd.integrator (finitediff.py:1479-1512) applied to the transported field (sim.py:2249-2252).
__global__ void k_ion_smooth(const KParams* __restrict__ Pp, const KArrays A, const int nxt, const int n_ions)
{
    const KParams& P = *Pp;
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
        A.scratch_env[(size_t)i * E + k] = v;
    }
}

// membrane -> env exchange (update_Co env branch + div_env), env charge, raw env voltage.
template <int NI>
__global__ void __launch_bounds__(256)
k_envacc(const KParams* __restrict__ Pp, const KArrays A, const int nxt, const int apply_flux)
{
    const KParams& P = *Pp;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int E = P.nx * P.ny;
    if (k >= E) return;
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
            for (int i = 0; i < NI; ++i) acc[i] = A.flux_slots[(size_t)s * NI + i];
        }
    } else {
        for (int j = s0; j < s1; ++j) {
            const int s = ldgi(A.slot_idx + j);
#pragma unroll
            for (int i = 0; i < NI; ++i) acc[i] += A.flux_slots[(size_t)s * NI + i];
        }
    }
    double rho = 0.0;
    const double msa = P.fast_update_ecm ? ldg(A.memsa_env + k) : 0.0;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        double c = A.cc_env[nxt][(size_t)i * E + k];
        double delta_env;
        if (P.fast_update_ecm) delta_env = ((-acc[i]) * msa) / P.ecm_vol;        // sim_toolbox.py:1222-1225
        else delta_env = (-acc[i]) / P.env_vol_div;                             // sim_toolbox.py:1229
        if (apply_flux) {
            c = c + delta_env * P.dt;
            A.cc_env[nxt][(size_t)i * E + k] = c;
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
#define FT 32
#define FH 5
__global__ void __launch_bounds__(256)
k_field(const KParams* __restrict__ Pp, const KArrays A)
{
    const KParams& P = *Pp;
    __shared__ double sA[FT + 2 * FH][FT + 2 * FH + 1];   // raw
    __shared__ double sB[FT + 2][FT + 2 * FH + 1];        // after the axis-0 (y) pass
    __shared__ double sC[FT + 2][FT + 2 + 1];             // screen * v_env
    const int nx = P.nx, ny = P.ny;
    const int x0 = blockIdx.x * FT, y0 = blockIdx.y * FT;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int gny = P.ny_global;
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
            if (ly >= 1 && ly <= FT && lx >= 1 && lx <= FT) A.v_env[y * nx + x] = v;
        }
        sC[ly][lx] = P.screen * v;
    }
    __syncthreads();
    const double d = P.delta;
    for (int t = tid; t < FT * FT; t += 256) {
        const int ly = t / FT + 1, lx = t % FT + 1;
        const int y = y0 + ly - 1, x = x0 + lx - 1;
        if (x >= nx || y >= ny) continue;
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
k_diag(const KParams* __restrict__ Pp, const KArrays A, const int newb)
{
    __shared__ double s_a[BT_TPB], s_b[BT_TPB];
    __shared__ double s_c[BT_MAX_CTA_CELLS], s_d[BT_MAX_CTA_CELLS];
    const KParams& P = *Pp;
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
        const double Jmem = -dm + (A.extra_J_mem ? ldg(A.extra_J_mem + m) : 0.0);
        A.Jmem[m] = Jmem; A.Jgj[m] = dg;
        Jn0 = Jmem + dg;
        double phi = 0.0;
        if (P.has_phi) phi = ldg(A.phi_b + ldgi(A.map_mem2ecm + m));
        vm = A.vm_cell[newb][c] - phi;
        vmo = A.vm_cell[newb ^ 1][c] - phi;
        A.vm_mem[m] = vm;
        A.dvm[m] = (vm - vmo) / P.dt;
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
        A.vm_ave[cc] = sv / ldg(A.num_mems + cc);
    }
    __syncthreads();
    double Jn = 0.0;
    if (act) {
        const double nmem = ldg(A.num_mems + c);
        const double wm = (P.smooth_cells * nmem - 1.0) / (P.smooth_cells * nmem);   // sim.py:782-786
        const double wo = 1.0 / (P.smooth_cells * nmem);
        Jn = wm * Jn0 + s_c[c - c0] * wo;
        A.Jn[m] = Jn;
        A.I_mem[m] = -Jn * sa;
    }
    __syncthreads();
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
        const double Ecx = Jx / sg, Ecy = Jy / sg;
        A.E_cell_x[cc] = Ecx; A.E_cell_y[cc] = Ecy;
        s_c[tid] = Jx; s_d[tid] = Jy;
    }
    __syncthreads();
    if (act) {
        const double Jx = s_c[c - c0], Jy = s_d[c - c0];
        A.Jc[m] = Jx * nxv + Jy * nyv;
        const double sg = A.sigma_cell[c];
        A.Emc[m] = (Jx / sg) * nxv + (Jy / sg) * nyv;
    }
}

// per-membrane Vmem for download: vm = vm_cell[cell] - Phi_b[map_mem2ecm]
__global__ void k_expand_vm(const KParams* __restrict__ Pp, const KArrays A, const int cur)
{
    const KParams& P = *Pp;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P.n_mems_owned) return;
    double phi = 0.0;
    if (P.has_phi) phi = A.phi_b[A.map_mem2ecm[m]];
    A.vm_mem[m] = A.vm_cell[cur][A.mem_to_cells[m]] - phi;
}

// vm_ave = np.dot(M_sum_mems, vm)/num_mems (sim.py:2038) for download
__global__ void k_vm_ave(const KParams* __restrict__ Pp, const KArrays A)
{
    const KParams& P = *Pp;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_cells_owned) return;
    double s = 0.0;
    for (int m = A.cell_mem_ptr[c]; m < A.cell_mem_ptr[c + 1]; ++m) s += A.vm_mem[m];
    A.vm_ave[c] = s / A.num_mems[c];
}

// ---------------------------------------------------------------------------- launchers
template <int NI>
static void launch_mem_t(const KParams* dP, const KArrays& A, int n_ctas, int cur, int diag, cudaStream_t st)
{
    const size_t smem = (size_t)(2 * NI * BT_TPB + 2 * NI * BT_MAX_CTA_CELLS) * sizeof(double);
    static_assert((2 * 8 * BT_TPB + 2 * 8 * BT_MAX_CTA_CELLS) * sizeof(double) <= 48 * 1024, "fits default smem");
    k_mem<NI><<<n_ctas, BT_TPB, smem, st>>>(dP, A, cur, diag);
}

void launch_mem(int ni, const KParams* dP, const KArrays& A, int n_ctas, int cur, int diag, cudaStream_t st)
{
    switch (ni) {
        case 4: launch_mem_t<4>(dP, A, n_ctas, cur, diag, st); break;
        case 5: launch_mem_t<5>(dP, A, n_ctas, cur, diag, st); break;
        case 6: launch_mem_t<6>(dP, A, n_ctas, cur, diag, st); break;
        case 7: launch_mem_t<7>(dP, A, n_ctas, cur, diag, st); break;
        default: launch_mem_t<8>(dP, A, n_ctas, cur, diag, st); break;
    }
}

void launch_ion(int ni, const KParams* dP, const KArrays& A, int ny, int nx, int cur, int diag, cudaStream_t st)
{
    dim3 b(32, 8), g((nx + 31) / 32, (ny + 7) / 8);
    switch (ni) {
        case 4: k_ion<4><<<g, b, 0, st>>>(dP, A, cur, diag); break;
        case 5: k_ion<5><<<g, b, 0, st>>>(dP, A, cur, diag); break;
        case 6: k_ion<6><<<g, b, 0, st>>>(dP, A, cur, diag); break;
        case 7: k_ion<7><<<g, b, 0, st>>>(dP, A, cur, diag); break;
        default: k_ion<8><<<g, b, 0, st>>>(dP, A, cur, diag); break;
    }
}

void launch_ion_smooth(int ni, const KParams* dP, const KArrays& A, int ny, int nx, int nxt, cudaStream_t st)
{
    dim3 b(32, 8), g((nx + 31) / 32, (ny + 7) / 8);
    k_ion_smooth<<<g, b, 0, st>>>(dP, A, nxt, ni);
}

void launch_envacc(int ni, const KParams* dP, const KArrays& A, int E, int nxt, int apply, cudaStream_t st)
{
    const int g = (E + 255) / 256;
    switch (ni) {
        case 4: k_envacc<4><<<g, 256, 0, st>>>(dP, A, nxt, apply); break;
        case 5: k_envacc<5><<<g, 256, 0, st>>>(dP, A, nxt, apply); break;
        case 6: k_envacc<6><<<g, 256, 0, st>>>(dP, A, nxt, apply); break;
        case 7: k_envacc<7><<<g, 256, 0, st>>>(dP, A, nxt, apply); break;
        default: k_envacc<8><<<g, 256, 0, st>>>(dP, A, nxt, apply); break;
    }
}

// rho_cells and Vmem from the current concentrations (ion_current.py:19; sim.py:2027-2029)
__global__ void k_cell_charge(const KParams* __restrict__ Pp, const KArrays A, const int cur)
{
    const KParams& P = *Pp;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_cells_owned) return;
    double rho = 0.0;
    for (int i = 0; i < P.n_ions; ++i) rho = fma(P.zF[i], A.cc_cells[(size_t)i * P.n_cells + c], rho);
    if (A.extra_rho_cells) rho += A.extra_rho_cells[c];
    A.rho_cells[c] = rho;
    A.vm_cell[cur][c] = P.inv_cm * (rho * A.diviterm[c]);
}

void launch_cell_charge(const KParams* dP, const KArrays& A, int C, int cur, cudaStream_t st)
{
    k_cell_charge<<<(C + 255) / 256, 256, 0, st>>>(dP, A, cur);
}

void launch_field(const KParams* dP, const KArrays& A, int ny, int nx, cudaStream_t st)
{
    dim3 b(32, 8), g((nx + FT - 1) / FT, (ny + FT - 1) / FT);
    k_field<<<g, b, 0, st>>>(dP, A);
}

void launch_envmix(int ni, const KParams* dP, const KArrays& A, int cur, cudaStream_t st)
{
    k_envmix<<<ni, 256, 0, st>>>(dP, A, cur);
}

void launch_diag(int ni, const KParams* dP, const KArrays& A, int n_ctas, int newb, cudaStream_t st)
{
    switch (ni) {
        case 4: k_diag<4><<<n_ctas, BT_TPB, 0, st>>>(dP, A, newb); break;
        case 5: k_diag<5><<<n_ctas, BT_TPB, 0, st>>>(dP, A, newb); break;
        case 6: k_diag<6><<<n_ctas, BT_TPB, 0, st>>>(dP, A, newb); break;
        case 7: k_diag<7><<<n_ctas, BT_TPB, 0, st>>>(dP, A, newb); break;
        default: k_diag<8><<<n_ctas, BT_TPB, 0, st>>>(dP, A, newb); break;
    }
}

void launch_expand_vm(const KParams* dP, const KArrays& A, int M, int C, int cur, cudaStream_t st)
{
    k_expand_vm<<<(M + 255) / 256, 256, 0, st>>>(dP, A, cur);
    k_vm_ave<<<(C + 255) / 256, 256, 0, st>>>(dP, A);
}
