// General network / gene regulatory network (the part implemented so far):
//   MasterOfNetworks.run_loop                 betse/science/chemistry/networks.py:2805-2982
//   Molecule.transport -> stb.molecule_mover  networks.py:5670-5700, sim_toolbox.py:909-1153 (gap-junction branch)
//   write_growth_and_decay / write_reactions  networks.py:1187-1572 (the strings; compiled by betse_b200/ratelaw.py)
//
// k_net       one thread per cell: every growth/decay and reaction rate from the OLD concentrations
//             (the reference evals all strings first, networks.py:2826-2853), delta = reaction_matrix . rates,
//             c += delta*dt (networks.py:2856-2914); substances that do not pass gap junctions are checked for
//             negative values here (sim_toolbox.py:1124-1150 raises, networks.py:2945 clamps).
// k_net_gj    substances that pass gap junctions: GHK flux between the two cells of every membrane from the
//             updated concentrations, zero at the cluster boundary (sim_toolbox.py:976-1006), summed per cell;
// k_net_gj_apply  c += dt*delta*time_dilation_factor, negative check.
#include "../../include/betse_b200.h"
#include "kparams.cuh"
#include "network.cuh"

#define FLOAT_NONCE 1.0e-25
#define ST_NEG_NET 16u

// (two kernels: one thread per (rate law, cell) — a 50 k-cell tissue has too few cells to hide the interpreter's latency
// with one thread per cell walking all ~20 programs —, then one thread per cell for np.dot(reaction_matrix, all_rates) and
// the update, in the order of the single loop this replaces)
__global__ void __launch_bounds__(128)
k_net_rates(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int cur)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (c >= P.n_cells_owned) return;
    const int C = P.n_cells, M = P.n_mems_owned;
    double v = rl_eval(N, j, c, -1, A, C, M, cur, 0.0);
    if (j < N.K && N.gmask && !N.gmask[(size_t)j * C + c]) v = 0.0;         // mat[trgs] = rts[trgs], networks.py:2844-2846
    N.rates[(size_t)j * C + c] = v;
}

__global__ void __launch_bounds__(128)
k_net(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int cur)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_cells_owned) return;
    const int C = P.n_cells;
    double r[NET_MAX_RATES];
    for (int j = 0; j < N.n_rates; ++j) r[j] = N.rates[(size_t)j * C + c];
    unsigned int flags = 0;
    for (int k = 0; k < N.K; ++k) {
        double d = 0.0;
        for (int j = 0; j < N.n_rates; ++j) d += __ldg(N.stoich + k * N.n_rates + j) * r[j];   // np.dot(reaction_matrix, all_rates)
        double cn = N.c[(size_t)k * C + c] + d * P.dt;                      // networks.py:2914
        if (N.clamp) { const double cl = N.clamp[k]; if (cl == cl) cn = cl; }   // cell_clamp_method: before the transport (networks.py:2932-2936)
        // update_Co cell branch, sim_toolbox.py:1177-1181 (a pumped substance gets its membrane leg after the pump)
        if (N.mem_delta && __ldg(N.Dm + k) != 0.0 && !(N.pumped && N.pumped[k])) cn = cn + N.mem_delta[(size_t)k * C + c] * P.dt;
        if (__ldg(N.Dgj + k) < 0.0 && cn < 0.0) { flags |= ST_NEG_NET; cn = 0.0; }
        N.c[(size_t)k * C + c] = cn;
    }
    if (flags) atomicOr(A.status, flags);
}

// Reactions outside the cells (write_reactions_env, networks.py:1830-2088): one thread per env square evaluates the
// extracellular-zone rate laws from the env concentrations the step found and applies reaction_matrix_env . rates * dt to
// the substances that live there — the top of run_loop (networks.py:2872-2889), before any transport of the step.
#define NET_MAX_ENV_RX 16
__global__ void __launch_bounds__(128)
k_net_env_rx(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int cur)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N.E) return;
    double r[NET_MAX_ENV_RX];
    for (int j = 0; j < N.n_env_rx; ++j) r[j] = rl_eval(N, __ldg(N.env_rx_prog + j), 0, -2 - e, A, P.n_cells, P.n_mems_owned, cur, 0.0);
    for (int k = 0; k < N.K; ++k) {
        if (!N.env_on_d[k]) continue;
        double d = 0.0;
        for (int j = 0; j < N.n_env_rx; ++j) d += __ldg(N.stoich_env + k * N.n_env_rx + j) * r[j];      // np.dot(reaction_matrix_env, rates)
        double* c = N.c_env + (size_t)k * N.E + e;
        *c = *c + d * P.dt;
    }
}

// gap-junction transport of substance k: per-cell sum of -f_gj*mem_sa (one warp per tile, the packing of k_mem)
__global__ void __launch_bounds__(BT_TPB)
k_net_gj(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int k, const int cur,
         const int nonces, const int has_mem, const int use_cmem)
{
    __shared__ double s_all[(BT_TPB / 32) * 32];
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (BT_TPB / 32) + (threadIdx.x >> 5);
    if (tile >= P.n_tiles) return;
    double* s_f = s_all + (threadIdx.x >> 5) * 32;
    const int4 td = __ldg(reinterpret_cast<const int4*>(A.tile_desc) + tile);
    const int c0 = td.x, nc = td.y, m0 = td.z, nm = td.w;
    const int C = P.n_cells;
    const double* __restrict__ cc = N.c + (size_t)k * C;
    double fsa = 0.0;
    if (lane < nm) {
        const int m = m0 + lane;
        const int c = __ldg(A.mem_to_cells + m);
        const int nnp = __ldg(A.nn_cell_flag + m);
        const int cn = nnp & 0x7fffffff;
        double fgj = 0.0;
        if (nnp >= 0) {                                                    // fgj_X[cells.bflags_mems] = 0
            const double gjb = A.gj_block ? __ldg(A.gj_block + m) : P.gj_block;
            const double D = (__ldg(N.Dgj + k) * gjb) * A.gjopen[m];          // Dgj*sim.gj_block*sim.gjopen
            double va = A.vm_cell[cur][cn], vb = A.vm_cell[cur][c];
            if (P.polar) { va = A.vm_pol[cur][__ldg(A.nn_i + m)]; vb = A.vm_pol[cur][m]; }
            else if (P.has_phi) { va -= __ldg(A.phi_b_old + __ldg(A.map_mem2ecm + __ldg(A.nn_i + m))); vb -= __ldg(A.phi_b_old + __ldg(A.map_mem2ecm + m)); }
            double vBA = va - vb;                                          // sim.vgj (sim.py:2166) ...
            for (int q = 0; q < nonces; ++q) vBA += FLOAT_NONCE;           // ... after the in-place `vBA += 1e-25` of every earlier electroflux call
            const double zc = __ldg(N.z + k) + FLOAT_NONCE;
            const double alpha = ((zc * vBA) * P.F) / P.RT_p;              // p.T (sim_toolbox.py:986)
            const double ex = exp(-alpha), deno = -expm1(-alpha);
            double cA = cc[c], cB = cc[cn];                                // cX_mems[mem_i], cX_mems[nn_i]
            if (use_cmem) {                                                // 'update intracellular': the transported membrane values
                const double* __restrict__ cmv = N.cmem + (size_t)k * P.n_mems_owned;
                cA = cmv[m]; cB = cmv[__ldg(A.nn_i + m)];
            }
            const double f = -((D * alpha) / P.gj_len) * ((cB - cA * ex) / deno);
            fsa = -f * __ldg(A.mem_sa + m);
            fgj = f;
        }
        if (use_cmem) N.gjf[m] = fgj;
        if (N.affect) {                                                    // networks.py:2946
            const double z = __ldg(N.z + k), sc = __ldg(N.scale + k);
            const double fm = has_mem ? N.fmem_tmp[m] : 0.0;
            A.chanJ[m] += (((-z) * fm) * P.F) * sc + ((z * fgj) * P.F) * sc;
        }
    }
    s_f[lane] = fsa;
    __syncwarp();
    if (lane < nc) {
        const int c = c0 + lane;
        const int jb = __ldg(A.cell_mem_ptr + c) - m0, je = __ldg(A.cell_mem_ptr + c + 1) - m0;
        double S = 0.0;
        for (int j = jb; j < je; ++j) S += s_f[j];
        N.gj_delta[c] = S / __ldg(A.cell_vol + c);                         // delta_cco, sim_toolbox.py:994
    }
}

__global__ void __launch_bounds__(256)
k_net_gj_apply(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int k)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_cells_owned) return;
    double cn = N.c[(size_t)k * P.n_cells + c] + (P.dt * N.gj_delta[c]) * __ldg(N.tdf + k);   // sim_toolbox.py:999
    if (cn < 0.0) { atomicOr(A.status, ST_NEG_NET); cn = 0.0; }
    N.c[(size_t)k * P.n_cells + c] = cn;
}

// ---------------------------------------------------------------------------- transporters
// run_loop_transporters (networks.py:2985-3107), cell zone with extracellular spaces.
//   k_trans_flux   flux = rho_pump * eval(transporter_eval_string) on every membrane (one warp per tile of whole cells),
//                  f*mem_sa into the exchange slots, the per-cell sum of it, extra_J_mem += net_z*flux*F
//   k_trans_cell   X[cell] += coeff * ((sign * sum_mems(flux*mem_sa)) / cell_vol) * dt on the target cells
//   k_trans_env    X[env]  += coeff * div_env(sign*flux) * dt on the target env squares
//   k_trans_tweak  the reference also nudges mem_concs[X] at the target membranes by -/+flux*(mem_sa/mem_vol)*dt
//                  (networks.py:3020-3022, 3070-3072); the value lives until the next update_intra / update_Co and is read
//                  by LATER membrane-zone rate laws of the same step: kept as rows of membrane values (KNet.tw, gathered
//                  from the cells by k_tw_gather) that rl_eval reads instead of the cell arrays.
__global__ void __launch_bounds__(BT_TPB)
k_trans_flux(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int prog, const double net_z,
             const int cur)
{
    __shared__ double s_all[(BT_TPB / 32) * 32];
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (BT_TPB / 32) + (threadIdx.x >> 5);
    if (tile >= P.n_tiles) return;
    double* s_f = s_all + (threadIdx.x >> 5) * 32;
    const int4 td = __ldg(reinterpret_cast<const int4*>(A.tile_desc) + tile);
    const int c0 = td.x, nc = td.y, m0 = td.z, nm = td.w;
    double fsa = 0.0;
    if (lane < nm) {
        const int m = m0 + lane;
        const int c = __ldg(A.mem_to_cells + m);
        double vm = A.vm_cell[cur][c];
        if (P.polar) vm = A.vm_pol[cur][m];
        else if (P.has_phi) vm -= __ldg(A.phi_b_old + __ldg(A.map_mem2ecm + m));
        const double f = P.rho_pump * rl_eval(N, prog, c, m, A, P.n_cells, P.n_mems_owned, cur, vm);
        fsa = f * __ldg(A.mem_sa + m);
        A.chan_slots[m] = fsa;
        N.tr_flux[m] = f;
        if (N.affect) A.chanJ[m] += (net_z * f) * P.F;                      // networks.py:3004
    }
    s_f[lane] = fsa;
    __syncwarp();
    if (lane < nc) {
        const int c = c0 + lane;
        const int jb = __ldg(A.cell_mem_ptr + c) - m0, je = __ldg(A.cell_mem_ptr + c + 1) - m0;
        double S = 0.0;
        for (int j = jb; j < je; ++j) S += s_f[j];
        N.gj_delta[c] = S;                                                  // scratch [C]: sum_mems(flux*mem_sa)
    }
}

__global__ void __launch_bounds__(256)
k_trans_cell(const __grid_constant__ KParams P, const KArrays A, const double* __restrict__ S, double* __restrict__ dst,
             const double sign, const double coeff, const unsigned char* __restrict__ mask)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_cells_owned || (mask && !mask[c])) return;
    dst[c] = dst[c] + (coeff * ((sign * S[c]) / __ldg(A.cell_vol + c))) * P.dt;
}

__global__ void __launch_bounds__(256)
k_trans_env(const __grid_constant__ KParams P, const KArrays A, double* __restrict__ dst, const double sign, const double coeff,
            const unsigned char* __restrict__ mask)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.nx * P.ny || (mask && !mask[q])) return;
    const int s0 = __ldg(A.slot_ptr + q), s1 = __ldg(A.slot_ptr + q + 1);
    if (s1 == s0) return;
    double acc = 0.0;
    for (int j = s0; j < s1; ++j) acc += sign * A.chan_slots[__ldg(A.slot_idx + j)];
    dst[q] = dst[q] + (coeff * (acc / P.env_vol_div)) * P.dt;              // stb.div_env, sim_toolbox.py:1229
}

__global__ void __launch_bounds__(256)
k_tw_gather(const __grid_constant__ KParams P, const KArrays A, double* __restrict__ row, const double* __restrict__ src)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < P.n_mems_owned) row[m] = src[__ldg(A.mem_to_cells + m)];
}

void launch_tw_gather(const KParams& P, const KArrays& A, double* row, const double* src, cudaStream_t st)
{
    k_tw_gather<<<(P.n_mems_owned + 255) / 256, 256, 0, st>>>(P, A, row, src);
}

__global__ void __launch_bounds__(256)
k_trans_tweak(const __grid_constant__ KParams P, const __grid_constant__ KNet N, const int row, const double sign,
              const unsigned char* __restrict__ mask)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P.n_mems_owned || (mask && !mask[m])) return;
    double* t = N.tw + (size_t)row * P.n_mems_owned + m;
    *t = *t + sign * ((N.tr_flux[m] * __ldg(N.sa_over_vol + m)) * P.dt);
}

void launch_transporter(const KParams& P, const KArrays& A, const KNet& N, const betse_transporter& T,
                        const unsigned char* d_cell_mask, const unsigned char* d_env_mask, const unsigned char* d_mem_mask,
                        int cur, cudaStream_t st)
{
    const int tgrid = (P.n_tiles + (BT_TPB / 32) - 1) / (BT_TPB / 32);
    const int E = P.nx * P.ny, C = P.n_cells;
    k_trans_flux<<<tgrid, BT_TPB, 0, st>>>(P, A, N, T.prog, T.net_z, cur);
    for (int q = 0; q < T.n_terms; ++q) {
        const betse_transporter_term& t = T.terms[q];
        if (t.kind == 0 || t.kind == 2) {
            double* dst = t.kind == 0 ? A.cc_cells + (size_t)t.index * C : N.c + (size_t)t.index * C;
            k_trans_cell<<<(P.n_cells_owned + 255) / 256, 256, 0, st>>>(P, A, N.gj_delta, dst, (double)t.sign, t.coeff, d_cell_mask);
            const int row = t.kind == 0 ? N.tw_i[t.index] : N.tw_s[t.index];
            if (N.tw && row >= 0) k_trans_tweak<<<(P.n_mems_owned + 255) / 256, 256, 0, st>>>(P, N, row, (double)t.sign, d_mem_mask);
        } else {
            double* dst = t.kind == 1 ? A.cc_env[cur ^ 1] + (size_t)t.index * E : N.c_env + (size_t)t.index * E;
            k_trans_env<<<(E + 255) / 256, 256, 0, st>>>(P, A, dst, (double)t.sign, t.coeff, d_env_mask);
        }
    }
}

// run_loop_modulators (networks.py:3282-3325): sim.gj_block / sim.NaKATP_block = max_val * eval(alpha_eval_string),
// membrane zone; in force from the gap-junction transport of the substances of this step onwards.
__global__ void __launch_bounds__(256)
k_net_mod(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int prog, const double max_val,
          double* __restrict__ dst, const int cur)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P.n_mems_owned) return;
    const int c = __ldg(A.mem_to_cells + m);
    double vm = A.vm_cell[cur][c];
    if (P.polar) vm = A.vm_pol[cur][m];
    else if (P.has_phi) vm -= __ldg(A.phi_b_old + __ldg(A.map_mem2ecm + m));
    dst[m] = max_val * rl_eval(N, prog, c, m, A, P.n_cells, P.n_mems_owned, cur, vm);
}

// Tight-junction modulator (run_loop_modulators, target 'TJ', networks.py:3301-3317): modulator = max_val * eval(extracellular-
// zone rate law) on the env squares of the barrier (sim.TJ_targets) becomes sim.TJ_modulator there — for one ion or for all —,
// i.e. the effective diffusion constant the next step's transport reads is D_env * modulator (sim.py:2231-2233).
__global__ void __launch_bounds__(256)
k_net_mod_tj(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int prog, const double max_val,
             const int ion, const int* __restrict__ tj, const int n_tj, const double* __restrict__ denv_raw, double* __restrict__ tj_mod,
             const int cur)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tj) return;
    const int k = __ldg(tj + t);
    const int E = P.nx * P.ny;
    const double v = max_val * rl_eval(N, prog, 0, -2 - k, A, P.n_cells, P.n_mems_owned, cur, 0.0);
    double* Denv = const_cast<double*>(A.Denv);
    for (int i = (ion >= 0 ? ion : 0); i < (ion >= 0 ? ion + 1 : P.n_ions); ++i) {
        tj_mod[(size_t)i * E + k] = v;                                          // sim.TJ_modulator itself (what the host reads back)
        Denv[(size_t)i * E + k] = __ldg(denv_raw + (size_t)i * E + k) * v;
    }
}

void launch_net_mod_tj(const KParams& P, const KArrays& A, const KNet& N, int prog, double max_val, int ion, const int* tj, int n_tj,
                       const double* denv_raw, double* tj_mod, int cur, cudaStream_t st)
{
    if (n_tj > 0) k_net_mod_tj<<<(n_tj + 255) / 256, 256, 0, st>>>(P, A, N, prog, max_val, ion, tj, n_tj, denv_raw, tj_mod, cur);
}

void launch_net_mod(const KParams& P, const KArrays& A, const KNet& N, int prog, double max_val, double* dst, int cur, cudaStream_t st)
{
    k_net_mod<<<(P.n_mems_owned + 255) / 256, 256, 0, st>>>(P, A, N, prog, max_val, dst, cur);
}

// Ligand-gated channel (Molecule.gating, networks.py:5847-5916), one warp per tile of whole cells (the packing of
// k_mem): Hill opening from the ligand at the membrane, GHK flux of the conducted ion, which the reference both ADDS to
// sim.fluxes_mem (so update_all_concs applies it: deferred cell sums, membrane->env exchange slots, Jmem) and applies
// IMMEDIATELY through update_Co (cells here; env squares / the bath by k_chan_env / k_chan_mix).
__device__ __forceinline__ double np_pow(const double a, const double b)      // NumPy's scalar-exponent fast paths, as in rl_eval
{
    return (b == 1.0) ? a : (b == 2.0) ? a * a : (b == 0.5) ? sqrt(a) : (b == -1.0) ? 1.0 / a : (b == 0.0) ? 1.0 : pow(a, b);
}

// The opening of a gate is formed from the ligand as the step FOUND it (cc_at_mem is refreshed by update_intra before
// the substance's growth/decay, its c_env moves only in its own transport, which follows the gating): k_lig_prep runs
// before the handler's substances are advanced, k_net_lig — which only touches ions, as nothing of a substance's
// transport reads them — after it, so the rate laws still see the ions untouched by the gates (networks.py:2826-2853).
__global__ void __launch_bounds__(256)
k_lig_prep(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int sp, const int extracell,
           const double Kn, const double n, const double max_val, double* __restrict__ Dm_mod)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P.n_mems_owned) return;
    const int E = P.ny * P.nx;
    const double x = extracell ? N.c_env[(size_t)sp * E + __ldg(A.map_mem2ecm + m)] : N.c[(size_t)sp * P.n_cells + __ldg(A.mem_to_cells + m)];
    const double xn = np_pow(x, n);
    const double hill = xn / (Kn + xn);                                       // tb.hill, math/toolbox.py:322
    Dm_mod[m] = extracell ? (max_val * hill) : ((P.rho_channel * max_val) * hill);   // networks.py:5859-5873
}

__global__ void __launch_bounds__(BT_TPB)
k_net_lig(const __grid_constant__ KParams P, const KArrays A, const int ion, const double* __restrict__ Dm_mod_arr,
          const double mod, const int cur, const int diag)
{
    __shared__ double s_all[(BT_TPB / 32) * 32];
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (BT_TPB / 32) + (threadIdx.x >> 5);
    if (tile >= P.n_tiles) return;
    double* s_f = s_all + (threadIdx.x >> 5) * 32;
    const int4 td = __ldg(reinterpret_cast<const int4*>(A.tile_desc) + tile);
    const int c0 = td.x, nc = td.y, m0 = td.z, nm = td.w;
    const int C = P.n_cells, E = P.ny * P.nx, NI = P.n_ions, Mo = P.n_mems_owned;
    double* __restrict__ ccell = A.cc_cells + (size_t)ion * C;
    double fsa = 0.0;
    if (lane < nm) {
        const int m = m0 + lane;
        const int c = __ldg(A.mem_to_cells + m);
        const int e = __ldg(A.map_mem2ecm + m);
        double vm = A.vm_cell[cur][c];
        if (P.polar) vm = A.vm_pol[cur][m];
        else if (P.has_phi) vm -= __ldg(A.phi_b_old + e);
        const double Dchan = (P.rho_channel * Dm_mod_arr[m]) * mod;
        const double alpha = ((P.z[ion] + FLOAT_NONCE) * (vm + FLOAT_NONCE) * P.F) / P.RT_sim;
        const double ex = exp(-alpha), deno = -expm1(-alpha);
        const double cB = ccell[c], cA = P.is_ecm ? A.cc_env[cur ^ 1][(size_t)ion * E + e] : A.cenv_u[cur * 8 + ion];
        const double f = -((Dchan * alpha) / P.tm) * ((cB - cA * ex) / deno) * P.rho_channel;
        fsa = f * __ldg(A.mem_sa + m);
        // sim.fluxes_mem[ion] += chan_flx: this step's update_all_concs moves it a second time
        if (P.is_ecm) { A.chan_slots[m] = fsa; A.flux_slots[(size_t)m * NI + ion] += fsa; }
        if (diag) A.fl_mem[(size_t)ion * Mo + m] += f;
    }
    s_f[lane] = fsa;
    __syncwarp();
    if (!P.is_ecm && lane == 0) {
        double S = 0.0;
        for (int j = 0; j < nm; ++j) S += s_f[j];
        A.chan_part[tile] = S;
        A.cenv_part[tile * 8 + ion] += S;                                     // deferred share of the bath (k_envmix)
    }
    if (lane < nc) {
        const int c = c0 + lane;
        const int jb = __ldg(A.cell_mem_ptr + c) - m0, je = __ldg(A.cell_mem_ptr + c + 1) - m0;
        double S = 0.0;
        for (int j = jb; j < je; ++j) S += s_f[j];
        A.dsum_m[(size_t)ion * C + c] += S;                                   // deferred: update_all_concs (k_cell_update)
        ccell[c] = ccell[c] + (S / __ldg(A.cell_vol + c)) * P.dt;              // immediate: update_Co cell branch
    }
}

void launch_lig_prep(const KParams& P, const KArrays& A, const KNet& N, int sp, int extracell, double Kn, double n,
                     double max_val, double* Dm_mod, cudaStream_t st)
{
    k_lig_prep<<<(P.n_mems_owned + 255) / 256, 256, 0, st>>>(P, A, N, sp, extracell, Kn, n, max_val, Dm_mod);
}

void launch_net_lig(const KParams& P, const KArrays& A, int ion, const double* Dm_mod, double mod, int cur, int diag, cudaStream_t st)
{
    const int grid = (P.n_tiles + (BT_TPB / 32) - 1) / (BT_TPB / 32);
    k_net_lig<<<grid, BT_TPB, 0, st>>>(P, A, ion, Dm_mod, mod, cur, diag);
}

// ---------------------------------------------------------------------------- membrane + extracellular legs
// Molecule.transport -> stb.molecule_mover (networks.py:5670-5700, sim_toolbox.py:909-1153) for substances with a
// membrane permeability and/or a presence in the environment:
//   k_net_mem       GHK flux between the env square of each membrane and its cell (sim_toolbox.py:943-954) from the
//                   concentration the step STARTED with (cmems = cc_at_mem, refreshed by update_intra before this
//                   step's growth/decay, networks.py:2832 + 5722-5724); per-cell sums for k_net, exchange slots for
//   k_net_env_acc   update_Co env branch: c_env += div_env(-f)*dt (sim_toolbox.py:1189-1195, 1209-1234)
//   k_sub_flux / k_sub_div   Dirichlet fill with c_bound, Nernst-Planck flux with D = Do*D_env_weight(*TJ factor),
//                   divergence with fd.diff's edge rows, forward Euler times the time-dilation factor
//                   (sim_toolbox.py:1061-1112); same stencil conventions as kernels.cu:k_ion.
__global__ void __launch_bounds__(BT_TPB)
k_net_mem(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int k, const int cur,
          const double* __restrict__ csrc, const int use_cmem)
{
    __shared__ double s_all[(BT_TPB / 32) * 32];
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (BT_TPB / 32) + (threadIdx.x >> 5);
    if (tile >= P.n_tiles) return;
    double* s_f = s_all + (threadIdx.x >> 5) * 32;
    const int4 td = __ldg(reinterpret_cast<const int4*>(A.tile_desc) + tile);
    const int c0 = td.x, nc = td.y, m0 = td.z, nm = td.w;
    const int C = P.n_cells, E = P.ny * P.nx;
    const double Dm = __ldg(N.Dm + k);
    const double zc = __ldg(N.z + k) + FLOAT_NONCE;
    double fsa = 0.0;
    if (lane < nm) {
        const int m = m0 + lane;
        const int c = __ldg(A.mem_to_cells + m);
        const int e = __ldg(A.map_mem2ecm + m);
        double vm = A.vm_cell[cur][c];
        if (P.polar) vm = A.vm_pol[cur][m];
        else if (P.has_phi) vm -= __ldg(A.phi_b_old + e);
        const double alpha = ((zc * (vm + FLOAT_NONCE)) * P.F) / P.RT_sim;    // sim.T (sim_toolbox.py:947)
        const double ex = exp(-alpha), deno = -expm1(-alpha);
        double* cmv = use_cmem ? N.cmem + (size_t)k * P.n_mems_owned + m : nullptr;
        const double cA = N.c_env[(size_t)k * E + e], cB = cmv ? *cmv : csrc ? csrc[c] : N.c[(size_t)k * C + c];
        double f = -((Dm * alpha) / P.tm) * ((cB - cA * ex) / deno) * P.rho_channel;
        if (!P.cluster_open && __ldg(A.nn_cell_flag + m) < 0) f = 0.0;        // f_X_ED[cells.bflags_mems] = 0
        if (cmv) *cmv = cB + (f * (__ldg(N.sa_over_vol + m) / 0.75)) * P.dt;  // update_Co(update_at_mems=True), sim_toolbox.py:1183-1185
        fsa = f * __ldg(A.mem_sa + m);
        A.chan_slots[m] = fsa;
        if (N.affect) {
            if (__ldg(N.Dgj + k) >= 0.0) N.fmem_tmp[m] = f;                  // joined with f_gj in k_net_gj
            else {                                                          // networks.py:2946 with f_gj == 0
                const double z = __ldg(N.z + k), sc = __ldg(N.scale + k);
                A.chanJ[m] += (((-z) * f) * P.F) * sc + ((z * 0.0) * P.F) * sc;
            }
        }
    }
    s_f[lane] = fsa;
    __syncwarp();
    if (lane < nc) {
        const int c = c0 + lane;
        const int jb = __ldg(A.cell_mem_ptr + c) - m0, je = __ldg(A.cell_mem_ptr + c + 1) - m0;
        double S = 0.0;
        for (int j = jb; j < je; ++j) S += s_f[j];
        N.mem_delta[(size_t)k * C + c] = S / __ldg(A.cell_vol + c);
    }
}

// extra_rho_cells / extra_rho_env of the handler: F*c*z*scale_factor over its substances (networks.py:2945, 2950)
__global__ void __launch_bounds__(256)
k_net_charge(const __grid_constant__ KParams P, const __grid_constant__ KNet N, const int n, const int stride, const int env)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const double* __restrict__ src = env ? N.c_env : N.c;
    double rho = 0.0;
    for (int k = 0; k < N.K; ++k) rho += ((P.F * src[(size_t)k * stride + q]) * __ldg(N.z + k)) * __ldg(N.scale + k);
    (env ? N.rho_env : N.rho_cells)[q] = rho;
}

// Molecule.pump (networks.py:5809-5844) -> stb.molecule_pump / stb.molecule_transporter (sim_toolbox.py:658-907) with n = 1,
// Km_ATP = 1 and Keq = 1 as Molecule.pump passes them: flux from the substance AFTER its growth/decay and its env square,
// immediate update_Co (cells here, env squares by k_net_env_acc).
__global__ void __launch_bounds__(BT_TPB)
k_net_pump(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int k, const int into_cell,
           const int uses_ATP, const double alpha_max, const double Km, const double dG_RT, const int cur)
{
    __shared__ double s_all[(BT_TPB / 32) * 32];
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (BT_TPB / 32) + (threadIdx.x >> 5);
    if (tile >= P.n_tiles) return;
    double* s_f = s_all + (threadIdx.x >> 5) * 32;
    const int4 td = __ldg(reinterpret_cast<const int4*>(A.tile_desc) + tile);
    const int c0 = td.x, nc = td.y, m0 = td.z, nm = td.w;
    const int C = P.n_cells, E = P.ny * P.nx;
    const double z = __ldg(N.z + k);
    double* __restrict__ cc = N.c + (size_t)k * C;
    double fsa = 0.0;
    if (lane < nm) {
        const int m = m0 + lane;
        const int c = __ldg(A.mem_to_cells + m);
        const int e = __ldg(A.map_mem2ecm + m);
        double vm = A.vm_cell[cur][c];
        if (P.polar) vm = A.vm_pol[cur][m];
        else if (P.has_phi) vm -= __ldg(A.phi_b_old + e);
        const double xe = N.c_env[(size_t)k * E + e], xc = cc[c];
        double f;
        if (uses_ATP) {
            const double cATP = P.cATP;
            if (!into_cell) {
                const double Qn = (P.cADP * P.cPi) * xe;
                double Qd = cATP * xc;
                if (Qd == 0.0) Qd = 1.0e-10;
                const double Keq = exp(-dG_RT + ((((1.0 * z) * P.F) * vm) / P.RT_sim));
                const double alpha = alpha_max * (1.0 - ((Qn / Qd) / Keq));
                const double numo = (xc / Km) * (cATP / 1.0), deno = (1.0 + (xc / Km)) * (1.0 + (cATP / 1.0));
                f = ((-P.rho_pump) * alpha) * (numo / deno);
            } else {
                const double Qn = (P.cADP * P.cPi) * xc;
                double Qd = cATP * xe;
                if (Qd == 0.0) Qd = 1.0e-10;
                const double Keq = exp(-dG_RT - (((z * P.F) * vm) / P.RT_sim));
                const double alpha = alpha_max * (1.0 - ((Qn / Qd) / Keq));
                const double numo = (xe / Km) * (cATP / 1.0), deno = (1.0 + (xe / Km)) * (1.0 + (cATP / 1.0));
                f = (P.rho_pump * alpha) * (numo / deno);
            }
        } else if (!into_cell) {
            double Qd = xc;
            if (Qd == 0.0) Qd = 1.0e-15;
            const double Keq = 1.0 * exp(((z * P.F) * vm) / P.RT_sim);
            const double alpha = alpha_max * (1.0 - ((xe / Qd) / Keq));
            f = ((-P.rho_pump) * alpha) * ((xc / Km) / (1.0 + (xc / Km)));
        } else {
            double Qd = xe;
            if (Qd == 0.0) Qd = 1.0e-15;
            const double Keq = 1.0 * exp(-((z * P.F) * vm) / P.RT_sim);
            const double alpha = alpha_max * (1.0 - ((xc / Qd) / Keq));
            f = (P.rho_pump * alpha) * ((xe / Km) / (1.0 + (xe / Km)));
        }
        if (!P.cluster_open && __ldg(A.nn_cell_flag + m) < 0) f = 0.0;
        fsa = f * __ldg(A.mem_sa + m);
        A.chan_slots[m] = fsa;
    }
    s_f[lane] = fsa;
    __syncwarp();
    if (lane < nc) {
        const int c = c0 + lane;
        const int jb = __ldg(A.cell_mem_ptr + c) - m0, je = __ldg(A.cell_mem_ptr + c + 1) - m0;
        double S = 0.0;
        for (int j = jb; j < je; ++j) S += s_f[j];
        cc[c] = cc[c] + (S / __ldg(A.cell_vol + c)) * P.dt;                   // update_Co cell branch
    }
}

// the deferred membrane leg of a pumped substance: c += sum_mems(f*sa)/vol * dt
__global__ void __launch_bounds__(256)
k_net_apply_mem(const __grid_constant__ KParams P, const __grid_constant__ KNet N, const int k)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_cells_owned) return;
    N.c[(size_t)k * P.n_cells + c] = N.c[(size_t)k * P.n_cells + c] + N.mem_delta[(size_t)k * P.n_cells + c] * P.dt;
}

__global__ void __launch_bounds__(256)
k_net_env_acc(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int k)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int E = P.nx * P.ny;
    if (q >= E) return;
    const int s0 = __ldg(A.slot_ptr + q), s1 = __ldg(A.slot_ptr + q + 1);
    if (s1 == s0) return;
    double acc = 0.0;
    for (int j = s0; j < s1; ++j) acc += A.chan_slots[__ldg(A.slot_idx + j)];
    double* c = N.c_env + (size_t)k * E + q;
    *c = *c + ((-acc) / P.env_vol_div) * P.dt;
}

__device__ __forceinline__ double sub_cval(const double* __restrict__ c, const int y, const int x, const int ny, const int nx, const double cb)
{
    return (y == 0 || y == ny - 1 || x == 0 || x == nx - 1) ? cb : c[(size_t)y * nx + x];
}

__global__ void __launch_bounds__(256)
k_sub_flux(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int k)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int nx = P.nx, ny = P.ny;
    if (x >= nx) return;
    const int E = nx * ny;
    const size_t q = (size_t)y * nx + x;
    const double* __restrict__ c = N.c_env + (size_t)k * E;
    const double cb = __ldg(N.c_bound + k);
    const double inv_d = P.inv_delta, inv_2d = P.inv_2delta;
    const double cc = sub_cval(c, y, x, ny, nx, cb);
    double gcx, gcy;                                       // fd.gradient, finitediff.py:1236-1266
    if (x == 0) gcx = (sub_cval(c, y, 1, ny, nx, cb) - cc) * inv_d;
    else if (x == nx - 1) gcx = (cc - sub_cval(c, y, nx - 2, ny, nx, cb)) * inv_d;
    else gcx = -(sub_cval(c, y, x - 1, ny, nx, cb) - sub_cval(c, y, x + 1, ny, nx, cb)) * inv_2d;
    if (y == 0) gcy = (sub_cval(c, 1, x, ny, nx, cb) - cc) * inv_d;
    else if (y == ny - 1) gcy = (cc - sub_cval(c, ny - 2, x, ny, nx, cb)) * inv_d;
    else gcy = -(sub_cval(c, y - 1, x, ny, nx, cb) - sub_cval(c, y + 1, x, ny, nx, cb)) * inv_2d;
    const double Dk = __ldg(N.D_env + (size_t)k * E + q);
    const double al = (Dk * (__ldg(N.z + k) * P.q)) * P.inv_kbT_sim;          // nernst_planck_flux, sim_toolbox.py:409-411
    const double fx = -Dk * gcx - (al * (-A.E_x[q])) * cc;
    const double fy = -Dk * gcy - (al * (-A.E_y[q])) * cc;
    N.env_tmp[q] = fx;
    N.env_tmp[E + q] = fy;
    if (N.affect) {                                                        // networks.py:2953-2954
        const double z = __ldg(N.z + k), sc = __ldg(N.scale + k);
        A.extra_Jenv_x[q] += ((fx * z) * P.F) * sc;
        A.extra_Jenv_y[q] += ((fy * z) * P.F) * sc;
    }
}

__global__ void __launch_bounds__(256)
k_sub_div(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int k)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int nx = P.nx, ny = P.ny;
    if (x >= nx) return;
    const int E = nx * ny;
    const size_t q = (size_t)y * nx + x;
    const double* __restrict__ Fx = N.env_tmp;
    const double* __restrict__ Fy = N.env_tmp + E;
    const double inv_d = P.inv_delta, inv_2d = P.inv_2delta;
    double dx, dy;                                         // fd.divergence(-fx, -fy) with fd.diff's edge rows (finitediff.py:1268-1311)
    if (x == 0) dx = ((-Fx[q]) - (-Fx[q + 1])) * inv_d;
    else if (x == nx - 1) dx = ((-Fx[q - 1]) - (-Fx[q])) * inv_d;
    else dx = -((-Fx[q - 1]) - (-Fx[q + 1])) * inv_2d;
    if (y == 0) dy = -((-Fy[q + nx]) - (-Fy[q])) * inv_d;
    else if (y == ny - 1) dy = -((-Fy[q]) - (-Fy[q - nx])) * inv_d;
    else dy = -((-Fy[q - nx]) - (-Fy[q + nx])) * inv_2d;
    double* c = N.c_env + (size_t)k * E;
    double cn = sub_cval(c, y, x, ny, nx, __ldg(N.c_bound + k)) + ((dx + dy) * P.dt) * __ldg(N.tdf + k);
    if (cn < 0.0) { atomicOr(A.status, ST_NEG_NET); cn = 0.0; }            // sim_toolbox.py:1140-1150 raises
    c[q] = cn;
}

// Molecule.update_intra with intracellular transport (networks.py:5727-5795; transmem False, no motor transport): the
// membrane value relaxes implicitly towards the cell value the step started with; charged substances also drift in the
// cell's field normal to the membrane (sim.Emc of the previous step's update_V)
// (one thread per membrane walks ALL substances with intracellular transport: the membrane's geometry, cell and field are
// loaded once, and a network with three such substances is one launch of ~9 us instead of three)
#define NET_INTRA_MAX 32
struct KIntraList { int n; int k[NET_INTRA_MAX]; };

__global__ void __launch_bounds__(256)
k_net_intra(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const __grid_constant__ KIntraList L)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P.n_mems_owned) return;
    const double g = __ldg(N.sa_over_vol + m) / 0.75;                        // gamma = mem_sa/((3/4)*mem_vol)
    const double Rr = __ldg(N.R_rads + m);
    const int c = __ldg(A.mem_to_cells + m);
    const double En = A.Emc ? A.Emc[m] : 0.0;
    bool neg = false;
    for (int j = 0; j < L.n; ++j) {
        const int k = L.k[j];
        const double Do = __ldg(N.Do + k), dt = P.dt * __ldg(N.tdf + k);
        const double cav = N.c[(size_t)k * P.n_cells + c];
        double* cm = N.cmem + (size_t)k * P.n_mems_owned + m;
        const double z = __ldg(N.z + k), mu = N.mu_mem ? __ldg(N.mu_mem + k) : 0.0;
        double alpha_tot = 0.0;
        if (z != 0.0 || mu != 0.0) alpha_tot = ((0.0 + (((Do * P.q) * z) / P.kbT_sim) * En) + mu * En) + 0.0;
        const double v = (((((g * Do) * dt) * cav) / Rr) + ((((g * alpha_tot) * cav) * dt) / 2.0) + *cm) /
                         ((1.0 + (((g * Do) * dt) / Rr)) - (((g * alpha_tot) * dt) / 2.0));
        neg = neg || v < 0.0;                                                // networks.py:5798-5804
        *cm = v;
    }
    if (neg) atomicOr(A.status, ST_NEG_NET);
}

// the membrane values of an 'update intracellular' substance after its gap-junction flux (sim_toolbox.py:1001-1005),
// and molecule_mover's closing check of them (sim_toolbox.py:1139-1142)
__global__ void __launch_bounds__(256)
k_net_cmem_gj(const __grid_constant__ KParams P, const KArrays A, const __grid_constant__ KNet N, const int k)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P.n_mems_owned) return;
    const double g = __ldg(N.sa_over_vol + m) / 0.75;
    double* cm = N.cmem + (size_t)k * P.n_mems_owned + m;
    const double v = *cm + (((-N.gjf[m]) * g) * __ldg(N.tdf + k)) * P.dt;
    if (v < 0.0) atomicOr(A.status, ST_NEG_NET);
    *cm = v;
}

void launch_net(const KParams& P, const KArrays& A, const KNet& N, const double* h_Dgj, const double* h_Dm,
                const unsigned char* h_env_on, const betse_substance_pump* pumps, int n_pumps, double dG_RT,
                const unsigned char* h_intra, int n_ions, int cur, cudaStream_t st)
{
    if (N.K <= 0) return;
    const int tgrid = (P.n_tiles + (BT_TPB / 32) - 1) / (BT_TPB / 32);
    const int E = P.nx * P.ny;
    auto pump_of = [&](int k) { for (int j = 0; j < n_pumps; ++j) if (pumps[j].species == k) return j; return -1; };
    if (N.n_env_rx > 0 && N.c_env) k_net_env_rx<<<(E + 127) / 128, 128, 0, st>>>(P, A, N, cur);
    if (N.cmem && h_intra) {
        KIntraList L;
        L.n = 0;
        for (int k = 0; k <= N.K; ++k) {
            if (k < N.K && h_intra[k]) L.k[L.n++] = k;
            if (L.n == NET_INTRA_MAX || (k == N.K && L.n > 0)) {
                k_net_intra<<<(P.n_mems_owned + 255) / 256, 256, 0, st>>>(P, A, N, L);
                L.n = 0;
            }
        }
    }
    if (N.c_env)
        for (int k = 0; k < N.K; ++k) {
            if (!h_env_on[k] || h_Dm[k] == 0.0 || pump_of(k) >= 0) continue;
            k_net_mem<<<tgrid, BT_TPB, 0, st>>>(P, A, N, k, cur, nullptr, (N.cmem && h_intra && h_intra[k]) ? 1 : 0);
            k_net_env_acc<<<(E + 255) / 256, 256, 0, st>>>(P, A, N, k);
        }
    // pumped substances: pump (from the concentration after growth/decay), then the membrane leg (from the
    // concentration before it: cc_at_mem), each with its immediate update_Co (networks.py:2914-2935)
    for (int j = 0; j < n_pumps; ++j)
        if (h_Dm[pumps[j].species] != 0.0)
            cudaMemcpyAsync(N.c_save + (size_t)j * P.n_cells, N.c + (size_t)pumps[j].species * P.n_cells,
                            (size_t)P.n_cells * sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (N.n_rates > 0) k_net_rates<<<dim3((unsigned)((P.n_cells_owned + 127) / 128), (unsigned)N.n_rates), 128, 0, st>>>(P, A, N, cur);
    k_net<<<(P.n_cells_owned + 127) / 128, 128, 0, st>>>(P, A, N, cur);
    for (int j = 0; j < n_pumps; ++j) {
        const betse_substance_pump& q = pumps[j];
        const int k = q.species;
        k_net_pump<<<tgrid, BT_TPB, 0, st>>>(P, A, N, k, q.into_cell, q.uses_ATP, q.max_val, q.Km, dG_RT, cur);
        k_net_env_acc<<<(E + 255) / 256, 256, 0, st>>>(P, A, N, k);
        if (h_Dm[k] != 0.0) {
            k_net_mem<<<tgrid, BT_TPB, 0, st>>>(P, A, N, k, cur, N.c_save + (size_t)j * P.n_cells, 0);
            k_net_apply_mem<<<(P.n_cells_owned + 255) / 256, 256, 0, st>>>(P, N, k);
            k_net_env_acc<<<(E + 255) / 256, 256, 0, st>>>(P, A, N, k);
        }
    }
    const int grid = (P.n_tiles + (BT_TPB / 32) - 1) / (BT_TPB / 32);
    int nonces = n_ions;
    for (int k = 0; k < N.K; ++k) {
        if (h_Dgj[k] < 0.0) continue;
        ++nonces;
        const int use_cmem = (N.cmem && h_intra && h_intra[k]) ? 1 : 0;
        k_net_gj<<<grid, BT_TPB, 0, st>>>(P, A, N, k, cur, nonces, (N.c_env && h_env_on[k] && h_Dm[k] != 0.0) ? 1 : 0, use_cmem);
        k_net_gj_apply<<<(P.n_cells_owned + 255) / 256, 256, 0, st>>>(P, A, N, k);
        if (use_cmem) k_net_cmem_gj<<<(P.n_mems_owned + 255) / 256, 256, 0, st>>>(P, A, N, k);
    }
    if (N.c_env) {
        dim3 gE((P.nx + 255) / 256, P.ny);
        for (int k = 0; k < N.K; ++k) {
            if (!h_env_on[k]) continue;
            k_sub_flux<<<gE, 256, 0, st>>>(P, A, N, k);
            k_sub_div<<<gE, 256, 0, st>>>(P, A, N, k);
        }
    }
    if (N.affect) {
        k_net_charge<<<(P.n_cells_owned + 255) / 256, 256, 0, st>>>(P, N, P.n_cells_owned, P.n_cells, 0);
        if (N.c_env) k_net_charge<<<(E + 255) / 256, 256, 0, st>>>(P, N, E, E, 1);
    }
}
