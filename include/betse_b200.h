/* betse_b200.h — C ABI of the B200-native BETSE tissue-update engine.
 *
 * Drop-in boundary: the reference has no FFI; its seam is the bound-method
 * dispatch `solver_method = self._run_sim_core_loop` in Simulator.run_sim_core
 * (betse/science/sim.py:1064-1075).  A Python shim (betse_b200/simloop.py) replaces that
 * method and drives this library through ctypes.  Every entry point below states the
 * reference code it stands in for.
 *
 * Conventions: plain pointers + sizes, no torch types.  Host arrays are borrowed for the
 * duration of the call only.  All floating point is IEEE fp64, all indices int32.
 * Return value 0 = OK, non-zero = argument/CUDA error (text via betse_last_error).
 * Numerical instability is NOT an error code: it is reported in the status word
 * (BETSE_STATUS_*), which the shim turns into BetseSimUnstableException
 * (betse/exceptions.py:625; raised by stb.check_v / stb.no_negs, sim_toolbox.py:332-344,475-503).
 * One ctx per GPU and simulation; calls on one ctx must come from one host thread at a time.
 * There is no CPU fallback: every entry point fails if no CUDA device is usable.
 */
#ifndef BETSE_B200_H
#define BETSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BETSE_MAX_IONS 8
#define BETSE_ABI_VERSION 3

typedef struct betse_ctx betse_ctx;

/* Mesh / index arrays the loop consumes — the `Cells` attributes of SURVEY §2 ★(data),
 * built by Cells.make_world (betse/science/cells.py:414-526).  All [M] arrays are in the
 * reference's membrane order (membranes of a cell contiguous, cells.py:1095-1146). */
typedef struct betse_mesh {
    int32_t n_cells;              /* C  = len(cells.cell_i)                                  */
    int32_t n_mems;               /* M  = len(cells.mem_i)                                   */
    int32_t ny, nx;               /* cells.X.shape; E = ny*nx env points (row-major k=y*nx+x) */
    const int32_t *mem_to_cells;  /* [M]   cells.mem_to_cells            cells.py:1126       */
    const int32_t *cell_mem_ptr;  /* [C+1] CSR form of cells.cell_to_mems cells.py:1129-1146 */
    const int32_t *nn_i;          /* [M]   GJ partner membrane            cells.py:1498-1533 */
    const int32_t *map_mem2ecm;   /* [M]   nearest env point              cells.py:1758      */
    const uint8_t *bflags_mems;   /* [M]   1 on cluster-boundary membranes cells.py:1402-1457 */
    const double *mem_sa;         /* [M]   cells.mem_sa                   cells.py:1110      */
    const double *mem_nx;         /* [M]   cells.mem_vects_flat[:,2]                         */
    const double *mem_ny;         /* [M]   cells.mem_vects_flat[:,3]                         */
    const double *R_rads;         /* [M]   cells.R_rads                   cells.py:1173      */
    const double *cell_vol;       /* [C]   cells.cell_vol                 cells.py:1385      */
    const double *cell_sa;        /* [C]   cells.cell_sa                                     */
    const double *diviterm;       /* [C]   cells.diviterm                 cells.py:1390      */
    const double *num_mems;       /* [C]   cells.num_mems (as fp64)       cells.py:1309      */
    const double *memSa_per_envSquare; /* [E] cells.py:1801-1816 (fast_update_ecm only)      */
    const double *gj_default_weights;  /* [M] cells.py:1995-2008 (static-GJ mode only)       */
    double delta;                 /* cells.delta: env grid spacing                           */
    double gj_len;                /* cells.gj_len                                            */
    double ecm_vol;               /* cells.ecm_vol = cell_height*delta^2                     */
    double memsa_mean;            /* cells.memSa_per_envSquare[cells.map_mem2ecm].mean()
                                     (ion_current.py:94-95) – computed by the host in NumPy   */
    /* Domain decomposition.  Single GPU: own everything, y0 = 0, ny_global = ny. */
    int32_t n_cells_owned;        /* cells [0,n_cells_owned) are stepped; the rest are ghosts */
    int32_t n_mems_owned;         /* membranes of owned cells                                 */
    int32_t n_flux_slots;         /* >= n_mems_owned: env-exchange slots incl. remote membranes */
    int32_t y0, ny_global;        /* first local grid row in the global grid; global rows      */
    int32_t y_own0, y_own1;       /* local rows [y_own0,y_own1) are owned (others are halo)     */
    const int32_t *ecm_slot_ptr;  /* [E+1] env point -> flux slots CSR (NULL: build from map_mem2ecm) */
    const int32_t *ecm_slot_idx;  /* [..]                                                      */
} betse_mesh;

/* Scalars the loop reads: physical constants (betse/science/parameters.py:1105-1122), pump
 * constants (:1064-1081), feature flags, and the per-step schedule that
 * TissueHandler.fire_events (tissue/tishandler.py:709-917) rewrites: T, c_env_bound,
 * bound_V.  Re-sent whole through betse_set_schedule whenever any of it changes. */
typedef struct betse_params {
    int32_t abi_version;
    int32_t n_ions;                    /* I                                              */
    int32_t iNa, iK, iCa, iP;          /* ion indices, -1 if the ion is disabled          */
    double z[BETSE_MAX_IONS];          /* sim.zs                                         */
    double D_free[BETSE_MAX_IONS];     /* sim.D_free                                     */
    double D_gj[BETSE_MAX_IONS];       /* sim.D_gj[i] (uniform over membranes)           */
    double c_env_bound[BETSE_MAX_IONS];/* sim.c_env_bound                                */
    double cenv_uniform[BETSE_MAX_IONS];/* no-ECM only: the well-mixed bath value        */
    double F, R, q, kb, eo, er, cm, tm, NAv, mu;
    double T_sim;                      /* sim.T  (membrane GHK, pumps, env Nernst-Planck) */
    double T_p;                        /* p.T    (gap-junction GHK, sigma_cell)           */
    double dt;
    double alpha_NaK, alpha_Ca, KmNK_Na, KmNK_K, KmNK_ATP, KmCa_Ca, KmCa_ATP;
    double cATP, cADP, cPi, deltaGATP;
    double gj_surface, gj_vthresh, gj_min;
    double rho_pump, rho_channel;      /* hard-wired 1 in the reference (sim.py:964-965)  */
    double cell_height, vol_env, cell_radius, true_cell_size;
    double ko_env;                     /* frozen at init_dynamics (sim.py:974)            */
    double sharpness;                  /* p.sharpness (<1: fd.integrator smoothing)       */
    double cell_polarizability;        /* != 0: Vmem is per-membrane state integrated from Jn (sim.py:2048-2080) */
    double smooth_cells;               /* p.smooth_cells (Jn smoothing weights)           */
    double bound_V[4];                 /* T, B, L, R (sim.bound_V)                        */
    double gauss_w[5];                 /* taps w0..w4 of scipy's gaussian_filter(sigma=1) kernel
                                          (ion_current.py:104), formed by the host in NumPy   */
    double NaKATP_block_scalar;        /* used when no per-membrane block array is set    */
    double gj_block_scalar;
    int32_t is_ecm, v_sensitive_gj, cluster_open, fast_update_ecm;
    double sigma_env;                  /* sim.sigma: bath conductivity of the no-ECM field diagnostics (sim.py:986-996,
                                          ion_current.py:142-143)                            */
    double reserved_d;
} betse_params;

/* Host-side view of the Simulator state.  NULL members are skipped.  Shapes in brackets;
 * [I,·] arrays are C-contiguous with the ion index slowest, as np.asarray(sim.cc_cells). */
typedef struct betse_state_host {
    double *cc_cells;        /* [I,C] sim.cc_cells                                         */
    double *cc_at_mem_cell;  /* [I,C] per-cell content of sim.cc_at_mem (it is a gather of a
                                       per-cell quantity: sim_toolbox.py:1182, sim.py:2310)   */
    double *cc_env;          /* [I,E] sim.cc_env (ECM); ignored without ECM               */
    double *vm;              /* [M]   sim.vm                                              */
    double *gjopen;          /* [M]   sim.gjopen                                          */
    double *Dm_cells;        /* [I,M] sim.Dm_cells                                        */
    double *D_env_eff;       /* [I,E] sim.D_env * sim.TJ_modulator (sim.py:2231-2233)     */
    double *E_env_x;         /* [E]   sim.E_env_x                                         */
    double *E_env_y;         /* [E]   sim.E_env_y                                         */
    double *v_env;           /* [E]   sim.v_env      (download only)                      */
    double *rho_env;         /* [E]   sim.rho_env    (download only)                      */
    double *rho_cells;       /* [C]   sim.rho_cells  (download only)                      */
    double *vm_ave;          /* [C]   sim.vm_ave     (download only)                      */
    double *Phi_b;           /* [E]   sim.Phi_b (boundary-voltage potential; NULL = 0)    */
    double *extra_rho_cells; /* [C]   sim.extra_rho_cells (NULL = 0)                      */
    double *extra_rho_env;   /* [E]   sim.extra_rho_env   (NULL = 0)                      */
    double *extra_J_mem;     /* [M]   sim.extra_J_mem     (NULL = 0)                      */
    double *NaKATP_block;    /* [M]   sim.NaKATP_block when it is an array                */
    double *gj_block;        /* [M]   sim.gj_block when it is an array                    */
    /* diagnostics of the last step run with BETSE_STEP_DIAG (download only) */
    double *fluxes_mem;      /* [I,M] */
    double *fluxes_gj;       /* [I,M] */
    double *fluxes_env_x;    /* [I,E] */
    double *fluxes_env_y;    /* [I,E] */
    double *rate_NaKATP;     /* [M]   */
    double *Jmem, *Jgj, *Jn, *I_mem, *Jc, *Emc, *dvm;   /* [M] */
    double *J_cell_x, *J_cell_y, *E_cell_x, *E_cell_y, *sigma_cell;  /* [C] */
    double *E_gj_x, *E_gj_y; /* [M]   gap-junction field of the step (sim.py:2168-2172; stored by write2storage) */
    double *J_env_x, *J_env_y, *B_field, *Jtx, *Jty;  /* [E] Helmholtz-Hodge decomposition of the env current
                                (ion_current.py:50-73, sim_toolbox.py:1236-1290); undivided ECM tissues */
    double *cenv_uniform;    /* [I]   no-ECM bath concentrations (download only)          */
    double *vm_cell;         /* [C]   per-cell Vmem incl. ghost cells (upload only; overrides vm)  */
    double *D_env_weight;    /* [E]   sim.D_env_weight (no-ECM field diagnostics, ion_current.py:142-143; upload only) */
} betse_state_host;

#define BETSE_STATUS_NAN_VM     1u  /* stb.check_v would raise          */
#define BETSE_STATUS_NAN_CONC   2u  /* stb.no_negs would raise          */
#define BETSE_STATUS_NEG_CLAMP  4u  /* informational: a negative concentration was clamped */

#define BETSE_STEP_DIAG   1  /* last step of the call also produces the sampled-step diagnostics */

/* Number of distinct kernels a step may launch (for betse_step_profile). */
#define BETSE_NKERNELS 8

/* Library / device probing (no ctx needed). */
int  betse_abi_version(void);
int  betse_device_count(void);

/* Replaces: state set-up at the top of _run_sim_core_loop — takes the mesh and constants the
 * loop closes over (cells.*, p.*), builds device-side SoA + the cell/CTA packing + env CSR. */
int  betse_create(betse_ctx **out, const betse_mesh *mesh, const betse_params *params, int device);
void betse_destroy(betse_ctx *ctx);
int  betse_last_error(betse_ctx *ctx, char *buf, size_t n);
int  betse_create_error(char *buf, size_t n);   /* message of the last failed betse_create */

/* Replaces: nothing in the reference (its state already lives in host NumPy); moves the
 * Simulator attributes created by init_core/init_dynamics (sim.py:452-1012) into HBM. */
int  betse_upload_state(betse_ctx *ctx, const betse_state_host *state);

/* Replaces: the Simulator.update_V call that precedes the loop (sim.py:1041 -> 2007-2083 ->
 * get_current, ion_current.py:13-114): charge, Vmem and the environmental field from the
 * uploaded concentrations.  Needed only when the caller did not run the reference's update_V. */
int  betse_update_v(betse_ctx *ctx);

/* Replaces: the array/scalar side effects of TissueHandler.fire_events + makeAllChanges
 * (tishandler.py:709-917,1321-1332): new T / c_env_bound / bound_V / blocks. */
int  betse_set_schedule(betse_ctx *ctx, const betse_params *params);

/* Replaces: `nsteps` iterations of the loop body, sim.py:1169-1365 (steps 0-11 of SURVEY §3.2).
 * status_out (nullable) receives the OR of BETSE_STATUS_* over all steps of this call. */
int  betse_step(betse_ctx *ctx, int nsteps, int flags, uint32_t *status_out);

/* Ensembles of small tissues (last bullet; BASELINE configs[0]-[1] are launch-latency bound): the reference
 * runs one Simulator per process and parameter set (simrunner.py:93-296 per configuration file); here `n` independent
 * contexts of ONE device advance `nsteps` timesteps each per graph launch, `launches` launches back to back — each member's
 * kernels on its own stream inside ONE CUDA graph.  Every member must have been stepped at least once through betse_step
 * (kernels loaded) and all members must be in lockstep.  Bit-identical to stepping each member alone.  status_out
 * (nullable) receives one status word per member, device_ms (nullable) the device time of the launches. */
int  betse_ensemble_step(betse_ctx **ctxs, int n, int nsteps, int launches, uint32_t *status_out, float *device_ms);

/* ---------------------------------------------------------------------------------------------
 * The FAST (equivalent-circuit) solver: Simulator._run_fast_sim_core_loop (betse/science/sim.py:1454-1640), selected by
 * `solver options: type: fast` (sim.py:1068-1070), without networks.  The constants come from Simulator.fast_sim_init
 * (sim.py:1393-1452), which the reference runs on the host before the loop.  A ctx steps EITHER solver. */
typedef struct betse_fast_host {
    /* betse_fast_setup reads, betse_fast_download fills (NULL members are skipped) */
    double *vm_ave;          /* [C] sim.vm_ave                                                         */
    double *gjopen;          /* [M] sim.gjopen                                                         */
    /* betse_fast_setup only */
    const double *G_Leak;    /* [C] sim.G_Leak       sim.py:1441                                       */
    const double *E_Leak;    /* [C] sim.E_Leak = sim.vm_GHK, sim.py:1434                               */
    const double *G_gj;      /* [C] sim.G_gj         sim.py:1445                                       */
    const double *sigma_cell;/* [C] sim.sigma_cell (constant in this solver)                           */
    const double *extra_J_mem; /* [M] sim.extra_J_mem, NULL = 0                                        */
    /* betse_fast_download only */
    double *vgj;             /* [M] sim.vgj of the last step, sim.py:1547                              */
    double *Jn, *Emx, *Emy;  /* [M] sim.py:1566-1573; valid after a step with BETSE_STEP_DIAG          */
    double *J_cell_x, *J_cell_y, *E_cell_x, *E_cell_y;   /* [C] sim.py:1579-1584; same                 */
} betse_fast_host;
int  betse_fast_setup(betse_ctx *ctx, const betse_fast_host *state);
/* Replaces `nsteps` iterations of the fast loop body (sim.py:1547-1592). */
int  betse_fast_step(betse_ctx *ctx, int nsteps, int flags, uint32_t *status_out);
int  betse_fast_download(betse_ctx *ctx, betse_fast_host *out);
/* Voltage-gated channels under the fast solver: MasterOfNetworks.run_fast_loop_channels (networks.py:3217-3280).  The
 * channels are those of betse_set_channels (unmodulated ones); every step their gates advance at the cell's potential and
 * J_ED = G (vm - E_rev) joins extra_J_mem, G = DChan * cond_coef[ion]:
 *   cond_coef[ion] = q z^2 F cbar_dic[ion] / (tm kb p.T) * geo_conv     stb.get_conductivity, sim_toolbox.py:1363-1372
 *   rev_E[ion]     = sim.rev_E_dic[ion]                                 fast_sim_init, sim.py:1393-1452
 * Both [n_ions], in ion-index order.  Call after betse_set_channels and before betse_fast_setup. */
int  betse_fast_set_channels(betse_ctx *ctx, const double *cond_coef, const double *rev_E);

/* Same as betse_step but timed with CUDA events on the ctx's stream; per-kernel mean
 * durations (ms per launch) and launch counts are returned for the roofline in bench.py. */
int  betse_step_profile(betse_ctx *ctx, int nsteps, float *total_ms,
                        float kernel_ms[BETSE_NKERNELS], int kernel_launches[BETSE_NKERNELS]);
const char *betse_kernel_name(int k);

/* Replaces: the attribute reads of Simulator.write2storage (sim.py:1789-1884). */
int  betse_download_sample(betse_ctx *ctx, betse_state_host *state);

/* ---------------------------------------------------------------------------------------------
 * Voltage-gated channels of the general network:
 * MasterOfNetworks.run_loop_channels (betse/science/chemistry/networks.py:3115-3213) driving the
 * Hodgkin-Huxley classes of betse/science/channels/vg_{na,k,ca,cl}.py through
 * ChannelsABC.update_mh (channels/channelsabc.py:40-60).  A model is DATA: the four quantities
 * mInf, mTau, hInf, hTau as terms of U = 1000*vm + v_shift [mV] (betse_b200/channels.py documents
 * the term types and is held to the reference classes by tests/test_channels_table.py). */
typedef struct betse_gate_term { int32_t type; int32_t pad; double p[4]; } betse_gate_term;

typedef struct betse_channel {
    int32_t ion;                  /* index of the conducted ion (channel_core.ions[0])            */
    int32_t mpower, hpower;       /* P = m^mpower * h^hpower (vg_na.py:104)                        */
    int32_t kind[4];              /* mInf, mTau, hInf, hTau: 0 = a, 1 = a/(a+b), 2 = 1/(a+b)       */
    int32_t handler;              /* network handler the channel belongs to: 0 general network, 1 gene network */
    int32_t mod_prog;             /* program index of chan.alpha_eval_string (networks.py:3147) in that handler's
                                     betse_network, < 0: no modulation (moddy == 1)                */
    int32_t same_gates;           /* 1: a further conducted ion (channel_core.ions[j > 0], rel_perm[j]) of the
                                     PREVIOUS entry's channel — only `ion` and `rel_perm` are read (vg_funny, cation) */
    betse_gate_term a[4], b[4];
    double time_unit;             /* channel_core.time_unit (1e3 for models in ms)                 */
    double max_Dm;                /* Channel.maxDm (networks.py:6559)                              */
    double rel_perm;              /* channel_core.rel_perm[0]                                      */
    double v_shift;               /* mV added to V inside the model (e.g. vg_ca.py:315: V - 10)    */
    const uint8_t *target_mask;   /* [M] 1 on channel_core.targets; NULL = every membrane          */
    const double *m0, *h0;        /* [M] gate states at loop entry (read on targets only)          */
} betse_channel;

/* Replaces: the channel objects the loop closes over (Channel.init_channel, networks.py:6550-6629).
 * Channels are applied in list order, each with an immediate concentration update, between the
 * flux computation of the ion loop and update_all_concs (sim.py:1290-1357).  n = 0 removes them.
 * affect_charge = p.substances_affect_charge (Jmem takes the channels' currents, networks.py:2971). */
int  betse_set_channels(betse_ctx *ctx, int n, const betse_channel *channels, int affect_charge);
/* Gate states / open probability / last flux / DChan (networks.py:3164) of channel k, [M] each (NULL members are skipped). */
int  betse_channel_state(betse_ctx *ctx, int k, double *m, double *h, double *P, double *flux, double *DChan);

/* Dynamic noise (sim.py:1322-1339): protein_noise_flux [M] = p.dynamic_noise_level*(np.random.random(mdl) - 0.5), drawn by
 * the host from the reference's own random stream, applied to ion `ion` (P) by the NEXT timestep only — update_Co after
 * the networks and before update_all_concs. */
int  betse_set_noise_flux(betse_ctx *ctx, int ion, const double *flux);

/* Page-locked host staging for sampled-step downloads (the buffers write2storage copies from, sim.py:1789-1884):
 * device->host copies into these run at PCIe speed without the driver's bounce buffer. */
int  betse_host_alloc(size_t bytes, void **out);
/* the same from a helper thread (binds `device` first): staging can be pinned while the engine is being built */
int  betse_host_alloc_on(int device, size_t bytes, void **out);
/* dst[0..bytes) = src[0..bytes) split over a few threads: a fresh NumPy destination is first-touch page faults */
void betse_host_copy(void *dst, const void *src, size_t bytes);
/* dst[i][m] = src[i][idx[m]] (rows of n_src -> rows of n_dst doubles) over a few threads: sim.cc_at_mem [I][M] from the
 * per-cell array the device keeps (cc_at_mem is a gather of a per-cell quantity, sim_toolbox.py:1181-1182) */
void betse_host_expand(double *dst, const double *src, const int32_t *idx, int n_rows, size_t n_src, size_t n_dst);
void betse_host_free(void *p);

/* ---------------------------------------------------------------------------------------------
 * General network / gene regulatory network (MasterOfNetworks, networks.py).
 * The reference writes every rate law as a Python expression string and evals it per timestep
 * (write_growth_and_decay networks.py:1187-1310, write_reactions 1312-1572, channel
 * alpha_eval_string 1102-1110); the host compiles those strings into postfix programs
 * (betse_b200/ratelaw.py) that the device interprets per cell / per membrane.
 *
 * Implemented: substances with growth/decay and cell-zone reactions (run_loop, networks.py:2805-2914),
 * gap-junction transport of substances (molecule_mover, sim_toolbox.py:976-1006), channel
 * modulation by substances.  Refused by the host shim: membrane-permeable substances (Dm != 0),
 * substances present in the environment, pumps, ligand gating, transporters, modulators,
 * mitochondria, boundary / clamp events. */
#define BETSE_STATUS_NEG_NET 16u  /* a network substance went negative (sim_toolbox.py:1124-1150 raises) */

/* One sim modulator (Modulator, networks.py:6655-6700; run_loop_modulators, networks.py:3282-3325):
 * target = max_val * program(membrane), written over sim.gj_block (target 0) or sim.NaKATP_block (target 1);
 * target 2 = tight junctions (networks.py:3301-3317): max_val * program(env square) — an EXTRACELLULAR-zone program —
 * written over sim.TJ_modulator on the squares betse_network.tj_targets, for ion `ion` or for all ions (ion < 0); the
 * transport of the NEXT step then sees D_env_raw * TJ_modulator there (sim.py:2231-2233). */
typedef struct betse_modulator { int32_t target; int32_t prog; double max_val; int32_t ion; int32_t pad; } betse_modulator;

/* One ligand-gated channel (Molecule.gating, networks.py:5847-5916): substance `species` opens a channel for ion `ion`,
 * Dchan = rho_channel * Dm_mod * mod, Dm_mod = rho_channel*max*hill(c at the membrane) (intracellular ligand) or
 * max*hill(c_env at the membrane's env square) (extracellular); its GHK flux is added to fluxes_mem AND applied
 * immediately, like the reference does. */
typedef struct betse_ligand_gate {
    int32_t species, ion, extracell, pad;
    double K, n, max_val, mod;            /* hill(x) = x^n / (K^n + x^n); mod = the folded gating_mod_eval_string */
} betse_ligand_gate;

/* Molecule.pump (networks.py:5809-5844): the substance's own active pump (stb.molecule_pump, sim_toolbox.py:658-775,
 * with the fixed cATP/cADP/cPi of the parameters) or facilitated transporter (stb.molecule_transporter, :777-907),
 * applied right after the substance's growth/decay and before its membrane / gap-junction / extracellular transport. */
typedef struct betse_substance_pump {
    int32_t species, into_cell, uses_ATP, pad;
    double max_val, Km;
} betse_substance_pump;

/* One transporter (run_loop_transporters, networks.py:2985-3107; strings from write_transporters, :2090-2614), cell zone,
 * with extracellular spaces: flux[m] = rho_pump * program(membrane) on every membrane; every reactant (sign -1) and
 * product (sign +1) then moves by coeff * sign * sum_mems(flux*mem_sa)/cell_vol in the cells of cell_mask, or by
 * coeff * div_env(sign*flux) on the env squares of env_mask; extra_J_mem += net_z*flux*F. */
#define BETSE_TR_MAX_TERMS 12
typedef struct betse_transporter_term {
    int32_t kind;                 /* 0 ion in the cells, 1 ion outside, 2 substance in the cells, 3 substance outside */
    int32_t index;                /* ion / substance index                                           */
    int32_t sign;                 /* -1 reactant, +1 product                                         */
    int32_t pad;
    double coeff;
} betse_transporter_term;
typedef struct betse_transporter {
    int32_t prog;                 /* membrane-zone program of transporter_eval_string                */
    int32_t n_terms;
    double net_z;
    const uint8_t *cell_mask;     /* [C] transporter_targets_cell; NULL = every cell                 */
    const uint8_t *env_mask;      /* [E] transporter_targets_env;  NULL = every env square           */
    const uint8_t *mem_mask;      /* [M] transporter_targets_mem;  NULL = every membrane: where the reference also nudges
                                     mem_concs[X] -/+= flux*(mem_sa/mem_vol)*dt (networks.py:3020-3022, 3070-3072), a value
                                     that lives until the next update_intra / update_Co and that LATER rate laws of the same
                                     step (transporters, channel modulation, sim modulators) read                  */
    betse_transporter_term terms[BETSE_TR_MAX_TERMS];
} betse_transporter;

typedef struct betse_network {
    int32_t n_species;            /* K substances, MasterOfNetworks.molecules order                 */
    int32_t n_rates;              /* K growth/decay rates + R cell-zone reactions = columns of reaction_matrix */
    int32_t n_programs;           /* n_rates rate programs followed by channel-modulator programs   */
    int32_t n_consts, n_cell_arrays, n_mem_arrays;
    const double  *c_cells;       /* [K][C] concentrations at loop entry                            */
    const int32_t *code;          /* (opcode, argument) pairs of all programs                       */
    const int32_t *prog_ptr;      /* [n_programs + 1] program ranges, in pairs                      */
    const double  *consts;        /* [n_consts]                                                     */
    const double  *cell_arrays;   /* [n_cell_arrays][C] per-cell constants (growth_mod_function_cells ...) */
    const double  *mem_arrays;    /* [n_mem_arrays][M]                                              */
    const uint8_t *growth_mask;   /* [K][C] 1 on growth_targets_cell (networks.py:2844-2846); NULL = all */
    const double  *stoich;        /* [K][n_rates] substance rows of reaction_matrix (networks.py:2686-2725) */
    const double  *Dgj;           /* [K] gap-junction diffusion constant; < 0: ignoreGJ             */
    const double  *z;             /* [K] charge                                                     */
    const double  *time_factor;   /* [K] modify_time_factor                                         */
    /* Membrane and extracellular legs of Molecule.transport -> stb.molecule_mover (networks.py:5670-5700,
     * sim_toolbox.py:909-1153).  env_on == NULL: every substance lives in the cells only. */
    const uint8_t *env_on;        /* [K] 1: the substance crosses the membrane and/or exists in the environment */
    const double  *Dm;            /* [K] membrane diffusion constant (0: no trans-membrane flux)    */
    const double  *c_bound;       /* [K] concentration at the global boundary (Molecule.c_bound)     */
    const double  *c_env;         /* [K][E] env concentrations at loop entry (rows with env_on == 0 are ignored) */
    const double  *D_env;         /* [K][E] Do * D_env_weight (* TJ_factor on sim.TJ_targets) or Do where the
                                     substance passes tight junctions (sim_toolbox.py:1075-1085)     */
    /* p.substances_affect_charge (networks.py:2942-2977): charged substances add F*c*z*scale to the charge of cells
     * and env squares, -z*f_mem*F*scale + z*f_gj*F*scale to the membrane current, z*F*scale*f_env to J_env. */
    const double  *scale_factor;  /* [K] Molecule.scale_factor (NULL: 1)                             */
    int32_t affect_charge;
    int32_t n_modulators;
    const betse_modulator *modulators;   /* programs are membrane-zone programs (index >= n_rates)    */
    const betse_ligand_gate *ligand_gates;
    int32_t n_ligand_gates;
    int32_t n_pumps;
    const betse_substance_pump *pumps;   /* at most one per substance; the substance needs env_on      */
    const betse_transporter *transporters;   /* applied in order, before the handler's channels (sim.py:1293-1297) */
    int32_t n_transporters;
    int32_t reserved;
    const double *mem_sa_over_vol;       /* [M] cells.mem_sa / cells.mem_vol (needed with transporters)  */
    /* Molecule.update_intra with 'update intracellular' (networks.py:5714-5806) for neutral substances whose membrane
     * value nothing reads back: cc_at_mem relaxes towards the cell value, (g*Do*dt*c/R + cc)/(1 + g*Do*dt/R),
     * g = mem_sa/(3/4 mem_vol).  intra_on == NULL: every membrane value is the cell value. */
    const uint8_t *intra_on;             /* [K]                                                          */
    const double  *Do;                   /* [K] Molecule.Do                                              */
    const double  *c_mems;               /* [K][M] cc_at_mem at loop entry (rows with intra_on == 0 ignored) */
    const double  *R_rads;               /* [M] cells.R_rads                                             */
    const int32_t *map_cell2ecm;         /* [C] cells.map_cell2ecm: cell-zone rate laws reading env concentrations (NULL: none do) */
    /* intra_on substances that are charged / membrane-permeable / gap-junction permeable: the membrane value is state that
     * feeds back — update_intra adds electrophoresis in the cell's field ((Do*q*z/(kb*T) + Mu_mem)*Emc, Emc as the previous
     * step's update_V left it; at loop entry from betse_state_host.Emc), and molecule_mover's membrane and gap-junction legs
     * read and move the membrane values (sim_toolbox.py:962-1005, 1183-1185). */
    const double  *mu_mem;               /* [K] Molecule.Mu_mem (NULL: zeros)                            */
    const double  *Emc;                  /* [M] sim.Emc at loop entry (NULL: zeros)                      */
    /* tight-junction modulators (betse_modulator.target == 2) */
    const int32_t *tj_targets;           /* [n_tj] sim.TJ_targets: env squares of the barrier (sim.py:2394-2395) */
    int32_t n_tj;
    int32_t reserved2;
    const double  *D_env_raw;            /* [I][E] sim.D_env WITHOUT sim.TJ_modulator (betse_state_host.D_env_eff is their product) */
    const double  *TJ_modulator;         /* [I][E] sim.TJ_modulator at loop entry (NULL: ones)            */
    /* reactions OUTSIDE the cells (write_reactions_env, networks.py:1830-2088; the top of run_loop, networks.py:2872-2889):
     * extracellular-zone programs, c_env += stoich_env . rates * dt before anything else of the handler's run_loop */
    const int32_t *env_rx_prog;          /* [n_env_rx] program indices                                    */
    int32_t n_env_rx;
    int32_t reserved3;
    const double  *stoich_env;           /* [K][n_env_rx] substance rows of reaction_matrix_env            */
} betse_network;

/* handler 0 = sim.molecules.core, 1 = sim.grn.core (run in that order, sim.py:1290-1318).  net == NULL
 * removes the handler.  Call before betse_set_channels when channels carry mod_prog >= 0. */
int  betse_set_network(betse_ctx *ctx, int handler, const betse_network *net);
/* Substance concentrations [K][C] and the last rates [n_rates][C] (NULL members are skipped). */
int  betse_network_state(betse_ctx *ctx, int handler, double *c_cells, double *rates);
/* Env concentrations [K][E] of the substances (rows with env_on == 0 come back as zeros). */
int  betse_network_env_state(betse_ctx *ctx, int handler, double *c_env);
/* sim.TJ_modulator [I][E] as the tight-junction modulators left it (networks.py:3301-3317); an error without such modulators. */
int  betse_network_tj_modulator(betse_ctx *ctx, double *tj_modulator);
/* Membrane values [K][M] (Molecule.cc_at_mem): the cell value gathered, or the transported value with intra_on. */
int  betse_network_mem_state(betse_ctx *ctx, int handler, double *c_mems);
/* The substances' own timed events, evaluated by the host for the step about to run (scalar schedule logic like
 * fire_events): c_bound [K] (Molecule.update_boundary, networks.py:6043-6066; NULL: unchanged) and clamp [K] — the value
 * every cell takes right after growth/decay, NaN where no clamp is in force (Molecule.cell_clamp_method,
 * networks.py:6069-6088; NULL: none). */
int  betse_network_set_events(betse_ctx *ctx, int handler, const double *c_bound, const double *clamp);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: the tissue is cut into strips of env-grid rows; each rank owns the
 * cells whose centre lies in its rows.  The reference has no counterpart (one process, one
 * thread); what crosses a strip edge every timestep is exactly what the reference's index arrays
 * couple: gap-junction partners (cells.nn_i, cells.py:1498-1533), membrane->env-square exchange
 * (cells.map_mem2ecm, cells.py:1758) and the env-grid stencils (update_ecm, sim.py:2209-2254;
 * gaussian_filter + gradient, ion_current.py:101-109).
 *
 * Every buffer a neighbour writes lives in ONE device allocation per ctx, the "window"; ranks
 * exchange its CUDA-IPC handle (or the raw pointer when they share a process) and then PUSH their
 * boundary values straight into the neighbour's window with ordinary stores over NVLink, followed
 * by a release-store of an epoch flag; the consumer spins on its own flag (bounded; a timeout sets
 * BETSE_STATUS_XCHG_TIMEOUT).  No host round trip and no NCCL call on the data path.
 * Exchange points of one step:  X1 after the membrane/cell kernel (ghost-cell concentrations and
 * Vmem + membrane fluxes into squares the neighbour owns), X2 after the env accumulation (env
 * concentration rows + raw env voltage rows). */
#define BETSE_STATUS_XCHG_TIMEOUT 8u   /* a neighbour's flag did not arrive (multi-GPU only) */

#define BETSE_XCHG_X1 0
#define BETSE_XCHG_X2 1
#define BETSE_XCHG_PUSH 1   /* mode bits of betse_exchange */
#define BETSE_XCHG_WAIT 2

typedef struct betse_window_info {
    void    *base;             /* device pointer of the window in THIS process                  */
    uint64_t bytes;
    uint8_t  ipc_handle[64];   /* cudaIpcMemHandle_t of the window                               */
    uint64_t off_cc_mid[2];    /* [I,C] x2  byte offsets inside the window                       */
    uint64_t off_vm_cell[2];   /* [C]   x2                                                       */
    uint64_t off_flux;         /* [n_flux_slots, I]                                              */
    uint64_t off_cc_env[2];    /* [I,E] x2                                                       */
    uint64_t off_v_raw;        /* [E]                                                            */
    uint64_t off_flags;        /* uint64 [2 exchange points][2 sides]                            */
    int32_t  n_cells, n_env, nx, n_ions;   /* leading dimensions of the arrays above             */
    int32_t  n_flux_slots, reserved;       /* rows of the flux array: betse_attach_neighbor checks recv_slot0 + n_send_flux against it */
} betse_window_info;

/* One neighbouring strip.  side 0 = the strip below (smaller y), 1 = above. */
typedef struct betse_neighbor {
    int32_t side;
    int32_t same_process;          /* 1: info.base is directly usable; 0: open info.ipc_handle   */
    betse_window_info info;        /* the NEIGHBOUR's window                                     */
    /* X1: my owned cells that are ghosts there, and my membranes whose env square it owns */
    int32_t n_send_cells;  const int32_t *send_cells;  int32_t recv_cell0;   /* first ghost index there */
    int32_t n_send_flux;   const int32_t *send_flux;   int32_t recv_slot0;   /* first remote slot there */
    /* X2: contiguous row blocks (local row numbers on each side) */
    int32_t cc_rows, cc_src_row0, cc_dst_row0;     /* cc_env rows                                */
    int32_t v_rows,  v_src_row0,  v_dst_row0;      /* raw env voltage rows                       */
} betse_neighbor;

/* Kernel row ranges of a strip (local rows): ion transport [yi0,yi1), env accumulation
 * [ya0,ya1), field E rows [yf0,yf1).  Single GPU: all [0,ny). */
int  betse_set_row_ranges(betse_ctx *ctx, int yi0, int yi1, int ya0, int ya1, int yf0, int yf1);
int  betse_window(betse_ctx *ctx, betse_window_info *out);
int  betse_attach_neighbor(betse_ctx *ctx, const betse_neighbor *nb);
/* Enqueue one exchange point on the ctx's stream.  buf_next = 1: the double-buffered arrays are
 * the ones the running step is producing (cur^1); 0: the current ones (initial exchange). */
int  betse_exchange(betse_ctx *ctx, int which, int buf_next, int mode);
/* A step split at the exchange points: phase 0 = env transport + membrane/cell update,
 * phase 1 = env accumulation, phase 2 = env field (+ buffer flip).  betse_step on a ctx with
 * neighbours == phase 0, X1, phase 1, X2, phase 2 (captured in one CUDA graph). */
int  betse_step_phase(betse_ctx *ctx, int phase, int flags);
/* update_V split the same way: phase 0 = cell charge + env charge, phase 1 = env field. */
int  betse_update_v_phase(betse_ctx *ctx, int phase);
int  betse_stream(betse_ctx *ctx, void **cuda_stream);
int  betse_sync(betse_ctx *ctx, uint32_t *status_out);

#ifdef __cplusplus
}
#endif
#endif /* BETSE_B200_H */
