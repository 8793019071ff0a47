// Device-side parameter block and array table shared by all kernels of the tissue step.
#pragma once
#include <stdint.h>

#define BT_MAX_IONS 8
#define BT_TPB 256          // threads per CTA of the membrane kernel == max membranes per CTA
#define BT_MAX_CTA_CELLS 96 // max cells packed into one membrane-kernel CTA
#define BT_TILE_MAXC 8      // max cells of one warp tile (<= 32 membranes)

// Scalars (passed BY VALUE as a __grid_constant__ kernel argument: lives in the constant bank).  Derived products are formed on
// the host in the same operand order as the reference's NumPy expressions.
struct KParams {
    int n_ions, iNa, iK, iCa;
    int zi[BT_MAX_IONS];          // integer valence class (+-1, +-2) or 0 = generic path
    int ia[BT_MAX_IONS], ib[BT_MAX_IONS];  // slots of (A, B) in the GHK table {A1,B1,A2,B2} for this valence
    double z[BT_MAX_IONS];
    double zF[BT_MAX_IONS];       // sim.zs * p.F
    double Dgj_surf[BT_MAX_IONS]; // sim.D_gj[i] * p.gj_surface
    double Dgj_len[BT_MAX_IONS];  // Dgj_surf[i] / cells.gj_len
    double cbound[BT_MAX_IONS];   // sim.c_env_bound
    double sig_k[BT_MAX_IONS];    // z^2 * F^2 (sigma_cell, diagnostics)
    double D_free[BT_MAX_IONS];
    double F, RT_sim, RT_p, kbT_sim, q, cm, inv_cm, tm, dt;
    double alpha_NaK, alpha_Ca, KmNK_Na, KmNK_K, KmNK_ATP, KmCa_Ca, KmCa_ATP;
    double cATP, cADP, cPi, K0;             // K0 = exp(-deltaGATP/(R*T_sim))
    double gj_vthresh, gj_min, gj_len;
    double rho_pump, rho_channel;
    double NaK_block, gj_block;             // scalar blocks (used when the arrays are null)
    double delta, env_vol_div;              // env_vol_div = cell_height*delta^2
    double ecm_vol, memsa_mean, ko_eo_er;   // ko_env*eo*er
    double true_cell_size, cell_radius;     // p.true_cell_size, p.cell_radius (sim.py:2068)
    double screen;                          // (2/(ko_env*delta))*(cell_radius/true_cell_size)
    double vol_env;                         // no-ECM bath volume
    double sharpness;
    double gw[5];                           // gaussian sigma=1 taps w0..w4 (normalised)
    double smooth_cells, R_T_p;             // R*T_p
    int is_ecm, v_sensitive_gj, cluster_open, fast_update_ecm;
    int polar;                              // cell_polarizability != 0: Vmem is per-membrane state (sim.py:2048-2080)
    int has_phi;                            // Phi_b != 0 somewhere
    // local grid geometry (domain decomposition: rows [y0, y0+ny) of a ny_global-row grid)
    int ny, nx, y0, ny_global, y_own0, y_own1;
    int n_cells, n_cells_owned, n_mems_owned, n_ctas, n_tiles;
    int pf_tiles;                           // k_mem: L2 prefetch distance in tiles (0 = off)
    int n_blocks, ell_rows, kb_max, kb_min; // k_cell: blocks of 32 cells, rows of the cell pack, rows of its widest / narrowest block
    int n_patches, n_out_sq, ptab_max;      // k_cell_patch: patches of 4 blocks (0 = off); env squares no patch owns; ints of the longest patch table
    int pf_dist;                            // k_cell, register build: a block pulls the streams of block + pf_dist into L2 (0 = off)
    int kc_persist;                         // k_cell, register build: persistent warps drawing tickets (1) or one block per warp (0)
    int defer;                              // the membrane kernel stores its membrane->cell sums instead of applying them (channels, networks)
    int defer_slots;                        // ... and a later kernel adds to its membrane->env fluxes in flux_slots: k_mem only
    int chan_charge;                        // p.substances_affect_charge: Jmem takes the channels' extra_J_mem
    // kernel row ranges in local rows (single GPU: all [0, ny)): ion transport, membrane->env
    // accumulation, env field (E rows; v_env is written on the accumulation rows)
    int yi0, yi1, ya0, ya1, yf0, yf1;
    // halo exchange of a decomposed tissue (xchg.cuh): xwait = the consumers wait inside the kernel; xsides bit s = a neighbour on
    // side s (0 below, 1 above); env squares in local rows < xw_lo or >= xw_hi take remote flux slots
    int xwait, xsides, xw_lo, xw_hi;
    int xpush;                              // the producers push their boundary values to the neighbours themselves (k_cell, k_envacc_ell)
    unsigned long long x_timeout_ns;
    // reciprocals of constants (formed once on the host; x*inv differs from x/c by <= 1 ulp)
    double inv_RT_sim, inv_RT_p, inv_tm, inv_gjl, inv_kbT_sim, inv_delta, inv_2delta;
    double inv_KmNK_Na, inv_KmNK_K, inv_KmCa_Ca, tNK, tCa;   // tNK = cATP/KmNK_ATP, tCa = cATP/KmCa_ATP
    double QnNK0, QdNK0, QnCa0;    // (cADP*1e-3)*(cPi*1e-3), cATP*1e-3, cADP*cPi
    double dtm;                    // dt*1e3 (gap_junction.py:68)
    double inv_K0;
};

// Device array table (kernel argument).
struct KArrays {
    // mesh
    const int *mem_to_cells, *cell_mem_ptr, *nn_cell_flag, *nn_i, *map_mem2ecm;
    const int *cta_cell_start;   // CTA packing (k_diag): whole cells, <= BT_TPB membranes
    const int *tile_desc;        // warp packing (k_mem): int4 {c0, nc, m0, nm} per tile of whole cells, <= 32 membranes
    const char *tile_pack;       // k_mem_pipe: fixed-size per-tile constant blocks (layout: kmem_pipe.cu header)
    const int *slot_ptr, *slot_idx;
    // cell pack of k_cell (SELL-32: block b = cells 32b..32b+31; row blk_row0[b] + k holds membrane k of each cell)
    const int *blk_row0;         // [n_blocks + 1] int2 {first row, first membrane} of every block
    int *ticket;                 // k_cell: next ticket (zeroed before every launch)
    const int *pcell;            // k_cell_patch: cell of every (block, lane), -1 = none; null = block b is cells 32b..32b+31
    const int *ptab_ptr, *ptab;  //   per patch: the env squares it owns and where their membranes' fluxes sit in shared memory (kcell.cu)
    const int *slot_ptrp;        //   slot_ptr with bit 31 set on the env squares a patch owns: their fluxes are summed in sq_sum
    double *sq_sum;              //   [I][E] summed membrane -> env fluxes of the owned squares
    const int *out_sq;           //   the other squares (shared between patches, or without membranes), as a list
    const int *bslot;            //   [rows][32] compact border slot (8 doubles each in flux_slots) of the membranes of shared squares
    const int *blk_x;            // k_cell, strips: int2 per block {row of ghost_tab or -1, offset into rslot_tab or -1}
    const int *ghost_tab;        //   int2 per (flagged block, lane): ghost index of the cell on the neighbour of side 0 / 1, or -1
    const int *rslot_tab;        //   int per (flagged block, row, lane): side << 30 | remote flux slot on that neighbour, or -1
    const char *cpack;           // [rows] x {DmS[I][32] doubles = (Dm*(-rho_channel/tm))*mem_sa, mem_sa[32] doubles,
                                 //            partner cell | boundary bit [32] ints, env square | KC_FIRST/LAST/INV bits [32] ints}
    double *flux_ell;            // [rows][I][32] membrane -> env exchange, written by k_cell
    const int *slot_off;         // [slots] position of every env-square slot in flux_ell (>= 0) or in flux_slots (-(s*I)-1)
    const double *mem_sa, *mem_nx, *mem_ny, *cell_vol, *cell_sa, *diviterm, *num_mems;
    const double *memsa_env, *gj_w;
    // state
    double *cc_cells;        // [I,C]
    double *cc_mid[2];       // [I,C] x2
    double *cc_env[2];       // [I,E] x2
    double *vm_cell[2];      // [C]   x2
    double *vm_pol[2];       // [M]   x2: Vmem itself when cell_polarizability != 0 (then vm_cell is vm_o, sim.py:2057)
    const double *R_rads;    // [M]   cells.R_rads (polarizability branch only)
    double *gjopen;          // [M]
    const double *Dm;        // [I,M]
    const double *Denv;      // [I,E]
    double *E_x, *E_y, *v_env, *v_raw, *rho_env, *rho_cells;
    const double *phi_b_old;     // Phi_b the Vmem of the previous step was formed with (differs from phi_b while bound_V ramps)
    const double *phi_b, *extra_rho_cells, *extra_rho_env, *extra_J_mem;
    const double *NaK_block, *gj_block;
    double *flux_slots;      // [slots, I]
    double *cenv_u;          // [2][8] no-ECM bath concentrations (device-resident, double buffered)
    double *cenv_part;       // [n_tiles, 8] per-tile partial sums
    unsigned int *status;
    const unsigned long long *xflags, *xepoch;   // this rank's exchange flags [2 points][2 sides] and epochs [2 points] (xchg.cuh)
    // diagnostics
    double *fl_mem, *fl_gj, *fl_env_x, *fl_env_y, *rate_NaK;
    double *Jmem, *Jgj, *Jn, *I_mem, *Jc, *Emc, *dvm, *vm_mem, *vm_ave;
    double *J_cell_x, *J_cell_y, *E_cell_x, *E_cell_y, *sigma_cell;
    double *E_gj_x, *E_gj_y; // [M] gap-junction field (sim.py:2168-2172)
    double *scratch_env;     // [I,E] temp for the sharpness<1 smoothing pass
    // voltage-gated channels (channels.cu)
    double *dsum_m, *dsum_g; // [I,C] deferred sums of f_mem*sa and f_gj*sa per cell
    double *chan_slots;      // [M]   f*sa of the channel being applied (membrane -> env exchange)
    double *chan_part;       // [n_tiles] per-tile sums of it (no-ECM: the channel's share of the well-mixed bath)
    double *chanJ;           // [M]   extra_J_mem accumulated over the channels of this step
    double *extra_Jenv_x, *extra_Jenv_y;   // [E] charged network substances moving through the env grid (networks.py:2953-2954)
};
