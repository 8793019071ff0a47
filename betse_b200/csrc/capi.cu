// C-ABI of the B200 BETSE engine (include/betse_b200.h): context, device SoA, CTA packing,
// env CSR, state upload/download, step scheduling (direct launches or CUDA graphs), profiling.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <atomic>
#include <thread>
#include <vector>

#include "../../include/betse_b200.h"
#include "kparams.cuh"
#include "xchg.cuh"
#include "fast.cuh"
#include <climits>
#include "channels.cuh"
#include "network.cuh"
#include "hh.cuh"

// launchers (kernels.cu)
int launch_mem(int ni, const KParams& P, const KArrays& A, int n_ctas, int cur, int diag, cudaStream_t st);   // 1: the membrane -> env fluxes went to flux_ell (k_cell)
int mem_kernel_kind(int ni, const KParams& P, const KArrays& A, int diag);
void launch_ion(const KParams& P, const KArrays& A, int nx, int cur, int diag, int ion0, int n, cudaStream_t st);
void launch_ion_smooth(int ni, const KParams& P, const KArrays& A, int ny, int nx, int nxt, cudaStream_t st);
void launch_envacc(int ni, const KParams& P, const KArrays& A, int E, int nxt, int apply, cudaStream_t st);
void launch_cell_charge(const KParams& P, const KArrays& A, int C, int cur, cudaStream_t st);
void launch_field(const KParams& P, const KArrays& A, int ny, int nx, cudaStream_t st);
void launch_envmix(int ni, const KParams& P, const KArrays& A, int cur, cudaStream_t st);
void launch_diag(int ni, const KParams& P, const KArrays& A, int n_ctas, int newb, cudaStream_t st);
void launch_expand_vm(const KParams& P, const KArrays& A, int M, int C, int cur, cudaStream_t st);
cudaError_t prepare_kernels(int ni);
// kmem_pipe.cu: the per-tile constant blocks of the pipelined membrane kernel
unsigned tile_pack_size(int ni);
void launch_hh_setup(const HHBuf& H, int ny, int nx, cudaStream_t st);
void launch_hh(const KParams& P, const KArrays& A, const HHBuf& H, cudaStream_t st);
void launch_phi_b(const KParams& P, const HHBuf& H, const double bound[4], double* phi, cudaStream_t st);
void launch_noecm_field(const KParams& P, const KArrays& A, const HHBuf& H, double sigma, const double* D_env_weight, int old, cudaStream_t st);
void launch_pack_const(const KParams& P, const KArrays& A, cudaStream_t st);
void launch_pack_dm(const KParams& P, const KArrays& A, cudaStream_t st);
// kcell.cu: the lane-per-cell membrane kernel, its cell pack and the env accumulation that reads its fluxes
bool kcell_enabled();
void launch_pack_cell_const(const KParams& P, const KArrays& A, int* mem_ell, const int* mem_bidx, cudaStream_t st);
bool kcell_patch_fits(int ni, int kb_max);
void launch_pack_cell_dm(const KParams& P, const KArrays& A, cudaStream_t st);
size_t cell_pack_row_bytes(int ni);
cudaError_t prepare_cell(int ni, int kb_max, int ptab_max);
void launch_slot_off(const int* slot_idx, const int* mem_ell, int* slot_off, int n, int Mo, int ni, cudaStream_t st);
void launch_gather_int(int* dst, const int* src, const int* idx, int n, cudaStream_t st);
void launch_envacc_ell(int ni, const KParams& P, const KArrays& A, int nxt, cudaStream_t st);
void launch_envacc_ell_x(int ni, const KParams& P, const KArrays& A, const XPlan& X, int nxt, cudaStream_t st);
void envacc_end_ctas(const KParams& P, int lo, int hi, int* n_lower, int* n_upper);
void launch_cell_x(int ni, const KParams& P, const KArrays& A, const XPlan& X, int cur, cudaStream_t st);
void launch_chan(const KParams& P, const KArrays& A, const KChan& ch, const KNet& N, int cur, cudaStream_t st);
void launch_chan_cell(const KParams& P, const KArrays& A, const KChan* chs, int n, double* ell, int cur, int diag, int fuse_update, cudaStream_t st);
void launch_chan_expand(const KChan& ch, const int* mem_to_cells, const int* mem_ell, int Mo, int ni, cudaStream_t st);
void launch_fast_chan(const KParams& P, const KArrays& A, const KChan* chs, int n, const double* coef, const double* rev_E,
                      const double* vm_ave, double* extra_J, cudaStream_t st);
void launch_net(const KParams& P, const KArrays& A, const KNet& N, const double* h_Dgj, const double* h_Dm,
                const unsigned char* h_env_on, const betse_substance_pump* pumps, int n_pumps, double dG_RT,
                const unsigned char* h_intra, int n_ions, int cur, cudaStream_t st);
void launch_lig_prep(const KParams& P, const KArrays& A, const KNet& N, int sp, int extracell, double Kn, double n,
                     double max_val, double* Dm_mod, cudaStream_t st);
void launch_net_lig(const KParams& P, const KArrays& A, int ion, const double* Dm_mod, double mod, int cur, int diag, cudaStream_t st);
void launch_chan_env(const KParams& P, const KArrays& A, int ion, int cur, cudaStream_t st);
void launch_flux_apply(const KParams& P, const KArrays& A, int ion, const double* flux, cudaStream_t st);
void launch_transporter(const KParams& P, const KArrays& A, const KNet& N, const betse_transporter& T,
                        const unsigned char* d_cell_mask, const unsigned char* d_env_mask, const unsigned char* d_mem_mask,
                        int cur, cudaStream_t st);
void launch_tw_gather(const KParams& P, const KArrays& A, double* row, const double* src, cudaStream_t st);
void launch_net_mod(const KParams& P, const KArrays& A, const KNet& N, int prog, double max_val, double* dst, int cur, cudaStream_t st);
void launch_net_mod_tj(const KParams& P, const KArrays& A, const KNet& N, int prog, double max_val, int ion, const int* tj, int n_tj,
                       const double* denv_raw, double* tj_mod, int cur, cudaStream_t st);
void launch_cell_update(const KParams& P, const KArrays& A, int cur, cudaStream_t st);
void launch_fast(const KParams& P, const KArrays& A, const KFast& Fz, int cur, cudaStream_t st);
void launch_fast_diag(const KParams& P, const KArrays& A, const KFast& Fz, cudaStream_t st);
void launch_xchg(const KParams& P, const KArrays& A, const XPlan& X, int which, int buf, int mode, int flux_ell, cudaStream_t st);

enum { K_ION = 0, K_MEM, K_ENVACC, K_FIELD, K_ENVMIX, K_SMOOTH, K_DIAG, K_XCHG };
static const char* kKernelNames[BETSE_NKERNELS] = {
    "k_ion", "k_mem", "k_envacc", "k_field", "k_envmix", "k_ion_smooth", "k_diag", "k_xchg"};

struct betse_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;          // env transport of the non-Ca ions, next to k_mem
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool overlap = true;
    KParams P;
    KArrays A;
    betse_params hp;          // last host params
    int cur = 0;
    int C = 0, Co = 0, M = 0, Mo = 0, E = 0, ny = 0, nx = 0, I = 0, n_ctas = 0, n_tiles = 0, n_slots = 0;
    bool diag_valid = false;
    int* mem_ell = nullptr;                  // [Mo] position of every membrane's fluxes in flux_ell (cell pack of k_cell)
    int* kc_sync = nullptr;                  // k_cell: the ticket counter, zeroed before every launch
    int flux_is_ell = 0;                     // layout the last membrane kernel wrote its membrane -> env fluxes in
    // exchange window (every buffer a neighbouring rank writes) and the halo-exchange plan
    char* win = nullptr;
    betse_window_info winfo;
    XPlan X;
    // exchange points fused into k_cell / k_envacc_ell: host copies of the plans, tables built by prepare_xfuse
    std::vector<int> h_cmp;                  // cell_mem_ptr
    std::vector<int> xh_cells[2], xh_flux[2];
    bool xfuse_dirty = true;
    bool last_step_fused = false;
    int xfuse_now = 0;                       // this step's k_cell / k_envacc_ell push to the neighbours themselves
    int xwait_now = 0;                       // this step's env accumulation / field kernels wait for the neighbours themselves
    bool xwait = false;                      // the consumers of an exchange wait inside their kernels (neighbours in other processes)
    std::vector<void*> ipc_opened;
    std::vector<KChan> chans;                // voltage-gated channels, applied in order
    bool chan_cell_mode = false;             // the channels run on the per-cell path (channels.cu:k_chan_cell): gate state in mc/hc/Pc/Dc
    double* chan_ell = nullptr;              // [rows][I][32] f*sa of the channels of a pass, in the order of the cell pack
    HHBuf hh;                                // Helmholtz-Hodge diagnostics (sampled steps, undivided ECM tissues)
    bool hh_on = false;
    bool want_hh = true;
    const double* denv_w = nullptr;          // [E] sim.D_env_weight (no-ECM field diagnostics)
    bool poisson_on = false;                 // sine matrices + work buffers of the Dirichlet Poisson solve (HH, Phi_b)
    // boundary-voltage potential Phi_b (ion_current.py:84-90): [new, old] while sim.bound_V ramps, see update_phi_b
    double* phi[2] = {nullptr, nullptr};
    int phi_new = 0;
    bool phi_lag = false, phi_nz = false;
    double bound_prev[4] = {0, 0, 0, 0};
    KNet nets[2];                            // network handlers: 0 general network, 1 gene regulatory network
    bool net_on[2] = {false, false};
    int net_nprog[2] = {0, 0};
    std::vector<double> net_Dgj[2];          // host copy (which substances pass gap junctions)
    std::vector<double> net_Dm[2];           // host copy (which substances cross the membrane)
    std::vector<unsigned char> net_env_on[2];
    bool net_affect[2] = {false, false};
    std::vector<betse_modulator> net_mods[2];   // sim modulators of each handler (run_loop_modulators)
    int* net_tj[2] = {nullptr, nullptr};         // tight-junction modulators: env squares of the barrier (sim.TJ_targets)
    int net_ntj[2] = {0, 0};
    double* denv_raw = nullptr;                  // [I][E] sim.D_env without TJ_modulator
    double* tj_mod = nullptr;                    // [I][E] sim.TJ_modulator as the modulators leave it
    std::vector<betse_ligand_gate> net_gates[2];   // ligand-gated channels (Molecule.gating)
    std::vector<betse_substance_pump> net_pumps[2]; // the substances' own pumps / transporters (Molecule.pump)
    std::vector<betse_transporter> net_trans[2];    // transporters (run_loop_transporters); masks below are device copies
    std::vector<const unsigned char*> net_trans_cm[2], net_trans_em[2], net_trans_mm[2];
    int net_tw_rows[2] = {0, 0};
    // dynamic noise (sim.py:1322-1339): the host's draw for the NEXT step
    double* noise_flux = nullptr;
    int noise_ion = -1;
    bool noise_on = false, noise_pending = false;
    bool need_emc = false;          // a network substance reads sim.Emc every step (update_intra of charged substances)
    std::vector<unsigned char> net_intra[2];        // substances with intracellular transport (membrane values of their own)
    double* lig_tmp[2] = {nullptr, nullptr};       // [n_gates][M] openings formed before the substances advance
    std::string err;
    std::vector<void*> allocs;
    // fast (equivalent-circuit) solver (csrc/fast.cu)
    KFast fast;
    bool fast_on = false;
    int fast_cur = 0;
    bool fast_chan_on = false;               // channels under the fast solver (betse_fast_set_channels)
    double fast_coef[BT_MAX_IONS] = {0}, fast_revE[BT_MAX_IONS] = {0};
    double* fast_extraJ = nullptr;           // [M] extra_J_mem the channels form each step
    cudaGraphExec_t fast_graph = nullptr;    // two steps (buffer parity returns)
    int fast_graph_cur = 0;
    // CUDA graphs of one plain step, for cur = 0 and cur = 1
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
    bool graphs_built = false;
    bool use_graphs = true;
    unsigned long long graph_epoch = 0;      // bumped whenever captured launches of this ctx went stale (destroy_graphs)
    // ensemble graphs led by this ctx (betse_ensemble_step): members, steps per launch, buffer parity, epochs at capture
    struct EnsGraph { std::vector<betse_ctx*> members; int nsteps; int cur; unsigned long long epochs; cudaGraphExec_t exec; };
    std::vector<EnsGraph> ens;
    cudaEvent_t ens_ev[2] = {nullptr, nullptr};
    // profiling
    cudaEvent_t ev[16];
    bool ev_init = false;
    // page-locked bounce buffers for copies from / to pageable host memory (xfer)
    char* xbuf[2] = {nullptr, nullptr};
    cudaEvent_t xev[2] = {nullptr, nullptr};
};

#define CK(call)                                                                        \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            char b_[512];                                                               \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                     __FILE__, __LINE__);                                               \
            ctx->err = b_;                                                              \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

static int fail(betse_ctx* ctx, const std::string& msg)
{
    ctx->err = msg;
    return 2;
}

template <typename T>
static int dev_alloc(betse_ctx* ctx, T** p, size_t n, bool zero = true)
{
    if (n == 0) n = 1;
    CK(cudaMalloc((void**)p, n * sizeof(T) + 16));   // 16 bytes of slack: k_mem_pipe fetches 16-byte aligned supersets of rows
    ctx->allocs.push_back((void*)*p);
    if (zero) CK(cudaMemsetAsync(*p, 0, n * sizeof(T), ctx->stream));
    return 0;
}

// Host <-> device copy of a caller's array.  Pageable host memory makes cudaMemcpy stage through the driver's small
// bounce buffer synchronously (~3 GB/s measured on the B200 box); pipelining 8 MB chunks through two page-locked
// buffers overlaps the DMA of chunk i+1 with the host memcpy of chunk i.  Page-locked callers (betse_host_alloc) and
// small arrays go straight through.  A download has landed in `dst` on return; an upload's source may be reused.
static const size_t XCHUNK = (size_t)8 << 20;

static bool host_is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// the destination of a download is usually a fresh NumPy array: first-touch page faults dominate a single-threaded
// copy, so the chunk is split over a few threads
static void par_memcpy(char* dst, const char* src, size_t n)
{
    static const int T = [] {
        const unsigned hw = std::thread::hardware_concurrency();
        const char* e = getenv("BETSE_COPY_THREADS");
        const int n = e ? atoi(e) : 0;
        return n >= 1 ? (n > 8 ? 8 : n) : (int)(hw >= 16 ? 8 : hw >= 8 ? 4 : 2);
    }();
    if (n < ((size_t)2 << 20)) { memcpy(dst, src, n); return; }
    std::thread th[8];
    const size_t per = ((n / T) + 4095) & ~(size_t)4095;
    for (int t = 1; t < T; ++t) {
        const size_t off = (size_t)t * per;
        const size_t len = off >= n ? 0 : (n - off < per ? n - off : per);
        th[t - 1] = std::thread([=] { if (len) memcpy(dst + off, src + off, len); });
    }
    memcpy(dst, src, per < n ? per : n);
    for (int t = 1; t < T; ++t) th[t - 1].join();
}

static int xfer(betse_ctx* ctx, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind)
{
    cudaStream_t st = ctx->stream;
    const bool down = kind == cudaMemcpyDeviceToHost;
    // uploads: the driver's own pageable path (~6 GB/s on the B200 box) measured faster than a host-side bounce through
    // the page-locked pair (BETSE_H2D_BOUNCE=1), single- or multi-threaded, and than slices issued from several threads
    // on streams of their own (the driver serialises them: 1 thread 62-82 ms, 4 threads 115-124 ms for the index arrays
    // of a 1 M-cell tissue, profiles/r02g_e2e_fixed_cost.txt)
    static const bool h2d_bounce = [] { const char* e = getenv("BETSE_H2D_BOUNCE"); return e && atoi(e) != 0; }();
    if ((!down && !h2d_bounce) || bytes < ((size_t)1 << 20) || host_is_pinned(down ? dst : src)) {
        CK(cudaMemcpyAsync(dst, src, bytes, kind, st));
        return 0;
    }
    if (!ctx->xbuf[0]) {
        for (int b = 0; b < 2; ++b) {
            CK(cudaHostAlloc((void**)&ctx->xbuf[b], XCHUNK, cudaHostAllocDefault));
            CK(cudaEventCreateWithFlags(&ctx->xev[b], cudaEventDisableTiming));
        }
    }
    const size_t nch = (bytes + XCHUNK - 1) / XCHUNK;
    if (down) {
        for (size_t i = 0; i <= nch; ++i) {
            if (i < nch) {
                const size_t off = i * XCHUNK, n = bytes - off < XCHUNK ? bytes - off : XCHUNK;
                CK(cudaMemcpyAsync(ctx->xbuf[i & 1], (const char*)src + off, n, kind, st));
                CK(cudaEventRecord(ctx->xev[i & 1], st));
            }
            if (i > 0) {
                const size_t off = (i - 1) * XCHUNK, n = bytes - off < XCHUNK ? bytes - off : XCHUNK;
                CK(cudaEventSynchronize(ctx->xev[(i - 1) & 1]));
                par_memcpy((char*)dst + off, ctx->xbuf[(i - 1) & 1], n);
            }
        }
    } else {
        for (size_t i = 0; i < nch; ++i) {
            const size_t off = i * XCHUNK, n = bytes - off < XCHUNK ? bytes - off : XCHUNK;
            CK(cudaEventSynchronize(ctx->xev[i & 1]));              // the buffer's previous DMA (if any) has drained
            par_memcpy(ctx->xbuf[i & 1], (const char*)src + off, n);
            CK(cudaMemcpyAsync((char*)dst + off, ctx->xbuf[i & 1], n, kind, st));
            CK(cudaEventRecord(ctx->xev[i & 1], st));
        }
    }
    return 0;
}

template <typename T>
static int dev_upload(betse_ctx* ctx, T** p, const T* host, size_t n)
{
    int r = dev_alloc(ctx, p, n, host == nullptr);
    if (r) return r;
    if (host) return xfer(ctx, *p, host, n * sizeof(T), cudaMemcpyHostToDevice);
    return 0;
}

static void fill_kparams(betse_ctx* ctx, const betse_params* hp)
{
    KParams& P = ctx->P;
    P.n_ions = hp->n_ions;
    P.iNa = hp->iNa; P.iK = hp->iK; P.iCa = hp->iCa;
    for (int i = 0; i < BT_MAX_IONS; ++i) {
        const double z = hp->z[i];
        P.z[i] = z;
        P.zi[i] = (z == 1.0) ? 1 : (z == -1.0) ? -1 : (z == 2.0) ? 2 : (z == -2.0) ? -2 : 0;
        // A(+1)=slot0, B(+1)=1, A(+2)=2, B(+2)=3; a negative valence swaps A and B (kernels.cu GhkAB)
        P.ia[i] = (z == 1.0) ? 0 : (z == -1.0) ? 1 : (z == 2.0) ? 2 : 3;
        P.ib[i] = (z == 1.0) ? 1 : (z == -1.0) ? 0 : (z == 2.0) ? 3 : 2;
        P.zF[i] = z * hp->F;                       // sim.zs * p.F
        P.Dgj_surf[i] = hp->D_gj[i] * hp->gj_surface;
        P.cbound[i] = hp->c_env_bound[i];
        P.sig_k[i] = (z * z) * (hp->F * hp->F);
        P.D_free[i] = hp->D_free[i];
    }
    P.F = hp->F;
    P.RT_sim = hp->R * hp->T_sim;
    P.RT_p = hp->R * hp->T_p;
    P.R_T_p = hp->R * hp->T_p;
    P.kbT_sim = hp->kb * hp->T_sim;
    P.q = hp->q;
    P.cm = hp->cm;
    P.inv_cm = 1.0 / hp->cm;
    P.tm = hp->tm;
    P.dt = hp->dt;
    P.alpha_NaK = hp->alpha_NaK; P.alpha_Ca = hp->alpha_Ca;
    P.KmNK_Na = hp->KmNK_Na; P.KmNK_K = hp->KmNK_K; P.KmNK_ATP = hp->KmNK_ATP;
    P.KmCa_Ca = hp->KmCa_Ca; P.KmCa_ATP = hp->KmCa_ATP;
    P.cATP = hp->cATP; P.cADP = hp->cADP; P.cPi = hp->cPi;
    P.K0 = exp(-(hp->deltaGATP / (hp->R * hp->T_sim)));
    P.gj_vthresh = hp->gj_vthresh; P.gj_min = hp->gj_min;
    P.rho_pump = hp->rho_pump; P.rho_channel = hp->rho_channel;
    P.NaK_block = hp->NaKATP_block_scalar; P.gj_block = hp->gj_block_scalar;
    P.env_vol_div = hp->cell_height * (P.delta * P.delta);   // p.cell_height*cells.delta**2
    P.ko_eo_er = (hp->ko_env * hp->eo) * hp->er;
    P.screen = (2.0 / (hp->ko_env * P.delta)) * (hp->cell_radius / hp->true_cell_size);
    P.true_cell_size = hp->true_cell_size; P.cell_radius = hp->cell_radius;
    P.polar = hp->cell_polarizability != 0.0 ? 1 : 0;
    P.vol_env = hp->vol_env;
    P.sharpness = hp->sharpness;
    P.smooth_cells = hp->smooth_cells;
    P.is_ecm = hp->is_ecm; P.v_sensitive_gj = hp->v_sensitive_gj;
    P.cluster_open = hp->cluster_open; P.fast_update_ecm = hp->fast_update_ecm;
    // scipy.ndimage._gaussian_kernel1d(sigma=1, radius=4) taps, formed by the host in NumPy
    for (int k = 0; k <= 4; ++k) P.gw[k] = hp->gauss_w[k];
    P.inv_RT_sim = 1.0 / P.RT_sim; P.inv_RT_p = 1.0 / P.RT_p;
    P.inv_tm = 1.0 / hp->tm; P.inv_gjl = 1.0 / P.gj_len;
    for (int i = 0; i < BT_MAX_IONS; ++i) P.Dgj_len[i] = P.Dgj_surf[i] * P.inv_gjl;
    P.inv_kbT_sim = 1.0 / P.kbT_sim;
    P.inv_delta = 1.0 / P.delta; P.inv_2delta = 1.0 / (2.0 * P.delta);
    P.inv_KmNK_Na = 1.0 / hp->KmNK_Na; P.inv_KmNK_K = 1.0 / hp->KmNK_K; P.inv_KmCa_Ca = 1.0 / hp->KmCa_Ca;
    P.tNK = hp->cATP / hp->KmNK_ATP; P.tCa = hp->cATP / hp->KmCa_ATP;
    P.QnNK0 = (hp->cADP * 1e-3) * (hp->cPi * 1e-3); P.QdNK0 = hp->cATP * 1e-3; P.QnCa0 = hp->cADP * hp->cPi;
    P.dtm = hp->dt * 1.0e3;
    P.inv_K0 = 1.0 / P.K0;
    ctx->hp = *hp;
}

extern "C" int betse_abi_version(void) { return BETSE_ABI_VERSION; }

extern "C" int betse_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" const char* betse_kernel_name(int k)
{
    return (k >= 0 && k < BETSE_NKERNELS) ? kKernelNames[k] : "";
}

extern "C" int betse_last_error(betse_ctx* ctx, char* buf, size_t n)
{
    if (!ctx || !buf || n == 0) return 1;
    snprintf(buf, n, "%s", ctx->err.c_str());
    return 0;
}

static std::string g_create_error;

extern "C" int betse_create_error(char* buf, size_t n)
{
    if (!buf || n == 0) return 1;
    snprintf(buf, n, "%s", g_create_error.c_str());
    return 0;
}

static void destroy_graphs(betse_ctx* ctx)
{
    for (int i = 0; i < 2; ++i)
        if (ctx->gexec[i]) { cudaGraphExecDestroy(ctx->gexec[i]); ctx->gexec[i] = nullptr; }
    ctx->graphs_built = false;
    if (ctx->fast_graph) { cudaGraphExecDestroy(ctx->fast_graph); ctx->fast_graph = nullptr; }
    ctx->graph_epoch++;
    for (auto& g : ctx->ens) if (g.exec) cudaGraphExecDestroy(g.exec);
    ctx->ens.clear();
}

extern "C" void betse_destroy(betse_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    destroy_graphs(ctx);
    if (ctx->ev_init) for (auto& e : ctx->ev) cudaEventDestroy(e);
    if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    for (auto& e : ctx->ens_ev) if (e) cudaEventDestroy(e);
    for (void* p : ctx->ipc_opened) cudaIpcCloseMemHandle(p);
    for (void* p : ctx->allocs) cudaFree(p);
    for (int b = 0; b < 2; ++b) { if (ctx->xbuf[b]) cudaFreeHost(ctx->xbuf[b]); if (ctx->xev[b]) cudaEventDestroy(ctx->xev[b]); }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// BETSE_TIMING=1: wall-clock marks of the set-up phases on stderr (tools/e2e_cprofile.py)
struct PhaseTimer {
    bool on; std::chrono::steady_clock::time_point t0; const char* what;
    explicit PhaseTimer(const char* w) : on(getenv("BETSE_TIMING") != nullptr), t0(std::chrono::steady_clock::now()), what(w) {}
    void mark(const char* name) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[betse timing] %s: %-28s %7.2f ms\n", what, name, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

static int create_impl(betse_ctx* ctx, const betse_mesh* mesh, const betse_params* hp)
{
    PhaseTimer T("create");
    if (hp->abi_version != BETSE_ABI_VERSION) return fail(ctx, "betse_params.abi_version mismatch");
    if (hp->n_ions < 4 || hp->n_ions > BETSE_MAX_IONS) return fail(ctx, "n_ions must be in [4,8]");
    if (hp->cell_polarizability != 0.0) {
        // Vmem becomes per-membrane state advanced from Jn every step (sim.py:2048-2080)
        if (!mesh->R_rads) return fail(ctx, "cell_polarizability != 0 needs cells.R_rads");
        if (mesh->n_cells_owned > 0 && mesh->n_cells_owned != mesh->n_cells)
            return fail(ctx, "cell_polarizability != 0 on a domain-decomposed tissue is not implemented");
    }
    if (hp->iNa < 0 || hp->iK < 0) return fail(ctx, "Na and K must be enabled");
    if (mesh->n_cells <= 0 || mesh->n_mems <= 0) return fail(ctx, "empty mesh");
    ctx->C = mesh->n_cells;
    ctx->Co = mesh->n_cells_owned > 0 ? mesh->n_cells_owned : mesh->n_cells;
    ctx->M = mesh->n_mems;
    ctx->Mo = mesh->n_mems_owned > 0 ? mesh->n_mems_owned : mesh->n_mems;
    ctx->ny = mesh->ny; ctx->nx = mesh->nx; ctx->E = mesh->ny * mesh->nx;
    ctx->I = hp->n_ions;
    const int C = ctx->C, Co = ctx->Co, Mo = ctx->Mo, E = ctx->E, I = ctx->I;
    if (ctx->M != Mo) return fail(ctx, "ghost membranes are not supported: pass owned membranes only");
    if (mesh->cell_mem_ptr[Co] != Mo) return fail(ctx, "cell_mem_ptr[n_cells_owned] != n_mems_owned");

    ctx->h_cmp.assign(mesh->cell_mem_ptr, mesh->cell_mem_ptr + Co + 1);
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    { const char* e = getenv("BETSE_OVERLAP"); ctx->overlap = !(e && e[0] == '0'); }
    memset(&ctx->A, 0, sizeof(KArrays));
    memset(&ctx->P, 0, sizeof(KParams));
    memset(ctx->nets, 0, sizeof ctx->nets);
    KArrays& A = ctx->A;
    KParams& P = ctx->P;
    P.delta = mesh->delta;
    P.gj_len = mesh->gj_len;
    P.ecm_vol = mesh->ecm_vol;
    P.memsa_mean = mesh->memsa_mean;
    P.ny = mesh->ny; P.nx = mesh->nx;
    P.y0 = mesh->y0;
    P.ny_global = mesh->ny_global > 0 ? mesh->ny_global : mesh->ny;
    P.y_own0 = mesh->y_own0;
    P.y_own1 = mesh->y_own1 > 0 ? mesh->y_own1 : mesh->ny;
    P.n_cells = C; P.n_cells_owned = Co; P.n_mems_owned = Mo;
    P.yi0 = P.ya0 = P.yf0 = 0;
    P.yi1 = P.ya1 = P.yf1 = mesh->ny;
    if ((long long)hp->n_ions * std::max(std::max(C, ctx->M), ctx->E) >= (1LL << 31) - 1)
        return fail(ctx, "I*max(C,M,E) must stay below 2^31 (32-bit device indices)");
    fill_kparams(ctx, hp);

    // ---- host-side index work on helper threads, next to the uploads of the raw arrays (the driver stages pageable
    //      memory on the calling thread; both are memory-bound loops over 1e6..1e7 entries)
    int r;
    std::vector<int> nnc(Mo);
    std::string err_nnc, err_pack;
    const bool par = [] { const char* e = getenv("BETSE_CREATE_THREADS"); return !(e && e[0] == '0'); }();
    auto job_nnc = [&] {
        for (int m = 0; m < Mo; ++m) {
            int nn = mesh->nn_i[m];
            int cn;
            if (nn >= 0 && nn < Mo) cn = mesh->mem_to_cells[nn];
            else if (nn <= -2) cn = -(nn + 2);      // multi-GPU: partner lives on a ghost cell: nn = -(ghost_cell+2)
            else { err_nnc = "nn_i out of range"; return; }
            if (cn < 0 || cn >= C) { err_nnc = "partner cell out of range"; return; }
            if (mesh->map_mem2ecm[m] < 0 || mesh->map_mem2ecm[m] >= E) { err_nnc = "map_mem2ecm out of range"; return; }
            if (mesh->mem_to_cells[m] < 0 || mesh->mem_to_cells[m] >= Co) { err_nnc = "mem_to_cells out of range"; return; }
            nnc[m] = cn | (mesh->bflags_mems[m] ? (int)0x80000000 : 0);
        }
    };
    // CTA packing (k_diag): contiguous runs of whole cells with <= BT_TPB membranes; warp packing (k_mem): contiguous runs
    // of whole cells with <= 32 membranes, <= BT_TILE_MAXC cells
    std::vector<int> cta_start, tile_start, tdesc;
    auto job_pack = [&] {
        cta_start.push_back(0);
        for (int c = 0; c < Co;) {
            int mstart = mesh->cell_mem_ptr[c];
            int cc = c;
            while (cc < Co && (mesh->cell_mem_ptr[cc + 1] - mstart) <= BT_TPB && (cc - c) < BT_MAX_CTA_CELLS) ++cc;
            if (cc == c) { err_pack = "a cell has more than 256 membranes"; return; }
            cta_start.push_back(cc);
            c = cc;
        }
        tile_start.push_back(0);
        for (int c = 0; c < Co;) {
            int mstart = mesh->cell_mem_ptr[c];
            int cc = c;
            while (cc < Co && (mesh->cell_mem_ptr[cc + 1] - mstart) <= 32 && (cc - c) < BT_TILE_MAXC) ++cc;
            if (cc == c) { err_pack = "a cell has more than 32 membranes (unsupported by the warp-tile kernel)"; return; }
            tile_start.push_back(cc);
            c = cc;
        }
        const int nt = (int)tile_start.size() - 1;
        tdesc.resize((size_t)nt * 4);
        for (int t = 0; t < nt; ++t) {
            const int a = tile_start[t], b = tile_start[t + 1];
            tdesc[4 * t] = a; tdesc[4 * t + 1] = b - a;
            tdesc[4 * t + 2] = mesh->cell_mem_ptr[a]; tdesc[4 * t + 3] = mesh->cell_mem_ptr[b] - mesh->cell_mem_ptr[a];
        }
    };
    // env point -> flux slot CSR (map_ecm2mem, cells.py:1793-1796), slots in membrane order
    const bool own_csr = !(mesh->ecm_slot_ptr && mesh->ecm_slot_idx);
    std::vector<int> csr_ptr, csr_idx;
    auto job_csr = [&] {
        if (!own_csr) return;
        csr_ptr.assign((size_t)E + 1, 0); csr_idx.resize(Mo);
        for (int m = 0; m < Mo; ++m) {
            const int k = mesh->map_mem2ecm[m];
            if (k >= 0 && k < E) csr_ptr[k + 1]++;
        }
        for (int k = 0; k < E; ++k) csr_ptr[k + 1] += csr_ptr[k];
        std::vector<int> fillp(csr_ptr.begin(), csr_ptr.end() - 1);
        for (int m = 0; m < Mo; ++m) {
            const int k = mesh->map_mem2ecm[m];
            if (k >= 0 && k < E) csr_idx[fillp[k]++] = m;
        }
    };
    // ---- patches of k_cell_patch (kcell.cu): undivided tissues with extracellular spaces.  Cells sorted into stripes of
    //      PATCH_TY env rows and along x inside a stripe, cut into chunks of 128 = compact rectangles of cells whatever
    //      the cell numbering; an env square all of whose membranes lie in one patch is OWNED by it and finished there
    struct PatchPlan {
        bool on = false;
        int n_patches = 0, kb_max = 0, kb_min = INT_MAX, tab_max = 0;
        std::vector<int> pcell, row0, ptab_ptr, ptab, slot_ptrp, out_sq;
        std::vector<int> bidx;                 // compact border slot per membrane, -1 = its env square is owned by a patch
        int n_border = 0;
        std::string err;
    } pp;
    const bool want_patches = own_csr && hp->is_ecm && hp->n_ions <= 7 && Co == C && E < (1 << 28) && kcell_enabled();
    auto job_patch = [&] {
        if (!want_patches) return;
        const int* cmp = mesh->cell_mem_ptr;
        const int nx = ctx->nx, ny = ctx->ny;
        int nm_max = 0;
        for (int c = 0; c < Co; ++c) nm_max = std::max(nm_max, cmp[c + 1] - cmp[c]);
        if (!kcell_patch_fits(I, nm_max)) return;
        constexpr int PATCH_TY = 8;
        const int n_str = (ny + PATCH_TY - 1) / PATCH_TY;
        // counting sort of the cells by (stripe, x) of the env square of their first membrane
        std::vector<int> cq(Co), cnt((size_t)n_str * nx + 1, 0), order(Co);
        for (int c = 0; c < Co; ++c) {
            const int q = (cmp[c + 1] > cmp[c]) ? mesh->map_mem2ecm[cmp[c]] : 0;
            if (q < 0 || q >= E) { pp.err = "map_mem2ecm out of range"; return; }
            cq[c] = q;
            cnt[(size_t)((q / nx) / PATCH_TY) * nx + (q % nx) + 1]++;
        }
        for (size_t b = 1; b < cnt.size(); ++b) cnt[b] += cnt[b - 1];
        for (int c = 0; c < Co; ++c) { const int q = cq[c]; order[cnt[(size_t)((q / nx) / PATCH_TY) * nx + (q % nx)]++] = c; }
        const int np = (Co + 127) / 128;
        pp.n_patches = np;
        pp.pcell.assign((size_t)np * 128, -1);
        std::vector<int> patch_of(Co), pos_of(Co);
        for (int p = 0; p < np; ++p) {
            const int a = p * 128, b = std::min(Co, a + 128);
            // inside a patch: row by row, so that the lanes of a warp are neighbours along x (coalesced cell state)
            std::sort(order.begin() + a, order.begin() + b, [&](int u, int v) { return cq[u] != cq[v] ? cq[u] < cq[v] : u < v; });
            for (int j = a; j < b; ++j) { pp.pcell[j] = order[j]; patch_of[order[j]] = p; pos_of[order[j]] = j - a; }
        }
        const int nb = np * 4;
        pp.row0.assign(2 * ((size_t)nb + 1), 0);
        for (int b = 0; b < nb; ++b) {
            int kb = 0;
            for (int l = 0; l < 32; ++l) { const int c = pp.pcell[(size_t)b * 32 + l]; if (c >= 0) kb = std::max(kb, cmp[c + 1] - cmp[c]); }
            pp.row0[2 * (b + 1)] = pp.row0[2 * b] + kb;
            pp.kb_max = std::max(pp.kb_max, kb);
            if (kb > 0) pp.kb_min = std::min(pp.kb_min, kb);
        }
        // owner of every env square: the patch that holds ALL its membranes, or -1
        std::vector<int> owner(E, -1), n_own(np + 1, 0);
        for (int q = 0; q < E; ++q) {
            const int s0 = csr_ptr[q], s1 = csr_ptr[q + 1];
            if (s1 == s0) continue;
            const int p0 = patch_of[mesh->mem_to_cells[csr_idx[s0]]];
            bool same = true;
            for (int j = s0 + 1; j < s1 && same; ++j) same = patch_of[mesh->mem_to_cells[csr_idx[j]]] == p0;
            if (same) { owner[q] = p0; n_own[p0 + 1]++; }
        }
        // border slots in pack order (patch, cell, membrane): a warp's border stores stay close together
        pp.bidx.assign(Mo, -1);
        for (size_t j = 0; j < pp.pcell.size(); ++j) {
            const int c = pp.pcell[j];
            if (c < 0) continue;
            for (int m = cmp[c]; m < cmp[c + 1]; ++m) if (owner[mesh->map_mem2ecm[m]] < 0) pp.bidx[m] = pp.n_border++;
        }
        pp.slot_ptrp.assign(csr_ptr.begin(), csr_ptr.end());
        for (int q = 0; q < E; ++q) if (owner[q] < 0) pp.out_sq.push_back(q);
        for (int q = 0; q < E; ++q) if (owner[q] >= 0) pp.slot_ptrp[q] = (int)((unsigned)pp.slot_ptrp[q] | 0x80000000u);
        for (int p = 0; p < np; ++p) n_own[p + 1] += n_own[p];
        std::vector<int> own_sq(n_own[np]), fill(n_own.begin(), n_own.end() - 1);
        for (int q = 0; q < E; ++q) if (owner[q] >= 0) own_sq[fill[owner[q]]++] = q;
        // tables: [n, q[n], (first slot | count << 16)[n], slots as u16 ...] per patch
        pp.ptab_ptr.assign((size_t)np + 1, 0);
        pp.ptab.reserve((size_t)n_own[np] * 2 + (size_t)Mo / 2 + np * 2);
        std::vector<unsigned short> sl;
        for (int p = 0; p < np; ++p) {
            pp.ptab_ptr[p] = (int)pp.ptab.size();
            const int n = n_own[p + 1] - n_own[p];
            pp.ptab.push_back(n);
            for (int t = 0; t < n; ++t) pp.ptab.push_back(own_sq[n_own[p] + t]);
            sl.clear();
            for (int t = 0; t < n; ++t) {
                const int q = own_sq[n_own[p] + t];
                const int s0 = csr_ptr[q], s1 = csr_ptr[q + 1];
                if (sl.size() > 65535 || s1 - s0 > 65535) { pp.err = "patch table overflow"; return; }
                pp.ptab.push_back((int)((unsigned)sl.size() | ((unsigned)(s1 - s0) << 16)));
                for (int j = s0; j < s1; ++j) {
                    const int m = csr_idx[j], c = mesh->mem_to_cells[m], pos = pos_of[c];
                    sl.push_back((unsigned short)((((pos >> 5) * pp.kb_max + (m - cmp[c])) * I) * 32 + (pos & 31)));
                }
            }
            if (sl.size() & 1) sl.push_back(0);
            for (size_t j = 0; j < sl.size(); j += 2) pp.ptab.push_back((int)((unsigned)sl[j] | ((unsigned)sl[j + 1] << 16)));
        }
        pp.ptab_ptr[np] = (int)pp.ptab.size();
        for (int p = 0; p < np; ++p) pp.tab_max = std::max(pp.tab_max, pp.ptab_ptr[p + 1] - pp.ptab_ptr[p]);
        pp.on = pp.tab_max <= 2048;
    };
    std::thread th_nnc, th_pack, th_csr;
    if (par) { th_nnc = std::thread(job_nnc); th_pack = std::thread(job_pack); th_csr = std::thread([&] { job_csr(); job_patch(); }); }
    else { job_nnc(); job_pack(); job_csr(); job_patch(); }
    struct Joiner { std::thread &a, &b, &c; ~Joiner() { if (a.joinable()) a.join(); if (b.joinable()) b.join(); if (c.joinable()) c.join(); } } joiner{th_nnc, th_pack, th_csr};
    T.mark("streams, params");
    // ---- index arrays and geometry: raw uploads
    if ((r = dev_upload(ctx, (int**)&A.mem_to_cells, mesh->mem_to_cells, Mo))) return r;
    if ((r = dev_upload(ctx, (int**)&A.cell_mem_ptr, mesh->cell_mem_ptr, Co + 1))) return r;
    if ((r = dev_upload(ctx, (int**)&A.nn_i, mesh->nn_i, Mo))) return r;
    if ((r = dev_upload(ctx, (int**)&A.map_mem2ecm, mesh->map_mem2ecm, Mo))) return r;
    if ((r = dev_upload(ctx, (double**)&A.mem_sa, mesh->mem_sa, Mo))) return r;
    if ((r = dev_upload(ctx, (double**)&A.mem_nx, mesh->mem_nx, Mo))) return r;
    if ((r = dev_upload(ctx, (double**)&A.mem_ny, mesh->mem_ny, Mo))) return r;
    if ((r = dev_upload(ctx, (double**)&A.cell_vol, mesh->cell_vol, Co))) return r;
    if ((r = dev_upload(ctx, (double**)&A.cell_sa, mesh->cell_sa, Co))) return r;
    if ((r = dev_upload(ctx, (double**)&A.diviterm, mesh->diviterm, Co))) return r;
    if ((r = dev_upload(ctx, (double**)&A.num_mems, mesh->num_mems, Co))) return r;
    if (mesh->memSa_per_envSquare) { if ((r = dev_upload(ctx, (double**)&A.memsa_env, mesh->memSa_per_envSquare, E))) return r; }
    else if (hp->fast_update_ecm) return fail(ctx, "fast_update_ecm needs memSa_per_envSquare");
    // cells.gj_default_weights is read by the static gap-junction mode only (sim.py:2175-2177)
    if (mesh->gj_default_weights && !hp->v_sensitive_gj) { if ((r = dev_upload(ctx, (double**)&A.gj_w, mesh->gj_default_weights, Mo))) return r; }
    else if (!hp->v_sensitive_gj) return fail(ctx, "static gap junctions need gj_default_weights");
    T.mark("raw uploads");
    if (th_nnc.joinable()) th_nnc.join();
    if (!err_nnc.empty()) return fail(ctx, err_nnc);
    if ((r = dev_upload(ctx, (int**)&A.nn_cell_flag, nnc.data(), Mo))) return r;
    if (th_pack.joinable()) th_pack.join();
    if (!err_pack.empty()) return fail(ctx, err_pack);
    ctx->n_ctas = (int)cta_start.size() - 1;
    P.n_ctas = ctx->n_ctas;
    if ((r = dev_upload(ctx, (int**)&A.cta_cell_start, cta_start.data(), cta_start.size()))) return r;
    ctx->n_tiles = (int)tile_start.size() - 1;
    P.n_tiles = ctx->n_tiles;
    { const char* e = getenv("BETSE_PF_TILES"); P.pf_tiles = e ? atoi(e) : 4096; }
    if ((r = dev_upload(ctx, (int**)&A.tile_desc, tdesc.data(), tdesc.size()))) return r;
    if (th_csr.joinable()) th_csr.join();
    ctx->n_slots = mesh->n_flux_slots > Mo ? mesh->n_flux_slots : Mo;
    if (!own_csr) {
        if ((r = dev_upload(ctx, (int**)&A.slot_ptr, mesh->ecm_slot_ptr, E + 1))) return r;
        if ((r = dev_upload(ctx, (int**)&A.slot_idx, mesh->ecm_slot_idx, mesh->ecm_slot_ptr[E]))) return r;
    } else {
        if ((r = dev_upload(ctx, (int**)&A.slot_ptr, csr_ptr.data(), E + 1))) return r;
        if ((r = dev_upload(ctx, (int**)&A.slot_idx, csr_idx.data(), Mo))) return r;
    }
    CK(cudaStreamSynchronize(ctx->stream));       // the host vectors above die with this scope
    T.mark("env slot CSR");
    T.mark("derived index arrays");
    // ---- cell pack (k_cell): SELL-32 rows of the per-membrane constants, built on the device like the tile pack
    if (th_csr.joinable()) th_csr.join();
    if (!pp.err.empty()) return fail(ctx, pp.err);
    if (hp->n_ions <= 7 && hp->is_ecm) {
        std::vector<int> row0;                           // int2 {first row, first membrane} per block
        int nb, kb_max = 0, kb_min = INT_MAX;
        if (pp.on) {
            // blocks composed from the patches' spatial order (their "first membrane" is not used)
            nb = pp.n_patches * 4; row0.swap(pp.row0); kb_max = pp.kb_max; kb_min = pp.kb_min;
        } else {
            nb = (Co + 31) / 32;
            row0.assign(2 * ((size_t)nb + 1), 0);
            for (int b = 0; b < nb; ++b) {
                int kb = 0;
                const int c1 = std::min(Co, (b + 1) * 32);
                for (int c = b * 32; c < c1; ++c) kb = std::max(kb, mesh->cell_mem_ptr[c + 1] - mesh->cell_mem_ptr[c]);
                row0[2 * (b + 1)] = row0[2 * b] + kb;
                row0[2 * b + 1] = mesh->cell_mem_ptr[b * 32];
                kb_max = std::max(kb_max, kb); kb_min = std::min(kb_min, kb);
            }
            row0[2 * nb + 1] = Mo;
        }
        const long long R32 = (long long)row0[2 * nb] * 32;
        // 32-bit flux positions and 28-bit env squares (the rows' flag bits); otherwise k_mem stays in charge
        if (R32 * hp->n_ions < (1LL << 31) - 64 && E < (1 << 28) && nb > 0) {
            P.n_blocks = nb; P.ell_rows = row0[2 * nb]; P.kb_max = kb_max; P.kb_min = kb_min;
            { const char* e = getenv("BETSE_KCELL_PF"); P.pf_dist = e ? atoi(e) : 1024; }
            { const char* e = getenv("BETSE_KCELL_PERSIST"); P.kc_persist = e ? atoi(e) : 1; }
            if ((r = dev_upload(ctx, (int**)&A.blk_row0, row0.data(), row0.size()))) return r;
            if ((r = dev_alloc(ctx, (char**)&A.cpack, (size_t)row0[2 * nb] * cell_pack_row_bytes(I)))) return r;
            if ((r = dev_alloc(ctx, &A.flux_ell, (size_t)R32 * I))) return r;
            if ((r = dev_alloc(ctx, &ctx->mem_ell, (size_t)Mo))) return r;
            const int n_sl = mesh->ecm_slot_ptr ? mesh->ecm_slot_ptr[E] : Mo;
            if ((r = dev_alloc(ctx, (int**)&A.slot_off, (size_t)n_sl))) return r;
            if ((r = dev_alloc(ctx, &ctx->kc_sync, (size_t)1))) return r;
            A.ticket = ctx->kc_sync;
            int* d_bidx = nullptr;
            if (pp.on) {
                P.n_patches = pp.n_patches; P.n_out_sq = (int)pp.out_sq.size(); P.ptab_max = pp.tab_max;
                { std::vector<int> lst(pp.out_sq); if (lst.empty()) lst.push_back(0); if ((r = dev_upload(ctx, (int**)&A.out_sq, lst.data(), lst.size()))) return r; }
                if ((r = dev_upload(ctx, (int**)&A.pcell, pp.pcell.data(), pp.pcell.size()))) return r;
                if ((r = dev_upload(ctx, (int**)&A.ptab_ptr, pp.ptab_ptr.data(), pp.ptab_ptr.size()))) return r;
                if ((r = dev_upload(ctx, (int**)&A.ptab, pp.ptab.data(), pp.ptab.size()))) return r;
                if ((r = dev_upload(ctx, (int**)&A.slot_ptrp, pp.slot_ptrp.data(), pp.slot_ptrp.size()))) return r;
                if ((r = dev_alloc(ctx, &A.sq_sum, (size_t)I * E))) return r;
                if ((r = dev_upload(ctx, &d_bidx, pp.bidx.data(), pp.bidx.size()))) return r;
                if ((r = dev_alloc(ctx, (int**)&A.bslot, (size_t)R32))) return r;
                if ((size_t)pp.n_border * 8 > (size_t)ctx->n_slots * I) return fail(ctx, "internal: border slots exceed the exchange-slot buffer");
            }
            launch_pack_cell_const(ctx->P, A, ctx->mem_ell, d_bidx, ctx->stream);
            launch_slot_off(A.slot_idx, ctx->mem_ell, const_cast<int*>(A.slot_off), n_sl, Mo, I, ctx->stream);
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(ctx->stream));           // the plan's host vectors die with this scope
        }
    }
    T.mark("cell pack");
    // ---- tile pack (k_mem_pipe): every tile's constant inputs as one fixed-size block, built on the device; the DmS
    //      rows are (re)built whenever Dm_cells or the schedule scalars are uploaded (launch_pack_dm)
    //      — only when k_cell (which supersedes k_mem_pipe) has no cell pack to run on
    if (hp->n_ions <= 7 && !(A.cpack && kcell_enabled())) {
        const size_t bytes = (size_t)ctx->n_tiles * tile_pack_size(hp->n_ions);
        if ((r = dev_alloc(ctx, (char**)&A.tile_pack, bytes))) return r;
        launch_pack_const(ctx->P, A, ctx->stream);
        CK(cudaGetLastError());
    }

    T.mark("tile pack");
    // ---- state
    const size_t IC = (size_t)I * C, IE = (size_t)I * E, IM = (size_t)I * Mo;
    if ((r = dev_alloc(ctx, &A.cc_cells, IC))) return r;
    {
        // the exchange window: one allocation (>= 4 MiB so that it is never sub-allocated and its
        // IPC handle maps exactly this block), carved at 256-byte boundaries
        betse_window_info& W = ctx->winfo;
        memset(&W, 0, sizeof W);
        size_t off = 0;
        auto carve = [&](size_t n_doubles) { size_t o = off; off += ((n_doubles * 8 + 255) / 256) * 256; return o; };
        const size_t nIE = hp->is_ecm ? IE : 1, nSl = hp->is_ecm ? (size_t)ctx->n_slots * I : 1;
        for (int b = 0; b < 2; ++b) { W.off_cc_mid[b] = carve(IC); W.off_vm_cell[b] = carve(C); W.off_cc_env[b] = carve(nIE); }
        W.off_flux = carve(nSl);
        W.off_v_raw = carve(E);
        W.off_flags = carve(16);
        if (off < ((size_t)4 << 20)) off = (size_t)4 << 20;
        W.bytes = off;
        if ((r = dev_alloc(ctx, &ctx->win, off))) return r;
        W.base = ctx->win;
        W.n_cells = C; W.n_env = E; W.nx = ctx->nx; W.n_ions = I;
        W.n_flux_slots = hp->is_ecm ? ctx->n_slots : 0;
        for (int b = 0; b < 2; ++b) {
            A.cc_mid[b] = (double*)(ctx->win + W.off_cc_mid[b]);
            A.vm_cell[b] = (double*)(ctx->win + W.off_vm_cell[b]);
            A.cc_env[b] = (double*)(ctx->win + W.off_cc_env[b]);
        }
        A.flux_slots = (double*)(ctx->win + W.off_flux);
        A.v_raw = (double*)(ctx->win + W.off_v_raw);
        memset(&ctx->X, 0, sizeof(XPlan));
        ctx->X.side_k[0] = ctx->X.side_k[1] = -1;
        ctx->X.my_flags = (unsigned long long*)(ctx->win + W.off_flags);
        ctx->X.epoch = ctx->X.my_flags + 4;
        ctx->X.done_ctr = (unsigned int*)(ctx->X.my_flags + 6);
        { const char* e = getenv("BETSE_XCHG_TIMEOUT_MS"); ctx->X.timeout_ns = (unsigned long long)(e ? atoll(e) : 10000) * 1000000ull; }
        A.xflags = ctx->X.my_flags; A.xepoch = ctx->X.epoch;
        P.x_timeout_ns = ctx->X.timeout_ns;
        // env squares that take flux slots a neighbour fills (slot >= Mo): local rows below xw_lo / from xw_hi on
        P.xw_lo = P.y_own0 + 1; P.xw_hi = P.y_own1 - 1;
        if (mesh->ecm_slot_ptr && mesh->ecm_slot_idx) {
            const int mid = (P.y_own0 + P.y_own1) / 2;
            for (int k = 0; k < E; ++k) {
                bool remote = false;
                for (int j = mesh->ecm_slot_ptr[k]; j < mesh->ecm_slot_ptr[k + 1] && !remote; ++j) remote = mesh->ecm_slot_idx[j] >= Mo;
                if (!remote) continue;
                const int y = k / ctx->nx;
                if (y < mid) P.xw_lo = std::max(P.xw_lo, y + 1); else P.xw_hi = std::min(P.xw_hi, y);
            }
        }
    }
    if ((r = dev_alloc(ctx, &A.gjopen, Mo))) return r;
    if ((r = dev_alloc(ctx, (double**)&A.Dm, IM))) return r;
    if ((r = dev_alloc(ctx, (double**)&A.Denv, hp->is_ecm ? IE : 1))) return r;
    if ((r = dev_alloc(ctx, &A.E_x, E))) return r;
    if ((r = dev_alloc(ctx, &A.E_y, E))) return r;
    if ((r = dev_alloc(ctx, &A.v_env, E))) return r;
    if ((r = dev_alloc(ctx, &A.rho_env, E))) return r;
    if ((r = dev_alloc(ctx, &A.rho_cells, C))) return r;
    if ((r = dev_alloc(ctx, &A.cenv_u, 16))) return r;
    if ((r = dev_alloc(ctx, &A.cenv_part, (size_t)ctx->n_tiles * 8))) return r;
    if ((r = dev_alloc(ctx, &A.status, 1))) return r;
    if ((r = dev_alloc(ctx, &A.vm_mem, Mo))) return r;
    if ((r = dev_alloc(ctx, &A.vm_ave, C))) return r;
    if (hp->cell_polarizability != 0.0) {
        if ((r = dev_alloc(ctx, &A.vm_pol[0], Mo))) return r;
        if ((r = dev_alloc(ctx, &A.vm_pol[1], Mo))) return r;
        if ((r = dev_upload(ctx, (double**)&A.R_rads, mesh->R_rads, Mo))) return r;
    }
    if (hp->is_ecm && hp->sharpness < 1.0) { if ((r = dev_alloc(ctx, &A.scratch_env, IE))) return r; }
    if (!hp->is_ecm) {
        double cu[16];
        for (int i = 0; i < 8; ++i) cu[i] = cu[8 + i] = hp->cenv_uniform[i];
        CK(cudaMemcpyAsync(A.cenv_u, cu, sizeof cu, cudaMemcpyHostToDevice, ctx->stream));
    }
    T.mark("state allocations");
    CK(prepare_kernels(I));
    CK(prepare_cell(I, ctx->P.kb_max, ctx->P.ptab_max));
    T.mark("prepare kernels");
    CK(cudaStreamSynchronize(ctx->stream));
    T.mark("stream sync");
    return 0;
}

extern "C" int betse_create(betse_ctx** out, const betse_mesh* mesh, const betse_params* params, int device)
{
    if (!out || !mesh || !params) { g_create_error = "null argument"; return 2; }
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        g_create_error = "no CUDA device available (this library has no CPU fallback)";
        return 3;
    }
    if (device < 0 || device >= n) { g_create_error = "device index out of range"; return 2; }
    betse_ctx* ctx = new betse_ctx();
    ctx->device = device;
    int r = create_impl(ctx, mesh, params);
    if (r) {
        g_create_error = ctx->err;
        betse_destroy(ctx);
        return r;
    }
    *out = ctx;
    return 0;
}

static int ensure_phi(betse_ctx* ctx);
static double nan_value() { return nan(""); }
static void publish_affect(betse_ctx* ctx);
static bool chan_cell_possible(betse_ctx* ctx);
static void chan_layout_sync(betse_ctx* ctx);
static void chan_expand_all(betse_ctx* ctx);

// sine matrices, eigenvalues and work buffers of the Dirichlet Poisson solve on the env grid (csrc/hh.cu)
static int ensure_poisson(betse_ctx* ctx)
{
    if (ctx->poisson_on) return 0;
    int r;
    HHBuf& H = ctx->hh;
    memset(&H, 0, sizeof H);
    const size_t my = ctx->ny - 2, mx = ctx->nx - 2;
    if ((r = dev_alloc(ctx, &H.Sy, my * my))) return r;
    if ((r = dev_alloc(ctx, &H.Sx, mx * mx))) return r;
    if ((r = dev_alloc(ctx, &H.ly, my))) return r;
    if ((r = dev_alloc(ctx, &H.lx, mx))) return r;
    if ((r = dev_alloc(ctx, &H.bB, (size_t)ctx->E))) return r;
    if ((r = dev_alloc(ctx, &H.R, 2 * my * mx))) return r;          // two solves at once (hh.cu:poisson)
    if ((r = dev_alloc(ctx, &H.T1, 2 * my * mx))) return r;
    {   // the folded (even / odd) sine transform: packed matrices and work
        const size_t hy = (my + 1) / 2, hx = (mx + 1) / 2;
        if ((r = dev_alloc(ctx, &H.PLy, hy * my))) return r;
        if ((r = dev_alloc(ctx, &H.PRx, mx * hx))) return r;
        if ((r = dev_alloc(ctx, &H.EO, 4 * std::max(hy * mx, my * hx)))) return r;
    }
    launch_hh_setup(H, ctx->ny, ctx->nx, ctx->stream);
    CK(cudaGetLastError());
    ctx->poisson_on = true;
    return 0;
}

static int ensure_diag_buffers(betse_ctx* ctx)
{
    KArrays& A = ctx->A;
    const size_t IM = (size_t)ctx->I * ctx->Mo, IE = (size_t)ctx->I * ctx->E;
    int r;
    const bool want_hh = (ctx->hp.is_ecm || ctx->denv_w) && ctx->X.n_nbr == 0 && ctx->ny > 2 && ctx->nx > 2;
    if (A.fl_mem && (ctx->hh_on || !want_hh)) return 0;
    if (!A.fl_mem) {
    if ((r = dev_alloc(ctx, &A.fl_mem, IM))) return r;
    if ((r = dev_alloc(ctx, &A.fl_gj, IM))) return r;
    if ((r = dev_alloc(ctx, &A.fl_env_x, ctx->hp.is_ecm ? IE : 1))) return r;
    if ((r = dev_alloc(ctx, &A.fl_env_y, ctx->hp.is_ecm ? IE : 1))) return r;
    if ((r = dev_alloc(ctx, &A.rate_NaK, ctx->Mo))) return r;
    double** mem_arrays[] = {&A.Jmem, &A.Jgj, &A.Jn, &A.I_mem, &A.Jc, &A.Emc, &A.dvm, &A.E_gj_x, &A.E_gj_y};
    for (auto p : mem_arrays) if ((r = dev_alloc(ctx, p, ctx->Mo))) return r;
    double** cell_arrays[] = {&A.J_cell_x, &A.J_cell_y, &A.E_cell_x, &A.E_cell_y, &A.sigma_cell};
    for (auto p : cell_arrays) if ((r = dev_alloc(ctx, p, ctx->C))) return r;
    }
    // Helmholtz-Hodge decomposition of the env current (ion_current.py:50-73) / of the bath current of a tissue without
    // extracellular spaces (ion_current.py:116-158): undivided tissues only
    if (want_hh && !ctx->hh_on) {
        if ((r = ensure_poisson(ctx))) return r;
        HHBuf& H = ctx->hh;
        double** eb[] = {&H.Jx, &H.Jy, &H.bA, &H.uA, &H.uB, &H.J_env_x, &H.J_env_y, &H.B_field, &H.Jtx, &H.Jty};
        for (auto pp : eb) if ((r = dev_alloc(ctx, pp, ctx->E))) return r;
        ctx->hh_on = true;
    }
    destroy_graphs(ctx);   // KArrays changed
    return 0;
}

template <typename T>
static int opt_array(betse_ctx* ctx, const T** slot, const T* host, size_t n)
{
    // optional device array: allocate on first use, then overwrite
    if (!host) return 0;
    if (!*slot) {
        T* p;
        int r = dev_alloc(ctx, &p, n, false);
        if (r) return r;
        *slot = p;
        destroy_graphs(ctx);
    }
    { int xr = xfer(ctx, (void*)*slot, host, n * sizeof(T), cudaMemcpyHostToDevice); if (xr) return xr; }
    return 0;
}

// every row of a[rows][n] one repeated value?  (the values go to vals)  Threads split each row; the first difference
// found anywhere stops all of them.
static bool rows_uniform(const double* a, int rows, size_t n, double* vals)
{
    if (n == 0) return false;
    const unsigned hw = std::thread::hardware_concurrency();
    const int T = (n * (size_t)rows < ((size_t)1 << 20)) ? 1 : (hw >= 16 ? 8 : hw >= 8 ? 4 : 2);
    std::atomic<bool> differs(false);
    for (int r = 0; r < rows; ++r) vals[r] = a[(size_t)r * n];
    auto scan = [&](int t) {
        const size_t lo = n * (size_t)t / T, hi = n * (size_t)(t + 1) / T;
        for (int r = 0; r < rows && !differs.load(std::memory_order_relaxed); ++r) {
            const double* p = a + (size_t)r * n;
            const double v = vals[r];
            for (size_t j0 = lo; j0 < hi; j0 += 4096) {
                const size_t j1 = j0 + 4096 < hi ? j0 + 4096 : hi;
                bool d = false;
                for (size_t j = j0; j < j1; ++j) d |= (p[j] != v);
                if (d) { differs.store(true, std::memory_order_relaxed); return; }
                if (differs.load(std::memory_order_relaxed)) return;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(scan, t);
    scan(0);
    for (auto& x : th) x.join();
    return !differs.load();
}

struct KRowVals { double v[BT_MAX_IONS]; };
static __global__ void k_fill_rows(double* __restrict__ dst, const KRowVals rv, const int n)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dst[(size_t)blockIdx.y * n + j] = rv.v[blockIdx.y];
}

extern "C" int betse_upload_state(betse_ctx* ctx, const betse_state_host* s)
{
    if (!ctx || !s) return 2;
    CK(cudaSetDevice(ctx->device));
    KArrays& A = ctx->A;
    const int C = ctx->C, Mo = ctx->Mo, E = ctx->E, I = ctx->I;
    const size_t IC = (size_t)I * C, IE = (size_t)I * E, IM = (size_t)I * Mo;
    const int cur = ctx->cur;
    cudaStream_t st = ctx->stream;
#define UP(dst, src, n) if (src) { int xr_ = xfer(ctx, (void*)(dst), (src), (n) * sizeof(double), cudaMemcpyHostToDevice); if (xr_) return xr_; }
    UP(A.cc_cells, s->cc_cells, IC);
    UP(A.cc_mid[cur], s->cc_at_mem_cell, IC);
    if (ctx->hp.is_ecm) {
        UP(A.cc_env[cur], s->cc_env, IE);
        UP(A.Denv, s->D_env_eff, IE);
        UP(A.E_x, s->E_env_x, E);
        UP(A.E_y, s->E_env_y, E);
    }
    UP(A.gjopen, s->gjopen, Mo);
    if (s->Dm_cells) {
        // sim.Dm_cells is one value per ion on every membrane unless tissue profiles or scheduled interventions differ
        // (sim.py:676-690, tishandler.py:1321-1332): I scalars and a device fill instead of I*M doubles over PCIe
        // (288 MB = ~48 ms of the pageable path at 1 M cells); the scan stops at the first membrane that differs
        double vals[BT_MAX_IONS];
        const char* eu = getenv("BETSE_DM_UNIFORM");                  // 0: always the full upload (A/B test)
        if (!(eu && eu[0] == '0') && rows_uniform(s->Dm_cells, I, (size_t)Mo, vals)) {
            KRowVals rv;
            for (int i = 0; i < BT_MAX_IONS; ++i) rv.v[i] = i < I ? vals[i] : 0.0;
            k_fill_rows<<<dim3((unsigned)((Mo + 255) / 256), (unsigned)I), 256, 0, st>>>(const_cast<double*>(A.Dm), rv, Mo);
        } else { UP(A.Dm, s->Dm_cells, IM); }
        launch_pack_dm(ctx->P, A, st); launch_pack_cell_dm(ctx->P, A, st);
    }
    if (ctx->P.polar && s->vm)
        CK(cudaMemcpyAsync(A.vm_pol[cur], s->vm, (size_t)Mo * sizeof(double), cudaMemcpyHostToDevice, st));
    if (s->vm_cell) {
        CK(cudaMemcpyAsync(A.vm_cell[cur], s->vm_cell, (size_t)C * sizeof(double), cudaMemcpyHostToDevice, st));
    } else if (s->vm) {
        // default mode: vm = vm_cell[cell] - Phi_b[map_mem2ecm] (sim.py:2029) -> keep the per-cell part
        std::vector<double> vc(C, 0.0);
        std::vector<int> ptr(ctx->Co + 1), m2e;
        CK(cudaMemcpy(ptr.data(), A.cell_mem_ptr, (ctx->Co + 1) * sizeof(int), cudaMemcpyDeviceToHost));
        const double* phi = s->Phi_b;
        if (phi) { m2e.resize(Mo); CK(cudaMemcpy(m2e.data(), A.map_mem2ecm, Mo * sizeof(int), cudaMemcpyDeviceToHost)); }
        for (int c = 0; c < ctx->Co; ++c) {
            const int m = ptr[c];
            vc[c] = s->vm[m] + (phi ? phi[m2e[m]] : 0.0);
        }
        // ghost cells (multi-GPU) are filled by the first exchange
        CK(cudaMemcpyAsync(A.vm_cell[cur], vc.data(), (size_t)ctx->Co * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
    }
    int r;
    if (s->Phi_b) {
        bool nz = false;
        for (int k = 0; k < E; ++k) if (s->Phi_b[k] != 0.0) { nz = true; break; }
        for (int q = 0; q < 4; ++q) ctx->bound_prev[q] = ctx->hp.bound_V[q];
        if (nz || ctx->phi[0]) {
            // the uploaded potential belongs to the bound_V the Simulator holds now: new == old
            if ((r = ensure_phi(ctx))) return r;
            CK(cudaMemcpyAsync(ctx->phi[ctx->phi_new], s->Phi_b, (size_t)E * sizeof(double), cudaMemcpyHostToDevice, st));
            A.phi_b = A.phi_b_old = ctx->phi[ctx->phi_new];
            ctx->phi_lag = false;
            ctx->phi_nz = nz;
            ctx->P.has_phi = nz ? 1 : 0;
            destroy_graphs(ctx);      // KParams is baked into the captured launches
        }
    }
    if (s->D_env_weight && !ctx->hp.is_ecm) {
        if ((r = opt_array(ctx, &ctx->denv_w, (const double*)s->D_env_weight, E))) return r;
    }
    if ((r = opt_array(ctx, &A.extra_rho_cells, (const double*)s->extra_rho_cells, C))) return r;
    if ((r = opt_array(ctx, &A.extra_rho_env, (const double*)s->extra_rho_env, E))) return r;
    if ((r = opt_array(ctx, &A.extra_J_mem, (const double*)s->extra_J_mem, Mo))) return r;
    if ((r = opt_array(ctx, &A.NaK_block, (const double*)s->NaKATP_block, Mo))) return r;
    if ((r = opt_array(ctx, &A.gj_block, (const double*)s->gj_block, Mo))) return r;
    if (s->cenv_uniform && !ctx->hp.is_ecm) {
        // NaN = keep what the device holds for that ion (a scheduled change of one ion's bath, tishandler.py:759-777)
        double cu[16];
        bool all = true;
        for (int i = 0; i < I; ++i) all = all && s->cenv_uniform[i] == s->cenv_uniform[i];
        if (all) {
            for (int i = 0; i < 8; ++i) cu[i] = cu[8 + i] = (i < I) ? s->cenv_uniform[i] : 0.0;
            CK(cudaMemcpy(A.cenv_u, cu, sizeof cu, cudaMemcpyHostToDevice));
        } else {
            for (int i = 0; i < I; ++i) {
                if (s->cenv_uniform[i] != s->cenv_uniform[i]) continue;
                CK(cudaMemcpy(A.cenv_u + i, s->cenv_uniform + i, sizeof(double), cudaMemcpyHostToDevice));
                CK(cudaMemcpy(A.cenv_u + 8 + i, s->cenv_uniform + i, sizeof(double), cudaMemcpyHostToDevice));
            }
        }
    }
#undef UP
    CK(cudaStreamSynchronize(st));
    return 0;
}

static int ensure_phi(betse_ctx* ctx)
{
    if (ctx->phi[0]) return 0;
    int r;
    if ((r = dev_alloc(ctx, &ctx->phi[0], (size_t)ctx->E))) return r;
    if ((r = dev_alloc(ctx, &ctx->phi[1], (size_t)ctx->E))) return r;
    ctx->A.phi_b = ctx->A.phi_b_old = ctx->phi[ctx->phi_new];
    return 0;
}

// sim.bound_V changed (the external-voltage event, tissue/event/tisevevolt.py:76-88): Phi_b = lapENVinv . (-div_Jb)
// (ion_current.py:84-90, 160-168) is re-solved on the device.  get_current runs inside update_V at the END of a step,
// so the step that starts now still reads the Vmem formed with the previous potential (phi_b_old) and closes with
// the new one (phi_b); betse_step retires the lag after that one step.
static int update_phi_b(betse_ctx* ctx, const double bound[4])
{
    bool nz = false;
    for (int q = 0; q < 4; ++q) nz = nz || bound[q] != 0.0;
    for (int q = 0; q < 4; ++q) ctx->bound_prev[q] = bound[q];
    if (!nz && !ctx->phi[0]) return 0;                 // never had a potential: Phi_b stays identically 0
    if (ctx->X.n_nbr > 0) return fail(ctx, "domain decomposition does not support a boundary-voltage potential (Phi_b)");
    if (ctx->ny <= 2 || ctx->nx <= 2) return fail(ctx, "env grid too small for the boundary-voltage solve");
    int r;
    if ((r = ensure_phi(ctx))) return r;
    if ((r = ensure_poisson(ctx))) return r;
    const int old = ctx->phi_new;
    if (ctx->phi_lag) {
        // two schedule changes without a step in between: the step still starts from `old`'s predecessor
        if (nz) launch_phi_b(ctx->P, ctx->hh, bound, ctx->phi[ctx->phi_new], ctx->stream);
        else CK(cudaMemsetAsync(ctx->phi[ctx->phi_new], 0, (size_t)ctx->E * sizeof(double), ctx->stream));
    } else {
        ctx->phi_new = old ^ 1;
        if (nz) launch_phi_b(ctx->P, ctx->hh, bound, ctx->phi[ctx->phi_new], ctx->stream);
        else CK(cudaMemsetAsync(ctx->phi[ctx->phi_new], 0, (size_t)ctx->E * sizeof(double), ctx->stream));
        ctx->A.phi_b = ctx->phi[ctx->phi_new];
        ctx->A.phi_b_old = ctx->phi[old];
        ctx->phi_lag = true;
    }
    CK(cudaGetLastError());
    ctx->phi_nz = nz;
    ctx->P.has_phi = 1;                                // the old or the new potential may be non-zero
    destroy_graphs(ctx);
    return 0;
}

extern "C" int betse_set_schedule(betse_ctx* ctx, const betse_params* hp)
{
    if (!ctx || !hp) return 2;
    CK(cudaSetDevice(ctx->device));
    if (hp->n_ions != ctx->I || hp->is_ecm != ctx->hp.is_ecm) return fail(ctx, "set_schedule cannot change n_ions/is_ecm");
    if ((hp->cell_polarizability != 0.0) != (ctx->P.polar != 0)) return fail(ctx, "set_schedule cannot switch cell_polarizability on or off");
    const int has_phi = ctx->P.has_phi;
    fill_kparams(ctx, hp);
    ctx->P.has_phi = has_phi;
    destroy_graphs(ctx);              // KParams is baked into the captured launches
    launch_pack_dm(ctx->P, ctx->A, ctx->stream);   // DmS carries rho_channel/tm
    launch_pack_cell_dm(ctx->P, ctx->A, ctx->stream);
    bool bv = false;
    for (int q = 0; q < 4; ++q) bv = bv || hp->bound_V[q] != ctx->bound_prev[q];
    if (bv) { int r = update_phi_b(ctx, hp->bound_V); if (r) return r; }
    return 0;
}

// One timestep.  Order: env transport of every ion (reads last step's E field and
// cc_env[cur], writes cc_env[nxt]) -> membranes+cells (reads cc_env[cur] for GHK/NaK, cc_env[nxt]
// for the Ca pump) -> env accumulation + charge -> env field for the next step.
// Tables of the fused exchange points (k_cell pushes X1, k_envacc_ell pushes X2; xchg.cuh): which blocks of the cell pack hold
// cells that are ghosts on a neighbour or membranes whose env square a neighbour owns, their slots there, the ticket order
// (those blocks first), and how many CTAs at each end of the accumulation rows push.  Not capturable: betse_step calls it.
static int prepare_xfuse(betse_ctx* ctx)
{
    if (!ctx->xfuse_dirty || ctx->X.n_nbr == 0 || !ctx->A.cpack) return 0;
    KArrays& A = ctx->A;
    const int nb = ctx->P.n_blocks;
    std::vector<int> h_row0(2 * ((size_t)nb + 1));
    CK(cudaMemcpy(h_row0.data(), A.blk_row0, h_row0.size() * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<char> has_g(nb, 0), has_r(nb, 0);
    for (int k = 0; k < ctx->X.n_nbr; ++k) {
        for (int c : ctx->xh_cells[k]) has_g[c / 32] = 1;
        for (int m : ctx->xh_flux[k]) {
            const int c = (int)(std::upper_bound(ctx->h_cmp.begin(), ctx->h_cmp.end(), m) - ctx->h_cmp.begin()) - 1;
            has_r[c / 32] = 1;
        }
    }
    std::vector<int> blk_x(2 * (size_t)nb, -1), ghost, rslot;
    int nbb = 0, n0 = 0, n1 = 0;
    for (int b = 0; b < nb; ++b) {
        if (has_g[b]) { blk_x[2 * b] = (int)ghost.size() / 2; ghost.resize(ghost.size() + 64, -1); }
        if (has_r[b]) { blk_x[2 * b + 1] = (int)rslot.size(); rslot.resize(rslot.size() + (size_t)(h_row0[2 * (b + 1)] - h_row0[2 * b]) * 32, -1); }
        if (has_g[b] || has_r[b]) {
            ++nbb;
            if (b < nb / 2) n0 = std::max(n0, b + 1); else n1 = std::max(n1, nb - b);
        }
    }
    for (int k = 0; k < ctx->X.n_nbr; ++k) {
        const XNbr& x = ctx->X.nb[k];
        for (size_t j = 0; j < ctx->xh_cells[k].size(); ++j) {
            const int c = ctx->xh_cells[k][j];
            ghost[2 * ((size_t)blk_x[2 * (c / 32)] + (c & 31)) + x.side] = x.recv_cell0 + (int)j;
        }
        for (size_t j = 0; j < ctx->xh_flux[k].size(); ++j) {
            const int m = ctx->xh_flux[k][j];
            const int c = (int)(std::upper_bound(ctx->h_cmp.begin(), ctx->h_cmp.end(), m) - ctx->h_cmp.begin()) - 1;
            const int kk = m - ctx->h_cmp[c];
            int& slot = rslot[(size_t)blk_x[2 * (c / 32) + 1] + (size_t)kk * 32 + (c & 31)];
            if (slot >= 0) return fail(ctx, "a membrane's env square cannot belong to two neighbours");
            slot = (x.side << 30) | (x.recv_slot0 + (int)j);
        }
    }
    if (ghost.empty()) ghost.resize(2, -1);
    if (rslot.empty()) rslot.resize(1, -1);
    int r;
    if ((r = dev_upload(ctx, (int**)&A.blk_x, blk_x.data(), blk_x.size()))) return r;
    if ((r = dev_upload(ctx, (int**)&A.ghost_tab, ghost.data(), ghost.size()))) return r;
    if ((r = dev_upload(ctx, (int**)&A.rslot_tab, rslot.data(), rslot.size()))) return r;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->X.n_bblocks = nbb;
    ctx->X.blk_n0 = n0; ctx->X.blk_n1 = n1;
    ctx->X.env_n0 = ctx->X.env_n1 = 0;
    for (int k = 0; k < ctx->X.n_nbr; ++k) {
        int lo = 0, up = 0;
        envacc_end_ctas(ctx->P, ctx->X.nb[k].push_lo, ctx->X.nb[k].push_hi, &lo, &up);
        ctx->X.env_n0 = std::max(ctx->X.env_n0, lo); ctx->X.env_n1 = std::max(ctx->X.env_n1, up);
    }
    const int g = ((ctx->P.ya1 - ctx->P.ya0) * ctx->P.nx + 255) / 256;
    if (ctx->X.env_n0 + ctx->X.env_n1 > g) ctx->X.env_n1 = g - ctx->X.env_n0;      // the two ends overlap on a very thin strip
    ctx->X.n_push_ctas = ctx->X.env_n0 + ctx->X.env_n1;
    ctx->xfuse_dirty = false;
    destroy_graphs(ctx);
    return 0;
}

// ligand-gated channels add their flux to the membrane -> env slots of the ion loop (k_net_lig, networks.py:5847-5916): the
// slots must then be the [slot][ion] array of k_mem, not the ELL rows of k_cell
static void refresh_defer_slots(betse_ctx* ctx)
{
    ctx->P.defer_slots = (!ctx->net_gates[0].empty() || !ctx->net_gates[1].empty()) ? 1 : 0;
}

static void enqueue_phase(betse_ctx* ctx, int phase, int diag, cudaEvent_t* evs)
{
    if (phase == 0) refresh_defer_slots(ctx);
    const int I = ctx->I, cur = ctx->cur, nxt = cur ^ 1;
    cudaStream_t st = ctx->stream;
    const KArrays& A = ctx->A;
    const bool ecm = ctx->hp.is_ecm != 0;
    if (phase == 0) {
        // The membrane kernel reads cc_env[nxt] of Ca only (the Ca-ATPase sees the transported value,
        // sim.py:1282 after 2254): the Ca row is transported first, the other ions run on a second
        // stream NEXT TO the membrane kernel (which is latency-bound and leaves issue slots free).
        const bool chans = !ctx->chans.empty() || ctx->net_on[0] || ctx->net_on[1] || ctx->noise_on;   // deferred-update mode
        // (also in the deferred-update mode: the channels and networks read the transported cc_env, but they are enqueued
        // behind the join below)
        const bool overlap = ecm && ctx->overlap && !evs && ctx->hp.sharpness >= 1.0;
        if (ctx->kc_sync && mem_kernel_kind(I, ctx->P, A, diag) == 2) cudaMemsetAsync(ctx->kc_sync, 0, sizeof(int), st);
        if (ecm && overlap) {
            const int iCa = ctx->hp.iCa;
            if (iCa >= 0) launch_ion(ctx->P, A, ctx->nx, cur, diag, iCa, 1, st);
            cudaEventRecord(ctx->ev_fork, st);
            cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0);
            // fused exchange: the halo rows of the other ions are written by the neighbours' env accumulation (X2) only —
            // their flag-ordered stores must not race with a redundant transport of the same rows running next to k_cell
            // (only the Ca row is read there before the exchange, and it is transported ahead of k_cell)
            KParams Pi = ctx->P;
            if (ctx->xfuse_now) { Pi.yi0 = std::max(Pi.yi0, Pi.y_own0); Pi.yi1 = std::min(Pi.yi1, Pi.y_own1); }
            if (iCa >= 0) {
                launch_ion(Pi, A, ctx->nx, cur, diag, 0, iCa, ctx->stream2);
                launch_ion(Pi, A, ctx->nx, cur, diag, iCa + 1, I - iCa - 1, ctx->stream2);
            } else launch_ion(Pi, A, ctx->nx, cur, diag, 0, I, ctx->stream2);
            cudaEventRecord(ctx->ev_join, ctx->stream2);
        } else if (ecm) {
            launch_ion(ctx->P, A, ctx->nx, cur, diag, 0, I, st);
            if (evs) cudaEventRecord(evs[1], st);
            if (ctx->hp.sharpness < 1.0) {
                launch_ion_smooth(I, ctx->P, A, ctx->ny, ctx->nx, nxt, st);
                cudaMemcpyAsync(A.cc_env[nxt], A.scratch_env, (size_t)I * ctx->E * sizeof(double),
                                cudaMemcpyDeviceToDevice, st);
            }
        } else if (evs) cudaEventRecord(evs[1], st);
        if (evs) cudaEventRecord(evs[2], st);
        if (ctx->xfuse_now) { launch_cell_x(I, ctx->P, A, ctx->X, cur, st); ctx->flux_is_ell = 1; }
        else ctx->flux_is_ell = launch_mem(I, ctx->P, A, ctx->n_ctas, cur, diag, st);
        if (overlap) cudaStreamWaitEvent(st, ctx->ev_join, 0);
        if (chans) {
            // between the ion loop's fluxes and update_all_concs (sim.py:1290-1357), per handler: run_loop_channels
            // (networks.py:3115-3213), then run_loop (networks.py:2805-2982)
            // per-cell channel passes with nothing behind them (no network, no noise): the last pass closes the step itself
            // (update_all_concs, charge, Vmem of its cells: k_cell_update's arithmetic) and k_cell_update is not launched
            const bool fuse_tail = ctx->chan_cell_mode && !ctx->chans.empty() && !ctx->net_on[0] && !ctx->net_on[1] && !ctx->noise_pending;
            bool tail_done = false;
            int hmax = 0;
            for (const KChan& ch : ctx->chans) hmax = std::max(hmax, ch.handler);
            for (int h = 0; h < 2; ++h) {
                // every MasterOfNetworks keeps its OWN extra_J_mem / extra_Jenv (clear_run_loop, networks.py:2790-2801) and
                // publishes them at the end of its run_loop (networks.py:2971-2977): with both handlers enabled the LAST one
                // wins, so the accumulators start from zero at every enabled handler's block
                bool enabled = ctx->net_on[h];
                for (const KChan& ch : ctx->chans) enabled = enabled || ch.handler == h;
                if (!enabled) continue;
                if (A.chanJ) cudaMemsetAsync(A.chanJ, 0, (size_t)ctx->Mo * sizeof(double), st);
                if (A.extra_Jenv_x) {
                    cudaMemsetAsync(A.extra_Jenv_x, 0, (size_t)ctx->E * sizeof(double), st);
                    cudaMemsetAsync(A.extra_Jenv_y, 0, (size_t)ctx->E * sizeof(double), st);
                }
                if (ctx->net_on[h] && !ctx->net_trans[h].empty()) {
                    const KNet& Nh = ctx->nets[h];
                    for (int k = 0; k < Nh.K && Nh.tw; ++k)
                        if (Nh.tw_s[k] >= 0) launch_tw_gather(ctx->P, A, Nh.tw + (size_t)Nh.tw_s[k] * ctx->Mo, Nh.c + (size_t)k * ctx->C, st);
                    for (int i = 0; i < I && Nh.tw; ++i)
                        if (Nh.tw_i[i] >= 0) launch_tw_gather(ctx->P, A, Nh.tw + (size_t)Nh.tw_i[i] * ctx->Mo, A.cc_cells + (size_t)i * ctx->C, st);
                    for (size_t j = 0; j < ctx->net_trans[h].size(); ++j)
                        launch_transporter(ctx->P, A, ctx->nets[h], ctx->net_trans[h][j], ctx->net_trans_cm[h][j],
                                           ctx->net_trans_em[h][j], ctx->net_trans_mm[h][j], cur, st);
                }
                // per-cell path (channels.cu:k_chan_cell): consecutive channels that conduct different ions go as ONE pass —
                // none of them sees another's update_Co.  Otherwise one kernel pair per channel.
                const bool tw = ctx->net_on[h] && ctx->nets[h].tw;
                const size_t nch = ctx->chans.size();
                for (size_t k0 = 0; k0 < nch;) {
                    if (ctx->chans[k0].handler != h) { ++k0; continue; }
                    size_t k1 = k0 + 1;
                    if (ctx->chan_cell_mode) {
                        const size_t cap = (size_t)std::min(KCH_PACK, I);
                        unsigned ions = 1u << ctx->chans[k0].ion;
                        while (k1 < nch && k1 - k0 < cap && ctx->chans[k1].handler == h && !(ions & (1u << ctx->chans[k1].ion))) {
                            ions |= 1u << ctx->chans[k1].ion; ++k1;
                        }
                        bool last = fuse_tail && h == hmax;                 // the last pass of the last handler that has channels
                        for (size_t k = k1; k < nch && last; ++k) if (ctx->chans[k].handler == h) last = false;
                        launch_chan_cell(ctx->P, A, &ctx->chans[k0], (int)(k1 - k0), ctx->chan_ell, cur, diag, last ? 1 : 0, st);
                        tail_done = tail_done || last;
                    } else {
                        const KChan& ch = ctx->chans[k0];
                        launch_chan(ctx->P, A, ch, ctx->nets[h], cur, st);
                        // the channel's update_Co renews sim.cc_at_mem[ion] (sim_toolbox.py:1182): a transporter's nudge of it ends here
                        if (tw && ctx->nets[h].tw_i[ch.ion] >= 0)
                            launch_tw_gather(ctx->P, A, ctx->nets[h].tw + (size_t)ctx->nets[h].tw_i[ch.ion] * ctx->Mo,
                                             A.cc_cells + (size_t)ch.ion * ctx->C, st);
                    }
                    k0 = k1;
                }
                if (ctx->net_on[h])
                    for (const betse_modulator& md : ctx->net_mods[h])
                        if (md.target == 2)
                            launch_net_mod_tj(ctx->P, A, ctx->nets[h], md.prog, md.max_val, md.ion, ctx->net_tj[h], ctx->net_ntj[h],
                                              ctx->denv_raw, ctx->tj_mod, cur, st);
                        else
                            launch_net_mod(ctx->P, A, ctx->nets[h], md.prog, md.max_val,
                                           const_cast<double*>(md.target == 0 ? A.gj_block : A.NaK_block), cur, st);
                if (ctx->net_on[h]) {
                    const std::vector<betse_ligand_gate>& gs = ctx->net_gates[h];
                    for (size_t j = 0; j < gs.size(); ++j)
                        launch_lig_prep(ctx->P, A, ctx->nets[h], gs[j].species, gs[j].extracell, pow(gs[j].K, gs[j].n), gs[j].n,
                                        gs[j].max_val, ctx->lig_tmp[h] + j * (size_t)ctx->Mo, st);
                    launch_net(ctx->P, A, ctx->nets[h], ctx->net_Dgj[h].data(), ctx->net_Dm[h].data(),
                               ctx->net_env_on[h].data(), ctx->net_pumps[h].data(), (int)ctx->net_pumps[h].size(),
                               ctx->hp.deltaGATP / (ctx->hp.R * ctx->hp.T_sim),
                               ctx->net_intra[h].empty() ? nullptr : ctx->net_intra[h].data(), I, cur, st);
                    for (size_t j = 0; j < gs.size(); ++j) {
                        launch_net_lig(ctx->P, A, gs[j].ion, ctx->lig_tmp[h] + j * (size_t)ctx->Mo, gs[j].mod, cur, diag, st);
                        launch_chan_env(ctx->P, A, gs[j].ion, cur, st);
                    }
                }
            }
            if (ctx->noise_pending) {
                // after the networks, before update_all_concs (sim.py:1322-1339)
                launch_flux_apply(ctx->P, A, ctx->noise_ion, ctx->noise_flux, st);
                launch_chan_env(ctx->P, A, ctx->noise_ion, cur, st);
                ctx->noise_pending = false;
            }
            if (!tail_done) launch_cell_update(ctx->P, A, cur, st);
        }
        if (evs) cudaEventRecord(evs[3], st);
    } else if (phase == 1) {
        if (ecm && ctx->flux_is_ell) {
            KParams Pw = ctx->P;
            Pw.xwait = ctx->xwait_now;
            if (ctx->xfuse_now) launch_envacc_ell_x(I, Pw, A, ctx->X, nxt, st);
            else launch_envacc_ell(I, Pw, A, nxt, st);
        }
        else if (ecm) launch_envacc(I, ctx->P, A, ctx->E, nxt, 1, st);
        else launch_envmix(I, ctx->P, A, cur, st);
        if (evs) cudaEventRecord(evs[4], st);
    } else {
        if (ecm) {
            KParams Pw = ctx->P;
            Pw.xwait = ctx->xwait_now;
            launch_field(Pw, A, ctx->ny, ctx->nx, st);
        }
        if (evs) cudaEventRecord(evs[5], st);
        if (diag) launch_diag(I, ctx->P, A, ctx->n_ctas, nxt, st);
        if (diag && ctx->want_hh && ecm && ctx->hh_on && ctx->X.n_nbr == 0) {
            ctx->hh.mu = ctx->hp.mu;
            for (int q = 0; q < 4; ++q) ctx->hh.bound[q] = ctx->hp.bound_V[q];
            launch_hh(ctx->P, A, ctx->hh, st);
        }
        if (diag && ctx->want_hh && !ecm && ctx->hh_on && ctx->denv_w) {
            ctx->hh.mu = ctx->hp.mu;
            for (int q = 0; q < 4; ++q) ctx->hh.bound[q] = 0.0;          // HH_Decomp(..., bounds=None)
            launch_noecm_field(ctx->P, A, ctx->hh, ctx->hp.sigma_env, ctx->denv_w, cur, st);
        }
        if (evs) cudaEventRecord(evs[6], st);
        ctx->cur = nxt;
    }
}

static void enqueue_step(betse_ctx* ctx, int diag, cudaEvent_t* evs)
{
    // cell_polarizability != 0: update_V integrates Vmem from Jn (sim.py:2059-2061), so get_current's membrane part
    // (k_diag) belongs to every step; the Helmholtz-Hodge part stays with the sampled steps
    ctx->want_hh = diag != 0;
    if (ctx->P.polar || ctx->need_emc) diag = 1;
    if (evs) cudaEventRecord(evs[0], ctx->stream);
    const bool nbr = ctx->X.n_nbr > 0;
    const int nxt = ctx->cur ^ 1;
    // exchange points of a decomposed tissue whose neighbours live in other processes: the wait sits inside the consuming
    // kernels' edge CTAs (k_envacc_ell for X1, k_field for X2; everything later in the step is ordered behind them by the
    // stream), and — fused — the push inside the producing kernels (k_cell: X1, k_envacc_ell: X2), no exchange kernel at
    // all.  Sampled steps (k_mem + k_envacc) and strips that share this process keep the stand-alone k_xchg.
    static const int fuse_ok = [] { const char* e = getenv("BETSE_XFUSE"); return (e && e[0] == '0') ? 0 : 1; }();
    const bool inwait = nbr && ctx->xwait && !diag && !ctx->P.has_phi && !ctx->P.polar &&
                        mem_kernel_kind(ctx->I, ctx->P, ctx->A, diag) == 2 && ctx->hp.is_ecm;
    const bool fused = inwait && fuse_ok && !ctx->xfuse_dirty && ctx->A.ticket && ctx->P.kc_persist;
    const int xmode = inwait ? BETSE_XCHG_PUSH : (BETSE_XCHG_PUSH | BETSE_XCHG_WAIT);
    ctx->xwait_now = inwait ? 1 : 0;
    ctx->xfuse_now = fused ? 1 : 0;
    enqueue_phase(ctx, 0, diag, evs);
    if (nbr && !fused) launch_xchg(ctx->P, ctx->A, ctx->X, BETSE_XCHG_X1, nxt, xmode, ctx->flux_is_ell, ctx->stream);
    if (evs) cudaEventRecord(evs[7], ctx->stream);
    enqueue_phase(ctx, 1, diag, evs);
    if (nbr && !fused) launch_xchg(ctx->P, ctx->A, ctx->X, BETSE_XCHG_X2, nxt, xmode, ctx->flux_is_ell, ctx->stream);
    if (evs) cudaEventRecord(evs[8], ctx->stream);
    enqueue_phase(ctx, 2, diag, evs);
    ctx->xwait_now = 0;
    ctx->xfuse_now = 0;
    ctx->last_step_fused = fused;
}

static int build_graphs(betse_ctx* ctx)
{
    const int saved = ctx->cur;
    for (int b = 0; b < 2; ++b) {
        ctx->cur = b;
        cudaGraph_t g;
        CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        enqueue_step(ctx, 0, nullptr);
        CK(cudaStreamEndCapture(ctx->stream, &g));
        CK(cudaGraphInstantiate(&ctx->gexec[b], g, 0));
        CK(cudaGraphDestroy(g));
    }
    ctx->cur = saved;
    ctx->graphs_built = true;
    return 0;
}

static int read_status(betse_ctx* ctx, uint32_t* status_out)
{
    unsigned int st = 0;
    CK(cudaMemcpyAsync(&st, ctx->A.status, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (st) CK(cudaMemsetAsync(ctx->A.status, 0, sizeof st, ctx->stream));
    if (status_out) *status_out = st;
    return 0;
}

extern "C" int betse_step(betse_ctx* ctx, int nsteps, int flags, uint32_t* status_out)
{
    if (!ctx || nsteps < 0) return 2;
    CK(cudaSetDevice(ctx->device));
    const bool want_diag = (flags & BETSE_STEP_DIAG) != 0;
    if (want_diag || ctx->P.polar) { int r = ensure_diag_buffers(ctx); if (r) return r; }
    if (ctx->xwait) { int r = prepare_xfuse(ctx); if (r) return r; }
    chan_layout_sync(ctx);
    const int asked = nsteps;
    if (ctx->phi_lag && nsteps > 0) {
        // first step after a change of sim.bound_V: starts from the old potential, closes with the new one
        enqueue_step(ctx, (nsteps == 1 && want_diag) ? 1 : 0, nullptr);
        ctx->A.phi_b_old = ctx->A.phi_b;
        ctx->phi_lag = false;
        ctx->P.has_phi = ctx->phi_nz ? 1 : 0;
        destroy_graphs(ctx);
        --nsteps;
    }
    // warm the launch path once without capture (cudaFuncSetAttribute is not capturable)
    if (ctx->use_graphs && !ctx->graphs_built && nsteps > 2) {
        enqueue_step(ctx, 0, nullptr);
        --nsteps;
        int r = build_graphs(ctx);
        if (r) return r;
    }
    for (int n = 0; n < nsteps; ++n) {
        const bool last = (n == nsteps - 1);
        if (last && want_diag) enqueue_step(ctx, 1, nullptr);
        else if (ctx->graphs_built) { CK(cudaGraphLaunch(ctx->gexec[ctx->cur], ctx->stream)); ctx->cur ^= 1; }
        else enqueue_step(ctx, 0, nullptr);
    }
    CK(cudaGetLastError());
    if (want_diag && asked > 0) ctx->diag_valid = true;
    else if (asked > 0) ctx->diag_valid = false;
    return read_status(ctx, status_out);
}

// Ensemble of small tissues (last bullet: 228-10 k cell tissues are launch-latency bound and run as parameter
// ensembles): n independent contexts of ONE device advance nsteps timesteps each inside ONE CUDA graph — every member's
// kernels on the member's own stream, forked from and joined to the leader's (ctxs[0]) stream, so the members' small
// kernels share the GPU and the host pays one launch for n x nsteps timesteps.  Each member runs exactly the kernels
// betse_step would launch for it: results are bit-identical to stepping it alone.
extern "C" int betse_ensemble_step(betse_ctx** ctxs, int n, int nsteps, int launches, uint32_t* status_out, float* device_ms)
{
    if (!ctxs || n <= 0 || !ctxs[0] || nsteps <= 0 || launches < 0) return 2;
    betse_ctx* ctx = ctxs[0];                     // leader: owns the graph, its stream carries the fork and the join
    CK(cudaSetDevice(ctx->device));
    unsigned long long epochs = 0;
    for (int j = 0; j < n; ++j) {
        betse_ctx* c = ctxs[j];
        if (!c) return fail(ctx, "betse_ensemble_step: null member");
        if (c->device != ctx->device) return fail(ctx, "betse_ensemble_step: members must live on one device");
        chan_layout_sync(c);
        if (c->cur != ctx->cur) return fail(ctx, "betse_ensemble_step: members are not in lockstep (step them only through the ensemble)");
        if (c->X.n_nbr > 0) return fail(ctx, "betse_ensemble_step: a member is part of a decomposed tissue");
        if (c->phi_lag || c->noise_on || c->P.polar || c->need_emc)
            return fail(ctx, "betse_ensemble_step: boundary-voltage ramps, dynamic noise and per-step diagnostics are stepped through betse_step");
        epochs += c->graph_epoch;
    }
    if (!ctx->ens_ev[0]) { CK(cudaEventCreate(&ctx->ens_ev[0])); CK(cudaEventCreate(&ctx->ens_ev[1])); }
    betse_ctx::EnsGraph* g = nullptr;
    for (auto& e : ctx->ens)
        if (e.nsteps == nsteps && e.cur == ctx->cur && e.epochs == epochs && (int)e.members.size() == n &&
            std::equal(e.members.begin(), e.members.end(), ctxs)) g = &e;
    if (!g) {
        // (the caller has stepped every member once through betse_step before the first capture: kernels are loaded)
        static thread_local std::vector<cudaEvent_t> evs;
        for (int j = 0; j < n; ++j) CK(cudaStreamSynchronize(ctxs[j]->stream));
        std::vector<int> saved(n);
        for (int j = 0; j < n; ++j) saved[j] = ctxs[j]->cur;
        while ((int)evs.size() < n) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); evs.push_back(e); }
        cudaGraph_t graph;
        CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        CK(cudaEventRecord(evs[0], ctx->stream));
        for (int j = 1; j < n; ++j) CK(cudaStreamWaitEvent(ctxs[j]->stream, evs[0], 0));
        for (int j = 0; j < n; ++j)
            for (int s = 0; s < nsteps; ++s) enqueue_step(ctxs[j], 0, nullptr);
        for (int j = 1; j < n; ++j) {
            CK(cudaEventRecord(evs[j], ctxs[j]->stream));
            CK(cudaStreamWaitEvent(ctx->stream, evs[j], 0));
        }
        CK(cudaStreamEndCapture(ctx->stream, &graph));
        for (int j = 0; j < n; ++j) ctxs[j]->cur = saved[j];
        betse_ctx::EnsGraph e;
        e.members.assign(ctxs, ctxs + n); e.nsteps = nsteps; e.cur = ctx->cur; e.epochs = epochs; e.exec = nullptr;
        CK(cudaGraphInstantiate(&e.exec, graph, 0));
        CK(cudaGraphDestroy(graph));
        ctx->ens.push_back(e);
        g = &ctx->ens.back();
    }
    CK(cudaEventRecord(ctx->ens_ev[0], ctx->stream));
    for (int l = 0; l < launches; ++l) {
        if (l > 0 && (nsteps & 1)) {
            // odd steps per launch flip the buffer parity: the other parity's graph
            for (int j = 0; j < n; ++j) ctxs[j]->cur ^= 1;
            int r = betse_ensemble_step(ctxs, n, nsteps, 0, nullptr, nullptr);
            for (int j = 0; j < n; ++j) ctxs[j]->cur ^= 1;
            if (r) return r;
        }
        const int want = ctx->cur ^ ((l & 1) && (nsteps & 1) ? 1 : 0);
        betse_ctx::EnsGraph* gl = nullptr;
        for (auto& e : ctx->ens)
            if (e.nsteps == nsteps && e.cur == want && e.epochs == epochs && (int)e.members.size() == n &&
                std::equal(e.members.begin(), e.members.end(), ctxs)) gl = &e;
        if (!gl) return fail(ctx, "betse_ensemble_step: graph of the other buffer parity is missing");
        CK(cudaGraphLaunch(gl->exec, ctx->stream));
    }
    CK(cudaEventRecord(ctx->ens_ev[1], ctx->stream));
    if ((launches * nsteps) & 1) for (int j = 0; j < n; ++j) ctxs[j]->cur ^= 1;
    for (int j = 0; j < n; ++j) if (launches > 0) ctxs[j]->diag_valid = false;
    CK(cudaGetLastError());
    if (launches == 0) return 0;
    CK(cudaEventSynchronize(ctx->ens_ev[1]));
    if (device_ms) CK(cudaEventElapsedTime(device_ms, ctx->ens_ev[0], ctx->ens_ev[1]));
    uint32_t all = 0;
    for (int j = 0; j < n; ++j) {
        uint32_t st = 0;
        betse_ctx* c = ctxs[j];
        { betse_ctx* ctx = c; int r = read_status(ctx, &st); if (r) return r; }
        if (status_out) status_out[j] = st;
        all |= st;
    }
    (void)all;
    return 0;
}

// ---------------------------------------------------------------------------- fast (equivalent-circuit) solver
extern "C" int betse_fast_setup(betse_ctx* ctx, const betse_fast_host* s)
{
    if (!ctx || !s) return 2;
    if (!s->vm_ave || !s->gjopen || !s->G_Leak || !s->E_Leak || !s->G_gj || !s->sigma_cell)
        return fail(ctx, "betse_fast_setup: vm_ave, gjopen, G_Leak, E_Leak, G_gj and sigma_cell are required");
    if (ctx->X.n_nbr > 0) return fail(ctx, "the fast solver on a domain-decomposed tissue is not implemented");
    if (ctx->net_on[0] || ctx->net_on[1]) return fail(ctx, "the fast solver with network substances is not implemented");
    if (!ctx->chans.empty() && !ctx->fast_chan_on)
        return fail(ctx, "the fast solver with channels needs betse_fast_set_channels (conductance factors, reversal potentials)");
    if (!ctx->hp.v_sensitive_gj && !ctx->A.gj_w) return fail(ctx, "static gap junctions need gj_default_weights");
    CK(cudaSetDevice(ctx->device));
    const int C = ctx->C, Mo = ctx->Mo;
    KFast& Fz = ctx->fast;
    int r;
    if (!ctx->fast_on) {
        memset(&Fz, 0, sizeof Fz);
        for (int b = 0; b < 2; ++b) if ((r = dev_alloc(ctx, &Fz.vm_ave[b], (size_t)C))) return r;
        if ((r = dev_alloc(ctx, (double**)&Fz.G_Leak, (size_t)C))) return r;
        if ((r = dev_alloc(ctx, (double**)&Fz.E_Leak, (size_t)C))) return r;
        if ((r = dev_alloc(ctx, (double**)&Fz.G_gj, (size_t)C))) return r;
        if ((r = dev_alloc(ctx, (double**)&Fz.sigma_cell, (size_t)C))) return r;
        if ((r = dev_alloc(ctx, &Fz.vgj, (size_t)Mo))) return r;
        if ((r = dev_alloc(ctx, &Fz.Jn, (size_t)Mo))) return r;
        if ((r = dev_alloc(ctx, &Fz.Emx, (size_t)Mo))) return r;
        if ((r = dev_alloc(ctx, &Fz.Emy, (size_t)Mo))) return r;
        if ((r = dev_alloc(ctx, &Fz.J_cell_x, (size_t)C))) return r;
        if ((r = dev_alloc(ctx, &Fz.J_cell_y, (size_t)C))) return r;
        if ((r = dev_alloc(ctx, &Fz.E_cell_x, (size_t)C))) return r;
        if ((r = dev_alloc(ctx, &Fz.E_cell_y, (size_t)C))) return r;
        ctx->fast_on = true;
    }
    destroy_graphs(ctx);
    ctx->fast_cur = 0;
#define UPF(dst, src, n) { int xr_ = xfer(ctx, (void*)(dst), (src), (size_t)(n) * sizeof(double), cudaMemcpyHostToDevice); if (xr_) return xr_; }
    UPF(Fz.vm_ave[0], s->vm_ave, C);
    UPF(ctx->A.gjopen, s->gjopen, Mo);
    UPF(Fz.G_Leak, s->G_Leak, C);
    UPF(Fz.E_Leak, s->E_Leak, C);
    UPF(Fz.G_gj, s->G_gj, C);
    UPF(Fz.sigma_cell, s->sigma_cell, C);
    if ((r = opt_array(ctx, &Fz.extra_J, s->extra_J_mem, (size_t)Mo))) return r;
    if (ctx->fast_chan_on && !ctx->chans.empty()) {
        // clear_run_loop zeroes extra_J_mem every step and the channels rebuild it (sim.py:1505-1512, networks.py:3272)
        if (!ctx->fast_extraJ && (r = dev_alloc(ctx, &ctx->fast_extraJ, (size_t)Mo))) return r;
        Fz.extra_J = ctx->fast_extraJ;
    }
#undef UPF
    double mean = 0.0;                                    // sim.sigma_cell.mean(): NumPy's pairwise sum is within an ulp
    for (int c = 0; c < C; ++c) mean += s->sigma_cell[c];
    mean /= (double)C;
    Fz.sm = 0.1 * mean;
    Fz.dt_cm = ctx->hp.dt * (1.0 / ctx->hp.cm);            // p.dt*(1/p.cm), sim.py:1561
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// one iteration of the fast loop body: the channels' currents from the potentials the step starts with, then the circuit
static void enqueue_fast(betse_ctx* ctx, int cur)
{
    if (ctx->fast_chan_on && !ctx->chans.empty())
        launch_fast_chan(ctx->P, ctx->A, ctx->chans.data(), (int)ctx->chans.size(), ctx->fast_coef, ctx->fast_revE,
                         ctx->fast.vm_ave[cur], ctx->fast_extraJ, ctx->stream);
    launch_fast(ctx->P, ctx->A, ctx->fast, cur, ctx->stream);
}

extern "C" int betse_fast_set_channels(betse_ctx* ctx, const double* cond_coef, const double* rev_E)
{
    if (!ctx || !cond_coef || !rev_E) return 2;
    CK(cudaSetDevice(ctx->device));
    for (const KChan& ch : ctx->chans)
        if (ch.mod_prog >= 0) return fail(ctx, "the fast solver with network-modulated channels is not implemented");
    if (ctx->chan_cell_mode) { chan_expand_all(ctx); ctx->chan_cell_mode = false; }     // this solver keeps the gates per membrane
    for (int i = 0; i < ctx->I; ++i) { ctx->fast_coef[i] = cond_coef[i]; ctx->fast_revE[i] = rev_E[i]; }
    ctx->fast_chan_on = true;
    if (ctx->fast_graph) { cudaGraphExecDestroy(ctx->fast_graph); ctx->fast_graph = nullptr; }
    return 0;
}

extern "C" int betse_fast_step(betse_ctx* ctx, int nsteps, int flags, uint32_t* status_out)
{
    if (!ctx || nsteps < 0) return 2;
    if (!ctx->fast_on) return fail(ctx, "betse_fast_step before betse_fast_setup");
    CK(cudaSetDevice(ctx->device));
    int n = nsteps;
    if (n >= 8 && ctx->use_graphs) {
        if (!ctx->fast_graph) {
            enqueue_fast(ctx, ctx->fast_cur);      // loads the kernel outside the capture
            ctx->fast_cur ^= 1; --n;
            cudaGraph_t g;
            const int c0 = ctx->fast_cur;
            CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            enqueue_fast(ctx, c0);
            enqueue_fast(ctx, c0 ^ 1);
            CK(cudaStreamEndCapture(ctx->stream, &g));
            CK(cudaGraphInstantiate(&ctx->fast_graph, g, 0));
            CK(cudaGraphDestroy(g));
            ctx->fast_graph_cur = c0;
        }
        if (ctx->fast_cur != ctx->fast_graph_cur && n > 0) { enqueue_fast(ctx, ctx->fast_cur); ctx->fast_cur ^= 1; --n; }
        for (; n >= 2; n -= 2) CK(cudaGraphLaunch(ctx->fast_graph, ctx->stream));
    }
    for (; n > 0; --n) { enqueue_fast(ctx, ctx->fast_cur); ctx->fast_cur ^= 1; }
    if (flags & BETSE_STEP_DIAG) launch_fast_diag(ctx->P, ctx->A, ctx->fast, ctx->stream);     // (nsteps == 0: of the last step run)
    CK(cudaGetLastError());
    return read_status(ctx, status_out);
}

extern "C" int betse_fast_download(betse_ctx* ctx, betse_fast_host* out)
{
    if (!ctx || !out) return 2;
    if (!ctx->fast_on) return fail(ctx, "betse_fast_download before betse_fast_setup");
    CK(cudaSetDevice(ctx->device));
    const int C = ctx->C, Mo = ctx->Mo;
    const KFast& Fz = ctx->fast;
#define DNF(dst, src, n) if (dst) { int xr_ = xfer(ctx, (void*)(dst), (src), (size_t)(n) * sizeof(double), cudaMemcpyDeviceToHost); if (xr_) return xr_; }
    DNF(out->vm_ave, Fz.vm_ave[ctx->fast_cur], C);
    DNF(out->gjopen, ctx->A.gjopen, Mo);
    DNF(out->vgj, Fz.vgj, Mo);
    DNF(out->Jn, Fz.Jn, Mo);
    DNF(out->Emx, Fz.Emx, Mo);
    DNF(out->Emy, Fz.Emy, Mo);
    DNF(out->J_cell_x, Fz.J_cell_x, C);
    DNF(out->J_cell_y, Fz.J_cell_y, C);
    DNF(out->E_cell_x, Fz.E_cell_x, C);
    DNF(out->E_cell_y, Fz.E_cell_y, C);
#undef DNF
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int betse_update_v_phase(betse_ctx* ctx, int phase)
{
    if (!ctx || phase < 0 || phase > 1) return 2;
    CK(cudaSetDevice(ctx->device));
    if (phase == 0) {
        launch_cell_charge(ctx->P, ctx->A, ctx->Co, ctx->cur, ctx->stream);
        if (ctx->hp.is_ecm) launch_envacc(ctx->I, ctx->P, ctx->A, ctx->E, ctx->cur, 0, ctx->stream);
    } else if (ctx->hp.is_ecm) launch_field(ctx->P, ctx->A, ctx->ny, ctx->nx, ctx->stream);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int betse_update_v(betse_ctx* ctx)
{
    if (!ctx) return 2;
    int r;
    if (ctx->P.polar) return fail(ctx, "betse_update_v with cell_polarizability != 0: upload the Vmem of the loop entry instead (sim.py:1041 has run)");
    if ((r = betse_update_v_phase(ctx, 0))) return r;
    if (ctx->X.n_nbr > 0) {
        // ghost-cell Vmem / concentrations and the env halo rows of the CURRENT buffers
        launch_xchg(ctx->P, ctx->A, ctx->X, BETSE_XCHG_X1, ctx->cur, BETSE_XCHG_PUSH | BETSE_XCHG_WAIT, ctx->flux_is_ell, ctx->stream);
        launch_xchg(ctx->P, ctx->A, ctx->X, BETSE_XCHG_X2, ctx->cur, BETSE_XCHG_PUSH | BETSE_XCHG_WAIT, ctx->flux_is_ell, ctx->stream);
    }
    if ((r = betse_update_v_phase(ctx, 1))) return r;
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int betse_step_phase(betse_ctx* ctx, int phase, int flags)
{
    if (!ctx || phase < 0 || phase > 2) return 2;
    CK(cudaSetDevice(ctx->device));
    const int diag = (flags & BETSE_STEP_DIAG) ? 1 : 0;
    if (diag || ctx->P.polar) { int r = ensure_diag_buffers(ctx); if (r) return r; }
    if (phase == 0) chan_layout_sync(ctx);
    enqueue_phase(ctx, phase, diag, nullptr);
    CK(cudaGetLastError());
    if (phase == 2) ctx->diag_valid = diag != 0;
    return 0;
}

extern "C" int betse_sync(betse_ctx* ctx, uint32_t* status_out)
{
    if (!ctx) return 2;
    CK(cudaSetDevice(ctx->device));
    return read_status(ctx, status_out);
}

extern "C" int betse_stream(betse_ctx* ctx, void** cuda_stream)
{
    if (!ctx || !cuda_stream) return 2;
    *cuda_stream = (void*)ctx->stream;
    return 0;
}

extern "C" int betse_step_profile(betse_ctx* ctx, int nsteps, float* total_ms,
                                  float kernel_ms[BETSE_NKERNELS], int kernel_launches[BETSE_NKERNELS])
{
    if (!ctx || nsteps <= 0) return 2;
    CK(cudaSetDevice(ctx->device));
    if (ctx->P.polar) { int r = ensure_diag_buffers(ctx); if (r) return r; }
    if (ctx->xwait) { int r = prepare_xfuse(ctx); if (r) return r; }
    chan_layout_sync(ctx);
    if (!ctx->ev_init) {
        for (auto& e : ctx->ev) CK(cudaEventCreate(&e));
        ctx->ev_init = true;
    }
    double acc[BETSE_NKERNELS] = {0};
    int cnt[BETSE_NKERNELS] = {0};
    const bool ecm = ctx->hp.is_ecm != 0;
    cudaEvent_t t0 = ctx->ev[14], t1 = ctx->ev[15];
    // (a) whole-run time, back-to-back launches exactly as betse_step issues them
    if (ctx->use_graphs && !ctx->graphs_built) { enqueue_step(ctx, 0, nullptr); int r = build_graphs(ctx); if (r) return r; }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventRecord(t0, ctx->stream));
    for (int n = 0; n < nsteps; ++n) {
        if (ctx->graphs_built) { CK(cudaGraphLaunch(ctx->gexec[ctx->cur], ctx->stream)); ctx->cur ^= 1; }
        else enqueue_step(ctx, 0, nullptr);
    }
    CK(cudaEventRecord(t1, ctx->stream));
    CK(cudaEventSynchronize(t1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, t0, t1));
    if (total_ms) *total_ms = ms;
    // (b) per-kernel durations: same steps again with an event after every kernel
    if (kernel_ms) {
        const int reps = nsteps < 20 ? nsteps : 20;
        for (int n = 0; n < reps; ++n) {
            enqueue_step(ctx, 0, ctx->ev);
            CK(cudaEventSynchronize(ctx->ev[6]));
            float d;
            if (ecm) {
                CK(cudaEventElapsedTime(&d, ctx->ev[0], ctx->ev[1])); acc[K_ION] += d; cnt[K_ION]++;
                if (ctx->hp.sharpness < 1.0) { CK(cudaEventElapsedTime(&d, ctx->ev[1], ctx->ev[2])); acc[K_SMOOTH] += d; cnt[K_SMOOTH]++; }
            }
            CK(cudaEventElapsedTime(&d, ctx->ev[2], ctx->ev[3])); acc[K_MEM] += d; cnt[K_MEM]++;
            if (ctx->X.n_nbr > 0) {      // both exchange kernels (push + wait for the neighbours)
                CK(cudaEventElapsedTime(&d, ctx->ev[3], ctx->ev[7])); acc[K_XCHG] += d;
                CK(cudaEventElapsedTime(&d, ctx->ev[4], ctx->ev[8])); acc[K_XCHG] += d; cnt[K_XCHG]++;
            }
            CK(cudaEventElapsedTime(&d, ctx->ev[7], ctx->ev[4]));
            if (ecm) { acc[K_ENVACC] += d; cnt[K_ENVACC]++; } else { acc[K_ENVMIX] += d; cnt[K_ENVMIX]++; }
            if (ecm) { CK(cudaEventElapsedTime(&d, ctx->ev[8], ctx->ev[5])); acc[K_FIELD] += d; cnt[K_FIELD]++; }
        }
        for (int k = 0; k < BETSE_NKERNELS; ++k) {
            kernel_ms[k] = cnt[k] ? (float)(acc[k] / cnt[k]) : 0.f;
            if (kernel_launches) kernel_launches[k] = cnt[k] ? (k == K_XCHG ? (ctx->last_step_fused ? 0 : 2) : 1) : 0;   // launches per step
        }
    }
    ctx->diag_valid = false;
    uint32_t st;
    return read_status(ctx, &st);
}

extern "C" int betse_download_sample(betse_ctx* ctx, betse_state_host* s)
{
    if (!ctx || !s) return 2;
    CK(cudaSetDevice(ctx->device));
    KArrays& A = ctx->A;
    const int C = ctx->C, Mo = ctx->Mo, E = ctx->E, I = ctx->I;
    const size_t IC = (size_t)I * C, IE = (size_t)I * E, IM = (size_t)I * Mo;
    const int cur = ctx->cur;
    cudaStream_t st = ctx->stream;
#define DN(dst, src, n) if (dst) { if (!(src)) return fail(ctx, "download of " #dst ": not available"); \
        int xr_ = xfer(ctx, (dst), (src), (n) * sizeof(double), cudaMemcpyDeviceToHost); if (xr_) return xr_; }
    DN(s->cc_cells, A.cc_cells, IC);
    DN(s->cc_at_mem_cell, A.cc_mid[cur], IC);
    if (ctx->hp.is_ecm) {
        DN(s->cc_env, A.cc_env[cur], IE);
        DN(s->D_env_eff, A.Denv, IE);
        DN(s->E_env_x, A.E_x, E);
        DN(s->E_env_y, A.E_y, E);
        DN(s->v_env, A.v_env, E);
        DN(s->rho_env, A.rho_env, E);
    } else if (s->E_env_x || s->E_env_y || s->v_env) {
        // the local field potential and its field: diagnostics of the last sampled step (ion_current.py:116-158)
        if (!ctx->hh_on || !ctx->denv_w || !ctx->diag_valid)
            return fail(ctx, "v_env / E_env without extracellular spaces: upload D_env_weight and run the step with BETSE_STEP_DIAG");
        DN(s->E_env_x, A.E_x, E);
        DN(s->E_env_y, A.E_y, E);
        DN(s->v_env, A.v_env, E);
    }
    if (s->vm || s->vm_ave) {
        if (ctx->P.polar) CK(cudaMemcpyAsync(A.vm_mem, A.vm_pol[cur], (size_t)Mo * sizeof(double), cudaMemcpyDeviceToDevice, st));
        launch_expand_vm(ctx->P, A, Mo, ctx->Co, cur, st);
        DN(s->vm, A.vm_mem, Mo);
        DN(s->vm_ave, A.vm_ave, C);
    }
    if (s->Phi_b) {                                    // sim.Phi_b of the last update_V (ion_current.py:114, 171)
        if (A.phi_b) CK(cudaMemcpyAsync(s->Phi_b, A.phi_b, (size_t)E * sizeof(double), cudaMemcpyDeviceToHost, st));
        else memset(s->Phi_b, 0, (size_t)E * sizeof(double));
    }
    // what the network handlers published last (networks.py:2971-2977): the host's update_V of the NEXT phase reads them
    if (s->extra_rho_cells) {
        if (A.extra_rho_cells) CK(cudaMemcpyAsync(s->extra_rho_cells, A.extra_rho_cells, (size_t)C * sizeof(double), cudaMemcpyDeviceToHost, st));
        else memset(s->extra_rho_cells, 0, (size_t)C * sizeof(double));
    }
    if (s->extra_rho_env) {
        if (A.extra_rho_env) CK(cudaMemcpyAsync(s->extra_rho_env, A.extra_rho_env, (size_t)E * sizeof(double), cudaMemcpyDeviceToHost, st));
        else memset(s->extra_rho_env, 0, (size_t)E * sizeof(double));
    }
    if (s->extra_J_mem) {
        const double* src = (ctx->P.chan_charge && A.chanJ) ? A.chanJ : A.extra_J_mem;
        if (src) CK(cudaMemcpyAsync(s->extra_J_mem, src, (size_t)Mo * sizeof(double), cudaMemcpyDeviceToHost, st));
        else memset(s->extra_J_mem, 0, (size_t)Mo * sizeof(double));
    }
    DN(s->gjopen, A.gjopen, Mo);
    DN(s->Dm_cells, A.Dm, IM);
    DN(s->rho_cells, A.rho_cells, C);
    if (s->cenv_uniform) CK(cudaMemcpyAsync(s->cenv_uniform, A.cenv_u + cur * 8, I * sizeof(double), cudaMemcpyDeviceToHost, st));
    const bool any_diag = s->fluxes_mem || s->fluxes_gj || s->fluxes_env_x || s->fluxes_env_y || s->rate_NaKATP ||
                          s->Jmem || s->Jgj || s->Jn || s->I_mem || s->Jc || s->Emc || s->dvm || s->J_cell_x ||
                          s->J_cell_y || s->E_cell_x || s->E_cell_y || s->sigma_cell || s->E_gj_x || s->E_gj_y || s->J_env_x || s->J_env_y || s->B_field || s->Jtx || s->Jty;
    if (any_diag) {
        if (!ctx->diag_valid) return fail(ctx, "diagnostics requested but the last step was not run with BETSE_STEP_DIAG");
        DN(s->fluxes_mem, A.fl_mem, IM);
        DN(s->fluxes_gj, A.fl_gj, IM);
        if (ctx->hp.is_ecm) { DN(s->fluxes_env_x, A.fl_env_x, IE); DN(s->fluxes_env_y, A.fl_env_y, IE); }
        DN(s->rate_NaKATP, A.rate_NaK, Mo);
        DN(s->Jmem, A.Jmem, Mo); DN(s->Jgj, A.Jgj, Mo); DN(s->Jn, A.Jn, Mo); DN(s->I_mem, A.I_mem, Mo);
        DN(s->Jc, A.Jc, Mo); DN(s->Emc, A.Emc, Mo); DN(s->dvm, A.dvm, Mo);
        DN(s->J_cell_x, A.J_cell_x, C); DN(s->J_cell_y, A.J_cell_y, C);
        DN(s->E_cell_x, A.E_cell_x, C); DN(s->E_cell_y, A.E_cell_y, C);
        DN(s->sigma_cell, A.sigma_cell, C);
        DN(s->E_gj_x, A.E_gj_x, Mo); DN(s->E_gj_y, A.E_gj_y, Mo);
        if (s->J_env_x || s->J_env_y || s->B_field || s->Jtx || s->Jty) {
            if (!ctx->hh_on) return fail(ctx, "J_env / B_field / Jtx: the Helmholtz-Hodge diagnostics need an undivided tissue (and sim.D_env_weight without extracellular spaces)");
            DN(s->J_env_x, ctx->hh.J_env_x, E); DN(s->J_env_y, ctx->hh.J_env_y, E); DN(s->B_field, ctx->hh.B_field, E);
            DN(s->Jtx, ctx->hh.Jtx, E); DN(s->Jty, ctx->hh.Jty, E);
        }
    }
#undef DN
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int betse_set_channels(betse_ctx* ctx, int n, const betse_channel* chs, int affect_charge)
{
    if (!ctx || n < 0 || (n > 0 && !chs)) return 2;
    CK(cudaSetDevice(ctx->device));
    if (n > 0 && ctx->hp.fast_update_ecm) return fail(ctx, "channels with fast_update_ecm are not implemented");
    if (n > 0 && ctx->X.n_nbr > 0) return fail(ctx, "channels on a domain-decomposed tissue are not implemented");
    KArrays& A = ctx->A;
    const int Mo = ctx->Mo;
    int r;
    ctx->chans.clear();
    destroy_graphs(ctx);
    ctx->P.defer = (n > 0 || ctx->net_on[0] || ctx->net_on[1]) ? 1 : 0;
    ctx->P.chan_charge = ((n > 0 && affect_charge) || ctx->net_affect[0] || ctx->net_affect[1]) ? 1 : 0;
    if (!ctx->P.defer) return 0;
    if (!A.dsum_m) {
        if ((r = dev_alloc(ctx, &A.dsum_m, (size_t)ctx->I * ctx->C))) return r;
        if ((r = dev_alloc(ctx, &A.dsum_g, (size_t)ctx->I * ctx->C))) return r;
        if ((r = dev_alloc(ctx, &A.chan_slots, (size_t)ctx->n_slots))) return r;
    if ((r = dev_alloc(ctx, &A.chan_part, (size_t)ctx->n_tiles))) return r;
        if ((r = dev_alloc(ctx, &A.chanJ, (size_t)Mo))) return r;
    }
    // per-cell path (channels.cu:k_chan_cell): every channel on every membrane, unmodulated, its initial gates uniform
    // within every cell (what the reference's initial conditions give: V is the cell's Vmem, vg_na.py:75-88)
    bool cellp = chan_cell_possible(ctx);
    { const char* e = getenv("BETSE_CHAN_CELL"); if (e && e[0] == '0') cellp = false; }      // A/B: one kernel pair per channel
    for (int k = 0; k < n && cellp; ++k) {
        const betse_channel& c = chs[k];
        if (c.same_gates) continue;
        cellp = !c.target_mask && c.mod_prog < 0 && c.m0 && c.h0;
        for (int cc = 0; cc < ctx->Co && cellp; ++cc)
            for (int m = ctx->h_cmp[cc] + 1; m < ctx->h_cmp[cc + 1]; ++m)
                if (c.m0[m] != c.m0[ctx->h_cmp[cc]] || c.h0[m] != c.h0[ctx->h_cmp[cc]]) { cellp = false; break; }
    }
    ctx->chan_cell_mode = cellp && n > 0;
    if (ctx->chan_cell_mode && !ctx->chan_ell)
        if ((r = dev_alloc(ctx, &ctx->chan_ell, (size_t)ctx->P.ell_rows * 32 * ctx->I))) return r;
    for (int k = 0; k < n; ++k) {
        const betse_channel& c = chs[k];
        if (c.ion < 0 || c.ion >= ctx->I) return fail(ctx, "channel ion index out of range");
        if (c.mpower < 0 || c.mpower > 8 || c.hpower < 0 || c.hpower > 8) return fail(ctx, "channel gate powers must be in [0,8]");
        if (c.same_gates) {
            // a further ion of the previous entry's channel (channel_core.ions[j], rel_perm[j]; networks.py:3158-3203)
            if (k == 0) return fail(ctx, "channel entry 0 cannot share the gates of a previous entry");
            KChan d = ctx->chans.back();
            d.ion = c.ion; d.rel_perm = c.rel_perm; d.frozen = 1;
            ctx->chans.push_back(d);
            continue;
        }
        if (!c.m0 || !c.h0) return fail(ctx, "channel needs initial gate states m0/h0");
        KChan d;
        memset(&d, 0, sizeof d);
        d.ion = c.ion; d.mpow = c.mpower; d.hpow = c.hpower;
        for (int q = 0; q < 4; ++q) {
            if (c.kind[q] < 0 || c.kind[q] > 2) return fail(ctx, "channel quantity kind must be 0, 1 or 2");
            d.kind[q] = c.kind[q];
            if (c.a[q].type < 0 || c.a[q].type > 9 || c.b[q].type < 0 || c.b[q].type > 9) return fail(ctx, "unknown gate term type");
            d.a[q].type = c.a[q].type; d.b[q].type = c.b[q].type;
            for (int j = 0; j < 4; ++j) { d.a[q].p[j] = c.a[q].p[j]; d.b[q].p[j] = c.b[q].p[j]; }
        }
        d.handler = c.handler; d.mod_prog = c.mod_prog;
        if (c.handler < 0 || c.handler > 1) return fail(ctx, "channel handler must be 0 or 1");
        if (c.mod_prog >= 0 && !ctx->net_on[c.handler]) return fail(ctx, "channel modulator without its network: call betse_set_network first");
        if (c.mod_prog >= 0 && (c.mod_prog < ctx->nets[c.handler].n_rates || c.mod_prog >= ctx->net_nprog[c.handler]))
            return fail(ctx, "channel mod_prog is not a membrane-zone program of its network");
        d.dt_tu = ctx->hp.dt * c.time_unit;
        d.maxDm = c.max_Dm; d.rel_perm = c.rel_perm; d.shift = c.v_shift;
        if (c.target_mask) { if ((r = dev_upload(ctx, (unsigned char**)&d.mask, c.target_mask, Mo))) return r; }
        if ((r = dev_upload(ctx, &d.m, c.m0, Mo))) return r;
        if ((r = dev_upload(ctx, &d.h, c.h0, Mo))) return r;
        if ((r = dev_alloc(ctx, &d.P, Mo))) return r;
        if ((r = dev_alloc(ctx, &d.flux, Mo))) return r;
        if ((r = dev_alloc(ctx, &d.D, Mo))) return r;
        if (ctx->chan_cell_mode) {
            std::vector<double> mc((size_t)ctx->C, 0.0), hc((size_t)ctx->C, 0.0);
            for (int cc = 0; cc < ctx->Co; ++cc)
                if (ctx->h_cmp[cc + 1] > ctx->h_cmp[cc]) { mc[cc] = c.m0[ctx->h_cmp[cc]]; hc[cc] = c.h0[ctx->h_cmp[cc]]; }
            if ((r = dev_upload(ctx, &d.mc, mc.data(), (size_t)ctx->C))) return r;
            if ((r = dev_upload(ctx, &d.hc, hc.data(), (size_t)ctx->C))) return r;
            if ((r = dev_alloc(ctx, &d.Pc, (size_t)ctx->C))) return r;
            if ((r = dev_alloc(ctx, &d.Dc, (size_t)ctx->C))) return r;
            if ((r = dev_alloc(ctx, &d.fell, (size_t)ctx->P.ell_rows * 32))) return r;
        }
        ctx->chans.push_back(d);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    publish_affect(ctx);
    return 0;
}

// What the per-cell channel path needs from the tissue (the channels' own conditions are checked in betse_set_channels):
// Vmem per cell (no polarizability, no boundary potential), extracellular spaces, the cell pack over consecutive cells, no
// transporter-nudged membrane values (their refresh follows every single channel).
static bool chan_cell_possible(betse_ctx* ctx)
{
    const KArrays& A = ctx->A;
    if (!ctx->hp.is_ecm || ctx->P.polar || ctx->P.has_phi || ctx->phi_lag || !A.cpack || A.pcell || !A.slot_off || !ctx->mem_ell) return false;
    if (ctx->X.n_nbr > 0 || ctx->P.n_blocks <= 0) return false;
    for (int h = 0; h < 2; ++h) if (ctx->net_on[h] && ctx->nets[h].tw) return false;
    return true;
}

// gate state of every channel back into the per-membrane arrays (betse_channel_state; leaving the per-cell path)
static void chan_expand_all(betse_ctx* ctx)
{
    for (const KChan& ch : ctx->chans)
        launch_chan_expand(ch, ctx->A.mem_to_cells, ctx->mem_ell, ctx->Mo, ctx->I, ctx->stream);
}

// Before anything is enqueued: the tissue may have left the conditions of the per-cell path since betse_set_channels (a
// boundary potential switched on, a network with transporters set afterwards) — the state moves to the per-membrane arrays
// and k_chan takes over, for good.
static void chan_layout_sync(betse_ctx* ctx)
{
    if (!ctx->chan_cell_mode || chan_cell_possible(ctx)) return;
    chan_expand_all(ctx);
    ctx->chan_cell_mode = false;
    destroy_graphs(ctx);
}

extern "C" int betse_channel_state(betse_ctx* ctx, int k, double* m, double* h, double* P, double* flux, double* DChan)
{
    if (!ctx) return 2;
    CK(cudaSetDevice(ctx->device));
    if (k < 0 || k >= (int)ctx->chans.size()) return fail(ctx, "channel index out of range");
    const KChan& d = ctx->chans[k];
    const size_t nb = (size_t)ctx->Mo * sizeof(double);
    if (ctx->chan_cell_mode) launch_chan_expand(d, ctx->A.mem_to_cells, ctx->mem_ell, ctx->Mo, ctx->I, ctx->stream);
    if (m) { int xr = xfer(ctx, m, d.m, nb, cudaMemcpyDeviceToHost); if (xr) return xr; }
    if (h) { int xr = xfer(ctx, h, d.h, nb, cudaMemcpyDeviceToHost); if (xr) return xr; }
    if (P) { int xr = xfer(ctx, P, d.P, nb, cudaMemcpyDeviceToHost); if (xr) return xr; }
    if (flux) { int xr = xfer(ctx, flux, d.flux, nb, cudaMemcpyDeviceToHost); if (xr) return xr; }
    if (DChan) { int xr = xfer(ctx, DChan, d.D, nb, cudaMemcpyDeviceToHost); if (xr) return xr; }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// sim.extra_rho_cells / sim.extra_rho_env are the arrays of the LAST enabled handler once its run_loop has published them
// (networks.py:2971-2977): zeros when that handler has no (charged) substances
static void publish_affect(betse_ctx* ctx)
{
    int hl = -1;
    for (int h = 0; h < 2; ++h) {
        bool enabled = ctx->net_on[h];
        for (const KChan& ch : ctx->chans) enabled = enabled || ch.handler == h;
        if (enabled) hl = h;
    }
    if (hl < 0 || !ctx->P.chan_charge) return;
    if (ctx->net_on[hl] && ctx->net_affect[hl]) {
        ctx->A.extra_rho_cells = ctx->nets[hl].rho_cells;
        ctx->A.extra_rho_env = ctx->nets[hl].rho_env;        // null: the handler has nothing outside the cells
    } else {
        ctx->A.extra_rho_cells = nullptr;
        ctx->A.extra_rho_env = nullptr;
    }
}

static int ensure_defer_buffers(betse_ctx* ctx)
{
    KArrays& A = ctx->A;
    int r;
    if (A.dsum_m) return 0;
    if ((r = dev_alloc(ctx, &A.dsum_m, (size_t)ctx->I * ctx->C))) return r;
    if ((r = dev_alloc(ctx, &A.dsum_g, (size_t)ctx->I * ctx->C))) return r;
    if ((r = dev_alloc(ctx, &A.chan_slots, (size_t)ctx->n_slots))) return r;
    if ((r = dev_alloc(ctx, &A.chan_part, (size_t)ctx->n_tiles))) return r;
    if ((r = dev_alloc(ctx, &A.chanJ, (size_t)ctx->Mo))) return r;
    return 0;
}

extern "C" int betse_set_network(betse_ctx* ctx, int handler, const betse_network* net)
{
    if (!ctx || handler < 0 || handler > 1) return 2;
    CK(cudaSetDevice(ctx->device));
    destroy_graphs(ctx);
    if (!net) {
        ctx->net_on[handler] = false;
        ctx->P.defer = (!ctx->chans.empty() || ctx->net_on[0] || ctx->net_on[1]) ? 1 : 0;
        return 0;
    }
    if (ctx->X.n_nbr > 0) return fail(ctx, "networks on a domain-decomposed tissue are not implemented");
    if (ctx->hp.fast_update_ecm) return fail(ctx, "networks with fast_update_ecm are not implemented");
    const int K = net->n_species, R = net->n_rates, C = ctx->C, Mo = ctx->Mo;
    if (K <= 0 || R < K || R > NET_MAX_RATES) return fail(ctx, "network: need 0 < n_species <= n_rates <= 48");
    if (net->n_programs < R) return fail(ctx, "network: n_programs < n_rates");
    if (!net->c_cells || !net->code || !net->prog_ptr || !net->stoich || !net->Dgj || !net->z || !net->time_factor)
        return fail(ctx, "network: null table");
    // validate the programs: operands in range, stack discipline
    const int n_pairs = net->prog_ptr[net->n_programs];
    for (int p = 0; p < net->n_programs; ++p) {
        int depth = 0;
        if (net->prog_ptr[p] > net->prog_ptr[p + 1]) return fail(ctx, "network: prog_ptr not monotone");
        for (int pc = net->prog_ptr[p]; pc < net->prog_ptr[p + 1]; ++pc) {
            const int op = net->code[2 * pc], arg = net->code[2 * pc + 1];
            const bool memzone = p >= R;
            if (op < 0 || op > RL_EXP) return fail(ctx, "network: unknown opcode");
            if (op == RL_PUSHC && (arg < 0 || arg >= net->n_consts)) return fail(ctx, "network: constant index out of range");
            if (op == RL_PUSHS && (arg < 0 || arg >= K)) return fail(ctx, "network: substance index out of range");
            if (op == RL_PUSHA && (arg < 0 || arg >= (memzone ? net->n_mem_arrays : net->n_cell_arrays))) return fail(ctx, "network: array index out of range");
            if ((op == RL_PUSHI || op == RL_PUSHM) && (arg < 0 || arg >= ctx->I)) return fail(ctx, "network: ion index out of range");
            if (op == RL_PUSHV && !memzone) return fail(ctx, "network: Vmem in a cell-zone program");
            if ((op == RL_PUSHE || op == RL_PUSHJ) && !ctx->hp.is_ecm) return fail(ctx, "network: env concentration in a rate law without extracellular spaces");
            if ((op == RL_PUSHE || op == RL_PUSHJ) && !memzone && !net->map_cell2ecm) return fail(ctx, "network: a cell-zone rate law reads an env concentration: map_cell2ecm needed");
            if (op == RL_PUSHE && (arg < 0 || arg >= K || !(net->env_on && net->env_on[arg]))) return fail(ctx, "network: env concentration of a substance without env_on");
            if (op == RL_PUSHJ && (arg < 0 || arg >= ctx->I)) return fail(ctx, "network: ion index out of range");
            if (op <= RL_LAST_PUSH) ++depth;
            else if (op != RL_NEG && op != RL_EXP) --depth;
            if (depth < 1 || depth > NET_STACK) return fail(ctx, "network: program violates the stack discipline");
        }
        if (depth != 1) return fail(ctx, "network: program does not leave exactly one value");
    }
    KNet N;
    memset(&N, 0, sizeof N);
    N.K = K; N.n_rates = R; N.E = ctx->E;
    if (net->map_cell2ecm) {
        for (int c = 0; c < ctx->Co; ++c)
            if (net->map_cell2ecm[c] < 0 || net->map_cell2ecm[c] >= ctx->E) return fail(ctx, "network: map_cell2ecm out of range");
        int rr = dev_upload(ctx, (int**)&N.cell2ecm, net->map_cell2ecm, (size_t)C);
        if (rr) return rr;
    }
    int r;
    if ((r = dev_upload(ctx, &N.c, net->c_cells, (size_t)K * C))) return r;
    if ((r = dev_alloc(ctx, &N.rates, (size_t)R * C))) return r;
    if ((r = dev_alloc(ctx, &N.gj_delta, (size_t)C))) return r;
    if ((r = dev_upload(ctx, (int**)&N.code, net->code, (size_t)2 * std::max(1, n_pairs)))) return r;
    if ((r = dev_upload(ctx, (int**)&N.ptr, net->prog_ptr, (size_t)net->n_programs + 1))) return r;
    if ((r = dev_upload(ctx, (double**)&N.consts, net->consts, (size_t)std::max(1, net->n_consts)))) return r;
    if (net->n_cell_arrays > 0) { if ((r = dev_upload(ctx, (double**)&N.cell_arrays, net->cell_arrays, (size_t)net->n_cell_arrays * C))) return r; }
    if (net->n_mem_arrays > 0) { if ((r = dev_upload(ctx, (double**)&N.mem_arrays, net->mem_arrays, (size_t)net->n_mem_arrays * Mo))) return r; }
    if (net->growth_mask) { if ((r = dev_upload(ctx, (unsigned char**)&N.gmask, net->growth_mask, (size_t)K * C))) return r; }
    if ((r = dev_upload(ctx, (double**)&N.stoich, net->stoich, (size_t)K * R))) return r;
    if ((r = dev_upload(ctx, (double**)&N.Dgj, net->Dgj, (size_t)K))) return r;
    if ((r = dev_upload(ctx, (double**)&N.z, net->z, (size_t)K))) return r;
    if ((r = dev_upload(ctx, (double**)&N.tdf, net->time_factor, (size_t)K))) return r;
    ctx->net_Dm[handler].assign((size_t)K, 0.0);
    ctx->net_env_on[handler].assign((size_t)K, 0);
    if (net->env_on) {
        bool any = false;
        for (int k = 0; k < K; ++k) any = any || net->env_on[k] != 0;
        if (any) {
            if (!net->Dm || !net->c_bound || !net->c_env || !net->D_env) return fail(ctx, "network: env_on needs Dm, c_bound, c_env and D_env");
            if (!ctx->hp.is_ecm) return fail(ctx, "network substances outside the cells need extracellular spaces");
            if (ctx->hp.sharpness < 1.0) return fail(ctx, "network substances in the environment with sharpness < 1 are not implemented");
            const size_t E = (size_t)ctx->E;
            if ((r = dev_upload(ctx, &N.c_env, net->c_env, (size_t)K * E))) return r;
            if ((r = dev_alloc(ctx, &N.env_tmp, 2 * E))) return r;
            if ((r = dev_alloc(ctx, &N.mem_delta, (size_t)K * C))) return r;
            if ((r = dev_upload(ctx, (double**)&N.Dm, net->Dm, (size_t)K))) return r;
            if ((r = dev_upload(ctx, (double**)&N.c_bound, net->c_bound, (size_t)K))) return r;
            if ((r = dev_upload(ctx, (double**)&N.D_env, net->D_env, (size_t)K * E))) return r;
            ctx->net_Dm[handler].assign(net->Dm, net->Dm + K);
            ctx->net_env_on[handler].assign(net->env_on, net->env_on + K);
            if ((r = dev_upload(ctx, (unsigned char**)&N.env_on_d, (const unsigned char*)net->env_on, (size_t)K))) return r;
        }
    }
    N.n_env_rx = 0;
    if (net->n_env_rx > 0) {
        // reactions outside the cells (networks.py:2872-2889)
        if (!N.c_env) return fail(ctx, "network: extracellular reactions need substances with env_on");
        if (net->n_env_rx > 16 || !net->env_rx_prog || !net->stoich_env) return fail(ctx, "network: at most 16 extracellular reactions, with programs and stoich_env");
        for (int j = 0; j < net->n_env_rx; ++j)
            if (net->env_rx_prog[j] < R || net->env_rx_prog[j] >= net->n_programs) return fail(ctx, "network: extracellular reaction program out of range");
        for (int k = 0; k < K; ++k)
            for (int j = 0; j < net->n_env_rx; ++j)
                if (net->stoich_env[(size_t)k * net->n_env_rx + j] != 0.0 && !net->env_on[k])
                    return fail(ctx, "network: an extracellular reaction moves a substance without env_on");
        if ((r = dev_upload(ctx, (int**)&N.env_rx_prog, (const int*)net->env_rx_prog, (size_t)net->n_env_rx))) return r;
        if ((r = dev_upload(ctx, (double**)&N.stoich_env, net->stoich_env, (size_t)K * net->n_env_rx))) return r;
        N.n_env_rx = net->n_env_rx;
    }
    if ((r = ensure_defer_buffers(ctx))) return r;
    {
        std::vector<double> nan((size_t)K, nan_value());
        if ((r = dev_upload(ctx, &N.clamp, (const double*)nan.data(), (size_t)K))) return r;
    }
    ctx->net_trans[handler].clear(); ctx->net_trans_cm[handler].clear(); ctx->net_trans_em[handler].clear();
    ctx->net_trans_mm[handler].clear();
    ctx->net_tw_rows[handler] = 0;
    for (int k = 0; k < NET_MAX_RATES; ++k) N.tw_s[k] = -1;
    for (int i = 0; i < 8; ++i) N.tw_i[i] = -1;
    if (net->n_transporters > 0) {
        if (!net->mem_sa_over_vol) return fail(ctx, "network: transporters need mem_sa_over_vol");
        int rows = 0;
        for (int j = 0; j < net->n_transporters; ++j)
            for (int q = 0; q < net->transporters[j].n_terms && q < BETSE_TR_MAX_TERMS; ++q) {
                const betse_transporter_term& t = net->transporters[j].terms[q];
                if (t.kind == 0 && t.index >= 0 && t.index < 8 && N.tw_i[t.index] < 0) N.tw_i[t.index] = (signed char)rows++;
                if (t.kind == 2 && t.index >= 0 && t.index < NET_MAX_RATES && N.tw_s[t.index] < 0) N.tw_s[t.index] = (signed char)rows++;
            }
        ctx->net_tw_rows[handler] = rows;
        if (rows > 0) { if ((r = dev_alloc(ctx, &N.tw, (size_t)rows * Mo))) return r; }
        if ((r = dev_alloc(ctx, &N.tr_flux, (size_t)Mo))) return r;
        if ((r = dev_upload(ctx, (double**)&N.sa_over_vol, net->mem_sa_over_vol, (size_t)Mo))) return r;
    }
    for (int j = 0; j < net->n_transporters; ++j) {
        const betse_transporter& T = net->transporters[j];
        if (!ctx->hp.is_ecm) return fail(ctx, "network: transporters without extracellular spaces are not implemented");
        if (T.prog < R || T.prog >= net->n_programs) return fail(ctx, "network: transporter program is not a membrane-zone program");
        if (T.n_terms < 0 || T.n_terms > BETSE_TR_MAX_TERMS) return fail(ctx, "network: transporter with too many terms");
        for (int q = 0; q < T.n_terms; ++q) {
            const betse_transporter_term& t = T.terms[q];
            const bool ion = t.kind == 0 || t.kind == 1, sub = t.kind == 2 || t.kind == 3;
            if ((!ion && !sub) || (ion && (t.index < 0 || t.index >= ctx->I)) || (sub && (t.index < 0 || t.index >= K)) ||
                (t.sign != 1 && t.sign != -1)) return fail(ctx, "network: bad transporter term");
            if (t.kind == 3 && !(net->env_on && net->env_on[t.index] && N.c_env)) return fail(ctx, "network: a transporter moves a substance outside the cells that has no env_on");
        }
        const unsigned char *cm = nullptr, *em = nullptr, *mm = nullptr;
        if (T.cell_mask) { if ((r = dev_upload(ctx, (unsigned char**)&cm, (const unsigned char*)T.cell_mask, (size_t)C))) return r; }
        if (T.env_mask) { if ((r = dev_upload(ctx, (unsigned char**)&em, (const unsigned char*)T.env_mask, (size_t)ctx->E))) return r; }
        if (T.mem_mask) { if ((r = dev_upload(ctx, (unsigned char**)&mm, (const unsigned char*)T.mem_mask, (size_t)Mo))) return r; }
        ctx->net_trans[handler].push_back(T);
        ctx->net_trans_cm[handler].push_back(cm);
        ctx->net_trans_em[handler].push_back(em);
        ctx->net_trans_mm[handler].push_back(mm);
    }
    ctx->net_intra[handler].clear();
    if (net->intra_on) {
        bool any = false;
        for (int k = 0; k < K; ++k) any = any || net->intra_on[k] != 0;
        if (any) {
            if (!net->Do || !net->c_mems || !net->R_rads || !net->mem_sa_over_vol) return fail(ctx, "network: intra_on needs Do, c_mems, R_rads and mem_sa_over_vol");
            if ((r = dev_upload(ctx, &N.cmem, net->c_mems, (size_t)K * Mo))) return r;
            if ((r = dev_upload(ctx, (double**)&N.Do, net->Do, (size_t)K))) return r;
            if ((r = dev_upload(ctx, (double**)&N.R_rads, net->R_rads, (size_t)Mo))) return r;
            if (!N.sa_over_vol) { if ((r = dev_upload(ctx, (double**)&N.sa_over_vol, net->mem_sa_over_vol, (size_t)Mo))) return r; }
            ctx->net_intra[handler].assign(net->intra_on, net->intra_on + K);
            if ((r = dev_alloc(ctx, &N.gjf, (size_t)Mo))) return r;
            if (net->mu_mem) { if ((r = dev_upload(ctx, (double**)&N.mu_mem, net->mu_mem, (size_t)K))) return r; }
            // charged substances drift in the cell's field: sim.Emc (k_diag) is then part of every step
            for (int k = 0; k < K; ++k)
                if (net->intra_on[k] && (net->z[k] != 0.0 || (net->mu_mem && net->mu_mem[k] != 0.0))) ctx->need_emc = true;
            if (ctx->need_emc) {
                if (ctx->X.n_nbr > 0) return fail(ctx, "network: 'update intracellular' of a charged substance on a domain-decomposed tissue is not implemented");
                if ((r = ensure_diag_buffers(ctx))) return r;
                if (net->Emc) { if ((r = xfer(ctx, ctx->A.Emc, net->Emc, (size_t)Mo * sizeof(double), cudaMemcpyHostToDevice))) return r; }
                destroy_graphs(ctx);
            }
        }
    }
    ctx->net_pumps[handler].clear();
    if (net->n_pumps > 0) {
        std::vector<unsigned char> pumped((size_t)K, 0);
        for (int j = 0; j < net->n_pumps; ++j) {
            const betse_substance_pump& q = net->pumps[j];
            if (q.species < 0 || q.species >= K || pumped[q.species]) return fail(ctx, "network: pump species out of range or pumped twice");
            if (!(net->env_on && net->env_on[q.species] && N.c_env)) return fail(ctx, "network: a pumped substance needs env_on and extracellular spaces");
            pumped[q.species] = 1;
            ctx->net_pumps[handler].push_back(q);
        }
        if ((r = dev_upload(ctx, (unsigned char**)&N.pumped, (const unsigned char*)pumped.data(), (size_t)K))) return r;
        if ((r = dev_alloc(ctx, &N.c_save, (size_t)net->n_pumps * C))) return r;
    }
    ctx->net_gates[handler].clear();
    for (int j = 0; j < net->n_ligand_gates; ++j) {
        const betse_ligand_gate& g = net->ligand_gates[j];
        if (g.species < 0 || g.species >= K || g.ion < 0 || g.ion >= ctx->I) return fail(ctx, "network: ligand gate species / ion out of range");
        if (g.extracell && !(net->env_on && net->env_on[g.species] && ctx->hp.is_ecm))
            return fail(ctx, "network: an extracellular ligand needs extracellular spaces and env_on for its substance");
        ctx->net_gates[handler].push_back(g);
    }
    if (net->n_ligand_gates > 0) {
        if ((r = dev_alloc(ctx, &ctx->lig_tmp[handler], (size_t)net->n_ligand_gates * Mo))) return r;
    }
    ctx->net_mods[handler].clear();
    bool tj_fresh = false;
    for (int j = 0; j < net->n_modulators; ++j) {
        const betse_modulator& md = net->modulators[j];
        if (md.target < 0 || md.target > 2) return fail(ctx, "network: modulator target must be 0 (gap junctions), 1 (Na/K-ATPase) or 2 (tight junctions)");
        if (md.prog < R || md.prog >= net->n_programs) return fail(ctx, "network: modulator program is not a membrane-zone program");
        if (md.target == 2) {
            // tight junctions: an extracellular-zone program over the squares of the barrier
            if (!ctx->hp.is_ecm || !net->tj_targets || net->n_tj <= 0 || !net->D_env_raw)
                return fail(ctx, "network: a tight-junction modulator needs extracellular spaces, tj_targets and D_env_raw");
            if (md.ion >= ctx->I) return fail(ctx, "network: tight-junction modulator ion out of range");
            if (ctx->X.n_nbr > 0) return fail(ctx, "network: tight-junction modulators on a domain-decomposed tissue are not implemented");
            for (int q = 0; q < net->n_tj; ++q)
                if (net->tj_targets[q] < 0 || net->tj_targets[q] >= ctx->E) return fail(ctx, "network: tj_targets out of range");
            if (!tj_fresh) {              // (a repeated betse_set_network replaces the list; the old one goes with the ctx)
                ctx->net_tj[handler] = nullptr;
                if ((r = dev_upload(ctx, &ctx->net_tj[handler], (const int*)net->tj_targets, (size_t)net->n_tj))) return r;
                ctx->net_ntj[handler] = net->n_tj;
                tj_fresh = true;
            }
            if (!ctx->denv_raw) { if ((r = dev_upload(ctx, &ctx->denv_raw, net->D_env_raw, (size_t)ctx->I * ctx->E))) return r; }
            if (!ctx->tj_mod) {
                std::vector<double> ones;
                if (!net->TJ_modulator) ones.assign((size_t)ctx->I * ctx->E, 1.0);
                if ((r = dev_upload(ctx, &ctx->tj_mod, net->TJ_modulator ? net->TJ_modulator : (const double*)ones.data(), (size_t)ctx->I * ctx->E))) return r;
                CK(cudaStreamSynchronize(ctx->stream));
            }
            CK(cudaStreamSynchronize(ctx->stream));
            ctx->net_mods[handler].push_back(md);
            continue;
        }
        // the block becomes a per-membrane array (it starts from the value in force now)
        const double** slot = md.target == 0 ? &ctx->A.gj_block : &ctx->A.NaK_block;
        if (!*slot) {
            std::vector<double> init((size_t)Mo, md.target == 0 ? ctx->P.gj_block : ctx->P.NaK_block);
            double* p;
            if ((r = dev_upload(ctx, &p, (const double*)init.data(), (size_t)Mo))) return r;
            CK(cudaStreamSynchronize(ctx->stream));
            *slot = p;
        }
        ctx->net_mods[handler].push_back(md);
    }
    ctx->net_affect[handler] = false;
    if (net->affect_charge) {
        // sim.extra_rho_cells / extra_rho_env / extra_J_mem / extra_Jenv are the handler's arrays from now on
        // (networks.py:2971-2977: published every step, whatever the substances' charges)
        std::vector<double> ones((size_t)K, 1.0);
        N.affect = 1;
        if ((r = dev_upload(ctx, (double**)&N.scale, net->scale_factor ? net->scale_factor : ones.data(), (size_t)K))) return r;
        if ((r = dev_alloc(ctx, &N.fmem_tmp, (size_t)Mo))) return r;
        if ((r = dev_alloc(ctx, &N.rho_cells, (size_t)C))) return r;
        if (N.c_env) {
            if ((r = dev_alloc(ctx, &N.rho_env, (size_t)ctx->E))) return r;
            if (!ctx->A.extra_Jenv_x) {
                if ((r = dev_alloc(ctx, &ctx->A.extra_Jenv_x, (size_t)ctx->E))) return r;
                if ((r = dev_alloc(ctx, &ctx->A.extra_Jenv_y, (size_t)ctx->E))) return r;
            }
        }
        ctx->net_affect[handler] = true;
        ctx->P.chan_charge = 1;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->nets[handler] = N;
    ctx->net_Dgj[handler].assign(net->Dgj, net->Dgj + K);
    ctx->net_on[handler] = true;
    ctx->net_nprog[handler] = net->n_programs;
    ctx->P.defer = 1;
    publish_affect(ctx);
    return 0;
}

extern "C" int betse_network_state(betse_ctx* ctx, int handler, double* c_cells, double* rates)
{
    if (!ctx || handler < 0 || handler > 1) return 2;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->net_on[handler]) return fail(ctx, "no network on this handler");
    const KNet& N = ctx->nets[handler];
    if (c_cells) { int xr = xfer(ctx, c_cells, N.c, (size_t)N.K * ctx->C * sizeof(double), cudaMemcpyDeviceToHost); if (xr) return xr; }
    if (rates) { int xr = xfer(ctx, rates, N.rates, (size_t)N.n_rates * ctx->C * sizeof(double), cudaMemcpyDeviceToHost); if (xr) return xr; }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int betse_network_tj_modulator(betse_ctx* ctx, double* tj_modulator)
{
    if (!ctx || !tj_modulator) return 2;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->tj_mod) return fail(ctx, "no tight-junction modulator is configured");
    int xr = xfer(ctx, tj_modulator, ctx->tj_mod, (size_t)ctx->I * ctx->E * sizeof(double), cudaMemcpyDeviceToHost);
    if (xr) return xr;
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int betse_network_env_state(betse_ctx* ctx, int handler, double* c_env)
{
    if (!ctx || handler < 0 || handler > 1 || !c_env) return 2;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->net_on[handler]) return fail(ctx, "no network on this handler");
    const KNet& N = ctx->nets[handler];
    const size_t nb = (size_t)N.K * ctx->E * sizeof(double);
    if (N.c_env) {
        CK(cudaMemcpyAsync(c_env, N.c_env, nb, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } else memset(c_env, 0, nb);
    return 0;
}

static __global__ void k_gather_mems(const double* __restrict__ c, const int* __restrict__ m2c, double* __restrict__ out, int M)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < M) out[m] = c[m2c[m]];
}

extern "C" int betse_network_mem_state(betse_ctx* ctx, int handler, double* c_mems)
{
    if (!ctx || handler < 0 || handler > 1 || !c_mems) return 2;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->net_on[handler]) return fail(ctx, "no network on this handler");
    const KNet& N = ctx->nets[handler];
    const int Mo = ctx->Mo;
    double* tmp = nullptr;
    CK(cudaMalloc((void**)&tmp, (size_t)Mo * sizeof(double)));
    for (int k = 0; k < N.K; ++k) {
        const bool intra = !ctx->net_intra[handler].empty() && ctx->net_intra[handler][k];
        const double* src = tmp;
        if (intra) src = N.cmem + (size_t)k * Mo;
        else k_gather_mems<<<(Mo + 255) / 256, 256, 0, ctx->stream>>>(N.c + (size_t)k * ctx->C, ctx->A.mem_to_cells, tmp, Mo);
        int xr = xfer(ctx, c_mems + (size_t)k * Mo, src, (size_t)Mo * sizeof(double), cudaMemcpyDeviceToHost);
        if (xr) { cudaFree(tmp); return xr; }
        CK(cudaStreamSynchronize(ctx->stream));
    }
    cudaFree(tmp);
    return 0;
}

extern "C" int betse_network_set_events(betse_ctx* ctx, int handler, const double* c_bound, const double* clamp)
{
    if (!ctx || handler < 0 || handler > 1) return 2;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->net_on[handler]) return fail(ctx, "no network on this handler");
    const KNet& N = ctx->nets[handler];
    // device arrays, not kernel parameters: the captured step graphs stay valid
    if (c_bound && N.c_bound) CK(cudaMemcpyAsync((void*)N.c_bound, c_bound, (size_t)N.K * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (clamp && N.clamp) CK(cudaMemcpyAsync(N.clamp, clamp, (size_t)N.K * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int betse_set_noise_flux(betse_ctx* ctx, int ion, const double* flux)
{
    if (!ctx || !flux) return 2;
    CK(cudaSetDevice(ctx->device));
    if (ion < 0 || ion >= ctx->I) return fail(ctx, "noise flux: ion index out of range");
    if (ctx->X.n_nbr > 0) return fail(ctx, "dynamic noise on a domain-decomposed tissue is not implemented");
    if (ctx->hp.fast_update_ecm) return fail(ctx, "dynamic noise with fast_update_ecm is not implemented");
    int r;
    if (!ctx->noise_on) {
        if ((r = ensure_defer_buffers(ctx))) return r;
        if ((r = dev_alloc(ctx, &ctx->noise_flux, (size_t)ctx->Mo))) return r;
        ctx->noise_on = true;
        ctx->P.defer = 1;
        ctx->use_graphs = false;         // the extra launch exists only in steps that carry a draw
        destroy_graphs(ctx);
    }
    if ((r = xfer(ctx, ctx->noise_flux, flux, (size_t)ctx->Mo * sizeof(double), cudaMemcpyHostToDevice))) return r;
    ctx->noise_ion = ion;
    ctx->noise_pending = true;
    return 0;
}

extern "C" int betse_host_alloc(size_t bytes, void** out)
{
    if (!out) return 2;
    *out = nullptr;
    return cudaHostAlloc(out, bytes ? bytes : 8, cudaHostAllocDefault) == cudaSuccess ? 0 : 1;
}

extern "C" int betse_host_alloc_on(int device, size_t bytes, void** out)
{
    if (!out) return 2;
    *out = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return 1;
    return cudaHostAlloc(out, bytes ? bytes : 8, cudaHostAllocDefault) == cudaSuccess ? 0 : 1;
}

extern "C" void betse_host_copy(void* dst, const void* src, size_t bytes)
{
    par_memcpy(static_cast<char*>(dst), static_cast<const char*>(src), bytes);
}

extern "C" void betse_host_expand(double* dst, const double* src, const int32_t* idx, int n_rows, size_t n_src, size_t n_dst)
{
    const unsigned hw = std::thread::hardware_concurrency();
    const int T = n_dst < ((size_t)1 << 16) ? 1 : (hw >= 16 ? 8 : hw >= 8 ? 4 : 2);
    auto work = [=](size_t m0, size_t m1) {
        for (int i = 0; i < n_rows; ++i) {
            const double* s = src + (size_t)i * n_src;
            double* d = dst + (size_t)i * n_dst;
            for (size_t m = m0; m < m1; ++m) d[m] = s[idx[m]];
        }
    };
    std::thread th[8];
    const size_t per = (n_dst + T - 1) / T;
    for (int t = 1; t < T; ++t) {
        const size_t a = (size_t)t * per, b = a + per < n_dst ? a + per : n_dst;
        th[t - 1] = std::thread([=] { if (a < b) work(a, b); });
    }
    work(0, per < n_dst ? per : n_dst);
    for (int t = 1; t < T; ++t) th[t - 1].join();
}

extern "C" void betse_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

extern "C" int betse_set_row_ranges(betse_ctx* ctx, int yi0, int yi1, int ya0, int ya1, int yf0, int yf1)
{
    if (!ctx) return 2;
    const int ny = ctx->ny;
    if (yi0 < 0 || yi1 > ny || ya0 < 0 || ya1 > ny || yf0 < 0 || yf1 > ny || yi0 > yi1 || ya0 > ya1 || yf0 > yf1)
        return fail(ctx, "row range outside the local grid");
    KParams& P = ctx->P;
    P.yi0 = yi0; P.yi1 = yi1; P.ya0 = ya0; P.ya1 = ya1; P.yf0 = yf0; P.yf1 = yf1;
    ctx->xfuse_dirty = true;
    destroy_graphs(ctx);
    return 0;
}

extern "C" int betse_window(betse_ctx* ctx, betse_window_info* out)
{
    if (!ctx || !out) return 2;
    CK(cudaSetDevice(ctx->device));
    *out = ctx->winfo;
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ctx->win);
    if (e == cudaSuccess) memcpy(out->ipc_handle, &h, sizeof h);
    else { cudaGetLastError(); memset(out->ipc_handle, 0, sizeof out->ipc_handle); }   // same-process use still works
    return 0;
}

extern "C" int betse_attach_neighbor(betse_ctx* ctx, const betse_neighbor* nb)
{
    if (!ctx || !nb) return 2;
    CK(cudaSetDevice(ctx->device));
    if (ctx->X.n_nbr >= 2) return fail(ctx, "a strip has at most two neighbours");
    if (nb->side != 0 && nb->side != 1) return fail(ctx, "neighbor.side must be 0 or 1");
    if (!ctx->hp.is_ecm) return fail(ctx, "domain decomposition needs extracellular spaces (no-ECM tissues run as replicas)");
    if (ctx->P.has_phi) return fail(ctx, "domain decomposition does not support a boundary-voltage potential (Phi_b)");
    if (!ctx->chans.empty()) return fail(ctx, "channels on a domain-decomposed tissue are not implemented");
    if (ctx->net_on[0] || ctx->net_on[1]) return fail(ctx, "networks on a domain-decomposed tissue are not implemented");
    if (ctx->noise_on) return fail(ctx, "dynamic noise on a domain-decomposed tissue is not implemented");
    const betse_window_info& W = nb->info;
    if (W.n_ions != ctx->I || W.nx != ctx->nx) return fail(ctx, "neighbour window has different n_ions / nx");
    char* base = (char*)W.base;
    if (!nb->same_process) {
        cudaIpcMemHandle_t h;
        memcpy(&h, W.ipc_handle, sizeof h);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->ipc_opened.push_back(p);
        base = (char*)p;
    }
    if (!base) return fail(ctx, "neighbour window pointer is null");
    if (nb->recv_cell0 < 0 || nb->recv_cell0 + nb->n_send_cells > W.n_cells) return fail(ctx, "ghost-cell range outside the neighbour");
    if (nb->n_send_flux < 0 || (nb->n_send_flux > 0 && (nb->recv_slot0 < 0 || nb->recv_slot0 + nb->n_send_flux > W.n_flux_slots)))
        return fail(ctx, "remote flux-slot range outside the neighbour's exchange slots");
    if (nb->cc_rows < 0 || nb->v_rows < 0 || (nb->cc_dst_row0 + nb->cc_rows) * W.nx > W.n_env ||
        (nb->v_dst_row0 + nb->v_rows) * W.nx > W.n_env || (nb->cc_src_row0 + nb->cc_rows) > ctx->ny ||
        (nb->v_src_row0 + nb->v_rows) > ctx->ny) return fail(ctx, "row block outside the grid");
    for (int j = 0; j < nb->n_send_cells; ++j)
        if (nb->send_cells[j] < 0 || nb->send_cells[j] >= ctx->Co) return fail(ctx, "send_cells entry is not an owned cell");
    for (int j = 0; j < nb->n_send_flux; ++j)
        if (nb->send_flux[j] < 0 || nb->send_flux[j] >= ctx->Mo) return fail(ctx, "send_flux entry is not an owned membrane");
    XNbr& x = ctx->X.nb[ctx->X.n_nbr];
    memset(&x, 0, sizeof x);
    for (int b = 0; b < 2; ++b) {
        x.cc_mid[b] = (double*)(base + W.off_cc_mid[b]);
        x.vm_cell[b] = (double*)(base + W.off_vm_cell[b]);
        x.cc_env[b] = (double*)(base + W.off_cc_env[b]);
    }
    x.flux = (double*)(base + W.off_flux);
    x.v_raw = (double*)(base + W.off_v_raw);
    x.flags = (unsigned long long*)(base + W.off_flags);
    x.Cn = W.n_cells; x.En = W.n_env; x.side = nb->side;
    x.n_send_cells = nb->n_send_cells; x.recv_cell0 = nb->recv_cell0;
    x.n_send_flux = nb->n_send_flux; x.recv_slot0 = nb->recv_slot0;
    x.cc_rows = nb->cc_rows; x.cc_src_row0 = nb->cc_src_row0; x.cc_dst_row0 = nb->cc_dst_row0;
    x.v_rows = nb->v_rows; x.v_src_row0 = nb->v_src_row0; x.v_dst_row0 = nb->v_dst_row0;
    int r;
    if ((r = dev_upload(ctx, (int**)&x.send_cells, nb->send_cells, nb->n_send_cells))) return r;
    if ((r = dev_upload(ctx, (int**)&x.send_flux, nb->send_flux, nb->n_send_flux))) return r;
    if (ctx->mem_ell) {          // where k_cell leaves the same fluxes
        if ((r = dev_alloc(ctx, (int**)&x.send_flux_ell, (size_t)nb->n_send_flux, false))) return r;
        launch_gather_int(const_cast<int*>(x.send_flux_ell), ctx->mem_ell, x.send_flux, nb->n_send_flux, ctx->stream);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    x.push_lo = std::min(nb->cc_rows > 0 ? nb->cc_src_row0 : INT_MAX, nb->v_rows > 0 ? nb->v_src_row0 : INT_MAX);
    x.push_hi = std::max(nb->cc_rows > 0 ? nb->cc_src_row0 + nb->cc_rows : -1, nb->v_rows > 0 ? nb->v_src_row0 + nb->v_rows : -1);
    ctx->xh_cells[ctx->X.n_nbr].assign(nb->send_cells, nb->send_cells + nb->n_send_cells);
    ctx->xh_flux[ctx->X.n_nbr].assign(nb->send_flux, nb->send_flux + nb->n_send_flux);
    ctx->X.side_k[nb->side] = ctx->X.n_nbr;
    ctx->xfuse_dirty = true;
    ctx->X.n_nbr++;
    ctx->P.xsides |= 1 << nb->side;
    // neighbours in other processes run on GPUs of their own: the consumers of an exchange may spin inside their kernels.
    // Strips that share this process's device are stepped by one host thread — a spinning kernel could keep the kernel it
    // waits for off the SMs — and keep the stand-alone wait.
    { const char* e = getenv("BETSE_XWAIT"); ctx->xwait = !nb->same_process && !(e && e[0] == '0'); }
    destroy_graphs(ctx);
    return 0;
}

extern "C" int betse_exchange(betse_ctx* ctx, int which, int buf_next, int mode)
{
    if (!ctx || (which != BETSE_XCHG_X1 && which != BETSE_XCHG_X2) || !(mode & 3)) return 2;
    CK(cudaSetDevice(ctx->device));
    if (ctx->X.n_nbr == 0) return 0;
    launch_xchg(ctx->P, ctx->A, ctx->X, which, buf_next ? (ctx->cur ^ 1) : ctx->cur, mode, ctx->flux_is_ell, ctx->stream);
    CK(cudaGetLastError());
    return 0;
}
