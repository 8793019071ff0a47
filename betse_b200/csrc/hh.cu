// Helmholtz-Hodge decomposition of the extracellular current density (sampled steps only):
//   get_current, env part          betse/science/physics/ion_current.py:50-73
//   stb.HH_Decomp                  betse/science/sim_toolbox.py:1236-1290
//   fd.divergence / fd.gradient    betse/science/math/finitediff.py:1236-1311
//   cells.lapENVinv (Dirichlet)    betse/science/cells.py:503-504, finitediff.py:252-431
// Outputs J_env_x/y (the divergence-free part: what write2storage stores as I_tot_*), B_field,
// Jtx/Jty.  They feed nothing back into the timestep, so this runs when a step is sampled.
//
// The reference multiplies by a dense (E x E) pseudo-inverse of the Laplacian whose border rows are
// the identity/d^2 and whose interior rows are the 5-point stencil/d^2.  That matrix is regular, so
// pinv == inverse and "lapENVinv . b" is the Dirichlet solve with border values u = b*d^2 lifted to
// the right-hand side (SURVEY §8 notes, checked to 7e-15 against the dense product).  The interior
// problem is diagonalised by the type-I sine transform in both axes; for a diagnostic evaluated
// every ~10th step a transform as two dense products with the sine matrices is plenty
// (u = Sy (Sy R Sx / lambda) Sx / ((my+1)(mx+1)), 4 x ~1e9 FMA at a 1000^2 grid).
#include <math.h>
#include <stdlib.h>
#include "kparams.cuh"

#include "hh.cuh"

__global__ void k_hh_sine(double* S, double* lam, const int n)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (k >= n) return;
    const long long prod = (long long)(j + 1) * (k + 1) % (2LL * (n + 1));     // exact argument reduction
    S[(size_t)j * n + k] = sinpi((double)prod / (double)(n + 1));
    if (j == 0) lam[k] = 2.0 * cospi((double)(k + 1) / (double)(n + 1)) - 2.0;
}

// Jx = np.dot(p.F*sim.zs, sim.fluxes_env_x) (ion_current.py:52-54)
__global__ void k_hh_J(const __grid_constant__ KParams P, const KArrays A, HHBuf H)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int E = P.nx * P.ny;
    if (k >= E) return;
    double jx = 0.0, jy = 0.0;
    for (int i = 0; i < P.n_ions; ++i) {
        jx = fma(P.zF[i], A.fl_env_x[(size_t)i * E + k], jx);
        jy = fma(P.zF[i], A.fl_env_y[(size_t)i * E + k], jy);
    }
    if (A.extra_Jenv_x) { jx += A.extra_Jenv_x[k]; jy += A.extra_Jenv_y[k]; }      // + sim.extra_Jenv_x (ion_current.py:53-54)
    H.Jx[k] = jx; H.Jy[k] = jy;
}

// fd.diff (finitediff.py:1268-1303): edge rows carry the opposite sign convention to fd.gradient
__device__ __forceinline__ double diff_x(const double* F, int y, int x, int nx, double d)
{
    const double* r = F + (size_t)y * nx;
    if (x == 0) return (r[0] - r[1]) / d;
    if (x == nx - 1) return (r[nx - 2] - r[nx - 1]) / d;
    return -(r[x - 1] - r[x + 1]) / (2.0 * d);
}
__device__ __forceinline__ double diff_y(const double* F, int y, int x, int ny, int nx, double d)
{
    if (y == 0) return -(F[(size_t)nx + x] - F[x]) / d;
    if (y == ny - 1) return -(F[(size_t)(ny - 1) * nx + x] - F[(size_t)(ny - 2) * nx + x]) / d;
    return -(F[(size_t)(y - 1) * nx + x] - F[(size_t)(y + 1) * nx + x]) / (2.0 * d);
}

// right-hand sides: bA = -div(-Jy, Jx) with a zero border, bB = div(Jx, Jy) with the border set from bound_V
__global__ void k_hh_rhs(const __grid_constant__ KParams P, HHBuf H)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int nx = P.nx, ny = P.ny;
    if (x >= nx) return;
    const double d = P.delta;
    const size_t k = (size_t)y * nx + x;
    const bool border = (x == 0 || x == nx - 1 || y == 0 || y == ny - 1);
    // divJr = diff(-Jy, axis 0 -> x) + diff(Jx, axis 1 -> y)
    double a = 0.0;
    if (!border) a = -(-diff_x(H.Jy, y, x, nx, d) + diff_y(H.Jx, y, x, ny, nx, d));
    H.bA[k] = a;
    double b;
    const double id2 = 1.0 / (d * d);
    if (y == ny - 1) b = -H.bound[0] * id2;          // T  (sim_toolbox.py:1271-1274: rows are assigned last)
    else if (y == 0) b = -H.bound[1] * id2;          // B
    else if (x == nx - 1) b = -H.bound[3] * id2;     // R
    else if (x == 0) b = -H.bound[2] * id2;          // L
    else b = diff_x(H.Jx, y, x, nx, d) + diff_y(H.Jy, y, x, ny, nx, d);
    H.bB[k] = b;
}

// border lift: u = b*d^2 on the border; interior rhs R = b*d^2 - (adjacent border values)
__global__ void k_hh_lift(const __grid_constant__ KParams P, const double* b, double* u, double* R)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int nx = P.nx, ny = P.ny;
    if (x >= nx) return;
    const double d2 = P.delta * P.delta;
    const size_t k = (size_t)y * nx + x;
    const bool border = (x == 0 || x == nx - 1 || y == 0 || y == ny - 1);
    if (border) { u[k] = b[k] * d2; return; }
    double r = b[k] * d2;
    if (y == 1) r -= b[x] * d2;
    if (y == ny - 2) r -= b[(size_t)(ny - 1) * nx + x] * d2;
    if (x == 1) r -= b[(size_t)y * nx] * d2;
    if (x == nx - 2) r -= b[(size_t)y * nx + nx - 1] * d2;
    R[(size_t)(y - 1) * (nx - 2) + (x - 1)] = r;
}

// C[M][N] = A[M][K] . B[K][N] (row major), batched over blockIdx.z (operand strides sA / sB / sC elements, 0 = shared).
// The Poisson solves are four such products each (2 GFLOP at a 1000^2 grid): bound by the fp64 pipe (36.7 TFLOP/s measured,
// tools/micro/fp64_peak.cu).  A warp issues a DFMA every other cycle at best, so the pipe wants >= 4 warps per scheduler:
// 128x64 tile per CTA of 256 threads, 8x4 accumulators per thread (~110 registers, two CTAs per SM), the k-tile of 16
// double-buffered in shared memory and filled with cp.async (no staging registers; out-of-range elements are zero-filled
// by the copy).  A thread owns rows {32g + 2ty + r} and columns {32h + 2tx + c} (g < 4, h < 2; r, c < 2): its 16-byte
// shared loads are conflict-free across tx and broadcast across ty.  (An 8x8 tile at 255 registers ran the pipe at 48 %:
// profiles/r02u_ncu_gemm_summary.csv.)
#define GM 128
#define GN 64
#define GK 16
#define GEMM_SMEM ((2 * GK * (GM + 2) + 2 * GK * GN) * 8)
__device__ __forceinline__ void cp_async8(double* dst, const double* src, const bool ok)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int n = ok ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

// one product per blockIdx.z: operands, inner dimension; element (row, k) of A at A[row*lda + k*ask], (k, col) of B at
// B[k*ldb + col], C[row*ldc + col] — the strides let a product read every other row / column of a matrix in place
struct GemmArgs { const double* A[4]; const double* B[4]; double* C[4]; int K[4]; };

__global__ void __launch_bounds__(256, 2)
k_hh_gemm(const __grid_constant__ GemmArgs g, const int M, const int N, const int lda, const int ask, const int ldb, const int ldc)
{
    extern __shared__ __align__(16) double sh_gemm[];                  // GEMM_SMEM bytes (opt-in size: launch_hh_setup)
    double (*shA)[GK][GM + 2] = reinterpret_cast<double (*)[GK][GM + 2]>(sh_gemm);      // + 2: the transposing copies hit 2 banks, not 16
    double (*shB)[GK][GN] = reinterpret_cast<double (*)[GK][GN]>(sh_gemm + 2 * GK * (GM + 2));
    const double* __restrict__ A = g.A[blockIdx.z];
    const double* __restrict__ B = g.B[blockIdx.z];
    double* __restrict__ C = g.C[blockIdx.z];
    const int K = g.K[blockIdx.z];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int r0 = blockIdx.y * GM, c0 = blockIdx.x * GN;
    // global -> shared, coalesced: A as 8 copies of (16 rows x 16 k) — a warp covers two rows of 128 bytes per copy;
    // B row k0 + t/16, 4 consecutive columns from c0 + 4 (t % 16)
    const int ak = t & 15, ar = t >> 4;
    const int bk = t >> 4, bc = (t & 15) * 4;
    auto fill = [&](const int buf, const int k0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int row = r0 + q * 16 + ar;
            const bool ok = row < M && k0 + ak < K;
            cp_async8(&shA[buf][ak][q * 16 + ar], ok ? A + (size_t)row * lda + (size_t)(k0 + ak) * ask : A, ok);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool ok = k0 + bk < K && c0 + bc + q < N;
            cp_async8(&shB[buf][bk][bc + q], ok ? B + (size_t)(k0 + bk) * ldb + c0 + bc + q : B, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    fill(0, 0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < K; k0 += GK) {
        const bool more = k0 + GK < K;
        if (more) fill(buf ^ 1, k0 + GK);
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            double a[8], b[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const double2 va = *reinterpret_cast<const double2*>(&shA[buf][kk][32 * g + 2 * ty]);
                a[2 * g] = va.x; a[2 * g + 1] = va.y;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double2 vb = *reinterpret_cast<const double2*>(&shB[buf][kk][32 * h + 2 * tx]);
                b[2 * h] = vb.x; b[2 * h + 1] = vb.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        if (more) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = r0 + 32 * (i >> 1) + 2 * ty + (i & 1);
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + 32 * (j >> 1) + 2 * tx + (j & 1);
            if (c < N) C[(size_t)r * ldc + c] = acc[i][j];
        }
    }
}

// spectral division (the two dstn factors of 2 folded in) / back-transform normalisation + scatter into u
__global__ void k_hh_scale(double* Tb, const double* ly, const double* lx, int my, int mx)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (l >= mx) return;
    double* T = Tb + (size_t)blockIdx.z * my * mx;
    T[(size_t)k * mx + l] = (4.0 * T[(size_t)k * mx + l]) / (ly[k] + lx[l]);
}
__global__ void k_hh_scatter(const double* T, double* u, int my, int mx)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (l >= mx) return;
    u[(size_t)(k + 1) * (mx + 2) + l + 1] = T[(size_t)k * mx + l] / ((double)(my + 1) * (double)(mx + 1));   // 4 S uh S / (4 (my+1)(mx+1))
}

// fd.gradient (finitediff.py:1236-1266)
__device__ __forceinline__ void grad_at(const double* F, int y, int x, int ny, int nx, double d, double& gx, double& gy)
{
    const double* r = F + (size_t)y * nx;
    if (x == 0) gx = (r[1] - r[0]) / d;
    else if (x == nx - 1) gx = (r[nx - 1] - r[nx - 2]) / d;
    else gx = -(r[x - 1] - r[x + 1]) / (2.0 * d);
    if (y == 0) gy = (F[(size_t)nx + x] - F[x]) / d;
    else if (y == ny - 1) gy = (F[(size_t)(ny - 1) * nx + x] - F[(size_t)(ny - 2) * nx + x]) / d;
    else gy = -(F[(size_t)(y - 1) * nx + x] - F[(size_t)(y + 1) * nx + x]) / (2.0 * d);
}

__global__ void k_hh_out(const __grid_constant__ KParams P, HHBuf H)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int nx = P.nx, ny = P.ny;
    if (x >= nx) return;
    const size_t k = (size_t)y * nx + x;
    double gAx, gAy, gBx, gBy;
    grad_at(H.uA, y, x, ny, nx, P.delta, gAx, gAy);
    grad_at(H.uB, y, x, ny, nx, P.delta, gBx, gBy);
    const double Fx = -gAy, Fy = gAx;                 // the rotational (divergence-free) part
    H.J_env_x[k] = Fx; H.J_env_y[k] = Fy;
    H.B_field[k] = H.uA[k] * H.mu;                    // ion_current.py:63
    H.Jtx[k] = Fx + gBx; H.Jty[k] = Fy + gBy;         // ion_current.py:67-68
}

static void gemm(const double* A, const double* B, double* C, int M, int N, int K, int batch, size_t sA, size_t sB, size_t sC, cudaStream_t st)
{
    GemmArgs a;
    for (int z = 0; z < 4; ++z) { const int q = z < batch ? z : 0; a.A[z] = A + q * sA; a.B[z] = B + q * sB; a.C[z] = C + q * sC; a.K[z] = K; }
    dim3 g((N + GN - 1) / GN, (M + GM - 1) / GM, batch);
    k_hh_gemm<<<g, 256, GEMM_SMEM, st>>>(a, M, N, K, 1, N, N);
}

// ---- the sine transform at half the arithmetic.  S[n-1-j][k] = (-1)^k S[j][k], so with the even and the odd rows of X
// transformed separately — E = Se . X[0::2], O = So . X[1::2], Se[j][k'] = S[j][2k'], So[j][k'] = S[j][2k'+1], j < h =
// ceil(n/2) — the transform is Y[j] = E[j] + O[j], Y[n-1-j] = E[j] - O[j]: two products of (h x h x N) instead of one of
// (n x n x N).  The same on the right (columns).  Packed matrices: PL [h][n] = [Se | So] for S . X, PR [n][h] = its
// transpose layout (rows 0..h-1: S[2k'][l'], rows h..n-1: S[2k'+1][l']) for X . S.
__global__ void k_hh_sine_packed(double* __restrict__ PL, double* __restrict__ PR, const int n)       // either may be null
{
    const int h = (n + 1) / 2;
    const int q = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;       // j < h: the row of PL / column of PR
    if (q >= n) return;
    const int k = q < h ? 2 * q : 2 * (q - h) + 1;                             // the column of S behind packed column q
    const long long prod = (long long)(j + 1) * (k + 1) % (2LL * (n + 1));     // exact argument reduction
    const double v = sinpi((double)prod / (double)(n + 1));
    if (PL) PL[(size_t)j * n + q] = v;
    if (PR) PR[(size_t)q * h + j] = v;                                          // S is symmetric
}

// Y[j] = E[j] + O[j], Y[n-1-j] = E[j] - O[j] (rows: left transform of an [n][N] array; E, O: [h][N])
__global__ void k_hh_fold_rows(const double* __restrict__ Eb, const double* __restrict__ Ob, double* __restrict__ Yb, const int n, const int N,
                               const size_t sEO, const size_t sY)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (c >= N) return;
    const double* E = Eb + blockIdx.z * sEO; const double* O = Ob + blockIdx.z * sEO; double* Y = Yb + blockIdx.z * sY;
    const double e = E[(size_t)j * N + c], o = O[(size_t)j * N + c];
    Y[(size_t)j * N + c] = e + o;
    if (n - 1 - j != j) Y[(size_t)(n - 1 - j) * N + c] = e - o;
}
// Z[r][l] = E[r][l] + O[r][l], Z[r][n-1-l] = E[r][l] - O[r][l] (columns: right transform of an [M][n] array; E, O: [M][h])
__global__ void k_hh_fold_cols(const double* __restrict__ Eb, const double* __restrict__ Ob, double* __restrict__ Zb, const int M, const int n,
                               const size_t sEO, const size_t sZ)
{
    const int h = (n + 1) / 2;
    const int l = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (l >= h) return;
    const double* E = Eb + blockIdx.z * sEO; const double* O = Ob + blockIdx.z * sEO; double* Z = Zb + blockIdx.z * sZ;
    const double e = E[(size_t)r * h + l], o = O[(size_t)r * h + l];
    Z[(size_t)r * n + l] = e + o;
    if (n - 1 - l != l) Z[(size_t)r * n + (n - 1 - l)] = e - o;
}

// Y = S . X for nb arrays X [n][N] (stride sX) -> Y (stride sY); EO: work, 2 * nb * h * N doubles
static void dst_left(const HHBuf& H, const double* X, double* Y, int n, int N, int nb, size_t sX, size_t sY, cudaStream_t st)
{
    const int h = (n + 1) / 2, l = n / 2;
    const size_t w = (size_t)h * N;
    GemmArgs a;
    for (int z = 0; z < 4; ++z) {
        const int sv = (z < 2 * nb ? z : 0) >> 1, par = z & 1;
        a.A[z] = H.PLy + (par ? h : 0);                       // [Se | So]
        a.B[z] = X + sv * sX + (par ? N : 0);                 // even / odd rows of X
        a.C[z] = H.EO + (size_t)(2 * sv + par) * w;
        a.K[z] = par ? l : h;
    }
    dim3 g((N + GN - 1) / GN, (h + GM - 1) / GM, 2 * nb);
    k_hh_gemm<<<g, 256, GEMM_SMEM, st>>>(a, h, N, n, 1, 2 * N, N);
    k_hh_fold_rows<<<dim3((N + 127) / 128, h, nb), 128, 0, st>>>(H.EO, H.EO + w, Y, n, N, 2 * w, sY);
}
// Z = X . S for nb arrays X [M][n]
static void dst_right(const HHBuf& H, const double* X, double* Z, int M, int n, int nb, size_t sX, size_t sZ, cudaStream_t st)
{
    const int h = (n + 1) / 2, l = n / 2;
    const size_t w = (size_t)M * h;
    GemmArgs a;
    for (int z = 0; z < 4; ++z) {
        const int sv = (z < 2 * nb ? z : 0) >> 1, par = z & 1;
        a.A[z] = X + sv * sX + (par ? 1 : 0);                 // even / odd columns of X
        a.B[z] = H.PRx + (par ? (size_t)h * h : 0);           // rows S[2k'][.] / S[2k'+1][.]
        a.C[z] = H.EO + (size_t)(2 * sv + par) * w;
        a.K[z] = par ? l : h;
    }
    dim3 g((h + GN - 1) / GN, (M + GM - 1) / GM, 2 * nb);
    k_hh_gemm<<<g, 256, GEMM_SMEM, st>>>(a, M, h, n, 2, h, h);
    k_hh_fold_cols<<<dim3((h + 127) / 128, M, nb), 128, 0, st>>>(H.EO, H.EO + w, Z, M, n, 2 * w, sZ);
}

void launch_hh_setup(const HHBuf& H, int ny, int nx, cudaStream_t st)
{
    const int my = ny - 2, mx = nx - 2;
    cudaFuncSetAttribute(k_hh_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    k_hh_sine<<<dim3((my + 127) / 128, my), 128, 0, st>>>(H.Sy, H.ly, my);
    k_hh_sine<<<dim3((mx + 127) / 128, mx), 128, 0, st>>>(H.Sx, H.lx, mx);
    if (H.PLy) {
        k_hh_sine_packed<<<dim3((my + 127) / 128, (my + 1) / 2), 128, 0, st>>>(H.PLy, nullptr, my);      // Sy . X: left
        k_hh_sine_packed<<<dim3((mx + 127) / 128, (mx + 1) / 2), 128, 0, st>>>(nullptr, H.PRx, mx);      // X . Sx: right
    }
}

// one Dirichlet solve (b -> u), or two that share the sine matrices (b2 -> u2 as well) as batched products: H.R / H.T1
// hold two [my][mx] work arrays
static void poisson(const KParams& P, const HHBuf& H, const double* b, double* u, cudaStream_t st, const double* b2 = nullptr, double* u2 = nullptr)
{
    const int ny = P.ny, nx = P.nx, my = ny - 2, mx = nx - 2;
    const int nb = b2 ? 2 : 1;
    const size_t w = (size_t)my * mx;
    dim3 gE((nx + 127) / 128, ny), gI((mx + 127) / 128, my), gIb((mx + 127) / 128, my, nb);
    k_hh_lift<<<gE, 128, 0, st>>>(P, b, u, H.R);
    if (b2) k_hh_lift<<<gE, 128, 0, st>>>(P, b2, u2, H.R + w);
    static const int sym = [] { const char* e = getenv("BETSE_HH_SYM"); return (e && e[0] == '0') ? 0 : 1; }();
    if (sym && H.PLy) {
        dst_left(H, H.R, H.T1, my, mx, nb, w, w, st);             // Sy . R at half the arithmetic (even / odd split)
        dst_right(H, H.T1, H.R, my, mx, nb, w, w, st);            // . Sx
        k_hh_scale<<<gIb, 128, 0, st>>>(H.R, H.ly, H.lx, my, mx);
        dst_left(H, H.R, H.T1, my, mx, nb, w, w, st);
        dst_right(H, H.T1, H.R, my, mx, nb, w, w, st);
    } else {
        gemm(H.Sy, H.R, H.T1, my, mx, my, nb, 0, w, w, st);           // Sy . R
        gemm(H.T1, H.Sx, H.R, my, mx, mx, nb, w, 0, w, st);           // . Sx
        k_hh_scale<<<gIb, 128, 0, st>>>(H.R, H.ly, H.lx, my, mx);
        gemm(H.Sy, H.R, H.T1, my, mx, my, nb, 0, w, w, st);
        gemm(H.T1, H.Sx, H.R, my, mx, mx, nb, w, 0, w, st);
    }
    k_hh_scatter<<<gI, 128, 0, st>>>(H.R, u, my, mx);
    if (b2) k_hh_scatter<<<gI, 128, 0, st>>>(H.R + w, u2, my, mx);
}

// -div_Jb of the boundary-voltage problem (ion_current.py:84-90 / 160-168): zero inside, +bound_V/d^2 on the border;
// assignment order R, L, T, B — the top/bottom rows own the corners
__global__ void k_phi_rhs(const __grid_constant__ KParams P, double* b, const double vT, const double vB, const double vL, const double vR)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int nx = P.nx, ny = P.ny;
    if (x >= nx) return;
    const double d2 = P.delta * P.delta;
    double v = 0.0;
    if (y == 0) v = -(-vB / d2);
    else if (y == ny - 1) v = -(-vT / d2);
    else if (x == 0) v = -(-vL / d2);
    else if (x == nx - 1) v = -(-vR / d2);
    b[(size_t)y * nx + x] = v;
}

// Phi_b = lapENVinv . (-div_Jb): the potential of an externally applied voltage (tissue/event/tisevevolt.py),
// solved when sim.bound_V changes.  Uses the sine matrices and work buffers of the HH diagnostics.
void launch_phi_b(const KParams& P, const HHBuf& H, const double bound[4], double* phi, cudaStream_t st)
{
    dim3 gE((P.nx + 127) / 128, P.ny);
    k_phi_rhs<<<gE, 128, 0, st>>>(P, H.bB, bound[0], bound[1], bound[2], bound[3]);
    poisson(P, H, H.bB, phi, st);
}

// ---------------------------------------------------------------------------- no-ECM field diagnostics
// get_current without extracellular spaces (ion_current.py:116-158): a local field potential from Vmem/2 scattered to
// the env squares of the membranes, smoothed by gaussian_filter(sigma=1) in scipy's default 'reflect' mode, its gradient
// smoothed through a Helmholtz-Hodge decomposition and reconstruction (stb.smooth_flux, sim_toolbox.py:1318-1333), the
// bath current sigma*E*D_env_weight decomposed again.  Nothing of it feeds back into the timestep (update_ecm only runs
// with ECM); write2storage stores v_env and J_env (sim.py:1826, 1866-1867), so it runs at sampled steps.
__global__ void k_ne_scatter(const __grid_constant__ KParams P, const KArrays A, double* __restrict__ phi, const int old)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int E = P.nx * P.ny;
    if (q >= E) return;
    const int s0 = A.slot_ptr[q], s1 = A.slot_ptr[q + 1];
    double v = 0.0;
    if (s1 > s0) {
        // vce[cells.map_mem2ecm] = vc[cells.mem_to_cells] (ion_current.py:121): fancy assignment, the last membrane of
        // the square wins; sic — a per-MEMBRANE array indexed by the membrane's CELL index
        // update_V calls get_current BEFORE it renews Vmem (sim.py:2010-2013): this is the Vmem the step started with
        const int m = A.slot_idx[s1 - 1];
        const int mm = A.mem_to_cells[m];                    // used as a MEMBRANE index
        double vm;
        if (P.polar) vm = A.vm_pol[old][mm];
        else {
            vm = A.vm_cell[old][A.mem_to_cells[mm]];
            if (P.has_phi) vm -= A.phi_b_old[A.map_mem2ecm[mm]];
        }
        v = vm / 2.0;
    }
    phi[q] = -v;
}

__device__ __forceinline__ int refl(int i, const int n)          // scipy 'reflect': d c b a | a b c d | d c b a
{
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - i - 1;
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

// one separable pass of gaussian_filter(sigma=1, mode='reflect'); axis 0 = rows (y), then axis 1 (x), like scipy
__global__ void k_ne_gauss(const __grid_constant__ KParams P, const double* __restrict__ in, double* __restrict__ out, const int axis)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int nx = P.nx, ny = P.ny;
    if (x >= nx) return;
    const int n = axis == 0 ? ny : nx, i0 = axis == 0 ? y : x;
    auto at = [&](int i) { i = refl(i, n); return axis == 0 ? in[(size_t)i * nx + x] : in[(size_t)y * nx + i]; };
    double s = at(i0) * P.gw[0];                       // scipy correlate1d, symmetric form (as kernels.cu:k_field)
    s += (at(i0 - 4) + at(i0 + 4)) * P.gw[4];
    s += (at(i0 - 3) + at(i0 + 3)) * P.gw[3];
    s += (at(i0 - 2) + at(i0 + 2)) * P.gw[2];
    s += (at(i0 - 1) + at(i0 + 1)) * P.gw[1];
    out[(size_t)y * nx + x] = s;
}

// F = delta*Phi (scaled in place of a copy), then fd.gradient(F, delta) -> (Jx, Jy) = the field HH smooths
__global__ void k_ne_scale(double* __restrict__ dst, const double* __restrict__ src, const double f, const int n)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) dst[q] = f * src[q];
}
__global__ void k_ne_grad(const __grid_constant__ KParams P, const double* __restrict__ F, double* __restrict__ gx, double* __restrict__ gy)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= P.nx) return;
    double a, b;
    grad_at(F, y, x, P.ny, P.nx, P.delta, a, b);
    gx[(size_t)y * P.nx + x] = a; gy[(size_t)y * P.nx + x] = b;
}
// E = -(smoothed gradient); J = sigma*E*D_env_weight (ion_current.py:138-143)
__global__ void k_ne_EJ(const __grid_constant__ KParams P, const KArrays A, HHBuf H, const double sigma, const double* __restrict__ W)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.nx * P.ny) return;
    const double ex = -H.Jtx[q], ey = -H.Jty[q];
    A.E_x[q] = ex; A.E_y[q] = ey;
    H.Jx[q] = (sigma * ex) * W[q];
    H.Jy[q] = (sigma * ey) * W[q];
}

static void hh_core(const KParams& P, const HHBuf& H, cudaStream_t st)
{
    dim3 gE((P.nx + 127) / 128, P.ny);
    k_hh_rhs<<<gE, 128, 0, st>>>(P, H);
    poisson(P, H, H.bA, H.uA, st, H.bB, H.uB);
    k_hh_out<<<gE, 128, 0, st>>>(P, H);
}

// H.bound must be zero (HH_Decomp without bounds, ion_current.py:134, 145); work: H.uA / H.uB double as Gaussian scratch
void launch_noecm_field(const KParams& P, const KArrays& A, const HHBuf& H, double sigma, const double* D_env_weight, int old, cudaStream_t st)
{
    const int E = P.nx * P.ny;
    dim3 gE((P.nx + 127) / 128, P.ny);
    k_ne_scatter<<<(E + 255) / 256, 256, 0, st>>>(P, A, H.uA, old);
    k_ne_gauss<<<gE, 128, 0, st>>>(P, H.uA, H.uB, 0);
    k_ne_gauss<<<gE, 128, 0, st>>>(P, H.uB, A.v_env, 1);                   // sim.v_env = Phi (ion_current.py:158)
    k_ne_scale<<<(E + 255) / 256, 256, 0, st>>>(H.uA, A.v_env, P.delta, E);
    k_ne_grad<<<gE, 128, 0, st>>>(P, H.uA, H.Jx, H.Jy);
    hh_core(P, H, st);                                                      // stb.smooth_flux: Jtx/Jty = Fr + Fd
    k_ne_EJ<<<(E + 255) / 256, 256, 0, st>>>(P, A, H, sigma, D_env_weight);
    hh_core(P, H, st);
}

void launch_hh(const KParams& P, const KArrays& A, const HHBuf& H, cudaStream_t st)
{
    const int E = P.nx * P.ny;
    dim3 gE((P.nx + 127) / 128, P.ny);
    k_hh_J<<<(E + 255) / 256, 256, 0, st>>>(P, A, H);
    k_hh_rhs<<<gE, 128, 0, st>>>(P, H);
    poisson(P, H, H.bA, H.uA, st, H.bB, H.uB);
    k_hh_out<<<gE, 128, 0, st>>>(P, H);
}
