// Arithmetic helpers shared by the membrane kernels (kernels.cu: k_mem, kmem_pipe.cu: k_mem_pipe).
#pragma once
#include "kparams.cuh"

#define FLOAT_NONCE 1.0e-25   // sim_toolbox.py:52

#define ST_NAN_VM 1u
#define ST_NAN_CONC 2u
#define ST_NEG 4u

__device__ __forceinline__ double ldg(const double* p) { return __ldg(p); }
__device__ __forceinline__ int ldgi(const int* p) { return __ldg(p); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// 1/x for finite, normal, non-zero x: MUFU.RCP64H seed (~20 bits) + two Newton steps; <= 1 ulp,
// branch-free (the compiler's IEEE division carries a slow-path call per use).
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
// a/b with one residual correction (<= 1 ulp)
__device__ __forceinline__ double fast_div(double a, double b)
{
    const double r = fast_rcp(b);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}

// GHK flux in "A/B form".  With alpha = z*a1, ex = exp(-alpha), rden = 1/(-expm1(-alpha)) the
// reference's  -((D*alpha)/d)*((cB - cA*ex)*rden)  (sim_toolbox.py:58-65) equals
// -(D/d)*(cB*A - cA*B) with A = alpha*rden, B = A*ex.  For the valences that occur (+-1, +-2)
// everything follows from ONE expm1 at -a1:  A(+1) = a1/(-em1), B(+1) = A(+1)*e1,
// A(+2) = 2*a1/((-em1)*(e1+1)), B(+2) = A(+2)*e1^2, and A(-z) = B(+z), B(-z) = A(+z).
struct GhkAB {
    double A1, B1, A2, B2;
    __device__ __forceinline__ void init(double a1, double e1, double em1) {
        const double e1p1 = e1 + 1.0;
        const double R = fast_rcp((-em1) * e1p1);     // 1/((-em1)(e1+1))
        A2 = (2.0 * a1) * R;
        A1 = (a1 * R) * e1p1;
        B1 = A1 * e1;
        B2 = A2 * (e1 * e1);
    }
};

// exp(x) and expm1(x) from ONE branch-free evaluation: x = k*ln2 + r, |r| <= ln2/2,
// e^r - 1 = r*q(r) with q the degree-12 Taylor polynomial (truncation 5e-16 relative), Estrin form
// (dependency depth 5 instead of 12); exp = s*(1 + r*q), expm1 = s*r*q + (s - 1), s = 2^k.
// Measured against NumPy over [-40, 40] and 1e-20..1: <= 2 ulp for exp, <= 2.5 ulp for expm1 —
// four to five orders of magnitude inside the 1e-10 parity bar.  The CUDA library's exp/expm1 carry a
// range branch each, which keeps the compiler from interleaving the independent evaluations.
// Arguments are clamped to [-708, 709] (NaN passes through).
static __constant__ double c_expq[13] = {
    1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
    1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0};

struct ExpParts {
    double s, rq;
    __device__ __forceinline__ double exp() const { return fma(s, rq, s); }
    __device__ __forceinline__ double expm1() const { return fma(s, rq, s - 1.0); }
};

__device__ __forceinline__ ExpParts exp_parts(double x)
{
    x = (x > 709.0) ? 709.0 : x;
    x = (x < -708.0) ? -708.0 : x;
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double p01 = fma(c_expq[1], r, c_expq[0]), p23 = fma(c_expq[3], r, c_expq[2]);
    const double p45 = fma(c_expq[5], r, c_expq[4]), p67 = fma(c_expq[7], r, c_expq[6]);
    const double p89 = fma(c_expq[9], r, c_expq[8]), pab = fma(c_expq[11], r, c_expq[10]);
    const double q0 = fma(p23, r2, p01), q1 = fma(p67, r2, p45), q2 = fma(pab, r2, p89);
    const double qh = fma(c_expq[12], r4, q2);
    const double q = fma(qh, r8, fma(q1, r4, q0));
    ExpParts o;
    o.rq = r * q;
    o.s = __hiloint2double((k + 1023) << 20, 0);
    return o;
}

// GHK table {A(+1), B(+1), A(+2), B(+2)} of one membrane side from alpha(z=+1); returns exp(-a1)
__device__ __forceinline__ double ghk_table(double a1, GhkAB& t)
{
    const ExpParts p = exp_parts(-a1);
    const double e1 = p.exp();
    t.init(a1, e1, p.expm1());
    return e1;
}

// Gap-junction gating (channels/gap_junction.py:56-72) as the affine map g' = g*c1 + c2 that one
// implicit-Euler sub-step applies (it is applied once per ion, sim.py:1272 -> 2180-2183):
//   al = k0*exp(-0.077 xg), be0 = k0*exp(0.14 xg), D1 = 1 + 50 be0,
//   g' = gjb*((g + dtm*al)*D1 + dtm*be0*gmin) / (D1*(1 + dtm*al) + dtm*be0).
// With u = exp(0.007 xg): be0 = k0*u^20, al = k0/u^11; numerator and denominator are multiplied by
// u^11 so that one exp and one reciprocal serve the whole membrane.
__device__ __forceinline__ void gj_gate_map(double vgj0, const KParams& P, double gjb, double& c1, double& c2)
{
    const double xg = 1.0e3 * fabs(vgj0) - P.gj_vthresh;
    const double u = exp_parts(0.007 * xg).exp();
    const double u2 = u * u, u4 = u2 * u2, u5 = u4 * u, u10 = u5 * u5;
    const double U = u10 * u;                        // u^11
    const double be0 = 0.0013 * (u10 * u10);         // u^20
    const double D1 = fma(50.0, be0, 1.0);
    const double ka = P.dtm * 0.0013;                // dtm*al*U
    const double kb = P.dtm * be0;
    const double den = fma(D1, U + ka, kb * U);
    const double rd = gjb * fast_rcp(den);
    c1 = (U * D1) * rd;
    c2 = fma(ka, D1, (kb * P.gj_min) * U) * rd;
}

// generic valence (incl. 0: the reference adds 1e-25 to z, sim_toolbox.py:56); rare, kept out of line
static __device__ __noinline__ double2 ghk_generic(double z, double a1)
{
    const double al = (z + FLOAT_NONCE) * a1;
    const double a = al / (-expm1(-al));
    return make_double2(a, a * exp(-al));
}

// The reference's ion order is fixed (Na, K, Cl, Ca, H, P, M; parameters.py:1296), so the shipped
// ion profiles give three valence signatures.  PROF = 1 builds bake the signature of the NI-ion
// profile plus the default feature switches (extracellular spaces, voltage-sensitive gap
// junctions, open cluster boundary, no per-membrane block arrays, no diagnostics) into the
// kernel; PROF = 0 reads everything from KParams at run time.
template <int NI> struct StdProf;
template <> struct StdProf<4> { static constexpr int iNa = 0, iK = 1, iCa = -1; __host__ __device__ static constexpr int z(int i) { constexpr int t[4] = {1, 1, -1, -1}; return t[i]; } };       // basic: Na K P M
template <> struct StdProf<5> { static constexpr int iNa = 0, iK = 1, iCa = 2; __host__ __device__ static constexpr int z(int i) { constexpr int t[5] = {1, 1, 2, -1, -1}; return t[i]; } };     // basic_Ca: Na K Ca P M
template <> struct StdProf<6> { static constexpr int iNa = 0, iK = 1, iCa = 3; __host__ __device__ static constexpr int z(int i) { constexpr int t[6] = {1, 1, -1, 2, -1, -1}; return t[i]; } };  // mammal/amphibian/custom: Na K Cl Ca P M
template <> struct StdProf<7> { static constexpr int iNa = 0, iK = 1, iCa = 3; __host__ __device__ static constexpr int z(int i) { constexpr int t[7] = {1, 1, -1, 2, 1, -1, -1}; return t[i]; } }; // + H
template <> struct StdProf<8> { static constexpr int iNa = 0, iK = 1, iCa = 3; __host__ __device__ static constexpr int z(int i) { return 0; } };

__device__ __forceinline__ void ghk_pick(const GhkAB& t, int zi, double& A, double& B)
{
    const bool two = (zi == 2) || (zi == -2);
    const double X = two ? t.A2 : t.A1, Y = two ? t.B2 : t.B1;
    A = (zi < 0) ? Y : X;
    B = (zi < 0) ? X : Y;
}

// ---------------------------------------------------------------------------- shared membrane math
// Every membrane kernel (kernels.cu:k_mem, kmem_pipe.cu:k_mem_pipe, kcell.cu:k_cell) forms its fluxes through the
// functions below, split into the part that depends on the CELL only (its Vmem and concentrations: computed once per
// membrane by the lane-per-membrane kernels, once per cell by k_cell) and the part that depends on the membrane (env
// square, gap-junction partner).  Same functions, same operand order => bit-identical results across the kernels;
// products that cross the cell/membrane boundary are rounded explicitly (__dmul_rn) so that the compiler's FMA
// contraction cannot differ between a value used once (per membrane) and six times (per cell).

// membrane side of a cell: GHK table at alpha(z=+1) of the cell's Vmem and the pumps' equilibrium constant
// Keq = exp(-dG/RT + F vm/RT) = K0/e1 (sim_toolbox.py:54-65, 96-100)
struct MemSide { GhkAB t; double keq; };
__device__ __forceinline__ double mem_side(double vm_own, const KParams& P, MemSide& s)
{
    const double a1 = ((vm_own + FLOAT_NONCE) * P.F) * P.inv_RT_sim;
    s.keq = P.K0 * fast_rcp(ghk_table(a1, s.t));
    return a1;
}

// Na/K-ATPase (sim_toolbox.py:71-122):  f_Na = -3 blk alpha fwd (1 - Q/Keq),
//   fwd = u3 w2 t / ((1+u3)(1+w2)(1+t)),  Q = Qn/Qd  =>  f_Na = -3 blk alpha (u3 t w2)(Qd Keq - Qn) / ((1+u3)(1+t)(1+w2) Qd Keq)
// cell part: everything of cNai, cKi;  membrane part: cNao, cKo (the membrane's env square)
struct NaKCell { double bb, Qd3, u3t, dct; };
__device__ __forceinline__ void nak_cell(double cNai, double cKi, const KParams& P, NaKCell& c)
{
    const double b = cKi * 1e-3;
    c.bb = __dmul_rn(b, b);
    const double a2 = cNai * 1e-3;
    c.Qd3 = __dmul_rn(P.QdNK0, __dmul_rn(__dmul_rn(a2, a2), a2));
    const double u = cNai * P.inv_KmNK_Na;
    const double u3 = __dmul_rn(__dmul_rn(u, u), u);
    c.u3t = __dmul_rn(u3, P.tNK);
    c.dct = __dmul_rn(1.0 + u3, 1.0 + P.tNK);
}
// returns f_Na before rho_pump (its negative is sim.rate_NaKATP)
__device__ __forceinline__ double nak_flux(const NaKCell& c, double keq, double cNao, double cKo, double blk, const KParams& P)
{
    const double a = cNao * 1e-3;
    const double Qn = __dmul_rn(__dmul_rn(P.QnNK0, __dmul_rn(__dmul_rn(a, a), a)), c.bb);
    const double b2 = cKo * 1e-3;
    double Qd = __dmul_rn(c.Qd3, __dmul_rn(b2, b2));
    if (Qd == 0.0) Qd = 1.0e-15;
    const double QdK = __dmul_rn(Qd, keq);
    const double w = cKo * P.inv_KmNK_K;
    const double w2 = __dmul_rn(w, w);
    const double num = __dmul_rn(__dmul_rn(c.u3t, w2), __dsub_rn(QdK, Qn));
    const double den = __dmul_rn(__dmul_rn(c.dct, 1.0 + w2), QdK);
    return ((-3.0 * blk) * P.alpha_NaK) * fast_div(num, den);
}

// Ca-ATPase (sim.py:2126-2155, sim_toolbox.py:124-182):  f = -alpha (x t/((1+x)(1+t))) (1 - Qn/(Qd Keq)),
// Keq = K0/e1^2; only Qn = cADP cPi cCao depends on the membrane.  cCai is the fresh cell value (after no_negs).
struct CaCell { double g1, g2; };                   // -alpha x t/((1+x)(1+t)),  1/(Qd Keq)
__device__ __forceinline__ void ca_cell(double cCai, double keq, const KParams& P, CaCell& c)
{
    double Qd = P.cATP * cCai;
    if (Qd == 0.0) Qd = 1.0e-16;
    const double QdK = Qd * ((keq * keq) * P.inv_K0);
    c.g2 = fast_rcp(QdK);
    const double x = cCai * P.inv_KmCa_Ca;
    c.g1 = -P.alpha_Ca * fast_div(x * P.tCa, (1.0 + x) * (1.0 + P.tCa));
}
__device__ __forceinline__ double ca_flux(const CaCell& c, double cCao, const KParams& P)
{
    const double Qn = __dmul_rn(P.QnCa0, cCao);
    return __dmul_rn(c.g1, fma(-Qn, c.g2, 1.0));
}

// one ion through one membrane, already times the membrane area: DmS = (Dm*(-rho_channel/tm))*mem_sa,
// cinAm = cin*A (cell part), co*B (membrane part)  (sim_toolbox.py:58-65 in A/B form)
__device__ __forceinline__ double ghk_mem_flux(double DmS, double cinAm, double co, double Bm)
{
    return __dmul_rn(DmS, fma(-co, Bm, cinAm));
}
// gap-junction flux times area; gsa = g*sa (0 at boundary membranes), cA = this cell, cB = partner (sim.py:2191-2197)
__device__ __forceinline__ double ghk_gj_flux(double Dgj_len, double gsa, double cnb, double Ag, double cin, double Bg)
{
    return __dmul_rn(-__dmul_rn(Dgj_len, gsa), fma(cnb, Ag, -__dmul_rn(cin, Bg)));
}

// update_Co + update_all_concs of one (cell, ion) (sim_toolbox.py:1177-1181, sim.py:2105-2108): cm = the stale
// cc_at_mem (after the membrane fluxes), cn = the new cell concentration (after the gap-junction fluxes, before no_negs)
__device__ __forceinline__ void cell_conc_update(double cc, double Sm, double Sg, double rvol, double dt, double& cm_new, double& cn_new)
{
    cm_new = fma(__dmul_rn(Sm, rvol), dt, cc);
    cn_new = fma(dt, __dmul_rn(-Sg, rvol), cm_new);
}
