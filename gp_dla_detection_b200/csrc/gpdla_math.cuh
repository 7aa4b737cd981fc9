// Device math for the DLA hot path: Lyman-series line data, FP64 Voigt/Faddeeva evaluation,
// fast reciprocal.  sm_100a only.
//
// Replaces the arithmetic of the reference's voigt.c:277-292 (per-pixel sum over lines of
// leading_constant * voigt(velocity, sigma, gamma), libcerf call at voigt.c:288) with a
// branch-light scheme suited to SIMT: see tools/gen_voigt_tables.py for the derivation and
// the fitted tables (voigt_tables.h).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "voigt_tables.h"

namespace gpdla {

constexpr int MAX_LINES = 31;  // voigt.c:16

// Per-line constants, filled on the host by init_line_constants() (gpdla_capi.cu) from the
// Lyman-series data; layout shared with the host.
struct LineConstants {
  double tw[MAX_LINES];     // transition wavelengths (cm)              voigt.c:31-64
  double lc[MAX_LINES];     // leading constants                         voigt.c:141-184
  double gam[MAX_LINES];    // Lorentzian widths (cm/s)                  voigt.c:186-220
  double y[MAX_LINES];      // gam / (sqrt2 sigma)
  double y2[MAX_LINES];     // y^2
  double kcore[MAX_LINES];  // lc / (sqrt(2 pi) sigma)
  double kwing[MAX_LINES];  // kcore * y / sqrt(pi)
  double ip[7];             // instrument profile                        voigt.c:242-251
  double c;                 // speed of light (cm/s)                     voigt.c:22
  double inv_s2s;           // 1 / (sqrt2 sigma)
};

__constant__ LineConstants c_lines;
__constant__ double c_wing_a[GPDLA_VOIGT_DEG_A + 1] = GPDLA_VOIGT_WING_A;
__constant__ double c_wing_b[GPDLA_VOIGT_DEG_B + 1] = GPDLA_VOIGT_WING_B;
// core table lives in global memory (divergent indexing; constant cache would serialise)
__device__ const double g_core_table[GPDLA_VOIGT_NINT * GPDLA_VOIGT_CORE_STRIDE] = GPDLA_VOIGT_CORE_TABLE;

// 1/x to ~1 ulp for normal positive x: MUFU.RCP64H seed (>= 20 good bits) + one cubically convergent
// step r (1 + e + e^2), e = 1 - x r  (3 DFMA; the MUFU runs on its own pipe).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  const double t = fma(e, e, e);
  return fma(r, t, r);
}

// Comparisons of non-negative doubles through their high words: integer pipe instead of the FP64 pipe
// (exact unless the operands agree in their top 32 bits, where the caller's threshold is soft anyway).
__device__ __forceinline__ bool hi_less(double a, double b) { return __double2hiint(a) < __double2hiint(b); }

__device__ __forceinline__ double wing_poly_a(double u) {
  double a = c_wing_a[GPDLA_VOIGT_DEG_A];
#pragma unroll
  for (int i = GPDLA_VOIGT_DEG_A - 1; i >= 0; --i) a = fma(a, u, c_wing_a[i]);
  return a;
}
__device__ __forceinline__ double wing_poly_b(double u) {
  double b = c_wing_b[GPDLA_VOIGT_DEG_B];
#pragma unroll
  for (int i = GPDLA_VOIGT_DEG_B - 1; i >= 0; --i) b = fma(b, u, c_wing_b[i]);
  return b;
}

// tau_j / N for |x| >= X0 given u = 1/x^2.
__device__ __forceinline__ double tau_wing(int j, double u) {
  double a = wing_poly_a(u);
  double b = wing_poly_b(u);
  double t = c_lines.y2[j] * u;
  return c_lines.kwing[j] * u * fma(-t, b, a);
}

// exp(x) for x <= 0 (and small positive x), branch-free.  The argument is clamped at -708.
//   GPDLA_EXP_TABLE (default): x = (256 n + j) ln2 / 256 + r, |r| <= ln2 / 512: exp(x) = 2^n T[j] (1 + q(r)) with the
//   256-entry table T[j] = 2^(j/256) (correctly rounded, 2 KB, L1-resident) and q = r + r^2/2 + r^3/6 + r^4/24
//   (truncation 3.8e-17 relative): 9 FP64 instructions instead of 15 -- FP64 issue is what bounds the producers.
//   Otherwise: Cody-Waite reduction to |r| <= ln2 / 2 and a degree-11 polynomial (tools/gen_voigt_tables.py).
// Both assemble the exponent in integer registers and agree with libm to ~1 ulp (tests/test_gpu_parity.py, 5e-15).
__constant__ double c_exp_poly[12] = GPDLA_EXP_POLY;
__device__ double g_exp2_tab[256];   // 2^(j/256), filled by upload_device_constants
#ifndef GPDLA_EXP_TABLE
#define GPDLA_EXP_TABLE 1
#endif
// TABLE = false selects the polynomial: the persistent INT8 kernel is bound by issue slots and dependency chains, not by
// FP64 throughput, and measured 0.5 % (table in shared memory) to 2.5 % (table through L1) slower with the table, while
// the FP64 DMMA kernels gain 1.6 % from it.
template <bool TABLE = (GPDLA_EXP_TABLE != 0)>
__device__ __forceinline__ double exp_nonpos(double x) {
  const double SHIFT = 6755399441055744.0;   // 1.5 * 2^52
  // x <= 0 (or tiny positive): "below -708" is an unsigned compare of the high word (integer pipe)
  const bool tiny = (unsigned)__double2hiint(x) > 0xC0862000u;   // x < -708 (hi word of -708.0 is 0xC0862000)
  const double xc = tiny ? -708.0 : x;
  if (TABLE) {
  double kd = fma(xc, 256.0 * 1.4426950408889634074, SHIFT);
  const int k = __double2loint(kd);
  kd -= SHIFT;
  double r = fma(kd, -6.93147180369123816490e-01 / 256.0, xc);
  r = fma(kd, -1.90821492927058770002e-10 / 256.0, r);
  const double T = __ldg(&g_exp2_tab[k & 255]);
  double a = fma(r, 1.0 / 24.0, 1.0 / 6.0);
  a = fma(a, r, 0.5);
  a = fma(a, r, 1.0);
  const double p = fma(T, a * r, T);
  return __hiloint2double(__double2hiint(p) + ((k >> 8) << 20), __double2loint(p));
  }
  double kd = fma(xc, 1.4426950408889634074, SHIFT);
  const int k = __double2loint(kd);
  kd -= SHIFT;
  double r = fma(kd, -6.93147180369123816490e-01, xc);
  r = fma(kd, -1.90821492927058770002e-10, r);
  double p = c_exp_poly[11];
#pragma unroll
  for (int i = 10; i >= 0; --i) p = fma(p, r, c_exp_poly[i]);
  // below -708 the result is exp(-708) = 3.3e-308 instead of the reference's libm value in [0, 2.3e-308]
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// tau_j / N for |x| < X0 (rare: <= 7 pixels per line per sample).  The warp that lands here delays its whole
// cluster (the chunk hand-over is lock-step), so the dependency chains are kept short: H1 as even + odd halves,
// H3 and the exponential as independent chains.
__device__ __noinline__ double tau_core(int j, double x) {
  double ax = fabs(x);
  int idx = (int)(ax * GPDLA_VOIGT_INV_H);
  idx = idx < GPDLA_VOIGT_NINT ? idx : GPDLA_VOIGT_NINT - 1;
  double t = fma(ax, 2.0 * GPDLA_VOIGT_INV_H, -(2.0 * idx + 1.0));   // (ax - centre)/(H/2)
  const double t2 = t * t;
  const double* tab = g_core_table + idx * GPDLA_VOIGT_CORE_STRIDE;
  static_assert(GPDLA_VOIGT_DEG_H1 % 2 == 1, "even/odd split of H1");
  double h1e = tab[GPDLA_VOIGT_DEG_H1 - 1], h1o = tab[GPDLA_VOIGT_DEG_H1];
#pragma unroll
  for (int i = GPDLA_VOIGT_DEG_H1 - 3; i >= 0; i -= 2) { h1e = fma(h1e, t2, tab[i]); h1o = fma(h1o, t2, tab[i + 1]); }
  const double h1 = fma(h1o, t, h1e);
  const double* tab3 = tab + GPDLA_VOIGT_DEG_H1 + 1;
  double h3 = tab3[GPDLA_VOIGT_DEG_H3];
#pragma unroll
  for (int i = GPDLA_VOIGT_DEG_H3 - 1; i >= 0; --i) h3 = fma(h3, t, tab3[i]);
  double x2 = x * x;
  double e = exp_nonpos(-x2);
  double y = c_lines.y[j], y2 = c_lines.y2[j];
  double p4 = fma(x2, fma(x2, 4.0, -12.0), 3.0) * (1.0 / 6.0);
  double even = fma(y2, fma(y2, p4, fma(-2.0, x2, 1.0)), 1.0);   // 1 + y^2 (1-2x^2) + y^4 p4
  double rew = fma(e, even, y * fma(y2, h3, h1));
  return c_lines.kcore[j] * rew;
}

// multiplier of voigt.c:279, same operation order: c / (tw * (1 + z)) / 1e8
__device__ __forceinline__ double line_multiplier(int j, double z) {
  return c_lines.c / (c_lines.tw[j] * (1.0 + z)) / 1e8;
}

// Sum over lines of tau_j / N at wavelength `lambda` (Angstrom) -- generic line count, one
// reciprocal per line.  mult[j * mstride] are the per-sample multipliers.
__device__ __forceinline__ double tau_sum_generic(double lambda, const double* mult, int mstride, int num_lines) {
  double total = 0.0;
  for (int j = 0; j < num_lines; ++j) {
    double v = __dsub_rn(__dmul_rn(lambda, mult[j * mstride]), c_lines.c);   // voigt.c:287 (no FMA contraction)
    double x = v * c_lines.inv_s2s;
    double x2 = x * x;
    double tj;
    if (x2 >= GPDLA_VOIGT_X0 * GPDLA_VOIGT_X0) tj = tau_wing(j, fast_rcp(x2));
    else tj = tau_core(j, x);
    total += tj;
  }
  return total;
}

// Three-line specialisation (num_lines = 3, set_parameters.m:63), branch-free wing part.
// Works in velocity units: with q = 1/v^2 (v in cm/s), u = 2 sigma^2 q and
//   tau_j / N = q [ A_j(q) - yy_j q B(u) ],   A_j(q) = 2 sigma^2 kwing_j A(2 sigma^2 q),  yy_j = 2 sigma^2 y_j^2 kwing_j 2 sigma^2
// with the per-line coefficients c_wing3 precomputed on the host (one multiply less per line, no
// separate x).  In the wings the velocity may use a fused multiply-add (relative effect on tau
// < 2e-13 at |x| = X0, far below the parity budget); the exact routine keeps voigt.c:287's rounding.
// Sets `core` when some line is within X0 Doppler widths, in which case the caller re-evaluates
// with tau_sum_3_exact().  Free of control flow so that several samples interleave.
struct Wing3 {
  double a[3][GPDLA_VOIGT_DEG_A + 1];   // A_j coefficients in q
  double yy[3];                          // coefficient of the y^3 H3 correction
  double b1;                             // B(u) ~ 1 + b1 q  (q units)
  double v2min;                          // X0^2 * 2 sigma^2: clamp / core threshold on v^2
};
__constant__ Wing3 c_wing3;

__device__ __forceinline__ double tau_wing3_line(int j, double q) {
  double a = c_wing3.a[j][GPDLA_VOIGT_DEG_A];
#pragma unroll
  for (int i = GPDLA_VOIGT_DEG_A - 1; i >= 0; --i) a = fma(a, q, c_wing3.a[j][i]);
  const double b = fma(c_wing3.b1, q, 1.0);
  const double t = c_wing3.yy[j] * q;
  return q * fma(-t, b, a);
}

__device__ __forceinline__ double tau_sum_3_wing(double lambda, double m0, double m1, double m2, bool& core) {
  const double v0 = fma(lambda, m0, -c_lines.c), v1 = fma(lambda, m1, -c_lines.c), v2 = fma(lambda, m2, -c_lines.c);
  const double s0 = v0 * v0, s1 = v1 * v1, s2 = v2 * v2;
  const double lim = c_wing3.v2min;
  // a lane inside a line core is re-evaluated by the caller, so its wing value may be anything (inf / NaN when
  // v = 0): no clamp of s before the reciprocal
  core = hi_less(s0, lim) | hi_less(s1, lim) | hi_less(s2, lim);
  const double q0 = fast_rcp(s0), q1 = fast_rcp(s1), q2 = fast_rcp(s2);
  return (tau_wing3_line(0, q0) + tau_wing3_line(1, q1)) + tau_wing3_line(2, q2);
}

// Three-line sum for lanes with a line core in reach (rare path; same summation order as voigt.c:285-290).  The wing
// terms of all three lines are evaluated first, branch-free (three independent chains, same formula as the fast
// path); then only the lines that are within X0 Doppler widths -- normally one, the same for every lane of the warp
// -- are replaced by the core evaluation, with voigt.c:287's exact rounding of the velocity.
__device__ __noinline__ double tau_sum_3_exact(double lambda, double m0, double m1, double m2) {
  const double v0 = fma(lambda, m0, -c_lines.c), v1 = fma(lambda, m1, -c_lines.c), v2 = fma(lambda, m2, -c_lines.c);
  const double s0 = v0 * v0, s1 = v1 * v1, s2 = v2 * v2;
  const double lim = c_wing3.v2min;
  const bool k0 = hi_less(s0, lim), k1 = hi_less(s1, lim), k2 = hi_less(s2, lim);
  double t0 = tau_wing3_line(0, fast_rcp(s0)), t1 = tau_wing3_line(1, fast_rcp(s1)), t2 = tau_wing3_line(2, fast_rcp(s2));
  if (k0) t0 = tau_core(0, __dsub_rn(__dmul_rn(lambda, m0), c_lines.c) * c_lines.inv_s2s);
  if (k1) t1 = tau_core(1, __dsub_rn(__dmul_rn(lambda, m1), c_lines.c) * c_lines.inv_s2s);
  if (k2) t2 = tau_core(2, __dsub_rn(__dmul_rn(lambda, m2), c_lines.c) * c_lines.inv_s2s);
  return (t0 + t1) + t2;
}

// ------------------------------------------------------------------------------------------
// Rest-frame table of tau / N (gpdla_rest_table.h): the raw absorption of voigt.c:282-291 depends on the pixel and
// the sample only through the absorber rest-frame wavelength, so away from the line cores the three (or num_lines)
// Voigt evaluations per (sample, pixel) collapse into one polynomial per 1e-4-dex cell.
//   u = ln(lambda_obs / lambda_ref) / h - ln((1 + z) lambda_lo / lambda_ref) / h = lh_pixel - K_sample
//   tau / N = sum_p coef[p][cell] (u - cell)^p        for any cell with |u - cell| <= 1
// The samples a producer warp interleaves are neighbours in redshift (K within RT_MAX_SPREAD of each other), so one
// cell per (warp, pixel) -- the one nearest to lh - K_mid -- serves them all: its coefficients are fetched once,
// a chunk ahead, and each sample evaluates the polynomial at its own s.  NaN coefficients mark the cells near a line
// centre (and the two end cells, which catch clamped indices): the warp then evaluates directly.
constexpr int RT_DEG_DEV = 8;
// Device layout of the table: five planes of coefficient PAIRS, [5][ncell] double2 = (c0,c1), (c2,c3), (c4,c5), (c6,c7),
// (c8,-).  A warp's lanes read neighbouring cells, so a 128-bit load of one plane is one contiguous 512-byte run: 4
// LSU wavefronts.  (Round 2 first kept the 9 coefficients of a cell side by side -- one address, immediate offsets, but
// an 80-byte stride between lanes: 20 wavefronts per load, 100 per cell fetch against 18 now; the kernel is bound by
// LSU wavefronts, shared and global together: profiles/r02f_ncu_loglik_i8_full.json.)
constexpr int RT_PLANES = 5;
constexpr double RT_MAX_SPREAD = 1.0;   // = 2 (RT_HALF_WIDTH - 1/2): largest K_max - K_min one cell can serve
struct RestTable {
  const double* coef;   // [RT_PLANES][ncell] double2; nullptr = table disabled
  int ncell;
  double inv_h;         // 1 / (pixel spacing in ln units)
  double lam_lo;        // rest wavelength of cell 0 (Angstrom)
};
struct RestCell {       // one lane's cell for the current chunk
  double cf[RT_DEG_DEV + 1];
  double base;          // lh - cell: s = base - K
};

// K of the formula above for a sample at redshift z in a quasar whose padded grid starts at lam_ref
__device__ __forceinline__ double rest_table_offset(const RestTable& rt, double z, double lam_ref) {
  return log((1.0 + z) * (rt.lam_lo / lam_ref)) * rt.inv_h;
}

// issue the loads of the cell nearest to lh - K_mid (no use of the loaded values here: the caller consumes them a
// chunk later, so the L2 / L1 latency stays off the critical path)
__device__ __forceinline__ void rest_table_fetch(const RestTable& rt, double lh, double K_mid, RestCell& rc) {
  const double RT_MAGIC = 6755399441055744.0;   // 1.5 * 2^52: x + MAGIC has ulp 1
  const double um = (lh - K_mid) + RT_MAGIC;
  int ci = __double2loint(um);                  // round-to-nearest-even integer of lh - K_mid
  rc.base = lh - (um - RT_MAGIC);               // exact
  ci = min(max(ci, 0), rt.ncell - 1);
#ifdef GPDLA_PROBE_RT_FIXED
  ci = 100 + (threadIdx.x & 31);   // timing probe: always the same (L1-resident) cells; results are wrong
#endif
  static_assert(RT_DEG_DEV == 8 && RT_PLANES == 5, "four 128-bit loads and one 64-bit load per cell");
  const double2* cp = reinterpret_cast<const double2*>(rt.coef) + ci;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const double2 c2 = __ldg(cp + (size_t)p * rt.ncell);
    rc.cf[2 * p] = c2.x; rc.cf[2 * p + 1] = c2.y;
  }
  rc.cf[8] = __ldg(reinterpret_cast<const double*>(cp + (size_t)4 * rt.ncell));
}
// tau / N of one sample in the fetched cell; NaN (hi word >= 0x7ff00000) where the caller must evaluate directly
__device__ __forceinline__ double rest_table_eval(const RestCell& rc, double K) {
  const double s = rc.base - K;
  double t = rc.cf[RT_DEG_DEV];
#pragma unroll
  for (int p = RT_DEG_DEV - 1; p >= 0; --p) t = fma(t, s, rc.cf[p]);
  return t;
}

}  // namespace gpdla
