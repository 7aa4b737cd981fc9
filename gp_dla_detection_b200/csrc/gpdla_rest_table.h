// Host-side builder of the rest-frame optical-depth table used by the fused log-likelihood kernels.
//
// The raw absorption of voigt.c:282-291 is exp(-N * T(w)) with
//   T(w) = sum_j leading_constant_j * voigt(c (w / lambda_j - 1), sigma, gamma_j),   w = lambda_obs / (1 + z_dla)
// (voigt.c:279,287: velocity = lambda * c / (lambda_j (1 + z)) / 1e8 - c), i.e. a function of ONE variable, the
// absorber rest-frame wavelength.  The kernels look T up in a table over u = ln(w / lambda_lo) / h (h = the pixel
// spacing in natural-log units, so one cell per 1e-4-dex pixel): cell c is centred at u = c and holds a degree-RT_DEG
// polynomial in s = u - c, valid for |s| <= RT_HALF_WIDTH = 1 (interpolant at the Chebyshev nodes of that interval, built
// here in long double).  The cells overlap on purpose: the four samples a producer warp interleaves are neighbours in
// redshift (the kernels walk the samples in sorted groups of four), so ONE cell -- the one nearest to the group's mean
// position -- serves all four, its coefficients are fetched once per (warp, pixel) and a chunk ahead, and each sample
// only evaluates the polynomial at its own s.  Cells closer than RT_NEAR_PIXELS to a line centre carry NaN: there the
// polynomial cannot follow T (tools/rest_table_design.py: degree 8 on |s| <= 1 is good to 4.5e-13 relative from 16
// cells outwards) and the kernels fall back to the direct evaluation of gpdla_math.cuh for the whole warp.  Far from
// the line cores T is the Lorentzian wing plus Doppler corrections, evaluated here from the asymptotic series of the
// Faddeeva function,
//   w(z) ~ i / (sqrt(pi) z) * sum_n (2n-1)!! / (2 z^2)^n,    z = (v + i gamma) / (sqrt2 sigma),
// which at |z| >= 80 (the nodes never get closer than 15 pixels = 80 Doppler widths to a line centre) reaches
// long-double rounding after a dozen terms (checked against mpmath: tests/test_rest_table.py).
#pragma once
#include <math.h>

#include <complex>
#include <vector>

namespace gpdla {

constexpr int RT_DEG = 8;                 // polynomial degree per cell
constexpr double RT_HALF_WIDTH = 1.0;     // the polynomial of a cell is valid for |s| <= this many cells
constexpr double RT_LAMBDA_LO = 880.0;    // table range in absorber rest-frame wavelength (Angstrom): a DLA between
constexpr double RT_LAMBDA_HI = 1720.0;   // z_min and z_max sees the modelled window at 903 .. 1605 Angstrom
constexpr double RT_NEAR_PIXELS = 16.0;   // cells closer than this to a line centre are left to direct evaluation

struct RestTableHost {
  std::vector<double> coef;   // [RT_DEG + 1][ncell]
  int ncell = 0;
  double h = 0;               // cell width in ln(wavelength)
  double ln_lo = 0;           // ln(RT_LAMBDA_LO)
};

// tw: transition wavelengths (cm), lc: leading constants, gam: Lorentzian widths (cm/s), as in LineConstants
inline long double rest_table_tau(long double x /* ln(rest wavelength / Angstrom) */, int num_lines, const double* tw,
                                  const double* lc, const double* gam, double sigma, double c) {
  const long double s2 = sqrtl(2.0L) * (long double)sigma;
  long double total = 0.0L;
  for (int j = 0; j < num_lines; ++j) {
    const long double v = (long double)c * expm1l(x - logl((long double)tw[j] * 1e8L));
    const std::complex<long double> z(v / s2, (long double)gam[j] / s2);
    const std::complex<long double> iz2 = 1.0L / (2.0L * z * z);
    std::complex<long double> term(1.0L, 0.0L), sum(1.0L, 0.0L);
    for (int n = 1; n < 40; ++n) {
      term *= (long double)(2 * n - 1) * iz2;
      sum += term;
      if (std::abs(term) < 1e-22L) break;
    }
    const std::complex<long double> w = std::complex<long double>(0.0L, 1.0L) / (sqrtl(M_PIl) * z) * sum;
    total += (long double)lc[j] * w.real() / (sqrtl(2.0L * M_PIl) * (long double)sigma);
  }
  return total;
}

inline RestTableHost build_rest_table(int num_lines, double pixel_spacing_dex, double near_pixels, const double* tw,
                                      const double* lc, const double* gam, double sigma, double c) {
  RestTableHost t;
  t.h = pixel_spacing_dex * log(10.0);
  t.ln_lo = log(RT_LAMBDA_LO);
  t.ncell = (int)ceil((log(RT_LAMBDA_HI) - t.ln_lo) / t.h) + 1;
  constexpr int n = RT_DEG + 1;
  t.coef.assign((size_t)n * t.ncell, 0.0);
  // Chebyshev nodes on [-1, 1] and the Chebyshev -> monomial conversion (T_j(t) = sum_p Tm[j][p] t^p)
  long double nodes[n], Tm[n][n] = {};
  for (int k = 0; k < n; ++k) nodes[k] = cosl(M_PIl * (k + 0.5L) / n);
  Tm[0][0] = 1.0L;
  Tm[1][1] = 1.0L;
  for (int j = 2; j < n; ++j)
    for (int p = 0; p < n; ++p) Tm[j][p] = (p > 0 ? 2.0L * Tm[j - 1][p - 1] : 0.0L) - Tm[j - 2][p];
  std::vector<double> centre(num_lines);
  for (int j = 0; j < num_lines; ++j) centre[j] = (log(tw[j] * 1e8) - t.ln_lo) / t.h;
  for (int cidx = 0; cidx < t.ncell; ++cidx) {
    bool near = (cidx == 0 || cidx == t.ncell - 1);   // the end cells catch clamped (out-of-range) lookups
    for (int j = 0; j < num_lines; ++j) near |= fabs((double)cidx - centre[j]) < near_pixels + 0.5;
    if (near) {
      for (int p = 0; p < n; ++p) t.coef[(size_t)p * t.ncell + cidx] = NAN;
      continue;
    }
    long double f[n], cj[n];
    for (int k = 0; k < n; ++k)
      f[k] = rest_table_tau((long double)t.ln_lo + (long double)t.h * ((long double)cidx + (long double)RT_HALF_WIDTH * nodes[k]),
                            num_lines, tw, lc, gam, sigma, c);
    for (int j = 0; j < n; ++j) {
      long double a = 0.0L;
      for (int k = 0; k < n; ++k) a += f[k] * cosl(M_PIl * j * (k + 0.5L) / n);
      cj[j] = a * (j == 0 ? 1.0L : 2.0L) / n;
    }
    long double scale = 1.0L;   // t = s / RT_HALF_WIDTH  ->  coefficient of s^p is that of t^p over RT_HALF_WIDTH^p
    for (int p = 0; p < n; ++p) {
      long double m = 0.0L;
      for (int j = 0; j < n; ++j) m += cj[j] * Tm[j][p];
      t.coef[(size_t)p * t.ncell + cidx] = (double)(m * scale);
      scale /= (long double)RT_HALF_WIDTH;
    }
  }
  return t;
}

}  // namespace gpdla
