/* Glue that lets the reference's voigt.c run outside MATLAB -- TEST INFRASTRUCTURE ONLY.
 *
 *   ref_set_wofz(ptr)   hand over scipy.special.cython_special's C entry point
 *                       `double complex wofz(double complex, int)` (Faddeeva package)
 *   ref_voigt(...)      marshal plain C arrays into the mx* shim, call the reference's
 *                       mexFunction (voigt.c:253), copy the result out
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "mex.h"
#include "cerf.h"

typedef double _Complex (*wofz_fn)(double _Complex, int);
static wofz_fn g_wofz = 0;

void ref_set_wofz(void *p) { g_wofz = (wofz_fn)p; }

/* libcerf's documented definition: voigt(x, sigma, gamma) =
 * Re[w((x + i|gamma|) / (sqrt2 |sigma|))] / (sqrt(2 pi) |sigma|), with the sigma = 0 and
 * gamma = 0 limits (pure Lorentzian / Gaussian). */
double voigt(double x, double sigma, double gamma) {
  double gam = fabs(gamma), sig = fabs(sigma);
  if (gam == 0.0) {
    if (sig == 0.0) return x == 0.0 ? INFINITY : 0.0;
    return exp(-x * x / 2.0 / (sig * sig)) / (sqrt(2.0 * M_PI) * sig);
  }
  if (sig == 0.0) return gam / (M_PI * (x * x + gam * gam));
  double _Complex z = (x + I * gam) / sqrt(2.0) / sig;
  return creal(g_wofz(z, 0)) / (sqrt(2.0 * M_PI) * sig);
}

double *mxGetPr(const mxArray *a) { return a->pr; }
double mxGetScalar(const mxArray *a) { return a->pr[0]; }
size_t mxGetNumberOfElements(const mxArray *a) { return a->m * a->n; }
mxArray *mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity flag) {
  (void)flag;
  mxArray *a = (mxArray *)malloc(sizeof(mxArray));
  a->m = m; a->n = n;
  a->pr = (double *)calloc(m * n ? m * n : 1, sizeof(double));
  return a;
}
void *mxMalloc(size_t n) { return malloc(n); }
void mxFree(void *p) { free(p); }

int ref_voigt(const double *lambdas, long num_points, double z, double N, int num_lines, double *profile) {
  if (!g_wofz || num_points < 7) return 1;
  mxArray l = {(double *)lambdas, (size_t)num_points, 1};
  double zz = z, nn = N, nl = (double)num_lines;
  mxArray az = {&zz, 1, 1}, an = {&nn, 1, 1}, al = {&nl, 1, 1};
  const mxArray *prhs[4] = {&l, &az, &an, &al};
  mxArray *plhs[1] = {0};
  mexFunction(1, plhs, num_lines > 0 ? 4 : 3, prhs);
  memcpy(profile, plhs[0]->pr, sizeof(double) * (size_t)(num_points - 6));
  free(plhs[0]->pr); free(plhs[0]);
  return 0;
}
