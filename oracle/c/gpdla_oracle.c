/* C restatement of the reference's per-sample hot loop -- TEST INFRASTRUCTURE ONLY.
 *
 * Used (a) as the multi-threaded CPU baseline timed by bench.py (the `parfor` analogue,
 * process_qsos.m:185) and (b) to evaluate full 10^4-sample quasars in tests in seconds.
 * parity unpinned (the reference ships no fixtures); checked in tests/ against the numpy
 * restatement, which is itself checked against the reference's compiled voigt.c.
 *
 * Follows, in the same operation order:
 *   voigt.c:277-299            multipliers, raw profile, 7-tap instrument convolution
 *   process_qsos.m:187-198     absorption(ind); dla_mu, dla_M, dla_omega2; d = dla_omega2 + v
 *   log_mvnpdf_low_rank.m:5-34 Woodbury + upper Cholesky, C = L\(L'\(D^-1 M)'), quadratic form
 * The line tables are passed in from oracle/process_qsos_oracle.py (one copy of the numbers).
 * libcerf's voigt() (voigt.c:288) = Re w((x + i gamma)/(sqrt2 sigma))/(sqrt(2 pi) sigma) with w the
 * Faddeeva-package wofz handed over from SciPy by gpdla_oracle_set_wofz().
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double _Complex (*wofz_fn)(double _Complex, int);
static wofz_fn g_wofz = 0;
static double g_c = 2.99792458e+10, g_sigma = 9.08537121627923800e+05;
static double g_tw[31], g_lc[31], g_gam[31], g_ip[7];
static const double LOG_2PI = 1.83787706640934534; /* log_mvnpdf_low_rank.m:7 */

void gpdla_oracle_set_wofz(void *p) { g_wofz = (wofz_fn)p; }
void gpdla_oracle_set_tables(const double *tw, const double *lc, const double *gam, const double *ip) {
  memcpy(g_tw, tw, sizeof g_tw); memcpy(g_lc, lc, sizeof g_lc);
  memcpy(g_gam, gam, sizeof g_gam); memcpy(g_ip, ip, sizeof g_ip);
}

static double cerf_voigt(double x, double sigma, double gamma) {
  double _Complex z = (x + I * fabs(gamma)) / sqrt(2.0) / fabs(sigma);
  return creal(g_wofz(z, 0)) / (sqrt(2.0 * M_PI) * fabs(sigma));
}

/* voigt.c:277-299; raw is scratch of num_points doubles */
static void voigt_profile(const double *lambdas, long num_points, double z, double N, int num_lines,
                          double *raw, double *profile) {
  double mult[31];
  for (int j = 0; j < num_lines; j++) mult[j] = g_c / (g_tw[j] * (1 + z)) / 1e8;
  for (long i = 0; i < num_points; i++) {
    double total = 0;
    for (int j = 0; j < num_lines; j++) {
      double velocity = lambdas[i] * mult[j] - g_c;
      total += -g_lc[j] * cerf_voigt(velocity, g_sigma, g_gam[j]);
    }
    raw[i] = exp(N * total);
  }
  long n_out = num_points - 6;
  for (long i = 0; i < n_out; i++) {
    double p = 0;
    for (int t = 0; t < 7; t++) p += raw[i + t] * g_ip[t];
    profile[i] = p;
  }
}

int gpdla_oracle_voigt(const double *lambdas, long num_points, double z, double N, int num_lines, double *profile) {
  if (!g_wofz || num_points < 7 || num_lines < 1 || num_lines > 31) return 1;
  double *raw = (double *)malloc(sizeof(double) * num_points);
  voigt_profile(lambdas, num_points, z, N, num_lines, raw, profile);
  free(raw);
  return 0;
}

/* log_mvnpdf_low_rank.m:5-34.  M is n x k row-major.  work: n*k (D_inv_M) + n*k (C, k x n) + n + n + k*k + 2k */
static double log_mvnpdf_low_rank(const double *y_in, const double *mu, const double *M, const double *d,
                                  long n, int k, double *work) {
  double *DM = work, *C = DM + n * k, *yc = C + n * k, *Dy = yc + n, *B = Dy + n, *Cy = B + k * k, *t = Cy + k;
  for (long i = 0; i < n; i++) {
    yc[i] = y_in[i] - mu[i];                      /* :11 */
    double di = 1.0 / d[i];                       /* :13 */
    Dy[i] = di * yc[i];                           /* :14 */
    for (int p = 0; p < k; p++) DM[i * k + p] = di * M[i * k + p];   /* :15 */
  }
  memset(B, 0, sizeof(double) * k * k);
  for (long i = 0; i < n; i++)                    /* :22  B = M' * D_inv_M */
    for (int p = 0; p < k; p++) {
      double mp = M[i * k + p];
      for (int q = 0; q < k; q++) B[p * k + q] += mp * DM[i * k + q];
    }
  for (int p = 0; p < k; p++) B[p * k + p] += 1.0;   /* :23 */
  /* :24  L = chol(B): upper triangular R with R'R = B, stored in B's upper triangle */
  for (int j = 0; j < k; j++) {
    for (int i = 0; i <= j; i++) {
      double s = B[i * k + j];
      for (int r = 0; r < i; r++) s -= B[r * k + i] * B[r * k + j];
      if (i < j) B[i * k + j] = s / B[i * k + i];
      else { if (!(s > 0)) return NAN; B[j * k + j] = sqrt(s); }
    }
  }
  /* :26  C = L \ (L' \ D_inv_M')  (k x n), column by column */
  for (long i = 0; i < n; i++) {
    for (int p = 0; p < k; p++) {                 /* forward: R' t = DM(i,:)' */
      double s = DM[i * k + p];
      for (int r = 0; r < p; r++) s -= B[r * k + p] * t[r];
      t[p] = s / B[p * k + p];
    }
    for (int p = k - 1; p >= 0; p--) {            /* backward: R c = t */
      double s = t[p];
      for (int r = p + 1; r < k; r++) s -= B[p * k + r] * C[r * n + i];
      C[p * n + i] = s / B[p * k + p];
    }
  }
  for (int p = 0; p < k; p++) {                   /* :28  C * y */
    double s = 0;
    for (long i = 0; i < n; i++) s += C[p * n + i] * yc[i];
    Cy[p] = s;
  }
  double quad = 0, logdet = 0;
  for (long i = 0; i < n; i++) {
    double s = 0;
    for (int p = 0; p < k; p++) s += DM[i * k + p] * Cy[p];
    quad += yc[i] * (Dy[i] - s);                  /* y' * K_inv_y */
    logdet += log(d[i]);                          /* :30 */
  }
  double ld = 0;
  for (int p = 0; p < k; p++) ld += log(B[p * k + p]);
  logdet += 2 * ld;
  return -0.5 * (quad + logdet + n * LOG_2PI);    /* :32 */
}

double gpdla_oracle_log_mvnpdf_low_rank(const double *y, const double *mu, const double *M, const double *d,
                                        long n, int k) {
  double *work = (double *)malloc(sizeof(double) * (2 * n * k + 2 * n + k * k + 2 * k));
  double r = log_mvnpdf_low_rank(y, mu, M, d, n, k, work);
  free(work);
  return r;
}

/* process_qsos.m:185-199 for one quasar (multi-DLA: `partners` gives, per sample, the indices of the
 * other DLAs whose profiles are multiplied in, ...meanflux.m:346-351; num_partners = 0 for single-DLA).
 *   padded      n_u + 6 padded wavelengths            keep    n_u flags (1 = pixel kept)
 *   y, mu, omega2, v   n kept-pixel vectors           M       n x k row-major
 *   sample_z, nhi      S sample parameters            sll     S outputs
 */
int gpdla_oracle_sample_loglik(const double *padded, long n_u, const unsigned char *keep,
                               const double *y, const double *mu, const double *M, const double *omega2,
                               const double *v, long n, int k,
                               const double *sample_z, const double *nhi, long S, int num_lines,
                               const int *partners, int num_partners,
                               double *sll, int nthreads) {
  if (!g_wofz) return 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    double *raw = (double *)malloc(sizeof(double) * (n_u + 6));
    double *prof = (double *)malloc(sizeof(double) * n_u);
    double *prof2 = (double *)malloc(sizeof(double) * n_u);
    double *dmu = (double *)malloc(sizeof(double) * n);
    double *dd = (double *)malloc(sizeof(double) * n);
    double *dM = (double *)malloc(sizeof(double) * n * k);
    double *work = (double *)malloc(sizeof(double) * (2 * n * k + 2 * n + k * k + 2 * k));
#pragma omp for schedule(dynamic, 8)
    for (long s = 0; s < S; s++) {
      voigt_profile(padded, n_u + 6, sample_z[s], nhi[s], num_lines, raw, prof);
      for (int j = 0; j < num_partners; j++) {
        long kk = partners[(long)j * S + s];
        voigt_profile(padded, n_u + 6, sample_z[kk], nhi[kk], num_lines, raw, prof2);
        for (long i = 0; i < n_u; i++) prof[i] = prof[i] * prof2[i];
      }
      long m = 0;
      for (long i = 0; i < n_u; i++) {
        if (!keep[i]) continue;
        double a = prof[i];
        dmu[m] = mu[m] * a;                                    /* :192 */
        for (int p = 0; p < k; p++) dM[m * k + p] = M[m * k + p] * a;   /* :193 */
        dd[m] = omega2[m] * (a * a) + v[m];                    /* :194,198 */
        m++;
      }
      sll[s] = log_mvnpdf_low_rank(y, dmu, dM, dd, n, k, work);
    }
    free(raw); free(prof); free(prof2); free(dmu); free(dd); free(dM); free(work);
  }
  return 0;
}
