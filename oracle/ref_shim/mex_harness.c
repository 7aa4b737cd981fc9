/* Harness for MEX shims written against libgpdla.so -- TEST INFRASTRUCTURE ONLY.
 * Implements the handful of mx / mex functions of ref_shim/mex.h and one entry point that calls a shim's
 * mexFunction(lambdas, z, N, num_lines) the way MATLAB would, catching mexErrMsgIdAndTxt. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mex.h"

static jmp_buf g_jmp;
static char g_msg[512];

double *mxGetPr(const mxArray *a) { return a->pr; }
double mxGetScalar(const mxArray *a) { return a->pr[0]; }
size_t mxGetNumberOfElements(const mxArray *a) { return a->m * a->n; }
mxArray *mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity flag) {
  (void)flag;
  mxArray *a = (mxArray *)malloc(sizeof(mxArray));
  a->m = m; a->n = n;
  const size_t count = m * n;
  a->pr = (double *)calloc(count > 0 ? count : 1, sizeof(double));
  return a;
}
void *mxMalloc(size_t n) { return malloc(n); }
void mxFree(void *p) { free(p); }
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...) {
  va_list ap;
  int n = snprintf(g_msg, sizeof g_msg, "%s: ", id);
  va_start(ap, fmt);
  vsnprintf(g_msg + n, sizeof g_msg - (size_t)n, fmt, ap);
  va_end(ap);
  longjmp(g_jmp, 1);
}

/* returns 0 and fills profile[num_points - 6], or 1 with the MEX error message in msg */
int harness_voigt(const double *lambdas, long num_points, double z, double N, int num_lines, double *profile, char *msg,
                  int msg_len) {
  mxArray l = {(double *)lambdas, (size_t)num_points, 1};
  double zz = z, nn = N, nl = (double)num_lines;
  mxArray az = {&zz, 1, 1}, an = {&nn, 1, 1}, al = {&nl, 1, 1};
  const mxArray *prhs[4] = {&l, &az, &an, &al};
  mxArray *plhs[1] = {0};
  if (setjmp(g_jmp)) {
    snprintf(msg, (size_t)msg_len, "%s", g_msg);
    return 1;
  }
  mexFunction(1, plhs, num_lines > 0 ? 4 : 3, prhs);
  memcpy(profile, plhs[0]->pr, sizeof(double) * (size_t)(num_points - 6));
  free(plhs[0]->pr); free(plhs[0]);
  return 0;
}
