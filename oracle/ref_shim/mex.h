/* Minimal stand-in for MATLAB's mex.h -- TEST INFRASTRUCTURE ONLY (oracle/).
 * Just enough of the mx* API for /root/reference/voigt.c (voigt.c:4,253-304) to
 * compile unmodified as plain C.  Written from the documented MATLAB C-API
 * semantics; contains no reference code. */
#ifndef GPDLA_ORACLE_MEX_SHIM_H
#define GPDLA_ORACLE_MEX_SHIM_H
#include <stddef.h>

typedef struct mxArray_tag {
  double *pr;
  size_t  m, n;
} mxArray;
typedef size_t mwSize;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

double  *mxGetPr(const mxArray *a);
double   mxGetScalar(const mxArray *a);
size_t   mxGetNumberOfElements(const mxArray *a);
mxArray *mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity flag); /* zero-filled */
void    *mxMalloc(size_t n);
void     mxFree(void *p);
/* error exit of a MEX function: the shim records the message and longjmps back to the harness (tests/test_integration_shims.py) */
void     mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#endif
