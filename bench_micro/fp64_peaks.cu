// Step-0 denominators for the FP64 roofline on B200 (sm_100a).
// Measures: DFMA peak, DMMA (mma.sync m8n8k4 / m16n8k8 / m16n8k16 f64) peak,
// concurrent DFMA+DMMA in different warps (do they share a pipe?), and the
// cost of FP64 rcp / exp / log per element.  Prints one JSON object.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// mode: 0 = all warps DFMA, 1 = all warps DMMA884, 2 = even warps DFMA / odd warps DMMA884,
//       3 = m16n8k8, 4 = m16n8k16
__global__ void __launch_bounds__(256) k_peak(double *out, int iters, int mode, double seed) {
  int warp = threadIdx.x >> 5;
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = seed * (i + 1) + threadIdx.x;
  double a = seed + 1e-3 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  bool do_fma = (mode == 0) || (mode == 2 && (warp & 1) == 0);
  if (mode == 3) {
    double c[4][4]; double av[4] = {a, a + 1, a + 2, a + 3}; double bv[2] = {b, b * 0.5};
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = acc[i * 4 + j];
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 4; i++) dmma1688(c[i], av, bv);
    }
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) acc[i * 4 + j] = c[i][j];
  } else if (mode == 4) {
    double c[4][4]; double av[8]; double bv[4] = {b, b * 0.5, b * 0.25, b * 2};
    for (int i = 0; i < 8; i++) av[i] = a + i;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = acc[i * 4 + j];
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 4; i++) dmma16816(c[i], av, bv);
    }
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) acc[i * 4 + j] = c[i][j];
  } else if (do_fma) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], b, a);
    }
  } else {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) dmma884(acc[2 * i], acc[2 * i + 1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i];
  if (s == 123.456) out[0] = s;
}

// mode 0: rcp (1.0/x), 1: exp, 2: log, 3: sqrt, 4: rcp.approx + 2 NR, 5: x/y
__global__ void __launch_bounds__(256) k_func(double *out, int iters, int mode, double seed) {
  double x[4];
  for (int i = 0; i < 4; i++) x[i] = seed + 1e-3 * (threadIdx.x + i * 7) + 0.5;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      double v = x[i];
      if (mode == 0) v = 1.0 / v + 0.25;
      else if (mode == 1) v = exp(-v) + 0.6;
      else if (mode == 2) v = log(v) + 1.7;
      else if (mode == 3) v = sqrt(v) + 0.3;
      else if (mode == 4) {
        double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
        double e = fma(-v, r, 1.0); r = fma(r, e, r); e = fma(-v, r, 1.0); r = fma(r, e, r);
        v = r + 0.25;
      } else v = seed / v + 0.25;
      x[i] = v;
    }
  }
  double s = x[0] + x[1] + x[2] + x[3];
  if (s == 123.456) out[0] = s;
}

template <class F> float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double *out; CK(cudaMalloc(&out, 64));
  int iters = 4096;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);
  for (int cps = 1; cps <= 8; cps *= 2) {   // CTAs (8 warps) per SM
    int grid = sms * cps;
    const char *names[5] = {"dfma", "dmma884", "mixed", "dmma1688", "dmma16816"};
    for (int mode = 0; mode < 5; mode++) {
      float ms = time_ms([&] { k_peak<<<grid, 256>>>(out, iters, mode, 1.0); });
      double flops;
      double thr = (double)grid * 256;
      if (mode == 0) flops = thr * iters * 16 * 2;
      else if (mode == 1) flops = (thr / 32) * iters * 8 * 512.0;
      else if (mode == 2) flops = (thr / 2) * iters * 16 * 2 + (thr / 64) * iters * 8 * 512.0;
      else if (mode == 3) flops = (thr / 32) * iters * 4 * (16.0 * 8 * 8 * 2);
      else flops = (thr / 32) * iters * 4 * (16.0 * 8 * 16 * 2);
      printf(", \"%s_cps%d_tflops\": %.3f", names[mode], cps, flops / ms * 1e-9);
    }
  }
  const char *fn[6] = {"rcp", "exp", "log", "sqrt", "rcp_approx_nr2", "div"};
  for (int mode = 0; mode < 6; mode++) {
    int grid = sms * 8;
    float ms = time_ms([&] { k_func<<<grid, 256>>>(out, 1024, mode, 1.0); });
    double n = (double)grid * 256 * 1024 * 4;
    printf(", \"%s_gops\": %.2f", fn[mode], n / ms * 1e-6);
  }
  printf("}\n");
  return 0;
}
