// How do DFMA warps progress when DMMA warps share their SM sub-partition?  (B200, sm_100a)
// One CTA per SM; warp w runs on sub-partition w % 4.  `ndmma` warps per sub-partition issue independent
// DMMA m8n8k4 streams for the whole kernel; `ndfma` warps per sub-partition run a fixed amount of DFMA work
// with `ILP` independent chains.  Reports the DFMA warps' completion time relative to running alone and
// the pipe utilisation.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while(0)

template <int ILP>
__global__ void __launch_bounds__(512, 1) k_mix(double* out, unsigned long long* res, int ndmma, int ndfma, int fma_iters, int* stop_flag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp >> 2;   // index of this warp within its sub-partition
  __shared__ int s_done;
  if (threadIdx.x == 0) s_done = 0;
  __syncthreads();
  double a = 1.0 + 1e-3 * lane, b = 0.999;
  if (slot < ndfma) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = a + i;
    long long t0 = clock64();
    for (int it = 0; it < fma_iters; it++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], b, a);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    if (s == 1.2345) out[0] = s;
    if (lane == 0) { atomicAdd(&s_done, 1); if (blockIdx.x == 0 && warp == 0) res[0] = t1 - t0; }
  } else if (slot < ndfma + ndmma) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = a; c[i][1] = b; }
    unsigned long long n = 0;
    long long t0 = clock64();
    while (*(volatile int*)&s_done < ndfma * 4) {
#pragma unroll
      for (int i = 0; i < 8; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
      n += 8;
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    if (s == 1.2345) out[0] = s;
    if (blockIdx.x == 0 && lane == 0 && warp == (ndfma * 4)) { res[1] = n; res[2] = t1 - t0; }
  }
}

template <int ILP> int run(double* out, unsigned long long* dres, int* flag, int ndmma, int ndfma) {
  const int iters = 20000;
  unsigned long long h[3] = {0, 0, 0};
  CK(cudaMemset(dres, 0, 24));
  k_mix<ILP><<<148, 512>>>(out, dres, ndmma, ndfma, iters, flag); CK(cudaDeviceSynchronize());
  CK(cudaMemset(dres, 0, 24));
  k_mix<ILP><<<148, 512>>>(out, dres, ndmma, ndfma, iters, flag); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h, dres, 24, cudaMemcpyDeviceToHost));
  double cyc_per_dfma = (double)h[0] / ((double)iters * ILP);            // per warp
  double dfma_pipe = ndfma * 2.0 / cyc_per_dfma;                          // fraction of pipe used by DFMA (2 cyc each)
  double dmma_pipe = h[2] ? ndmma * 16.0 * h[1] / (double)h[2] : 0.0;    // each DMMA warp: n DMMAs x 16 cycles
  printf("DMMA warps/SMSP %d, DFMA warps/SMSP %d, ILP %2d: %.2f cycles per DFMA per warp (pipe share DFMA %.2f, DMMA %.2f, total %.2f)\n",
         ndmma, ndfma, ILP, cyc_per_dfma, dfma_pipe, dmma_pipe, dfma_pipe + dmma_pipe);
  return 0;
}
int main() {
  double* out; unsigned long long* dres; int* flag;
  CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&dres, 64)); CK(cudaMalloc(&flag, 4));
  for (int ndmma : {0, 1, 2}) for (int ndfma : {1, 2}) {
    if (ndmma + ndfma > 4) continue;
    run<1>(out, dres, flag, ndmma, ndfma); run<2>(out, dres, flag, ndmma, ndfma);
    run<4>(out, dres, flag, ndmma, ndfma); run<8>(out, dres, flag, ndmma, ndfma);
  }
  return 0;
}
