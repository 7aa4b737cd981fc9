// DFMA / DMMA latency and issue-rate vs ILP and warps per SMSP on B200 (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while(0)

template <int ILP>
__global__ void k_dfma(double* out, long long* cyc, int iters, double b, double a0) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = a0 + i + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], b, a0);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  if (s == 1.2345) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
__global__ void k_dmma(double* out, long long* cyc, int iters, double b, double a0) {
  double acc[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; i++) { acc[i][0] = a0 + i; acc[i][1] = a0 - i; }
  double a = a0 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i][0] + acc[i][1];
  if (s == 1.2345) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP> int run(double* out, long long* dc, int warps) {
  int iters = 2048; long long c;
  k_dfma<ILP><<<1, 32 * warps>>>(out, dc, iters, 0.999, 1.0); CK(cudaDeviceSynchronize());
  k_dfma<ILP><<<1, 32 * warps>>>(out, dc, iters, 0.999, 1.0); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
  double per_dfma = (double)c / iters / ILP;
  k_dmma<ILP><<<1, 32 * warps>>>(out, dc, iters, 0.999, 1.0); CK(cudaDeviceSynchronize());
  k_dmma<ILP><<<1, 32 * warps>>>(out, dc, iters, 0.999, 1.0); CK(cudaDeviceSynchronize());
  long long c2; CK(cudaMemcpy(&c2, dc, 8, cudaMemcpyDeviceToHost));
  printf("warps/SM=%2d (per SMSP %.1f) ILP=%2d: DFMA %.2f cyc/instr/warp (dep latency if ILP=1)  SMSP DFMA rate %.3f instr/cyc | DMMA %.2f cyc/instr/warp, SMSP rate %.4f instr/cyc\n",
         warps, warps / 4.0, ILP, per_dfma, (warps / 4.0) / per_dfma, (double)c2 / iters / ILP, (warps / 4.0) / ((double)c2 / iters / ILP));
  return 0;
}
int main() {
  double* out; long long* dc; CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&dc, 64));
  for (int warps : {4, 8, 16}) {
    run<1>(out, dc, warps); run<2>(out, dc, warps); run<4>(out, dc, warps); run<8>(out, dc, warps); run<16>(out, dc, warps);
  }
  return 0;
}
