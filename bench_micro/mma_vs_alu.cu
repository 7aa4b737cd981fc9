// Does an executing tcgen05.mma kind::i8 slow down the ordinary arithmetic of the other warps of its SM, and which
// kind?  (in the fused kernel the producers lose about the MMAs' own execution time, whatever their
// shared-memory traffic.)  One CTA per SM: warp 0 issues a stream of 128 x N x 32 INT8 MMAs (or none), warps 1..3 run a
// fixed amount of independent FMA chains of one kind -- FP64 DFMA, FP32 FFMA or integer IMAD -- and the cycles they
// need are reported with and without the MMA stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_vs_alu mma_vs_alu.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while(0)

constexpr int M = 128, KMMA = 32, NMAX = 240, KSTEPS = 8;
constexpr uint32_t LBO = 128, SBO = 256;
constexpr int A_SLAB = M * KMMA, B_SLAB = NMAX * KMMA;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((LBO >> 4) & 0x3FFF) << 16) | ((uint64_t)((SBO >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind: 0 DFMA, 1 FFMA, 2 IMAD.  mma_n: 0 = no MMAs, else N of every MMA.  mma_count MMAs are issued (bounded).
__global__ void __launch_bounds__(128, 1) k_mix(int kind, int alu_iters, int mma_n, int mma_count, long long* alu_cycles,
                                                long long* mma_cycles, double* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)KSTEPS * A_SLAB;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < KSTEPS * (A_SLAB + B_SLAB) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (uint32_t)(i & 3);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_s;
  const long long t0 = clock64();
  if (warp == 0) {
    if (lane == 0 && mma_n > 0) {
      const uint32_t idesc = make_idesc(mma_n);
      for (int i = 0; i < mma_count; ++i) {
        const int s = i % KSTEPS;
        const uint64_t da = make_desc(smem_u32(sA + (size_t)s * A_SLAB));
        const uint64_t db = make_desc(smem_u32(sB + (size_t)s * B_SLAB));
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_base),
            "l"(da), "l"(db), "r"(idesc), "r"(i > 0 ? 1u : 0u), "r"(0u)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      // wait for the MMAs (bounded)
      for (int it = 0; it < (1 << 26); ++it) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        if (ok) break;
      }
      mma_cycles[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  } else {
    // 8 independent chains per thread
    if (kind == 0) {
      double a[8]; for (int j = 0; j < 8; ++j) a[j] = 1.0 + 1e-3 * (tid + j);
      const double m = 1.0000001, c = 1e-9;
      for (int i = 0; i < alu_iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fma(a[j], m, c);
      }
      double s = 0; for (int j = 0; j < 8; ++j) s += a[j];
      if (s == 12345.678) sink[tid] = s;
    } else if (kind == 1) {
      float a[8]; for (int j = 0; j < 8; ++j) a[j] = 1.0f + 1e-3f * (tid + j);
      const float m = 1.0000001f, c = 1e-9f;
      for (int i = 0; i < alu_iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], m, c);
      }
      float s = 0; for (int j = 0; j < 8; ++j) s += a[j];
      if (s == 12345.678f) sink[tid] = s;
    } else {
      int a[8]; for (int j = 0; j < 8; ++j) a[j] = tid + j;
      for (int i = 0; i < alu_iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = a[j] * 1664525 + 1013904223;
      }
      int s = 0; for (int j = 0; j < 8; ++j) s += a[j];
      if (s == 123456789) sink[tid] = s;
    }
    if (warp == 1 && lane == 0) alu_cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
}

int main() {
  long long *dalu, *dmma; double* dsink;
  CK(cudaMalloc(&dalu, 148 * 8)); CK(cudaMalloc(&dmma, 148 * 8)); CK(cudaMalloc(&dsink, 1024));
  const size_t smem = (size_t)KSTEPS * (A_SLAB + B_SLAB) + 1024;
  CK(cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const char* names[3] = {"DFMA", "FFMA", "IMAD"};
  const int iters[3] = {20000, 40000, 40000};
  printf("{");
  for (int kind = 0; kind < 3; ++kind) {
    for (int mode = 0; mode < 3; ++mode) {   // no MMAs, N = 80, N = 240
      const int n = mode == 0 ? 0 : (mode == 1 ? 80 : 240);
      const int count = mode == 0 ? 0 : 60000;     // long enough to cover the ALU work, bounded
      CK(cudaMemset(dalu, 0, 148 * 8)); CK(cudaMemset(dmma, 0, 148 * 8));
      k_mix<<<148, 128, smem>>>(kind, iters[kind], n, count, dalu, dmma, dsink);
      CK(cudaDeviceSynchronize());
      long long ca, cm; CK(cudaMemcpy(&ca, dalu, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&cm, dmma, 8, cudaMemcpyDeviceToHost));
      printf("%s\"%s_mmaN%d\": {\"alu_cycles\": %lld, \"cycles_per_fma_instr_per_warp\": %.2f, \"mma_stream_cycles\": %lld, \"cycles_per_mma\": %.1f}",
             (kind || mode) ? ", " : "", names[kind], n, ca, (double)ca / ((double)iters[kind] * 8), cm, count ? (double)cm / count : 0.0);
    }
  }
  printf("}\n");
  return 0;
}
