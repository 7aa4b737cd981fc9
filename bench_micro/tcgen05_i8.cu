// Groundwork for taking the Gram off the FP64 pipe ("next step"): an INT8 tcgen05.mma
// (kind::i8, s32 accumulators in TMEM) micro-kernel with hand-built shared-memory and instruction
// descriptors, checked against a host integer GEMM, then timed.   D[128 x 240] = A[128 x K] . B[240 x K]^T
//
// Layout (K-major, no swizzle): an operand slab for one MMA (K = 32 bytes) is made of 8-row x 16-byte core
// matrices; core matrices are 128 B apart along K (LBO) and 256 B apart along M/N (SBO).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while(0)

#ifndef NN
#define NN 240
#endif
constexpr int M = 128, N = NN, KMMA = 32;
constexpr int A_SLAB = M * KMMA, B_SLAB = N * KMMA;    // bytes per k-step
constexpr uint32_t LBO = 128, SBO = 256;
#ifndef A_LBO
#define A_LBO 128
#endif
#ifndef A_SBO
#define A_SBO 256
#endif
// A operand layout as in the fused kernel: K halves A_LBO apart, 8-row groups A_SBO apart (digit planes interleaved)
// -DA_MN: A is MN-major (no swizzle): 16 consecutive rows (M) are the 16 bytes of a granule, 8 consecutive k are 8
// granules (one 128-byte core matrix), k blocks of 8 are A_LBO apart, row blocks of 16 are A_SBO apart
// (canonical layout ((1,n),(8,k)):((X,SBO),(1,LBO)) in 16-byte units); instruction descriptor bit 15 = A MN-major.
#ifdef A_MN
constexpr int A_SLAB_BYTES = (M / 16) * A_SBO;
#else
constexpr int A_SLAB_BYTES = (M / 8) * A_SBO;
#endif
constexpr int TMEM_COLS = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo = LBO, uint32_t sbo = SBO) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46);   // version 1 (Blackwell), base offset 0, SWIZZLE_NONE
}
// kind::i8 instruction descriptor: D = s32, A = B = signed 8-bit, both K-major, N >> 3, M >> 4
__host__ __device__ constexpr uint32_t make_idesc() {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24)
#ifdef A_MN
         | (1u << 15)
#endif
      ;
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, int max_iter) {
  for (int it = 0; it < max_iter; ++it) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

// ksteps k-steps of operands are staged once; `reps` repetitions of the whole K loop are issued (timing)
__global__ void __launch_bounds__(128, 1) k_i8(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int K,
                                               int32_t* __restrict__ D, int reps, long long* cycles, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int ksteps = K / KMMA;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)ksteps * A_SLAB_BYTES;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // stage operands in the canonical layout (16-byte granules)
#ifdef A_MN
  for (int e = tid; e < ksteps * M * KMMA; e += 128) {       // A: byte by byte into the MN-major layout
    int s = e / (M * KMMA), r = (e / KMMA) % M, k = e % KMMA;
    sA[(size_t)s * A_SLAB_BYTES + (r / 16) * A_SBO + (k / 8) * A_LBO + (k % 8) * 16 + (r % 16)] = (uint8_t)A[(size_t)r * K + s * KMMA + k];
  }
#else
  for (int g = tid; g < ksteps * M * 2; g += 128) {          // A: granule = (kstep, row, half)
    int s = g / (M * 2), r = (g / 2) % M, h = g % 2;
    *reinterpret_cast<int4*>(sA + (size_t)s * A_SLAB_BYTES + (r / 8) * A_SBO + h * A_LBO + (r % 8) * 16) =
        *reinterpret_cast<const int4*>(A + (size_t)r * K + s * KMMA + h * 16);
  }
#endif
  for (int g = tid; g < ksteps * N * 2; g += 128) {
    int s = g / (N * 2), r = (g / 2) % N, h = g % 2;
    *reinterpret_cast<int4*>(sB + (size_t)s * B_SLAB + (r / 8) * SBO + h * LBO + (r % 8) * 16) =
        *reinterpret_cast<const int4*>(B + (size_t)r * K + s * KMMA + h * 16);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // make the generic-proxy smem writes visible to the async (tensor core) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_s;

  long long t0 = clock64();
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc();
      for (int rep = 0; rep < reps; ++rep) {
        for (int s = 0; s < ksteps; ++s) {
#ifdef A_DESC_SWAP
          const uint64_t da = make_desc(smem_u32(sA + (size_t)s * A_SLAB_BYTES), A_SBO, A_LBO);
#else
          const uint64_t da = make_desc(smem_u32(sA + (size_t)s * A_SLAB_BYTES), A_LBO, A_SBO);
#endif
          const uint64_t db = make_desc(smem_u32(sB + (size_t)s * B_SLAB));
          const uint32_t accumulate = (s > 0) ? 1u : 0u;
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_base),
              "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
              : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
  }
  const bool ok = mbar_wait_bounded(&bar, 0, 1 << 24);
  long long t1 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (!ok) { if (tid == 0) status[0] = 1; }
  if (ok && D != nullptr) {
    // thread (warp, lane) owns accumulator row 32 warp + lane; 8 columns per tcgen05.ld
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t v[8];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[(size_t)row * N + c0 + j] = (int32_t)v[j];
    }
  }
  if (tid == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

int main() {
#ifndef KTOT
#define KTOT 512
#endif
  const int K = KTOT;   // k-steps staged in shared memory
  std::vector<int8_t> hA((size_t)M * K), hB((size_t)N * K);
  srand(7);
  for (auto& x : hA) x = (int8_t)(rand() % 128 - 64);
  for (auto& x : hB) x = (int8_t)(rand() % 128 - 64);
  int8_t *dA, *dB; int32_t* dD; long long* dcyc; int* dst;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, (size_t)M * N * 4));
  CK(cudaMalloc(&dcyc, 148 * 8)); CK(cudaMalloc(&dst, 4)); CK(cudaMemset(dst, 0, 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  const size_t smem = (size_t)(K / KMMA) * (A_SLAB_BYTES + B_SLAB) + 1024;
  CK(cudaFuncSetAttribute(k_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // ---- correctness: one CTA
  k_i8<<<1, 128, smem>>>(dA, dB, K, dD, 1, dcyc, dst);
  CK(cudaDeviceSynchronize());
  int st = 0; CK(cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost));
  std::vector<int32_t> hD((size_t)M * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      int32_t ref = 0;
      for (int k = 0; k < K; ++k) ref += (int32_t)hA[(size_t)i * K + k] * (int32_t)hB[(size_t)j * K + k];
      if (ref != hD[(size_t)i * N + j]) { if (bad < 5) printf("mismatch D[%d][%d] = %d, expected %d\n", i, j, hD[(size_t)i * N + j], ref); ++bad; }
    }
  printf("{\"timeout\": %d, \"mismatches\": %lld of %d", st, bad, M * N);
  // ---- throughput: every SM repeats the K loop
  if (!st && !bad) {
    const int reps = 2000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_i8<<<148, 128, smem>>>(dA, dB, K, nullptr, reps, dcyc, dst); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); k_i8<<<148, 128, smem>>>(dA, dB, K, nullptr, reps, dcyc, dst); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long c; CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
    const double ops = 148.0 * reps * 2.0 * M * N * K;
    printf(", \"int8_tops_all_sms\": %.1f, \"cycles_per_mma_128x240x32\": %.1f", ops / ms * 1e-9, (double)c / reps / (K / KMMA));
  }
  printf("}\n");
  return 0;
}
