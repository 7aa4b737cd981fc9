// What limits a lone DMMA warp per SM sub-partition that also fetches its operands from shared memory?
// Replicates the consumer loop of dla_loglik_ws_kernel: per k4-step 2 A-fragment loads + NT B-fragment
// loads + NT DMMA m8n8k4; 4 warps per CTA (one per sub-partition), 1 CTA per SM.  Variants:
//   0: LDS.64 per B fragment (as shipped)        1: LDS.128 fetching two B fragments at once
//   2: no shared-memory loads at all (registers) -> pure issue limit
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while(0)
constexpr int NT = 30, KC = 32, BSTR = 244, ASTR = 36;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int VARIANT>
__global__ void __launch_bounds__(128, 1) k_loop(double* out, long long* cyc, int chunks) {
  extern __shared__ __align__(16) double sm[];
  double* Bt = sm;                       // [KC][BSTR]
  double* Wt = sm + KC * BSTR;           // [32][ASTR]
  for (int i = threadIdx.x; i < KC * BSTR + 32 * ASTR; i += 128) sm[i] = 1.0 + 1e-6 * i;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  double acc[NT][2];
#pragma unroll
  for (int ni = 0; ni < NT; ++ni) acc[ni][0] = acc[ni][1] = 0.0;
  const double* rW = Wt + (warp * 8 + gid) * ASTR + tig;
  long long t0 = clock64();
  for (int c = 0; c < chunks; ++c) {
    const double* Bc = Bt + tig * BSTR + gid;
#pragma unroll
    for (int ks = 0; ks < (VARIANT == 3 ? 0 : KC / 4); ++ks) {
      const double aw = rW[ks * 4];
      const double* brow = Bc + ks * 4 * BSTR;
      if (VARIANT == 0) {
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) dmma(acc[ni][0], acc[ni][1], aw, brow[ni * 8]);
      } else if (VARIANT == 1) {
        // layout assumption for the test only: two fragments adjacent -> one 16-byte load
        const double2* b2 = reinterpret_cast<const double2*>(Bt + (ks * 4 + tig) * BSTR) + gid;
#pragma unroll
        for (int ni = 0; ni < NT; ni += 2) {
          const double2 b = b2[ni * 4];
          dmma(acc[ni][0], acc[ni][1], aw, b.x);
          dmma(acc[ni + 1][0], acc[ni + 1][1], aw, b.y);
        }
      } else if (VARIANT == 2) {
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) dmma(acc[ni][0], acc[ni][1], aw, aw + ni);
      }
    }
    if (VARIANT == 3) {
      // explicit software pipeline: fragments of k4-step ks+1 are loaded (16-byte loads) into a second
      // register set while the DMMAs of step ks issue
      double2 cur[NT / 2], nxt[NT / 2];
      double awc = rW[0], awn = 0.0;
      {
        const double2* b2 = reinterpret_cast<const double2*>(Bt + tig * BSTR) + gid;
#pragma unroll
        for (int j = 0; j < NT / 2; ++j) cur[j] = b2[j * 8];
      }
#pragma unroll
      for (int ks = 0; ks < KC / 4; ++ks) {
        if (ks + 1 < KC / 4) {
          const double2* b2 = reinterpret_cast<const double2*>(Bt + ((ks + 1) * 4 + tig) * BSTR) + gid;
          awn = rW[(ks + 1) * 4];
#pragma unroll
          for (int j = 0; j < NT / 2; ++j) nxt[j] = b2[j * 8];
        }
#pragma unroll
        for (int j = 0; j < NT / 2; ++j) {
          dmma(acc[2 * j][0], acc[2 * j][1], awc, cur[j].x);
          dmma(acc[2 * j + 1][0], acc[2 * j + 1][1], awc, cur[j].y);
        }
#pragma unroll
        for (int j = 0; j < NT / 2; ++j) cur[j] = nxt[j];
        awc = awn;
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int ni = 0; ni < NT; ++ni) s += acc[ni][0] + acc[ni][1];
  if (s == 1.2345) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int V> int run(double* out, long long* dc, const char* name) {
  const int chunks = 400; const size_t smem = (KC * BSTR + 32 * ASTR) * 8;
  CK(cudaFuncSetAttribute(k_loop<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_loop<V><<<148, 128, smem>>>(out, dc, chunks); CK(cudaDeviceSynchronize());
  k_loop<V><<<148, 128, smem>>>(out, dc, chunks); CK(cudaDeviceSynchronize());
  long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
  printf("%-28s %.1f cycles per chunk (ideal %d) -> %.1f %% of DMMA peak\n", name, (double)c / chunks, NT * 8 * 16, 100.0 * NT * 8 * 16 / ((double)c / chunks));
  return 0;
}
int main() {
  double* out; long long* dc; CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&dc, 64));
  run<0>(out, dc, "LDS.64 per fragment");
  run<1>(out, dc, "LDS.128 per two fragments");
  run<2>(out, dc, "no smem loads");
  run<3>(out, dc, "LDS.128 + register double buffer");
  return 0;
}
