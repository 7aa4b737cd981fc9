// GP training objective on the device: objective.m:12-73 + spectrum_loss.m:14-74.
//   f(x) = -sum_i log N(y_i; 0, M M' + diag(sigma_i^2 + omega^2 (c_0 + 1 - exp(-tau_0 (1+z)^beta))^2)),  g = df/dx,
//   x = [vec M (column-major P x k); log omega (P); log c_0; log tau_0; log beta].
// With num_forest_lines > 0 the noise model is the Lyman-series one of multi_dlas/spectrum_loss_lyseries.m:20-47 and
// multi_dlas/objective_lyseries.m:13-87: the effective optical depth sums the series members that lie below the quasar,
//   tau(z) = tau_0 (1+z)^beta + sum_{l >= 2} tau_0 lambda_l f_l / (lambda_1 f_1) [(1+z_l)^beta if 1+z_l <= 1+z_qso],
// with 1 + z_qso taken from the last column of the quasar's lya_1pzs row (objective_lyseries.m:46).
// A CTA works through training spectra with the same Woodbury algebra as the hot path (B = I + M' D^-1 M, Cholesky):
//   C M = I - B^-1  =>  K^-1 M = D^-1 M B^-1,  diag K^-1 = d^-1 - rowdot(D^-1 M B^-1, D^-1 M)
// so no n x n or k x n intermediate is formed.  Two-stage reduction: every CTA of the (resident-sized) grid adds the
// gradients of its spectra into its own partial vector (one writer per vector, so the sums are in a fixed order and the
// result is the same on every evaluation), objective_reduce_kernel sums the partial vectors.
#pragma once
#include <stdint.h>
#include <math.h>

namespace gpdla {

struct ObjectiveArgs {
  const double* y;        // centered_rest_fluxes   [N x P], NaN = pixel not observed (objective.m:42)
  const double* lya_1pz;  // lya_1pzs               [N x P]
  const double* nv;       // rest_noise_variances   [N x P]
  const double* x;        // parameters
  double* f;              // scalar
  double* g;              // gradient, layout of x
  double* partial;        // [grid][nx + 1] per-CTA partial gradients, then the partial objective value
  int64_t N;
  int P;
  int num_forest_lines;   // 0: objective.m / spectrum_loss.m; > 0: the Lyman-series variant
  double tw[MAX_LINES];   // all_transition_wavelengths (any unit: only ratios enter)   set_parameters_multi.m:77-109
  double osc[MAX_LINES];  // all_oscillator_strengths                                   set_parameters_multi.m:110-142
};

constexpr int OBJ_THREADS = 256;
constexpr int OBJ_TILE = 32;

template <int K>
__host__ __device__ constexpr size_t objective_smem_bytes(int P) {
  // idx[P] (int), dinv[P], kiy[P] + tile sM[32][K+1] + B[K][K+1] + Binv[K][K+1] + vectors
  return (size_t)P * 4 + 8 + (size_t)P * 16 + (size_t)OBJ_TILE * (K + 1) * 8 + 2ull * K * (K + 1) * 8 + 4ull * K * 8 + 64 * 8;
}

template <int K>
__global__ void __launch_bounds__(OBJ_THREADS, (K <= 20) ? 2 : 1) objective_kernel(ObjectiveArgs a) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = a.P;
  const int64_t nx = (int64_t)P * (K + 1) + 3;
  double* part = a.partial + (int64_t)blockIdx.x * (nx + 1);
  for (int64_t e = tid; e <= nx; e += OBJ_THREADS) part[e] = 0.0;
  double f_cta = 0.0, c0_cta = 0.0, tau_cta = 0.0, beta_cta = 0.0;   // thread 0's running sums over this CTA's spectra
  const double* M = a.x;                          // M(i, j) = x[j P + i]
  const double* log_omega = a.x + (int64_t)P * K;
  const double c_0 = exp(a.x[(int64_t)P * (K + 1)]), tau_0 = exp(a.x[(int64_t)P * (K + 1) + 1]),
               beta = exp(a.x[(int64_t)P * (K + 1) + 2]);                                    // objective.m:29-32

  extern __shared__ __align__(16) unsigned char smem[];
  double* dinv = reinterpret_cast<double*>(smem);                 // [P] by compact index
  double* kiy = dinv + P;                                          // [P] K^-1 y
  double* sM = kiy + P;                                            // [TILE][K+1]
  double* B = sM + OBJ_TILE * (K + 1);                             // [K][K+1]  B, then its Cholesky factor (upper)
  double* Binv = B + K * (K + 1);                                  // [K][K+1]
  double* v0 = Binv + K * (K + 1);                                 // [K] M' D^-1 y
  double* v1 = v0 + K;                                             // [K] B^-1 M' D^-1 y  (= C y)
  double* wv = v1 + K;                                             // [K] (K^-1 y)' M
  double* red = wv + 2 * K;                                        // [64] reductions
  int* idx = reinterpret_cast<int*>(red + 64);                     // [P] valid pixels
  __shared__ int s_cnt[OBJ_THREADS / 32];

  for (int64_t s = blockIdx.x; s < a.N; s += gridDim.x) {
  const double* y = a.y + s * P;
  const double* z1 = a.lya_1pz + s * P;
  const double* nv = a.nv + s * P;
  const double zq1 = z1[P - 1];                   // 1 + z_qso   (objective_lyseries.m:46)
  __syncthreads();                                // the previous spectrum's shared arrays are dead
  // ---- valid pixels (objective.m:42), compacted in order
  int base = 0;
  for (int i0 = 0; i0 < P; i0 += OBJ_THREADS) {
    const int i = i0 + tid;
    const bool ok = i < P && !isnan(y[i]);
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += s_cnt[w];
    if (ok) idx[off + __popc(bal & ((1u << lane) - 1))] = i;
    for (int w = 0; w < OBJ_THREADS / 32; ++w) base += s_cnt[w];
    __syncthreads();
  }
  const int n = base;
  if (n == 0) continue;

  // per-pixel noise model (spectrum_loss.m:21-31)
  auto noise_terms = [&](int i, double& od, double& ab, double& sf, double& an) {
    od = tau_0 * pow(z1[i], beta);                 // lya_optical_depth
    for (int l = 1; l < a.num_forest_lines; ++l) { // spectrum_loss_lyseries.m:29-40
      double lyman_1pz = a.tw[0] * z1[i] / a.tw[l];
      lyman_1pz = (lyman_1pz <= zq1) ? lyman_1pz : 0.0;
      const double tau = tau_0 * a.tw[l] * a.osc[l] / (a.tw[0] * a.osc[0]);
      od = od + tau * pow(lyman_1pz, beta);
    }
    ab = exp(-od);                                 // lya_absorption
    sf = 1.0 - ab + c_0;                           // scaling_factor
    an = exp(2.0 * log_omega[i]) * (sf * sf);      // absorption_noise
  };
  double ld = 0.0;
  for (int t = tid; t < n; t += OBJ_THREADS) {
    const int i = idx[t];
    double od, ab, sf, an;
    noise_terms(i, od, ab, sf, an);
    const double d = nv[i] + an;
    dinv[t] = 1.0 / d;
    ld += log(d);
  }
  for (int t = tid; t < K * (K + 1); t += OBJ_THREADS) B[t] = 0.0;
  if (tid < K) v0[tid] = 0.0;
  __syncthreads();

  // ---- B = I + M' D^-1 M and v0 = M' D^-1 y, tiled over pixels (spectrum_loss.m:40-41)
  constexpr int NPAIR = K * (K + 1) / 2;
  constexpr int EPT = (NPAIR + K + OBJ_THREADS - 1) / OBJ_THREADS;   // entries (pairs p <= q, then v0 columns) per thread
  int bp[EPT], bq[EPT];
  double acc[EPT];
#pragma unroll
  for (int m = 0; m < EPT; ++m) {
    const int e = tid + m * OBJ_THREADS;
    bp[m] = 0; bq[m] = 0; acc[m] = 0.0;
    if (e < NPAIR) { int c = e, p = 0; while (c >= K - p) { c -= K - p; ++p; } bp[m] = p; bq[m] = p + c; }
    else if (e < NPAIR + K) bp[m] = e - NPAIR;
  }
  for (int t0 = 0; t0 < n; t0 += OBJ_TILE) {
    const int nt = min(OBJ_TILE, n - t0);
    for (int e = tid; e < OBJ_TILE * K; e += OBJ_THREADS) {
      const int r = e % OBJ_TILE, j = e / OBJ_TILE;
      sM[r * (K + 1) + j] = (r < nt) ? M[(int64_t)j * P + idx[t0 + r]] : 0.0;
    }
    __syncthreads();
    for (int r = 0; r < nt; ++r) {
      const double di = dinv[t0 + r];
      const double* row = sM + r * (K + 1);
#pragma unroll
      for (int m = 0; m < EPT; ++m) {
        const int e = tid + m * OBJ_THREADS;
        if (e < NPAIR) acc[m] = fma(di * row[bp[m]], row[bq[m]], acc[m]);
        else if (e < NPAIR + K) acc[m] = fma(di * row[bp[m]], y[idx[t0 + r]], acc[m]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int m = 0; m < EPT; ++m) {
    const int e = tid + m * OBJ_THREADS;
    if (e < NPAIR) {
      const double v = acc[m] + (bp[m] == bq[m] ? 1.0 : 0.0);
      B[bp[m] * (K + 1) + bq[m]] = v; B[bq[m] * (K + 1) + bp[m]] = v;
    } else if (e < NPAIR + K) {
      v0[bp[m]] = acc[m];
    }
  }
  __syncthreads();

  // ---- Cholesky B = R'R (upper, as MATLAB chol; :42), one warp
  if (warp == 0) {
    for (int p = 0; p < K; ++p) {
      double dpp = B[p * (K + 1) + p];
      for (int r = 0; r < p; ++r) dpp -= B[r * (K + 1) + p] * B[r * (K + 1) + p];
      const double rpp = sqrt(dpp);
      __syncwarp();
      for (int q = p + lane; q < K; q += 32) {
        double v = B[p * (K + 1) + q];
        for (int r = 0; r < p; ++r) v -= B[r * (K + 1) + p] * B[r * (K + 1) + q];
        B[p * (K + 1) + q] = (q == p) ? rpp : v / rpp;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // B^-1 column j: solve R'R x = e_j
  if (tid < K) {
    double xcol[K];
    for (int r = 0; r < K; ++r) {                   // R' w = e_j
      double v = (r == tid) ? 1.0 : 0.0;
      for (int c = 0; c < r; ++c) v -= B[c * (K + 1) + r] * xcol[c];
      xcol[r] = v / B[r * (K + 1) + r];
    }
    for (int r = K - 1; r >= 0; --r) {              // R x = w
      double v = xcol[r];
      for (int c = r + 1; c < K; ++c) v -= B[r * (K + 1) + c] * xcol[c];
      xcol[r] = v / B[r * (K + 1) + r];
    }
    for (int r = 0; r < K; ++r) Binv[r * (K + 1) + tid] = xcol[r];
  }
  __syncthreads();
  if (tid < K) {
    double v = 0.0;
    for (int c = 0; c < K; ++c) v = fma(Binv[tid * (K + 1) + c], v0[c], v);
    v1[tid] = v;                                    // C y (:44-46)
    wv[tid] = 0.0;
  }
  double logdetB = 0.0;
  for (int p = 0; p < K; ++p) logdetB += log(B[p * (K + 1) + p]);   // sum log diag L (:48)
  __syncthreads();

  // ---- pass 1 over pixels: K^-1 y (needs only v1), then w = (K^-1 y)' M
  double quad = 0.0;
  for (int t = tid; t < n; t += OBJ_THREADS) {
    const int i = idx[t];
    const double di = dinv[t], yi = y[i];
    double dot_v1 = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) dot_v1 = fma(di * M[(int64_t)j * P + i], v1[j], dot_v1);
    const double ky = di * yi - dot_v1;                                                    // K^-1 y (:46)
    kiy[t] = ky;
    quad = fma(yi, ky, quad);
  }
  __syncthreads();
  for (int j = warp; j < K; j += OBJ_THREADS / 32) {      // w = (K^-1 y)' M, one warp per column
    double t = 0.0;
    for (int u = lane; u < n; u += 32) t = fma(kiy[u], M[(int64_t)j * P + idx[u]], t);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) wv[j] = t;
  }
  __syncthreads();
  // ---- pass 2: K^-1 M = D^-1 M B^-1, diag K^-1, every gradient; one atomic per entry of dM:
  //      dM(i, :) = K^-1 M (i, :) - K^-1 y (i) w   (:54-55)
  double g_c0 = 0.0, g_tau = 0.0, g_beta = 0.0;
  double* gM = part;
  double* gom = part + (int64_t)P * K;
  for (int t = tid; t < n; t += OBJ_THREADS) {
    const int i = idx[t];
    const double di = dinv[t], ky = kiy[t];
    double r[K];
#pragma unroll
    for (int j = 0; j < K; ++j) r[j] = di * M[(int64_t)j * P + i];                       // D^-1 M row (:33)
    double diagk = di;
#pragma unroll 4
    for (int j = 0; j < K; ++j) {
      double tj = 0.0;                                                                     // (D^-1 M B^-1)(i, j)
#pragma unroll
      for (int c = 0; c < K; ++c) tj = fma(r[c], Binv[c * (K + 1) + j], tj);
      diagk = fma(-tj, r[j], diagk);                                                       // diag K^-1 (:58)
      // this CTA's own partial vector: no other writer, and one add per entry and spectrum in spectrum order -- the
      // reduction instruction (no load-add-store round trip through L2 in the inner loop) keeps the sum deterministic
      atomicAdd(&gM[(int64_t)j * P + i], fma(-ky, wv[j], tj));
    }
    double od, ab, sf, an;
    noise_terms(i, od, ab, sf, an);
    const double om2 = exp(2.0 * log_omega[i]);
    const double kk = ky * ky - diagk;
    atomicAdd(&gom[i], -(an * kk));                                                        // dlog_omega (:61)
    double da = c_0 * om2 * sf;                                                            // (:64-65)
    g_c0 -= da * kk;
    da = om2 * sf * od * ab;                                                               // (:68-69)
    g_tau -= da * kk;
    da = da * log(z1[i]) * beta;                                                           // (:72-73)
    g_beta -= da * kk;
  }
  auto block_reduce = [&](double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double tsum = 0.0;
    for (int w = 0; w < OBJ_THREADS / 32; ++w) tsum += red[w];
    return tsum;
  };
  quad = block_reduce(quad); ld = block_reduce(ld);
  g_c0 = block_reduce(g_c0); g_tau = block_reduce(g_tau); g_beta = block_reduce(g_beta);
  if (tid == 0) {
    const double log_2pi = 1.83787706640934534;                                            // (:17)
    f_cta += 0.5 * (quad + (ld + 2.0 * logdetB) + n * log_2pi);                            // (:48-52)
    c0_cta += g_c0; tau_cta += g_tau; beta_cta += g_beta;
  }
  }   // spectra of this CTA
  if (tid == 0) {
    part[(int64_t)P * (K + 1)] = c0_cta; part[(int64_t)P * (K + 1) + 1] = tau_cta; part[(int64_t)P * (K + 1) + 2] = beta_cta;
    part[nx] = f_cta;
  }
}

// second stage: g[e] = sum over the CTAs' partial vectors (fixed order), f likewise
__global__ void objective_reduce_kernel(const double* __restrict__ partial, int nparts, int64_t nx, double* __restrict__ g,
                                        double* __restrict__ f) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e > nx) return;
  double v = 0.0;
  for (int b = 0; b < nparts; ++b) v += partial[(int64_t)b * (nx + 1) + e];
  if (e < nx) g[e] = v; else *f = v;
}

// priors on tau_0 and beta (Kim et al. 2007), objective.m:59-71
__global__ void objective_prior_kernel(const double* x, double* g, int64_t off) {
  const double tau_0 = exp(x[off + 1]), beta = exp(x[off + 2]);
  const double tau_0_mu = 0.0023, tau_0_sigma = 0.0007, beta_mu = 3.65, beta_sigma = 0.21;
  g[off + 1] += tau_0 * (tau_0 - tau_0_mu) / (tau_0_sigma * tau_0_sigma);
  g[off + 2] += beta * (beta - beta_mu) / (beta_sigma * beta_sigma);
}

}  // namespace gpdla
