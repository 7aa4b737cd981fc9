// Spectrum preprocessing on the device: the step before the hot path.
//   read_spec.m:28-38      wavelengths = 10.^loglam, noise_variance = 1./ivar,
//                          pixel_mask = (ivar == 0) | bitget(and_mask, BRIGHTSKY = 24)
//   preload_qsos.m:26-67   median normalisation in the rest-frame window [1310, 1325] A of the unmasked pixels,
//                          filter flags (bit 3: cannot normalise, bit 4: fewer than min_num_pixels usable pixels),
//                          truncation to [910, 1217] A plus one unmasked pixel on either side
// One CTA per spectrum; outputs are the padded [Q x L_out] planes + lengths that gpdla_process_qsos_device reads.
#pragma once
#include <stdint.h>
#include <math.h>

namespace gpdla {

struct PreloadArgs {
  // raw coadd columns of the speclite FITS table, padded [Q x L_in] (read_spec.m:11-26)
  const double* flux;
  const double* loglam;
  const double* ivar;
  const int32_t* and_mask;
  const int32_t* lengths_in;
  const double* z_qsos;
  const uint8_t* filter_flags_in;   // nullable; > 0: the quasar is skipped (preload_qsos.m:19-21)
  int64_t L_in, L_out;
  // set_parameters.m:21-34
  double loading_min_lambda, loading_max_lambda, normalization_min_lambda, normalization_max_lambda, min_lambda, max_lambda;
  int min_num_pixels;
  // outputs
  double* wavelengths;       // [Q x L_out]
  double* out_flux;
  double* noise_variance;
  uint8_t* pixel_mask;       // padding pixels are masked
  int32_t* lengths;          // 0 for skipped / filtered quasars; -(needed length) if the spectrum does not fit L_out
  double* normalizers;       // all_normalizers (0 where not set, preload_qsos.m:16)
  uint8_t* filter_flags;     // input flags | 4 (bit 3) | 8 (bit 4)
};

constexpr int PRE_THREADS = 256;

__device__ __forceinline__ int block_sum_i32(int v, int* scratch) {
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
  for (int w = 0; w < PRE_THREADS / 32; ++w) t += scratch[w];
  return t;
}

__global__ void __launch_bounds__(PRE_THREADS) preload_qsos_kernel(PreloadArgs a) {
  const int q = blockIdx.x, tid = threadIdx.x;
  const int64_t L = a.lengths_in[q];
  const double z = a.z_qsos[q];
  const double* fl = a.flux + q * a.L_in;
  const double* ll = a.loglam + q * a.L_in;
  const double* iv = a.ivar + q * a.L_in;
  const int32_t* am = a.and_mask + q * a.L_in;
  double* ow = a.wavelengths + q * a.L_out;
  double* of = a.out_flux + q * a.L_out;
  double* ov = a.noise_variance + q * a.L_out;
  uint8_t* om = a.pixel_mask + q * a.L_out;

  __shared__ int s_red[PRE_THREADS / 32];
  __shared__ int s_scan[PRE_THREADS];
  __shared__ double s_med[2];
  __shared__ int s_first, s_last, s_lo, s_hi;

  uint8_t flags = a.filter_flags_in ? a.filter_flags_in[q] : 0;
  for (int64_t i = tid; i < a.L_out; i += PRE_THREADS) { ow[i] = 0.0; of[i] = 0.0; ov[i] = 1.0; om[i] = 1; }
  if (tid == 0) { a.lengths[q] = 0; a.normalizers[q] = 0.0; a.filter_flags[q] = flags; }
  if (flags > 0) return;                                                   // preload_qsos.m:19-21

  auto wavelength = [&](int64_t i) { return exp10(ll[i]); };               // read_spec.m:28
  auto masked = [&](int64_t i) { return (iv[i] == 0.0) || ((am[i] >> 23) & 1); };   // read_spec.m:35-37 (bit 24, 1-based)
  const double opz = 1.0 + z;
  auto rest = [&](int64_t i) { return wavelength(i) / opz; };              // emitted_wavelengths, set_parameters.m:14-15

  // ---- nanmedian of the flux in the normalisation window (preload_qsos.m:28-33): rank by counting
  auto in_norm = [&](int64_t i) {
    const double r = rest(i);
    return r >= a.normalization_min_lambda && r <= a.normalization_max_lambda && !masked(i) && !isnan(fl[i]);
  };
  int cnt = 0;
  for (int64_t i = tid; i < L; i += PRE_THREADS) cnt += in_norm(i) ? 1 : 0;
  const int n_norm = block_sum_i32(cnt, s_red);
  if (n_norm == 0) {                                                       // bit 2: cannot normalise (:36-39)
    if (tid == 0) a.filter_flags[q] = flags | 4;
    return;
  }
  if (tid == 0) { s_med[0] = 0.0; s_med[1] = 0.0; }
  __syncthreads();
  const int k_lo = (n_norm - 1) / 2, k_hi = n_norm / 2;                    // the two middle order statistics
  for (int64_t i = tid; i < L; i += PRE_THREADS) {
    if (!in_norm(i)) continue;
    const double x = fl[i];
    int rank = 0;
    for (int64_t j = 0; j < L; ++j) {
      if (!in_norm(j)) continue;
      const double y = fl[j];
      rank += (y < x) || (y == x && j < i);
    }
    if (rank == k_lo) s_med[0] = x;
    if (rank == k_hi) s_med[1] = x;
  }
  __syncthreads();
  // MATLAB median of an even count: mean of the two middle values
  const double median = (k_lo == k_hi) ? s_med[0] : (s_med[0] + s_med[1]) / 2.0;

  // ---- enough usable pixels in the modelled window? (preload_qsos.m:41-49)
  cnt = 0;
  for (int64_t i = tid; i < L; i += PRE_THREADS) {
    const double r = rest(i);
    cnt += (r >= a.min_lambda && r <= a.max_lambda && !masked(i)) ? 1 : 0;
  }
  const int n_use = block_sum_i32(cnt, s_red);
  if (n_use < a.min_num_pixels) {                                          // bit 3: not enough pixels (:46-49)
    if (tid == 0) a.filter_flags[q] = flags | 8;
    return;
  }

  // ---- loading window plus one unmasked pixel on either side (preload_qsos.m:56-62)
  if (tid == 0) { s_first = 1 << 30; s_last = -1; s_lo = -1; s_hi = 1 << 30; }
  __syncthreads();
  int first = 1 << 30, last = -1;
  for (int64_t i = tid; i < L; i += PRE_THREADS) {
    const double r = rest(i);
    if (r >= a.loading_min_lambda && r <= a.loading_max_lambda) { first = min(first, (int)i); last = max(last, (int)i); }
  }
  atomicMin(&s_first, first); atomicMax(&s_last, last);
  __syncthreads();
  first = s_first; last = s_last;
  int lo = -1, hi = 1 << 30;       // last available pixel before `first`, first available pixel after `last`
  for (int64_t i = tid; i < L; i += PRE_THREADS) {
    const double r = rest(i);
    const bool inw = r >= a.loading_min_lambda && r <= a.loading_max_lambda;
    if (!inw && !masked(i)) {
      if ((int)i < first) lo = max(lo, (int)i);
      if ((int)i > last) hi = min(hi, (int)i);
    }
  }
  atomicMax(&s_lo, lo); atomicMin(&s_hi, hi);
  __syncthreads();
  lo = s_lo; hi = s_hi;

  // ---- stream compaction of the kept pixels, normalised (preload_qsos.m:51-54,64-67)
  const double med2 = median * median;
  int base = 0;
  for (int64_t i0 = 0; i0 < L; i0 += PRE_THREADS) {
    const int64_t i = i0 + tid;
    bool keep = false;
    if (i < L) {
      const double r = rest(i);
      keep = (r >= a.loading_min_lambda && r <= a.loading_max_lambda) || (int)i == lo || (int)i == hi;
    }
    // inclusive scan of the keep flags over the block
    s_scan[tid] = keep ? 1 : 0;
    __syncthreads();
    for (int o = 1; o < PRE_THREADS; o <<= 1) {
      const int v = (tid >= o) ? s_scan[tid - o] : 0;
      __syncthreads();
      s_scan[tid] += v;
      __syncthreads();
    }
    const int pos = base + s_scan[tid] - 1;
    if (keep) {
      if (pos < a.L_out) {
        ow[pos] = wavelength(i);
        of[pos] = fl[i] / median;                       // :51
        ov[pos] = (1.0 / iv[i]) / med2;                 // read_spec.m:31, preload_qsos.m:52
        om[pos] = masked(i) ? 1 : 0;
      }
    }
    base += s_scan[PRE_THREADS - 1];
    __syncthreads();
  }
  if (tid == 0) { a.lengths[q] = base <= (int)a.L_out ? base : -base; a.normalizers[q] = median; }
}

}  // namespace gpdla
