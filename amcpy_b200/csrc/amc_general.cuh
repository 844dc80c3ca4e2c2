// General kernels: any frame length >= 1, any element strides, complex64 or complex128.
// One CTA (256 threads) per frame, float64 arithmetic with the reference's exact unwrap rules;
// used for every shape the fused kernel does not cover (ragged sizes, strided / sample-major
// views, the reference's 10-sample fixture) and for the helper value types.
#pragma once
#include "amc_device.cuh"

namespace amc {

constexpr int kGenThreads = 256;
constexpr int kGenWarps = kGenThreads / 32;

template <typename CT>
__device__ __forceinline__ void load_strided(const CT* __restrict__ base, int64_t n, int64_t sample_stride,
                                             double& a, double& b) {
  const CT v = base[n * sample_stride];
  a = static_cast<double>(v.x);
  b = static_cast<double>(v.y);
}

// Sum V doubles over the CTA; every thread receives the totals.  Fixed order (reproducible).
// red: shared double[kGenWarps * V].
template <int V>
__device__ __forceinline__ void block_sum(double (&v)[V], double* red, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    double x = v[i];
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) x += __shfl_xor_sync(0xffffffffu, x, w);
    if (lane == 0) red[warp * V + i] = x;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < V; ++i) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kGenWarps; ++w) s += red[w * V + i];
    v[i] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ double block_max(double x, double* red, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, w));
  if (lane == 0) red[warp] = x;
  __syncthreads();
  double m = red[0];
#pragma unroll
  for (int w = 1; w < kGenWarps; ++w) m = fmax(m, red[w]);
  __syncthreads();
  return m;
}

// In-place radix-2 FFT passes over `m` float2 in shared memory (m a power of two, every thread of the CTA calls).
// Forward: decimation in frequency, natural order in -> bit-reversed order out.
__device__ __forceinline__ void smem_fft_dif(float2* buf, int m, int tid) {
  for (int half = m >> 1; half >= 1; half >>= 1) {
    const float inv_half = 1.0f / static_cast<float>(half);
    for (int p = tid; p < (m >> 1); p += kGenThreads) {
      const int off = p & (half - 1);
      const int i0 = ((p - off) << 1) + off;
      const float2 u = buf[i0], w = buf[i0 + half];
      buf[i0] = make_float2(u.x + w.x, u.y + w.y);
      const float dx = u.x - w.x, dy = u.y - w.y;
      float sn, cs;
      sincospif(-static_cast<float>(off) * inv_half, &sn, &cs);
      buf[i0 + half] = make_float2(dx * cs - dy * sn, dx * sn + dy * cs);
    }
    __syncthreads();
  }
}
// Inverse (unscaled: m * IDFT): decimation in time with conjugate twiddles, bit-reversed order in -> natural order
// out - the stages of smem_fft_dif undone in reverse order.
__device__ __forceinline__ void smem_ifft_dit(float2* buf, int m, int tid) {
  for (int half = 1; half < m; half <<= 1) {
    const float inv_half = 1.0f / static_cast<float>(half);
    for (int p = tid; p < (m >> 1); p += kGenThreads) {
      const int off = p & (half - 1);
      const int i0 = ((p - off) << 1) + off;
      const float2 u = buf[i0], w = buf[i0 + half];
      float sn, cs;
      sincospif(static_cast<float>(off) * inv_half, &sn, &cs);
      const float tx = w.x * cs - w.y * sn, ty = w.x * sn + w.y * cs;
      buf[i0] = make_float2(u.x + tx, u.y + ty);
      buf[i0 + half] = make_float2(u.x - tx, u.y - ty);
    }
    __syncthreads();
  }
}

// fft_mode: 0 = direct DFT (float64, twiddle table of N double2 in dynamic smem),
//           1 = power-of-two in-place radix-2 DIF (float32, N float2 in dynamic smem),
//           2 = Bluestein (chirp-z) for other lengths: x_n c_n (c_n = exp(-i pi n^2 / N)) zero-padded to bl_m =
//               pow2 >= 2N-1, FFT, times the pre-transformed conjugate chirp `bl_bfft` (bit-reversed order, 1/bl_m
//               folded in), inverse FFT; |X_k| = |conv_k| for k < N, so the closing chirp multiply is not needed for
//               the spectral max (float32, bl_m float2 in dynamic smem; tables built once per N by the host side)
//           The float32 transforms (modes 1, 2) run on the frame scaled by an exact power of two taken from its mean
//           power, so that neither the samples nor |X_k|^2 leave the float32 range whatever the input's scale is; the
//           maximum is scaled back in float64.
// cache_off >= 0: byte offset in dynamic smem of 2 N doubles that keep every sample's phase and amplitude between the
//           passes (one libdevice atan2 + one hypot per sample instead of four + two; same values, same summation
//           order, so the results do not depend on it); it may alias the FFT buffer, which is only used afterwards.
//           -1 when it does not fit (N > 12800): the values are recomputed.
// One frame, all 256 threads of the CTA.  This is the CAREFUL path of the library: float64 statistics with the
// reference's own formulas (np.abs = hypot, np.angle = atan2, np.unwrap's rules, two-pass std / kurtosis, three passes for
// feature 4: std(|abs/mean(abs) - 1|), features.py:82-85).
template <typename CT>
__device__ __forceinline__ void general_frame(const CT* __restrict__ iq, int64_t f, int n, int64_t frame_stride,
                                              int64_t sample_stride, double* __restrict__ out, int64_t out_stride,
                                              int fft_mode, const float2* __restrict__ bl_chirp,
                                              const float2* __restrict__ bl_bfft, int bl_m, int cache_off,
                                              unsigned char* dyn, double* red, int tid, float2* fft_ws) {
  const bool cached = cache_off >= 0;
  double* ph_c = reinterpret_cast<double*>(dyn + (cached ? cache_off : 0));
  double* r_c = ph_c + n;
  const CT* base = iq + f * frame_stride;

  // ---- pass 1: raw sums -----------------------------------------------------------
  Monomials mono;
  mono.clear();
  double sr = 0.0, sph = 0.0, saph = 0.0, sfq = 0.0;
  for (int i = tid; i < n; i += kGenThreads) {
    double a, b;
    load_strided(base, i, sample_stride, a, b);
    mono.add(a, b);
    const double r0 = hypot(a, b);                  // np.abs == hypot (features.py:27)
    sr += r0;
    const double p0 = atan2_exact(b, a);            // np.angle  (features.py:28)
    sph += p0;
    saph += fabs(p0);
    if (cached) {
      ph_c[i] = p0;
      r_c[i] = r0;
    } else if (i + 1 < n) {
      double a1, b1;
      load_strided(base, i + 1, sample_stride, a1, b1);
      sfq += unwrap_step(atan2_exact(b1, a1) - p0) / kTwoPi;   // features.py:29-30
    }
  }
  if (cached) {
    __syncthreads();
    for (int i = tid; i + 1 < n; i += kGenThreads) sfq += unwrap_step(ph_c[i + 1] - ph_c[i]) / kTwoPi;
  }
  double v1[19];
#pragma unroll
  for (int i = 0; i < 15; ++i) v1[i] = mono.s[i];
  v1[15] = sr;
  v1[16] = sph;
  v1[17] = saph;
  v1[18] = sfq;
  block_sum<19>(v1, red, tid);
  const double dn = static_cast<double>(n);
  const double mu_r = v1[15] / dn, mu_ph = v1[16] / dn, mu_aph = v1[17] / dn;
  const double mu_f = v1[18] / (dn - 1.0);

  // ---- pass 2: centred sums ---------------------------------------------------------
  double v2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = tid; i < n; i += kGenThreads) {
    double r0, p0, p1 = 0.0;
    if (cached) {
      r0 = r_c[i];
      p0 = ph_c[i];
      if (i + 1 < n) p1 = ph_c[i + 1];
    } else {
      double a, b;
      load_strided(base, i, sample_stride, a, b);
      r0 = hypot(a, b);
      p0 = atan2_exact(b, a);
      if (i + 1 < n) {
        double a1, b1;
        load_strided(base, i + 1, sample_stride, a1, b1);
        p1 = atan2_exact(b1, a1);
      }
    }
    const double d = r0 - mu_r;
    const double d2 = d * d;
    v2[0] += fabs(d);
    v2[1] += d2;
    v2[2] += d2 * d2;
    v2[7] += fabs(r0 / mu_r - 1.0);                 // |cn_amplitude| exactly as features.py:31,84 forms it
    const double e = p0 - mu_ph, ea = fabs(p0) - mu_aph;
    v2[3] += e * e;
    v2[4] += ea * ea;
    if (i + 1 < n) {
      const double ef = unwrap_step(p1 - p0) / kTwoPi - mu_f;
      const double ef2 = ef * ef;
      v2[5] += ef2;
      v2[6] += ef2 * ef2;
    }
  }
  block_sum<8>(v2, red, tid);

  // ---- pass 3: spread of |cn_amplitude| about its own mean (feature 4, np.std's two passes) ----------
  double v3[1] = {0.0};
  {
    const double mu_acn = v2[7] / dn;
    for (int i = tid; i < n; i += kGenThreads) {
      double r0;
      if (cached) {
        r0 = r_c[i];
      } else {
        double a, b;
        load_strided(base, i, sample_stride, a, b);
        r0 = hypot(a, b);
      }
      const double e = fabs(r0 / mu_r - 1.0) - mu_acn;
      v3[0] += e * e;
    }
  }
  block_sum<1>(v3, red, tid);

  // ---- spectrum max -----------------------------------------------------------------
  // exact power-of-two scale for the float32 transforms: 2^-e with 4^e ~ mean |x|^2
  double scale = 1.0, unscale2 = 1.0;
  {
    const double pw = (v1[0] + v1[1]) / dn;
    if (pw > 0.0 && pw < 1.7976931348623157e308) {
      int e2 = 0;
      frexp(pw, &e2);
      const int e = e2 / 2;
      if (e > 40 || e < -40) {                      // ordinary data is left untouched (bitwise the unscaled result)
        scale = ldexp(1.0, -e);
        unscale2 = ldexp(1.0, 2 * e);
      }
    }
  }
  double smax = 0.0;
  if (fft_mode == 1) {
    // transform buffer: shared memory, or - frames too long for it - this CTA's slice of a global workspace (L2)
    float2* buf = fft_ws ? fft_ws : reinterpret_cast<float2*>(dyn);
    __syncthreads();                                // the phase / amplitude cache (which may alias buf) is dead now
    for (int i = tid; i < n; i += kGenThreads) {
      double a, b;
      load_strided(base, i, sample_stride, a, b);
      buf[i] = make_float2(static_cast<float>(a * scale), static_cast<float>(b * scale));
    }
    __syncthreads();
    smem_fft_dif(buf, n, tid);
    for (int i = tid; i < n; i += kGenThreads) {
      const float2 u = buf[i];
      smax = fmax(smax, static_cast<double>(u.x) * u.x + static_cast<double>(u.y) * u.y);
    }
    __syncthreads();
  } else if (fft_mode == 2) {
    float2* buf = fft_ws ? fft_ws : reinterpret_cast<float2*>(dyn);
    __syncthreads();
    for (int i = tid; i < bl_m; i += kGenThreads) {
      float2 v = make_float2(0.0f, 0.0f);
      if (i < n) {
        double a, b;
        load_strided(base, i, sample_stride, a, b);
        const float2 c = bl_chirp[i];
        const float af = static_cast<float>(a * scale), bf = static_cast<float>(b * scale);
        v = make_float2(af * c.x - bf * c.y, af * c.y + bf * c.x);
      }
      buf[i] = v;
    }
    __syncthreads();
    smem_fft_dif(buf, bl_m, tid);
    for (int i = tid; i < bl_m; i += kGenThreads) {
      const float2 u = buf[i], w = bl_bfft[i];
      buf[i] = make_float2(u.x * w.x - u.y * w.y, u.x * w.y + u.y * w.x);
    }
    __syncthreads();
    smem_ifft_dit(buf, bl_m, tid);
    for (int i = tid; i < n; i += kGenThreads) {
      const float2 u = buf[i];
      smax = fmax(smax, static_cast<double>(u.x) * u.x + static_cast<double>(u.y) * u.y);
    }
    __syncthreads();
  } else {
    const double2* tw = reinterpret_cast<const double2*>(dyn);
    for (int k = tid; k < n; k += kGenThreads) {
      double xr = 0.0, xi = 0.0;
      int idx = 0;
      for (int i = 0; i < n; ++i) {
        double a, b;
        load_strided(base, i, sample_stride, a, b);
        const double2 w = tw[idx];
        xr += a * w.x - b * w.y;
        xi += a * w.y + b * w.x;
        idx += k;
        if (idx >= n) idx -= n;
      }
      smax = fmax(smax, xr * xr + xi * xi);
    }
  }
  smax = block_max(smax, red, tid);
  if (fft_mode != 0) smax *= unscale2;

  if (tid == 0) {
    FrameSums fs;
#pragma unroll
    for (int i = 0; i < 15; ++i) fs.mono[i] = v1[i];
    fs.sum_r = v1[15];
    fs.c_abs1 = v2[0];
    fs.c_abs2 = v3[0];
    fs.c2 = v2[1];
    fs.c4 = v2[2];
    fs.ph_m2 = v2[3];
    fs.aph_m2 = v2[4];
    fs.f_m2 = v2[5];
    fs.f_m4 = v2[6];
    fs.mean_f = mu_f;
    fs.spec_max = smax;
    double res[18];
    finalize_features(fs, n, res, 0);
#pragma unroll
    for (int i = 0; i < 18; ++i) out[f * out_stride + i] = res[i];
  }
}

// redo_only != 0: careful-path pass behind a fused kernel - rows whose column 0 carries kRedoTagBits
// (finalize_features) are recomputed, every other row is left alone.
template <typename CT>
__global__ void __launch_bounds__(kGenThreads)
general_features_kernel(const CT* __restrict__ iq, int64_t n_frames, int n, int64_t frame_stride,
                        int64_t sample_stride, double* __restrict__ out, int64_t out_stride, int fft_mode,
                        const float2* __restrict__ bl_chirp, const float2* __restrict__ bl_bfft, int bl_m,
                        int cache_off, int redo_only, unsigned long long ticket, float2* fft_ws) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ double red[kGenWarps * 20];
  __shared__ int redo_rows[kGenThreads];
  __shared__ int redo_count;
  const int tid = threadIdx.x;

  if (redo_only) {
    pdl_launch_dependents();   // whatever follows in the stream may be scheduled; it waits for this grid in turn
    pdl_wait_primary();        // launched behind a fused kernel (possibly before it finished): its rows are visible from here
    // nothing tagged by launch `ticket` (the slot is still below it): done - the case for ordinary data
    if (ticket != 0 && *reinterpret_cast<volatile unsigned long long*>(&g_redo_ring[ticket & 255]) < ticket) return;
    // each CTA scans a contiguous block of rows, 256 at a time (one coalesced-ish load per thread, one barrier)
    const int64_t per = (n_frames + gridDim.x - 1) / gridDim.x;
    const int64_t lo = static_cast<int64_t>(blockIdx.x) * per;
    const int64_t hi = lo + per < n_frames ? lo + per : n_frames;
    for (int64_t r0 = lo; r0 < hi; r0 += kGenThreads) {
      if (tid == 0) redo_count = 0;
      __syncthreads();
      const int64_t r = r0 + tid;
      if (r < hi && static_cast<unsigned long long>(__double_as_longlong(out[r * out_stride])) == kRedoTagBits)
        redo_rows[atomicAdd(&redo_count, 1)] = tid;   // (order is irrelevant: frames are independent)
      __syncthreads();
      const int cnt = redo_count;
      for (int j = 0; j < cnt; ++j) {
        general_frame<CT>(iq, r0 + redo_rows[j], n, frame_stride, sample_stride, out, out_stride, fft_mode, bl_chirp,
                          bl_bfft, bl_m, cache_off, dyn, red, tid, nullptr);
        __syncthreads();
      }
    }
    return;
  }

  if (fft_mode == 0) {  // twiddle table once per CTA
    double2* tw = reinterpret_cast<double2*>(dyn);
    for (int m = tid; m < n; m += kGenThreads) {
      double s, c;
      sincospi(-2.0 * static_cast<double>(m) / static_cast<double>(n), &s, &c);
      tw[m] = make_double2(c, s);
    }
    __syncthreads();
  }
  for (int64_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
    // fft_ws: one slice of max(n, bl_m) float2 per CTA
    float2* ws = fft_ws ? fft_ws + static_cast<size_t>(blockIdx.x) * static_cast<size_t>(fft_mode == 2 ? bl_m : n) : nullptr;
    general_frame<CT>(iq, f, n, frame_stride, sample_stride, out, out_stride, fft_mode, bl_chirp, bl_bfft, bl_m,
                      cache_off, dyn, red, tid, ws);
    __syncthreads();
  }
}

// ------------------------------------------------------------------ MomentValues (features.py:39-58)
template <typename CT>
__global__ void __launch_bounds__(kGenThreads)
moments_kernel(const CT* __restrict__ iq, int64_t n_frames, int n, int64_t frame_stride, int64_t sample_stride,
               double* __restrict__ out) {
  __shared__ double red[kGenWarps * 15];
  const int tid = threadIdx.x;
  for (int64_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
    const CT* base = iq + f * frame_stride;
    Monomials mono;
    mono.clear();
    for (int i = tid; i < n; i += kGenThreads) {
      double a, b;
      load_strided(base, i, sample_stride, a, b);
      mono.add(a, b);
    }
    block_sum<15>(mono.s, red, tid);
    if (tid == 0) {
      const Moments m = moments_from_monomials(mono.s, 1.0 / static_cast<double>(n));
      double* o = out + f * 22;
      const Cplx list[11] = {m.m20, {m.m21, 0.0}, m.m22, m.m40, m.m41, {m.m42, 0.0}, m.m43,
                             m.m60, m.m61, {m.m62, 0.0}, m.m63};
#pragma unroll
      for (int i = 0; i < 11; ++i) {
        o[2 * i] = list[i].re;
        o[2 * i + 1] = list[i].im;
      }
    }
  }
}

// ------------------------------------------------------------------ InstantaneousValues (features.py:17-31)
// unwrapped = phase + cumsum(ph_correct): block-wide inclusive scan of the 2*pi corrections
// (warp shuffles + one cross-warp step), carried across 256-sample chunks.
template <typename CT>
__global__ void __launch_bounds__(kGenThreads)
instantaneous_kernel(const CT* __restrict__ iq, int64_t n_frames, int n, int64_t frame_stride,
                     int64_t sample_stride, double* __restrict__ abs_out, double* __restrict__ phase_out,
                     double* __restrict__ unwrapped_out, double* __restrict__ freq_out,
                     double* __restrict__ cna_out) {
  __shared__ double red[kGenWarps * 2];
  __shared__ double warp_tot[kGenWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
    const CT* base = iq + f * frame_stride;
    double carry = 0.0;      // cumsum of corrections over previous chunks
    double sum_abs = 0.0;
    for (int c0 = 0; c0 < n; c0 += kGenThreads) {
      const int i = c0 + tid;
      double p = 0.0, corr = 0.0, r = 0.0;
      if (i < n) {
        double a, b;
        load_strided(base, i, sample_stride, a, b);
        r = hypot(a, b);
        p = atan2_exact(b, a);
        if (i > 0) {
          double a0, b0;
          load_strided(base, i - 1, sample_stride, a0, b0);
          const double dd = p - atan2_exact(b0, a0);
          corr = unwrap_step(dd) - dd;    // ph_correct (0 when |dd| < pi)
        }
        sum_abs += r;
        if (abs_out) abs_out[f * n + i] = r;
        if (phase_out) phase_out[f * n + i] = p;
      }
      // inclusive scan of corr over the chunk
      double x = corr;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
      }
      if (lane == 31) warp_tot[warp] = x;
      __syncthreads();
      double off = carry;
      for (int w = 0; w < warp; ++w) off += warp_tot[w];
      double chunk_tot = 0.0;
      for (int w = 0; w < kGenWarps; ++w) chunk_tot += warp_tot[w];
      const double cum = off + x;        // cumsum(ph_correct) up to and including sample i
      if (i < n) {
        const double up = p + cum;
        if (unwrapped_out) unwrapped_out[f * n + i] = up;
        if (freq_out && i > 0) {
          // diff(unwrapped)/(2*pi): previous unwrapped value = p_prev + (cum - corr)
          double a0, b0;
          load_strided(base, i - 1, sample_stride, a0, b0);
          const double up_prev = atan2_exact(b0, a0) + (cum - corr);
          freq_out[f * (n - 1) + (i - 1)] = (up - up_prev) / kTwoPi;
        }
      }
      carry += chunk_tot;
      __syncthreads();
    }
    double v[2] = {sum_abs, 0.0};
    block_sum<2>(v, red, tid);
    if (cna_out) {
      const double mean = v[0] / static_cast<double>(n);
      for (int i = tid; i < n; i += kGenThreads) {
        double a, b;
        load_strided(base, i, sample_stride, a, b);
        cna_out[f * n + i] = hypot(a, b) / mean - 1.0;
      }
    }
  }
}

// ------------------------------------------------------------------ sample-major -> one row per frame
// src element (f, n) at f + n*src_sample_stride; dst (f, n) at f*N + n.  32x32 tiles through smem.
template <typename CT>
__global__ void __launch_bounds__(256)
frames_from_sample_major_kernel(const CT* __restrict__ src, int64_t n_frames, int n, int64_t src_sample_stride,
                                CT* __restrict__ dst) {
  __shared__ CT tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int64_t f0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int n0 = blockIdx.y * 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int nn = n0 + ty + 8 * k;
    const int64_t ff = f0 + tx;
    if (nn < n && ff < n_frames) tile[ty + 8 * k][tx] = src[static_cast<int64_t>(nn) * src_sample_stride + ff];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t ff = f0 + ty + 8 * k;
    const int nn = n0 + tx;
    if (nn < n && ff < n_frames) dst[ff * n + nn] = tile[tx][ty + 8 * k];
  }
}

// Planar (split real / imaginary planes, as stored in a Level-5 .mat file) and sample-major:
// plane element (f, n) at f + n*src_sample_stride; dst (f, n) at f*N + n, interleaved complex.
// `im` may be NULL (real input).  32x32 tiles through smem.
template <typename RT, typename CT>
__global__ void __launch_bounds__(256)
frames_from_planar_kernel(const RT* __restrict__ re, const RT* __restrict__ im, int64_t n_frames, int n,
                          int64_t src_sample_stride, CT* __restrict__ dst) {
  __shared__ CT tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int64_t f0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int n0 = blockIdx.y * 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int nn = n0 + ty + 8 * k;
    const int64_t ff = f0 + tx;
    if (nn < n && ff < n_frames) {
      const int64_t o = static_cast<int64_t>(nn) * src_sample_stride + ff;
      CT v;
      v.x = re[o];
      v.y = im ? im[o] : static_cast<RT>(0);
      tile[ty + 8 * k][tx] = v;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t ff = f0 + ty + 8 * k;
    const int nn = n0 + tx;
    if (nn < n && ff < n_frames) dst[ff * n + nn] = tile[tx][ty + 8 * k];
  }
}

}  // namespace amc
