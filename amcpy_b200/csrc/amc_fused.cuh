// Building blocks shared by the fused sm_100a kernels (amc_fused16.cuh / amc_fusedw.cuh / amc_large.cuh): small FP32
// DFTs, twiddle tables of the radix-8 stages, sample loads, the FP64 tie re-decision, group barriers.
//
// With -DAMC_EXPERIMENTS the file also compiles the round-1 first-generation kernel (8 samples per thread,
// AMC_FLAG_FUSED_SPT8) for A/B runs; it is NOT part of the product library:
//
//   * persistent CTAs; each "group" of N/8 threads owns one frame at a time (N=2048: 256 threads
//     = the whole CTA; N=256: one warp per frame, 8 frames per CTA);
//   * the complex frame is staged into shared memory by one bulk TMA copy per frame
//     (cp.async.bulk + mbarrier complete_tx), two slots per group, so the copy of frame i+2
//     is in flight while frames i and i+1 are being computed;
//   * pass 1 (registers): 15 monomial sums + sum|x| in FP64, atan2 in FP32, wrapped phase
//     differences (np.unwrap semantics; near-tie decisions re-done in FP64);
//   * pass 2 (registers, no re-read): centred sums for the std / kurtosis features;
//   * spectral max: Stockham radix-8 FFT in FP32 through the (now free) TMA slot, XOR-swizzled;
//   * warp 0 of the group turns the sums into the 18 float64 outputs (one 144-byte row).
#pragma once
#include "amc_device.cuh"

namespace amc {

// Lane-contiguous twiddle tables for the 8-samples-per-thread kernel (one or two 128-byte lines
// per warp load instead of 8-32 with the generic table; see profiles/r1_experiments.txt):
//   g_tw8_s2[q-1][k]          = W_64^(k q)    q=1..7, k=0..7
//   g_tw8_s3a[q-1][j]         = W_256^(j q)   q=1..3, j=0..63     (N = 256, last stage)
//   g_tw8_s3b[q-1][k]         = W_512^(k q)   q=1..7, k=0..63
//   g_tw8_s4[off(N)+q-1][j]   = W_N^(j q)     q=1..N/512-1, j=0..511
__device__ float2 g_tw8_s2[7 * 8];
__device__ float2 g_tw8_s3a[3 * 64];
__device__ float2 g_tw8_s3b[7 * 64];
__device__ float2 g_tw8_s4[11 * 512];
__host__ __device__ constexpr int tw8_s4_offset(int n) { return (n == 1024 ? 0 : (n == 2048 ? 1 : 4)) * 512; }

__device__ __forceinline__ float2 tw_exact(int num, int den) {
  double s, c;
  sincospi(-2.0 * static_cast<double>(num) / static_cast<double>(den), &s, &c);
  return make_float2(static_cast<float>(c), static_cast<float>(s));
}
__global__ void init_twiddle8_kernel() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 7 * 8) g_tw8_s2[i] = tw_exact((i % 8) * (i / 8 + 1), 64);
  if (i < 3 * 64) g_tw8_s3a[i] = tw_exact((i % 64) * (i / 64 + 1), 256);
  if (i < 7 * 64) g_tw8_s3b[i] = tw_exact((i % 64) * (i / 64 + 1), 512);
  if (i < 11 * 512) {
    const int row = i / 512, j = i % 512;
    const int n = row < 1 ? 1024 : (row < 4 ? 2048 : 4096);
    const int q = row < 1 ? 1 : (row < 4 ? row : row - 3);
    g_tw8_s4[i] = tw_exact(j * q, n);
  }
}

// ------------------------------------------------------------------ small complex FP32 FFT pieces
__device__ __forceinline__ float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 c_mul(float2 a, float2 w) {
  return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}
__device__ __forceinline__ float2 c_mul_mj(float2 a) { return make_float2(a.y, -a.x); }  // * (-j)
__device__ __forceinline__ void bfly2(float2& a, float2& b) {
  const float2 t = a;
  a = c_add(t, b);
  b = c_sub(t, b);
}
// forward DFTs, outputs left in the registers in the order given by the comment
__device__ __forceinline__ void dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
  bfly2(v0, v2);
  bfly2(v1, v3);
  v3 = c_mul_mj(v3);
  bfly2(v0, v1);
  bfly2(v2, v3);
  // X0=v0 X2=v1 X1=v2 X3=v3
}
// dft4 of (a0, W8^1 a1, W8^2 a2, W8^3 a3): the 1/sqrt(2) of the two odd twiddles is not applied to them but rides on the
// last two butterflies as fused multiply-adds (20 instructions; multiplying first takes 24).  Same output order as dft4.
__device__ __forceinline__ void dft4_w8(float2& a0, float2& a1, float2& a2, float2& a3) {
  constexpr float h = 0.70710678118654752440f;
  const float px = a1.x + a1.y, py = a1.y - a1.x;              // W8^1 a1 = h (px, py)
  const float s3 = a3.x + a3.y, d3 = a3.y - a3.x;              // W8^3 a3 = h (d3, -s3)
  const float Px = px + d3, Py = py - s3;                      // (W8^1 a1 + W8^3 a3) / h
  const float Qx = px - d3, Qy = py + s3;                      // (W8^1 a1 - W8^3 a3) / h
  const float2 b0 = make_float2(a0.x + a2.y, a0.y - a2.x);     // a0 + (-j) a2
  const float2 b2 = make_float2(a0.x - a2.y, a0.y + a2.x);
  a0 = make_float2(fmaf(h, Px, b0.x), fmaf(h, Py, b0.y));
  a1 = make_float2(fmaf(-h, Px, b0.x), fmaf(-h, Py, b0.y));
  a2 = make_float2(fmaf(h, Qy, b2.x), fmaf(-h, Qx, b2.y));     // b2 + (-j) h Q
  a3 = make_float2(fmaf(-h, Qy, b2.x), fmaf(h, Qx, b2.y));
}
// dft4 of (a0, a1, h (ux, uy), a3); a2 is output only
__device__ __forceinline__ void dft4_hu(float2& a0, float2& a1, float2& a2, float2& a3, float ux, float uy) {
  constexpr float h = 0.70710678118654752440f;
  const float2 t = a0;
  a0 = make_float2(fmaf(h, ux, t.x), fmaf(h, uy, t.y));
  a2 = make_float2(fmaf(-h, ux, t.x), fmaf(-h, uy, t.y));
  bfly2(a1, a3);
  a3 = c_mul_mj(a3);
  bfly2(a0, a1);
  bfly2(a2, a3);
}
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  bfly2(v[0], v[4]);
  bfly2(v[1], v[5]);
  bfly2(v[2], v[6]);
  bfly2(v[3], v[7]);
  dft4(v[0], v[1], v[2], v[3]);
  dft4_w8(v[4], v[5], v[6], v[7]);
  // X0=v0 X4=v1 X2=v2 X6=v3 X1=v4 X5=v5 X3=v6 X7=v7
}
// register index that holds output X_r after dft8 / dft4 / dft2
__device__ __forceinline__ constexpr int out8(int r) {
  return ((r & 1) << 2) | (r & 2) | ((r >> 2) & 1);
}
__device__ __forceinline__ constexpr int out4(int r) { return ((r & 1) << 1) | (r >> 1); }

// XOR swizzle of a float2 index: conflict-free for the stride-8 / stride-64 Stockham scatters
// (low 4 bits ^= [e4, e5, e6, e6]); reads of consecutive indices stay conflict-free.
__device__ __forceinline__ int swz(int e) { return e ^ (((e >> 4) & 7) | (((e >> 6) & 1) << 3)); }

// same decision, returned as the unwrapped phase step in radians (the fused16 kernel keeps frequency in radians)
template <typename CT>
__device__ __noinline__ float exact_phase_step(const CT* xs, int n) {
  const CT v0 = xs[n], v1 = xs[n + 1];
  const double p0 = atan2_exact(static_cast<double>(v0.y), static_cast<double>(v0.x));
  const double p1 = atan2_exact(static_cast<double>(v1.y), static_cast<double>(v1.x));
  return static_cast<float>(unwrap_step(p1 - p0));
}

#ifdef AMC_EXPERIMENTS
template <int N, typename CT>
struct FusedCfg {
  static constexpr int SPT = 8;                         // samples per thread
  static constexpr int GROUP = N / SPT;                 // threads per frame
  static constexpr int CTA = GROUP < 256 ? 256 : GROUP; // threads per CTA
  static constexpr int G = CTA / GROUP;                 // frames in flight per CTA
  static constexpr int W = GROUP / 32;                  // warps per frame
  static constexpr int STAGES = 2;
  static constexpr int SLOT_BYTES = N * static_cast<int>(sizeof(CT));
  static constexpr bool C128 = sizeof(CT) == 16;
  static constexpr int FFTB_BYTES = C128 ? 0 : N * 8;   // c128: both FFT buffers live in the slot
  // per (parity, warp): 16 f64 (pass 1) + 4 f64 (pass 2) + 4 f32 (pass 1) + 4 f32 (pass 2) + max + pad
  static constexpr int PART_D = 20, PART_F = 12;
  static constexpr int PART_BYTES = PART_D * 8 + PART_F * 4;             // 208
  static constexpr int GROUP_BYTES = STAGES * SLOT_BYTES + FFTB_BYTES + 2 * W * PART_BYTES + 64;
  static constexpr int SMEM_BYTES = G * GROUP_BYTES;
  static constexpr int MIN_BLOCKS = (CTA <= 256) ? 2 : 1;
  static_assert(GROUP % 32 == 0 && GROUP >= 32 && GROUP <= 1024, "frame size outside fused range");
  static_assert(GROUP_BYTES % 16 == 0, "group region must keep 16-byte alignment");
};

#endif  // AMC_EXPERIMENTS

template <int GROUP, int CTA>
__device__ __forceinline__ void group_sync(int g) {
  if constexpr (GROUP == CTA) {
    __syncthreads();
  } else if constexpr (GROUP == 32) {
    __syncwarp();
  } else {
    named_bar_sync(1 + g, GROUP);
  }
}

template <typename CT>
__device__ __forceinline__ void load_sample(const CT* p, double& a, double& b, float& af, float& bf);
template <>
__device__ __forceinline__ void load_sample<double2>(const double2* p, double& a, double& b, float& af,
                                                     float& bf) {
  const double2 v = *p;
  a = v.x;
  b = v.y;
  af = static_cast<float>(a);
  bf = static_cast<float>(b);
}
template <>
__device__ __forceinline__ void load_sample<float2>(const float2* p, double& a, double& b, float& af,
                                                    float& bf) {
  const float2 v = *p;
  af = v.x;
  bf = v.y;
  a = static_cast<double>(af);
  b = static_cast<double>(bf);
}

// FP64 re-decision of one wrapped phase difference (rare: |dd| within kTieEps of pi).
template <typename CT>
__device__ __noinline__ float exact_freq_step(const CT* xs, int n) {
  const CT v0 = xs[n], v1 = xs[n + 1];
  const double p0 = atan2_exact(static_cast<double>(v0.y), static_cast<double>(v0.x));
  const double p1 = atan2_exact(static_cast<double>(v1.y), static_cast<double>(v1.x));
  return static_cast<float>(unwrap_step(p1 - p0) / kTwoPi);
}

#ifdef AMC_EXPERIMENTS

template <int N, typename CT>
__global__ void __launch_bounds__(FusedCfg<N, CT>::CTA, FusedCfg<N, CT>::MIN_BLOCKS)
fused_features_kernel(const CT* __restrict__ iq, int64_t n_frames, int64_t frame_stride,
                      double* __restrict__ out, int64_t out_stride) {
  using Cfg = FusedCfg<N, CT>;
  constexpr int GROUP = Cfg::GROUP, W = Cfg::W, SPT = Cfg::SPT;
  extern __shared__ __align__(128) unsigned char smem_raw[];

  const int tid = threadIdx.x;
  const int g = tid / GROUP;          // group (frame lane) inside the CTA
  const int t = tid % GROUP;          // thread inside the group
  const int wg = t >> 5;              // warp inside the group
  const int lane = tid & 31;

  unsigned char* gbase = smem_raw + static_cast<size_t>(g) * Cfg::GROUP_BYTES;
  unsigned char* slots = gbase;
  float2* fft_b_extra = reinterpret_cast<float2*>(gbase + Cfg::STAGES * Cfg::SLOT_BYTES);
  unsigned char* part_base = gbase + Cfg::STAGES * Cfg::SLOT_BYTES + Cfg::FFTB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(part_base + 2 * W * Cfg::PART_BYTES);

  const int64_t gg = static_cast<int64_t>(blockIdx.x) * Cfg::G + g;   // global group id
  const int64_t tg = static_cast<int64_t>(gridDim.x) * Cfg::G;        // total groups
  const uint64_t policy = l2_evict_first_policy();

  if (t == 0) {
#pragma unroll
    for (int s = 0; s < Cfg::STAGES; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (t == 0) {
#pragma unroll
    for (int s = 0; s < Cfg::STAGES; ++s) {
      const int64_t f = gg + s * tg;
      if (f < n_frames) {
        mbar_arrive_expect_tx(&bars[s], Cfg::SLOT_BYTES);
        bulk_copy_g2s(slots + s * Cfg::SLOT_BYTES, iq + f * frame_stride, Cfg::SLOT_BYTES, &bars[s], policy);
      }
    }
  }

  int it = 0;
  for (int64_t f = gg; f < n_frames; f += tg, ++it) {
    const int slot = it & 1;
    const uint32_t parity = (it >> 1) & 1;
    unsigned char* slot_ptr = slots + slot * Cfg::SLOT_BYTES;
    const CT* xs = reinterpret_cast<const CT*>(slot_ptr);
    // partial layout per (parity, warp): PART_D doubles, then PART_F floats
    auto part_d = [&](int w) { return reinterpret_cast<double*>(part_base + (slot * W + w) * Cfg::PART_BYTES); };
    auto part_f = [&](int w) {
      return reinterpret_cast<float*>(part_base + (slot * W + w) * Cfg::PART_BYTES + Cfg::PART_D * 8);
    };

    mbar_wait(&bars[slot], parity);

    // ---------------------------------------------------------------- pass 1
    double acc[16];
    Monomials mono;
    mono.clear();
    double sum_r = 0.0;
    double r[SPT];
    float ph[SPT], xr[SPT], xi[SPT];
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      double a, b;
      load_sample<CT>(xs + t + GROUP * j, a, b, xr[j], xi[j]);
      const double s = mono.add(a, b);
      r[j] = sqrt_nr(s);
      sum_r += r[j];
      ph[j] = atan2_fast(xi[j], xr[j]);
    }
    // phase of the sample that follows this warp's 32-sample run, for every j (lane j computes it)
    float ph_edge = 0.0f;
    if (lane < SPT) {
      const int idx = 32 * (wg + 1) + GROUP * lane;
      if (idx < N) {
        double a, b;
        float af, bf;
        load_sample<CT>(xs + idx, a, b, af, bf);
        ph_edge = atan2_fast(bf, af);
      }
    }
    float fq[SPT];
    float s_ph = 0.0f, s_aph = 0.0f, s_f = 0.0f;
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      float nb = __shfl_down_sync(0xffffffffu, ph[j], 1);
      const float nb_edge = __shfl_sync(0xffffffffu, ph_edge, j);
      if (lane == 31) nb = nb_edge;
      const int n = t + GROUP * j;
      float dd = nb - ph[j];
      const float add = fabsf(dd);
      float fj;
      if (fabsf(add - kPiF) < kTieEps && n + 1 < N) {
        fj = exact_freq_step<CT>(xs, n);
      } else {
        if (add > kPiF) dd -= copysignf(kTwoPiF, dd);
        fj = dd * kInvTwoPiF;
      }
      if (n + 1 >= N) fj = 0.0f;
      fq[j] = fj;
      s_f += fj;
      s_ph += ph[j];
      s_aph += fabsf(ph[j]);
    }
#pragma unroll
    for (int i = 0; i < 15; ++i) acc[i] = mono.s[i];
    acc[15] = sum_r;
    warp_sum_multi<double, 16>(acc, lane);
    float accf[4] = {s_ph, s_aph, s_f, 0.0f};
    warp_sum_multi<float, 4>(accf, lane);
    if ((lane & 1) == 0) part_d(wg)[lane >> 1] = acc[0];
    if ((lane & 7) == 0) part_f(wg)[lane >> 3] = accf[0];

    group_sync<GROUP, Cfg::CTA>(g);   // (1) pass-1 partials visible, slot fully read

    double tot_r = 0.0;
    float tot_ph = 0.0f, tot_aph = 0.0f, tot_f = 0.0f;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      tot_r += part_d(w)[15];
      tot_ph += part_f(w)[0];
      tot_aph += part_f(w)[1];
      tot_f += part_f(w)[2];
    }
    const double mu_r = tot_r * (1.0 / N);
    const float mu_ph = tot_ph * (1.0f / N), mu_aph = tot_aph * (1.0f / N);
    const float mu_f = tot_f * (1.0f / (N - 1));

    // ---------------------------------------------------------------- pass 2 (registers only)
    double c2acc[4] = {0.0, 0.0, 0.0, 0.0};   // sum|d|, sum d^2, sum d^4, -
    float q2acc[4] = {0.0f, 0.0f, 0.0f, 0.0f}; // (phi-mu)^2, (|phi|-mu)^2, (f-mu)^2, (f-mu)^4
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      const double d = r[j] - mu_r;
      const double d2 = d * d;
      c2acc[0] += fabs(d);
      c2acc[1] += d2;
      c2acc[2] = fma(d2, d2, c2acc[2]);
      const float e = ph[j] - mu_ph;
      q2acc[0] = fmaf(e, e, q2acc[0]);
      const float ea = fabsf(ph[j]) - mu_aph;
      q2acc[1] = fmaf(ea, ea, q2acc[1]);
      if (t + GROUP * j + 1 < N) {
        const float ef = fq[j] - mu_f;
        const float ef2 = ef * ef;
        q2acc[2] += ef2;
        q2acc[3] = fmaf(ef2, ef2, q2acc[3]);
      }
    }
    warp_sum_multi<double, 4>(c2acc, lane);
    warp_sum_multi<float, 4>(q2acc, lane);
    if ((lane & 7) == 0) {
      part_d(wg)[16 + (lane >> 3)] = c2acc[0];
      part_f(wg)[4 + (lane >> 3)] = q2acc[0];
    }

    // ---------------------------------------------------------------- spectral max: FP32 Stockham FFT
    float2* buf_a = reinterpret_cast<float2*>(slot_ptr);
    float2* buf_b = Cfg::C128 ? reinterpret_cast<float2*>(slot_ptr + N * 8) : fft_b_extra;
    float vmax = 0.0f;
    {
      float2 v[8];
      // stage 1: Ns = 1, radix 8, inputs straight from registers (sample t + (N/8) r)
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = make_float2(xr[q], xi[q]);
      dft8(v);
#pragma unroll
      for (int q = 0; q < 8; ++q) buf_a[swz(8 * t + q)] = v[out8(q)];
      group_sync<GROUP, Cfg::CTA>(g);   // (2)

      // stage 2: Ns = 8, radix 8
      {
        const int k = t & 7;
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = buf_a[swz(t + GROUP * q)];
#pragma unroll
        for (int q = 1; q < 8; ++q) v[q] = c_mul(v[q], g_tw8_s2[(q - 1) * 8 + k]);
        dft8(v);
        const int base = (t >> 3) * 64 + k;
#pragma unroll
        for (int q = 0; q < 8; ++q) buf_b[swz(base + 8 * q)] = v[out8(q)];
      }
      group_sync<GROUP, Cfg::CTA>(g);   // (3)

      if constexpr (N == 256) {
        // stage 3 (last): Ns = 64, radix 4, two butterflies per thread
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) {
          const int jj = t + GROUP * bb;          // 0..63
          float2 u[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) u[q] = buf_b[swz(jj + 64 * q)];
#pragma unroll
          for (int q = 1; q < 4; ++q) u[q] = c_mul(u[q], g_tw8_s3a[(q - 1) * 64 + jj]);
          dft4(u[0], u[1], u[2], u[3]);
#pragma unroll
          for (int q = 0; q < 4; ++q) vmax = fmaxf(vmax, fmaf(u[q].x, u[q].x, u[q].y * u[q].y));
        }
      } else {
        // stage 3: Ns = 64, radix 8
        const int k = t & 63;
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = buf_b[swz(t + GROUP * q)];
#pragma unroll
        for (int q = 1; q < 8; ++q) v[q] = c_mul(v[q], g_tw8_s3b[(q - 1) * 64 + k]);
        dft8(v);
        if constexpr (N == 512) {
#pragma unroll
          for (int q = 0; q < 8; ++q) vmax = fmaxf(vmax, fmaf(v[q].x, v[q].x, v[q].y * v[q].y));
        } else {
          const int base = (t >> 6) * 512 + k;
#pragma unroll
          for (int q = 0; q < 8; ++q) buf_a[swz(base + 64 * q)] = v[out8(q)];
          group_sync<GROUP, Cfg::CTA>(g);   // (4)
          // stage 4 (last): Ns = 512, radix R = N/512, 8/R butterflies per thread
          constexpr int R = N / 512;
#pragma unroll
          for (int bb = 0; bb < 8 / R; ++bb) {
            const int jj = t + GROUP * bb;        // 0..511
            float2 u[R];
#pragma unroll
            for (int q = 0; q < R; ++q) u[q] = buf_a[swz(jj + 512 * q)];
#pragma unroll
            for (int q = 1; q < R; ++q) u[q] = c_mul(u[q], g_tw8_s4[tw8_s4_offset(N) + (q - 1) * 512 + jj]);
            if constexpr (R == 2) {
              bfly2(u[0], u[1]);
            } else if constexpr (R == 4) {
              dft4(u[0], u[1], u[2], u[3]);
            } else {
              float2(&u8)[8] = reinterpret_cast<float2(&)[8]>(u);
              dft8(u8);
            }
#pragma unroll
            for (int q = 0; q < R; ++q) vmax = fmaxf(vmax, fmaf(u[q].x, u[q].x, u[q].y * u[q].y));
          }
        }
      }
    }
    vmax = warp_max(vmax);
    if (lane == 0) part_f(wg)[8] = vmax;

    group_sync<GROUP, Cfg::CTA>(g);   // (5) everything of this frame is in the partial arrays; slot free

    if (t == 0) {
      const int64_t fn = f + Cfg::STAGES * tg;
      if (fn < n_frames) {
        fence_proxy_async_smem();   // our generic-proxy FFT scratch writes precede the async-proxy refill
        mbar_arrive_expect_tx(&bars[slot], Cfg::SLOT_BYTES);
        bulk_copy_g2s(slot_ptr, iq + fn * frame_stride, Cfg::SLOT_BYTES, &bars[slot], policy);
      }
    }
    if (wg == 0) {
      // ------------------------------------------------------------ finalize (warp 0, lanes redundant)
      FrameSums fs;
      double tot[20];
#pragma unroll
      for (int i = 0; i < 20; ++i) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < W; ++w) s += part_d(w)[i];
        tot[i] = s;
      }
      float totf[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < W; ++w) s += part_f(w)[i];
        totf[i] = s;
      }
      float mx = 0.0f;
#pragma unroll
      for (int w = 0; w < W; ++w) mx = fmaxf(mx, part_f(w)[8]);
#pragma unroll
      for (int i = 0; i < 15; ++i) fs.mono[i] = tot[i];
      fs.sum_r = tot[15];
      fs.c_abs1 = tot[16];
      fs.c2 = tot[17];
      fs.c4 = tot[18];
      fs.ph_m2 = static_cast<double>(totf[4]);
      fs.aph_m2 = static_cast<double>(totf[5]);
      fs.f_m2 = static_cast<double>(totf[6]);
      fs.f_m4 = static_cast<double>(totf[7]);
      fs.mean_f = static_cast<double>(totf[2]) / (N - 1);
      fs.spec_max = static_cast<double>(mx);
      double res[18];
      finalize_features(fs, N, res, kCheckAll);
      if (lane < 18) {
        double val = res[0];
#pragma unroll
        for (int i = 1; i < 18; ++i) val = (lane == i) ? res[i] : val;
        out[f * out_stride + lane] = val;
      }
    }
  }
}

#endif  // AMC_EXPERIMENTS

}  // namespace amc
