// Long frames (N = 8192, 16384): one 512-thread CTA per frame, persistent over frames.
// The frame does not fit the "registers hold the whole frame" scheme of the fused kernels, so:
//   pass 1 streams x from HBM (coalesced 16-byte loads): FP64 monomials + |x| (stashed in an L2-resident scratch),
//          FP32 atan2 and the raw power sums of phi and |phi| - pi/2; the FP32 copy of x goes to the FFT buffer and
//          the phase to a phase buffer, both in shared memory;
//   pass 1b forms the wrapped phase differences from the phase buffer (FP64 re-decision of ties) and their raw power
//          sums (sum f .. sum f^4);
//   pass 2 reads |x| back for the centred amplitude sums (the only statistics that need the mean first);
//   FFT    in-place Stockham, radix 16 x 16 x 16 x (N/4096), XOR-swizzled, lane-contiguous twiddle tables.
// Phase / frequency statistics are ONE-PASS (round 2: raw float32 sums centred in float64 at finalisation; a frame whose
// cancellation factor exceeds 4, or whose frequency mean exceeds 0.4 sigma, goes to the careful path) - the two-pass
// form cost a second trip through the phase buffer and a second evaluation of every wrapped difference:
// N = 8192 / 16384: 27.8 / 26.6 % -> 29.1 / 28.5 % of the measured HBM peak; with the |x| stash, its early read-back, the L2
// prefetch of the next frame, cache hints and the deeper unroll of the last FFT stage: 34.2 / 31.8 % (profiles/r2_experiments.txt).
// Same tolerance classes as the fused kernels.
#pragma once
#include "amc_fused16.cuh"

namespace amc {


// g_tw_l4[off(N) + q-1][j] = W_N^(j q), j < 4096, q = 1..N/4096-1
__device__ float2 g_tw_l4[4 * 4096];
__host__ __device__ constexpr int tw_l4_offset(int n) { return (n == 8192 ? 0 : 1) * 4096; }
__global__ void init_twiddle_large_kernel() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4 * 4096) {
    const int row = i / 4096, j = i % 4096;
    const int n = row < 1 ? 8192 : 16384;
    const int q = row < 1 ? 1 : row;
    g_tw_l4[i] = tw_exact(j * q, n);
  }
}

template <int N>
struct LargeCfg {
  static_assert(N == 8192 || N == 16384, "long-frame kernel sizes");
  static constexpr int THREADS = N == 8192 ? 256 : 512;            // 8192: two CTAs per SM (one 512-thread CTA
  static constexpr int WARPS = THREADS / 32;                       // was measured 25 % slower); 16384: one
  static constexpr int MIN_BLOCKS = N == 8192 ? 2 : 1;
  static constexpr int FFT_BYTES = N * 8;
  static constexpr int PHI_BYTES = N * 4;
  static constexpr int PART_BYTES = 2 * WARPS * 32 * 8;            // two parities x warps x 32 doubles
  static constexpr int BATCH = 32;                                 // frames finalised together, one lane each
  static constexpr int PEND_STRIDE = 29;                           // 28 parked values per frame (odd stride: conflict-free)
  static constexpr int PEND_BYTES = BATCH * PEND_STRIDE * 8;
  static constexpr int SMEM_BYTES = FFT_BYTES + PHI_BYTES + PART_BYTES + PEND_BYTES + 64;
  static constexpr int R4 = N / 4096;                              // radix of the last stage
  static constexpr int BPT = (N / 16) / THREADS;                   // radix-16 butterflies per thread: 2
};

// value already in registers -> (a, b) in FP64 and FP32
__device__ __forceinline__ void split_sample(double2 v, double& a, double& b, float& af, float& bf) {
  a = v.x;
  b = v.y;
  af = static_cast<float>(a);
  bf = static_cast<float>(b);
}
__device__ __forceinline__ void split_sample(float2 v, double& a, double& b, float& af, float& bf) {
  af = v.x;
  bf = v.y;
  a = static_cast<double>(af);
  b = static_cast<double>(bf);
}
// x is read once (streaming loads, evict-first) and the |x| stash lives in L2 (.cg): neither may push the FFT twiddle
// tables out of the little L1 that is left beside the shared memory (+1 % at both sizes)
#define AMC_LG_LDX(p) __ldcs(p)
#define AMC_LG_STR(p, v) __stcg(p, v)
#define AMC_LG_LDR(p) __ldcg(p)
constexpr int kLargeU = 4;   // samples per thread per software-pipelined group (loads of group g+1 fly during group g)

// one in-place radix-16 Stockham stage over the whole frame (BPT butterflies per thread).
// swz16(e) = e ^ ((e >> 4) & 15); for the three stage shapes the swizzle term is known in closed form, so
// every exchange address is base + immediate or one XOR of the byte address (the generic form cost
// three integer instructions per element):
//   reads  e = j + (N/16) q            : (e>>4)&15 = (j>>4)&15                 -> (j ^ ((j>>4)&15)) + (N/16) q
//   writes NS = 1   e = 16 j + q       : (e>>4)&15 = j & 15                    -> 16 j + (q ^ (j & 15))
//          NS = 16  e = 256 (j/16) + 16 q + k : (e>>4)&15 = q                  -> 256 (j/16) + 16 q + (k ^ q)
//          NS = 256 e = 4096 (j/256) + 256 q + k : (e>>4)&15 = (k>>4)&15       -> ... + 256 q + (k ^ ((k>>4)&15))
template <int N, int NS, int BPT, int THREADS>
__device__ __forceinline__ void large_stage16(float2* __restrict__ buf, const float2* __restrict__ tw, int tid) {
  static_assert((N / 256) % 16 == 0, "closed-form swizzle needs (N/16)/16 to be a multiple of 16");
  float2 v[BPT][16];
#pragma unroll
  for (int bb = 0; bb < BPT; ++bb) {
    const int j = tid + THREADS * bb;
    const float2* src = buf + (j ^ ((j >> 4) & 15));
#pragma unroll
    for (int q = 0; q < 16; ++q) v[bb][q] = src[(N / 16) * q];
  }
  // the first butterfly's twiddles are fetched BEFORE the barrier (L2 latency hides behind it)
  float2 tw0[15];
  if constexpr (NS > 1) {
#pragma unroll
    for (int q = 1; q < 16; ++q) tw0[q - 1] = tw[(q - 1) * NS + (tid % NS)];
  }
  __syncthreads();                                                 // all reads of this stage are done
  const uint32_t buf_s = smem_u32(buf);                            // 128-byte aligned
#pragma unroll
  for (int bb = 0; bb < BPT; ++bb) {
    const int j = tid + THREADS * bb;
    const int k = j % NS;
    if constexpr (NS > 1) {
#pragma unroll
      for (int q = 1; q < 16; ++q) v[bb][q] = c_mul(v[bb][q], bb == 0 ? tw0[q - 1] : tw[(q - 1) * NS + k]);
    }
    dft16(v[bb]);
    if constexpr (NS == 1) {
      const uint32_t a0 = (buf_s + 128u * j) ^ (8u * (j & 15));
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float2 o = v[bb][bitrev4(q)];
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a0 ^ (8u * q)), "f"(o.x), "f"(o.y) : "memory");
      }
    } else if constexpr (NS == 16) {
      const uint32_t a0 = buf_s + 2048u * (j >> 4) + 8u * k;         // bits 3..6 hold k
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float2 o = v[bb][bitrev4(q)];
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"((a0 ^ (8u * q)) + 128u * q), "f"(o.x), "f"(o.y)
                     : "memory");
      }
    } else {
      static_assert(NS == 256, "stage shapes: 1, 16, 256");
      float2* dst = buf + (j >> 8) * 4096 + (k ^ ((k >> 4) & 15));
#pragma unroll
      for (int q = 0; q < 16; ++q) dst[256 * q] = v[bb][bitrev4(q)];
    }
  }
  __syncthreads();
}

// STASH: |x| of every sample is written to this CTA's slice of a global scratch (`r_ws`, N doubles per CTA: 19 MB for
// the whole grid, L2-resident) in pass 1 and read back in pass 2, instead of re-reading x and recomputing the square
// root there (5 FP64 operations + 4 integer / MUFU per sample).
template <int N, typename CT, bool STASH>
__global__ void __launch_bounds__(LargeCfg<N>::THREADS, LargeCfg<N>::MIN_BLOCKS)
large_features_kernel(const CT* __restrict__ iq, int64_t n_frames, int64_t frame_stride,
                      double* __restrict__ out, int64_t out_stride, unsigned long long ticket, double* __restrict__ r_ws) {
  pdl_launch_dependents();   // the careful-path kernel may be launched now; it waits for this grid to complete
  using Cfg = LargeCfg<N>;
  constexpr int THREADS = Cfg::THREADS, WARPS = Cfg::WARPS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* buf = reinterpret_cast<float2*>(smem_raw);
  float* phi = reinterpret_cast<float*>(smem_raw + Cfg::FFT_BYTES);
  double* part = reinterpret_cast<double*>(smem_raw + Cfg::FFT_BYTES + Cfg::PHI_BYTES);
  double* pend = reinterpret_cast<double*>(smem_raw + Cfg::FFT_BYTES + Cfg::PHI_BYTES + Cfg::PART_BYTES);
  const int my_frames = blockIdx.x < n_frames ? static_cast<int>((n_frames - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  pdl_wait_primary();          // the predecessor in the stream has completed: global memory may be touched
  int it = 0;
  for (int64_t f = blockIdx.x; f < n_frames; f += gridDim.x, ++it) {
    const CT* x = iq + f * frame_stride;
    double* pw = part + ((it & 1) * WARPS + warp) * 32;       // this warp's 32 partial slots
    double* pall = part + (it & 1) * WARPS * 32;

    // ---------------------------------------------------------------- pass 1
    Monomials mono;
    mono.clear();
    double sum_r = 0.0;
    float s_ph = 0.0f, s_aph = 0.0f;
    // one-pass phase statistics (round 2): raw power sums of phi and of t = |phi| - pi/2 (s_aph holds sum t), centred in
    // float64 when the frame is finalised; the careful path takes frames whose cancellation factor exceeds 4
    float s_p2 = 0.0f, s_t2 = 0.0f;
    // software-pipelined with two register buffers (ping-pong, no loop-carried moves): the 16-byte loads of
    // the next group are issued before the current group is processed (ncu: 32 % of the stall samples
    // were long-scoreboard waits on these loads)
    {
      auto load_group = [&](CT (&g)[kLargeU], int i0) {
#pragma unroll
        for (int u = 0; u < kLargeU; ++u) g[u] = AMC_LG_LDX(x + i0 + THREADS * u);
      };
      float2* buf_t = buf + (tid ^ ((tid >> 4) & 15));      // swizzled position of sample tid (+ multiples of THREADS)
      [[maybe_unused]] double* r_mine = STASH ? r_ws + static_cast<size_t>(blockIdx.x) * N : nullptr;
      auto pass1_group = [&](const CT (&g)[kLargeU], int i0) {
#pragma unroll
        for (int u = 0; u < kLargeU; ++u) {
          const int i = i0 + THREADS * u;
          double a, b;
          float af, bf;
          split_sample(g[u], a, b, af, bf);
          const double s = mono.add(a, b);
          const double rr = sqrt_nr(s);
          sum_r += rr;
          if constexpr (STASH) AMC_LG_STR(r_mine + i, rr);
          const float p = atan2_fast(bf, af);
          buf_t[i - tid] = make_float2(af, bf);           // = buf[swz16(i)]: (i >> 4) & 15 == (tid >> 4) & 15
          phi[i] = p;
          s_ph += p;
          const float tt = fabsf(p) - kPiO2F;
          s_aph += tt;
          s_p2 = fmaf(p, p, s_p2);
          s_t2 = fmaf(tt, tt, s_t2);
        }
      };
      constexpr int STEP = THREADS * kLargeU;               // samples per group over the whole CTA
      static_assert(N % (2 * STEP) == 0, "ping-pong loop needs an even number of groups");
      static_assert(THREADS % 256 == 0, "closed-form swizzle of the FP32 copy needs THREADS/16 to be a multiple of 16");
      CT ga[kLargeU], gb[kLargeU];
      load_group(ga, tid);
#pragma unroll 1   // (unrolled 2 / 4 it loses the ~30 loop-carried register moves per trip but spills: -1.4 / -1.5 %)
      for (int k = 0; k < N / (2 * STEP); ++k) {
        const int i0 = tid + 2 * STEP * k;
        load_group(gb, i0 + STEP);
        pass1_group(ga, i0);
        if (k + 1 < N / (2 * STEP)) load_group(ga, i0 + 2 * STEP);
        pass1_group(gb, i0 + STEP);
      }
    }
    // the NEXT frame of this CTA starts its trip from HBM to L2 now: pass 1b, pass 2 and the FFT (half of the frame time)
    // leave the memory system idle, and pass 1 of the next frame then waits for L2 instead of HBM latency
    if (f + gridDim.x < n_frames) {
      const char* nx = reinterpret_cast<const char*>(iq + (f + gridDim.x) * frame_stride);
      constexpr int kBytes = N * static_cast<int>(sizeof(CT));
#pragma unroll
      for (int o = 128 * tid; o < kBytes + 128; o += 128 * THREADS)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + min(o, kBytes - 1)));
    }
    __syncthreads();
    // ---------------------------------------------------------------- pass 1b: sum of wrapped differences
    float s_f = 0.0f;
    float s_f2 = 0.0f, s_f3 = 0.0f, s_f4 = 0.0f;            // one-pass frequency statistics (cycles per sample)
#pragma unroll 4
    for (int i = tid; i < N - 1; i += THREADS) {
      const float dd = wrap_step_f32(phi[i + 1] - phi[i]);
      float fj = dd * kInvTwoPiF;
      if (fabsf(dd) > kPiF - kTieEps) fj = exact_freq_step<CT>(x, i);   // rare: re-decided in float64
      s_f += fj;
      const float f2 = fj * fj;
      s_f2 += f2;
      s_f3 = fmaf(f2, fj, s_f3);
      s_f4 = fmaf(f2, f2, s_f4);
    }
    // pass 2 reads back what THIS thread stashed (same index set): the first half of the loads is issued here, before the
    // reductions and the barrier, the second half when the first is consumed - one exposed L2 latency per frame, not four
    constexpr int RPT = N / THREADS;                          // |x| values per thread: 32
    [[maybe_unused]] double rv0[RPT / 2];
    if constexpr (STASH) {
      const double* r_mine = r_ws + static_cast<size_t>(blockIdx.x) * N + tid;
#pragma unroll
      for (int u = 0; u < RPT / 2; ++u) rv0[u] = AMC_LG_LDR(r_mine + THREADS * u);
    }
    {
      double acc[16];
#pragma unroll
      for (int i = 0; i < 15; ++i) acc[i] = mono.s[i];
      acc[15] = sum_r;
      warp_sum_multi<double, 16>(acc, lane);
      if ((lane & 1) == 0) pw[lane >> 1] = acc[0];                          // 0..15
      float accf[8] = {s_p2, s_t2, s_f2, s_f4, s_f, s_ph, s_aph, s_f3};
      warp_sum_multi<float, 8>(accf, lane);
      if ((lane & 3) == 0) pw[16 + (lane >> 2)] = static_cast<double>(accf[0]);   // 16..23
    }
    __syncthreads();
    double tot_r = 0.0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) tot_r += pall[w * 32 + 15];
    const double mu_r = tot_r * (1.0 / N);

    // ---------------------------------------------------------------- pass 2: centred sums
    double c2acc[4] = {0.0, 0.0, 0.0, 0.0};
    if constexpr (STASH) {                                     // |x| comes back from the L2-resident scratch
      const double* r_mine = r_ws + static_cast<size_t>(blockIdx.x) * N + tid;
      double rv1[RPT / 2];
#pragma unroll
      for (int u = 0; u < RPT / 2; ++u) rv1[u] = AMC_LG_LDR(r_mine + THREADS * (RPT / 2 + u));
#pragma unroll
      for (int u = 0; u < RPT; ++u) {
        const double d = (u < RPT / 2 ? rv0[u % (RPT / 2)] : rv1[u % (RPT / 2)]) - mu_r;
        const double d2 = d * d;
        c2acc[0] += fabs(d);
        c2acc[1] += d2;
        c2acc[2] = fma(d2, d2, c2acc[2]);
      }
    } else {                                                   // no scratch: re-read x (L2-resident), same software pipeline
      auto load_group = [&](CT (&g)[kLargeU], int i0) {
#pragma unroll
        for (int u = 0; u < kLargeU; ++u) g[u] = AMC_LG_LDX(x + i0 + THREADS * u);
      };
      auto pass2_group = [&](const CT (&g)[kLargeU]) {
#pragma unroll
        for (int u = 0; u < kLargeU; ++u) {
          const double a = static_cast<double>(g[u].x), b = static_cast<double>(g[u].y);
          // |x|^2 formed exactly as Monomials::add forms it in pass 1: the same bits as the stashed values
          const double d = sqrt_nr(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b))) - mu_r;
          const double d2 = d * d;
          c2acc[0] += fabs(d);
          c2acc[1] += d2;
          c2acc[2] = fma(d2, d2, c2acc[2]);
        }
      };
      constexpr int STEP = THREADS * kLargeU;
      CT ga[kLargeU], gb[kLargeU];
      load_group(ga, tid);
#pragma unroll 1
      for (int i0 = tid; i0 < N; i0 += 2 * STEP) {
        load_group(gb, i0 + STEP);
        pass2_group(ga);
        if (i0 + 2 * STEP < N) load_group(ga, i0 + 2 * STEP);
        pass2_group(gb);
      }
    }
    warp_sum_multi<double, 4>(c2acc, lane);
    if ((lane & 7) == 0) pw[24 + (lane >> 3)] = c2acc[0];                // 24..27 (27 unused)

    // ---------------------------------------------------------------- FFT: 16 x 16 x 16 x R4, in place
    large_stage16<N, 1, Cfg::BPT, THREADS>(buf, nullptr, tid);
    large_stage16<N, 16, Cfg::BPT, THREADS>(buf, g_tw_s2, tid);
    large_stage16<N, 256, Cfg::BPT, THREADS>(buf, g_tw_s3 + tw_s3_offset(4096), tid);
    float vmax = 0.0f;
    {
      constexpr int R4 = Cfg::R4;                                        // last stage: N/R4 = 4096 butterflies
#pragma unroll (Cfg::R4 == 2 ? 8 : 4)
      // (the twiddles come from L2 - the table does not fit beside 212 KB of shared memory; 2 -> 4: +1.9 % at N = 8192)
      for (int bb = 0; bb < 4096 / THREADS; ++bb) {
        const int j = tid + THREADS * bb;
        float2 u[R4];
#pragma unroll
        for (int q = 0; q < R4; ++q) u[q] = buf[(j ^ ((j >> 4) & 15)) + 4096 * q];   // = swz16(j + 4096 q)
#pragma unroll
        for (int q = 1; q < R4; ++q) u[q] = c_mul(u[q], g_tw_l4[tw_l4_offset(N) + (q - 1) * 4096 + j]);
        if constexpr (R4 == 2) {
          bfly2(u[0], u[1]);
        } else {
          dft4(u[0], u[1], u[2], u[3]);
        }
#pragma unroll
        for (int q = 0; q < R4; ++q) vmax = fmaxf(vmax, fmaf(u[q].x, u[q].x, u[q].y * u[q].y));
      }
    }
    vmax = warp_max(vmax);
    if (lane == 0) pw[28] = static_cast<double>(vmax);
    __syncthreads();                                                     // totals complete; buf/phi free again

    if (warp == 0) {                                                     // other warps start the next frame
      // park this frame's 25 totals (lane i sums value i over the warps); every BATCH frames the warp
      // finalises BATCH frames at once, one lane per frame - a single lane finalising every frame kept the
      // other 15 warps waiting at the next frame's first barrier for 17 % of their time (latency-bound FP64 chain)
      const int bi = it % Cfg::BATCH;
      double* pe = pend + bi * Cfg::PEND_STRIDE;
      // parked values: 0..15 FP64 sums, 16..18 centred amplitude sums, 19 sum phi^2, 20 sum t^2, 21 sum f^2, 22 sum f^4,
      // 23 sum f, 24 max, 25 sum phi, 26 sum t, 27 sum f^3; partial rows: 0..15 | 16..23 float sums | 24..26 | 28 (max)
      constexpr int kParked = 28;
      if (lane < kParked) {
        const int src = lane < 16 ? lane : (lane < 19 ? lane + 8 : (lane < 24 ? lane - 3 : (lane == 24 ? 28 : lane - 4)));
        double v = pall[src];
        if (lane == 24) {
#pragma unroll
          for (int w = 1; w < WARPS; ++w) v = fmax(v, pall[w * 32 + src]);
        } else {
#pragma unroll
          for (int w = 1; w < WARPS; ++w) v += pall[w * 32 + src];
        }
        pe[lane] = v;
      }
      if (bi == Cfg::BATCH - 1 || it == my_frames - 1) {
        __syncwarp();
        if (lane <= bi) {
          const double* pl = pend + lane * Cfg::PEND_STRIDE;
          FrameSums fs;
#pragma unroll
          for (int i = 0; i < 15; ++i) fs.mono[i] = pl[i];
          fs.sum_r = pl[15];
          fs.c_abs1 = pl[16];
          fs.c2 = pl[17];
          fs.c4 = pl[18];
          fs.spec_max = pl[24];
          int checks = kCheckAll;
          {   // centre the one-pass sums in float64 (frequency in cycles per sample)
            constexpr double dn = N, n1 = N - 1;
            const double mu_f = pl[23] / n1;
            fs.ph_m2 = pl[19] - pl[25] * pl[25] / dn;
            fs.aph_m2 = pl[20] - pl[26] * pl[26] / dn;
            fs.f_m2 = pl[21] - pl[23] * mu_f;
            fs.f_m4 = pl[22] - 4.0 * mu_f * pl[27] + 6.0 * mu_f * mu_f * pl[21] - 3.0 * n1 * mu_f * mu_f * mu_f * mu_f;
            fs.mean_f = mu_f;
            // cancellation factor of the float32 raw sums <= 4, frequency mean small against its spread
            if (!(pl[19] <= 4.0 * fs.ph_m2) || !(pl[20] <= 4.0 * fs.aph_m2) || !(mu_f * mu_f * n1 <= 0.16 * fs.f_m2))
              checks |= kCheckForce;
          }
          const int64_t fo = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(it - bi + lane) * gridDim.x;
          finalize_features(fs, N, out + fo * out_stride, checks, ticket);
        }
        __syncwarp();
      }
    }
  }
}

}  // namespace amc
