// N = 2048, FOUR CTAs per SM: the 16-samples-per-thread kernel (amc_fused16.cuh) re-cut so that a frame needs <= 128
// registers per thread and <= 55 KB of shared memory.  Same arithmetic for features 1, 4, 6, 7, 8, 10..18 (bitwise the
// same operations in the same order); features 2, 3, 5, 9 from ONE-PASS float32 sums (see below).
//
// What round 1 measured (profiles/r1p, r2_experiments.txt): the kernel issues on 78 % of its cycles with three warps per
// scheduler; every dependent chain is exposed.  A fourth resident CTA needs:
//   * registers: the per-sample arrays that pass 2 re-read (ph[16], fq[16]) and the FP32 copy of x kept for the FFT
//     (xr/xi[16]) are gone - phase / |phase| / frequency statistics are accumulated in pass 1 as raw power sums
//     (sum phi, sum phi^2, sum t, sum t^2 with t = |phi| - pi/2, sum f .. sum f^4) and centred when the totals are parked,
//     and FFT stage A re-reads its 16 samples from the TMA slot; only r[16] (FP64) survives pass 1;
//   * shared memory: the warp-private exchange buffer is gone - the FP64 transposition reduction uses the warp's own
//     rows of the stage-A buffer (XOR-swizzled, before stage A overwrites them) and FFT stage B writes its outputs back
//     IN PLACE (every lane to the 16 locations it has just read), from where stage C gathers them;
//     frames are finalised in batches of 16.
// MEASURED (round 2, B200): parity-green (117 of the 121 GPU parity / boundary / fuzz tests; the four others compare
// bitwise with the reduced feature profiles of the 3-CTA kernel or link the C example against the product path) and a
// clean 45 s soak - but 0.7106 ms against 0.6277 ms for the 3-CTA kernel: 13 % SLOWER.  Built only with -DAMC_F16_V2
// (tools/exp/build_variants.py v2=AMC_F16_V2); kept as the record of what a fourth CTA costs.
// One-pass centred moments cancel when |mean| is large against the spread: finalisation computes the cancellation
// factor from the same sums and hands the frame to the careful path (amc_device.cuh) when it exceeds 4 - a frame whose
// phases sit in a cluster away from 0, or whose frequency has a mean above 0.4 sigma (a carrier offset of more than
// ~0.1 cycles/sample on noisy data).  Those frames are correct, not fast.
#pragma once
#ifdef AMC_F16_AMP32
// a different experiment behind the same hook of amc_api.cu: the 3-CTA kernel with float32 amplitude statistics
#include "amc_fused16a.cuh"
#else
#include "../amc_fused16.cuh"

namespace amc {

constexpr int kXPart = 28;         // doubles per (parity, warp) partial row == parked values per frame
constexpr int kXPendStride = 29;   // odd: conflict-free lane-per-frame reads

template <typename CT>
struct Fused16xCfg {
  static constexpr int N = 2048, SPT = 16, CTA = 128, W = 4, M1 = 8, LOG_M1 = 3, F = 4;
  static constexpr int SLOT_BYTES = N * static_cast<int>(sizeof(CT));
  static constexpr int FFT_BYTES = N * 8;
  static constexpr int PART_BYTES = kXPart * 8;
  static constexpr int EDGE_BYTES = W * 16 * 4;
  static constexpr int BATCH = 16;
  static constexpr int PEND_BYTES = BATCH * kXPendStride * 8;
  static constexpr int SMEM_BYTES = (SLOT_BYTES + FFT_BYTES + 2 * W * PART_BYTES + EDGE_BYTES + PEND_BYTES + 16 + 127) / 128 * 128;
  static constexpr int MIN_BLOCKS = 4;
  static_assert(4 * (SMEM_BYTES + 1024) <= 227 * 1024, "four CTAs per SM");
};

// Rare: some step of this thread came within kTieEps of +-pi.  Recompute the thread's 16 steps exactly as pass 1 did
// (same float32 function on the same samples -> the same values), re-decide the near-ties in float64 (np.unwrap's
// rules) and return the corrections to the four frequency power sums.
template <typename CT>
__device__ __noinline__ void tie_corrections(const CT* xs, int t, float (&corr)[4]) {
  constexpr int N = 2048, GROUP = 128;
  corr[0] = corr[1] = corr[2] = corr[3] = 0.0f;
  for (int j = 0; j < 16; ++j) {
    const int idx = t + GROUP * j;
    if (idx + 1 >= N) break;
    double a, b;
    float af, bf, cf, df;
    load_sample<CT>(xs + idx, a, b, af, bf);
    load_sample<CT>(xs + idx + 1, a, b, cf, df);
    float dd = atan2_fast(df, cf) - atan2_fast(bf, af);
    if (fabsf(dd) - kPiF > 0.0f) dd -= copysignf(kTwoPiF, dd);
    if (fabsf(kPiF - fabsf(dd)) < 2.0f * kTieEps) {
      const float nw = exact_phase_step<CT>(xs, idx);
      const float o2 = dd * dd, n2 = nw * nw;
      corr[0] += nw - dd;
      corr[1] += n2 - o2;
      corr[2] += n2 * nw - o2 * dd;
      corr[3] += n2 * n2 - o2 * o2;
    }
  }
}

template <typename CT>
__global__ void __launch_bounds__(Fused16xCfg<CT>::CTA, Fused16xCfg<CT>::MIN_BLOCKS)
fused16x_features_kernel(const CT* __restrict__ iq, int64_t n_frames, int64_t frame_stride,
                         double* __restrict__ out, int64_t out_stride, unsigned long long ticket) {
  pdl_launch_dependents();
  using Cfg = Fused16xCfg<CT>;
  constexpr int N = Cfg::N, GROUP = Cfg::CTA, W = Cfg::W, SPT = Cfg::SPT, M1 = Cfg::M1, LOG_M1 = Cfg::LOG_M1;
  extern __shared__ __align__(128) unsigned char smem_raw[];

  const int t = threadIdx.x;
  const int wg = t >> 5;
  const int lane = t & 31;

  unsigned char* gbase = smem_raw;
  const CT* xs = reinterpret_cast<const CT*>(gbase);
  float2* buf_a = reinterpret_cast<float2*>(gbase + Cfg::SLOT_BYTES);
  unsigned char* part_base = gbase + Cfg::SLOT_BYTES + Cfg::FFT_BYTES;
  float* edge_s = reinterpret_cast<float*>(part_base + 2 * W * Cfg::PART_BYTES) + wg * 16;
  double* pend = reinterpret_cast<double*>(part_base + 2 * W * Cfg::PART_BYTES + Cfg::EDGE_BYTES);
  uint64_t* bar = reinterpret_cast<uint64_t*>(part_base + 2 * W * Cfg::PART_BYTES + Cfg::EDGE_BYTES + Cfg::PEND_BYTES);
  uint64_t* rbar = bar + 1;   // "every warp has finished a frame": its reads of buf_a (stages B, C) and of the partials

  const int gg = static_cast<int>(blockIdx.x);
  const int tg = static_cast<int>(gridDim.x);
  const int my_frames = (gg < n_frames) ? static_cast<int>((n_frames - gg + tg - 1) / tg) : 0;

  if (t == 0) {
    mbar_init(bar, 1);
    mbar_init(rbar, W);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait_primary();
  if (lane == 0) mbar_arrive(rbar);   // phase 0 = "nothing to wait for"
  if (t == 0 && my_frames > 0) {
    mbar_arrive_expect_tx(bar, Cfg::SLOT_BYTES);
    bulk_copy_g2s(gbase, iq + static_cast<int64_t>(gg) * frame_stride, Cfg::SLOT_BYTES, bar, l2_evict_first_policy());
  }

  auto part_d = [&](int par, int w) { return reinterpret_cast<double*>(part_base + (par * W + w) * Cfg::PART_BYTES); };

  // Parked values (lane i owns value i): 0..14 monomial sums, 15 sum|x|, 16 sum|r-mu|, 17/18 sum (r-mu)^2/^4,
  // 19 sum phi^2, 20 sum t^2, 21 sum f^2, 22 sum f^4, 23 sum f, 24 max|X|^2, 25 sum phi, 26 sum t, 27 sum f^3
  // (t = |phi| - pi/2; f = unwrapped phase step in radians).
  auto park_and_finalize = [&](int k) {
    const int par = k & 1, bi = k % Cfg::BATCH;
    double* pe = pend + bi * kXPendStride;
    if (lane < kXPart) {
      double s = part_d(par, 0)[lane];
      if (lane == 24) {
#pragma unroll
        for (int w = 1; w < W; ++w) s = fmax(s, part_d(par, w)[lane]);
      } else {
#pragma unroll
        for (int w = 1; w < W; ++w) s += part_d(par, w)[lane];
      }
      pe[lane] = s;
    }
    if (bi == Cfg::BATCH - 1 || k == my_frames - 1) {
      __syncwarp();
      if (lane <= bi) {
        const double* pl = pend + lane * kXPendStride;
        FrameSums fs;
#pragma unroll
        for (int i = 0; i < 15; ++i) fs.mono[i] = pl[i];
        fs.sum_r = pl[15];
        fs.c_abs1 = pl[16];
        fs.c2 = pl[17];
        fs.c4 = pl[18];
        // centre the one-pass sums in float64 (the cancellation happens here, on exact-enough totals)
        constexpr double dn = N, n1 = N - 1;
        constexpr double k1 = 0.15915494309189533577, k2 = k1 * k1;
        const double ph_m2 = pl[19] - pl[25] * pl[25] / dn;
        const double aph_m2 = pl[20] - pl[26] * pl[26] / dn;
        const double mu_f = pl[23] / n1;
        const double f_m2 = pl[21] - pl[23] * mu_f;
        const double f_m4 = pl[22] - 4.0 * mu_f * pl[27] + 6.0 * mu_f * mu_f * pl[21] - 3.0 * n1 * mu_f * mu_f * mu_f * mu_f;
        fs.ph_m2 = ph_m2;
        fs.aph_m2 = aph_m2;
        fs.f_m2 = f_m2 * k2;
        fs.f_m4 = f_m4 * k2 * k2;
        fs.mean_f = mu_f * k1;
        fs.spec_max = pl[24];
        // cancellation factor of the float32 raw sums: raw / centred <= 4, and the frequency mean small against its
        // spread (its fourth moment mixes four raw sums); otherwise the careful path recomputes the frame
        const bool unsafe = !(pl[19] <= 4.0 * ph_m2) || !(pl[20] <= 4.0 * aph_m2) ||
                            !(mu_f * mu_f * n1 <= 0.16 * f_m2);
        const int64_t fo = gg + static_cast<int64_t>(k - bi + lane) * tg;
        double* row = out + fo * out_stride;
        finalize_features(fs, N, row, kCheckAll | (unsafe ? kCheckForce : 0), ticket);
      }
      __syncwarp();
    }
  };

  for (int it = 0; it < my_frames; ++it) {
    const int par = it & 1;

    mbar_wait(bar, static_cast<uint32_t>(par));        // frame `it` has landed in the x slot

    // phase of the sample after this warp's run of 32, for every j: lane j evaluates it, lane 31 uses it
    if (lane < SPT) {
      const int idx = 32 * (wg + 1) + GROUP * lane;
      float pe = 0.0f;
      if (idx < N) {
        double a, b;
        float af, bf;
        load_sample<CT>(xs + idx, a, b, af, bf);
        pe = atan2_fast(bf, af);
      }
      edge_s[lane] = pe;
    }
    __syncwarp();

    // ---------------------------------------------------------------- pass 1 (the only pass over phase / frequency)
    Monomials mono;
    double sum_r;
    double r[SPT];
    float s_ph = 0.0f, s_t = 0.0f, s_p2 = 0.0f, s_t2 = 0.0f;
    float s_f = 0.0f, s_f2 = 0.0f, s_f3 = 0.0f, s_f4 = 0.0f;
    float tie_min = 1.0f;                                      // min | |dd| - pi | over this thread's steps
    const float last_keep = (t == GROUP - 1) ? 0.0f : 1.0f;   // sample N-1 has no successor
    float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      double a, b;
      float af, bf;
      load_sample<CT>(xs + t + GROUP * j, a, b, af, bf);
      const double s = (j == 0) ? mono.init(a, b) : mono.add(a, b);
      r[j] = sqrt_nr(s);
      sum_r = (j == 0) ? r[j] : sum_r + r[j];
      const float ph = atan2_fast(bf, af);
      float nb = __shfl_down_sync(0xffffffffu, ph, 1);
      if ((j & 3) == 0) e4 = reinterpret_cast<const float4*>(edge_s)[j >> 2];   // uniform address: broadcast
      const float ej = (j & 3) == 0 ? e4.x : ((j & 3) == 1 ? e4.y : ((j & 3) == 2 ? e4.z : e4.w));
      if (lane == 31) nb = ej;
      float dd = nb - ph;
      const float over = fabsf(dd) - kPiF;
      tie_min = fminf(tie_min, fabsf(over));
      if (over > 0.0f) dd -= copysignf(kTwoPiF, dd);
      if (j == SPT - 1) dd *= last_keep;
      s_ph += ph;
      const float tt = fabsf(ph) - kPiO2F;
      s_t += tt;
      s_p2 = fmaf(ph, ph, s_p2);
      s_t2 = fmaf(tt, tt, s_t2);
      const float d2 = dd * dd;
      s_f += dd;
      s_f2 += d2;
      s_f3 = fmaf(d2, dd, s_f3);
      s_f4 = fmaf(d2, d2, s_f4);
    }
    if (tie_min < kTieEps) {   // rare (about once per 10^5 samples on noisy data)
      float corr[4];
      tie_corrections<CT>(xs, t, corr);
      s_f += corr[0];
      s_f2 += corr[1];
      s_f3 += corr[2];
      s_f4 += corr[3];
    }

    // every warp has finished the previous frame (its stage-B / stage-C reads of buf_a, warp 0's parking of the
    // partials about to be overwritten); in steady state this completed long ago
    mbar_wait(rbar, static_cast<uint32_t>(par));
    {
      // 16 FP64 partials per lane -> 16 warp totals, transposed through the warp's OWN rows of the stage-A buffer
      // (rows 32 wg .. 32 wg + 31, 128 bytes each, overwritten by this warp's stage A right afterwards):
      // element (lane, i) at row lane, 8-byte slot i ^ (lane & 15)
      const uint32_t red_base = smem_u32(buf_a) + 4096u * wg;
      const uint32_t wrow = red_base + 128u * lane;
      const uint32_t wx = 8u * (lane & 15);
#pragma unroll
      for (int i = 0; i < 15; ++i)
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(wrow + ((8u * i) ^ wx)), "d"(mono.s[i]) : "memory");
      asm volatile("st.shared.f64 [%0], %1;" ::"r"(wrow + ((8u * 15) ^ wx)), "d"(sum_r) : "memory");
      float accf[8] = {s_p2, s_t2, s_f2, s_f4, s_f, s_ph, s_t, s_f3};   // -> parked 19, 20, 21, 22, 23, 25, 26, 27
      warp_sum_multi<float, 8>(accf, lane);
      __syncwarp();
      // reader lane: column c = lane & 15 of rows (lane >> 4) * 16 + i: slot c ^ (row & 15) = c ^ i
      const uint32_t rrow = red_base + 2048u * (lane >> 4);
      const uint32_t rx = 8u * (lane & 15);
      double cs[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(cs[i]) : "r"(rrow + 128u * i + ((8u * i) ^ rx)) : "memory");
#pragma unroll
      for (int w = 8; w >= 1; w >>= 1)                  // pairwise tree: same order as the 3-CTA kernel
#pragma unroll
        for (int i = 0; i < w; ++i) cs[i] += cs[i + w];
      const double tot = cs[0] + __shfl_xor_sync(0xffffffffu, cs[0], 16);
      if (lane < 16) part_d(par, wg)[lane] = tot;
      if ((lane & 3) == 0) {
        const int kk = lane >> 2;                        // 0..7
        part_d(par, wg)[19 + kk + (kk >= 5 ? 1 : 0)] = static_cast<double>(accf[0]);
      }
      __syncwarp();                                     // the reduction's reads are done: stage A may overwrite the rows
    }

    // ---------------------------------------------------------------- FFT stage A: radix 16 over this thread's own
    // samples x[t + GROUP j], re-read from the slot, then the twiddle W_N^(t k1); row t of the block-wide buffer
    float2 v[16];
    {
      const int rot_t = (((t & 15) << (4 - LOG_M1)) | ((t & 15) >> LOG_M1)) & 15;
      const uint32_t row_a = (smem_u32(buf_a) + 128u * t) ^ (8u * rot_t);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        double a, b;
        load_sample<CT>(xs + t + GROUP * q, a, b, v[q].x, v[q].y);
      }
      dft16(v);
      float4 tw[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) tw[p] = g_tw_a4[tw_a4_offset(N) + p * GROUP + t];
#pragma unroll
      for (int q = 1; q < 16; ++q) {
        const float4 w = tw[(q - 1) >> 1];
        v[bitrev4(q)] = c_mul(v[bitrev4(q)], (q & 1) ? make_float2(w.x, w.y) : make_float2(w.z, w.w));
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) {   // element (t, q) -> row t, slot q ^ rot_t
        const float2 o = v[bitrev4(q)];
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(row_a ^ (8u * q)), "f"(o.x), "f"(o.y) : "memory");
      }
    }

    __syncthreads();   // THE barrier: pass-1 partials + stage-A output visible; x slot fully read

    if (t == 0 && it + 1 < my_frames) {                 // refill the x slot: frame it+1 streams in during pass 2 + FFT
      fence_proxy_async_smem();
      mbar_arrive_expect_tx(bar, Cfg::SLOT_BYTES);
      bulk_copy_g2s(gbase, iq + (gg + static_cast<int64_t>(it + 1) * tg) * frame_stride, Cfg::SLOT_BYTES, bar,
                    l2_evict_first_policy());
    }
    if (wg == 0 && it > 0) park_and_finalize(it - 1);   // the previous frame's totals are complete now

    // ---------------------------------------------------------------- pass 2: centred amplitude sums (registers only)
    {
      double tot_r = 0.0;
#pragma unroll
      for (int w = 0; w < W; ++w) tot_r += part_d(par, w)[15];
      const double mu_r = tot_r * (1.0 / N);
      double c2acc[4] = {0.0, 0.0, 0.0, 0.0};            // sum |r-mu|, sum (r-mu)^2, sum (r-mu)^4, -
#pragma unroll
      for (int j = 0; j < SPT; ++j) {
        const double d = r[j] - mu_r;
        const double d2 = d * d;
        c2acc[0] += fabs(d);
        c2acc[1] += d2;
        c2acc[2] = fma(d2, d2, c2acc[2]);
      }
      warp_sum_multi<double, 4>(c2acc, lane);
      if ((lane & 7) == 0 && lane < 24) part_d(par, wg)[16 + (lane >> 3)] = c2acc[0];
    }

    // ---------------------------------------------------------------- FFT stage B: this warp owns the sub-transforms
    // k1 = 4 wg .. 4 wg + 3 (each 128 points, n1 = m1 + 8 m2); lane (f, m1) does the radix-16 over m2, applies
    // W_128^(m1 q) and writes output q back to where input m2 = q came from (only this lane ever touches those 16
    // locations, and only this warp its four k1 columns: no hazard, no buffer)
    float vmax = 0.0f;
    {
      const int m1 = lane & (M1 - 1);
      const int k1b = wg * Cfg::F + (lane >> LOG_M1);
      const int kk_b = k1b ^ ((m1 << (4 - LOG_M1)) & 15);
      float2* cell = buf_a + m1 * 16;
      float4 tw[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) tw[p] = g_tw_b4[tw_b4_offset(N) + p * M1 + m1];
#pragma unroll
      for (int m2 = 0; m2 < 16; ++m2) v[m2] = cell[(kk_b ^ (m2 & 1)) + 16 * M1 * m2];
      dft16(v);
#pragma unroll
      for (int q = 1; q < 16; ++q) {
        const float4 w = tw[(q - 1) >> 1];
        v[bitrev4(q)] = c_mul(v[bitrev4(q)], (q & 1) ? make_float2(w.x, w.y) : make_float2(w.z, w.w));
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) cell[(kk_b ^ (q & 1)) + 16 * M1 * q] = v[bitrev4(q)];
    }
    __syncwarp();
    // ---------------------------------------------------------------- FFT stage C (last): radix 8 over m1 for the
    // pairs (k1, q) of this warp: lane -> f = lane & 3, q = (lane >> 2) + 8 bb; element (k1, q, m1) sits at row
    // m1 + 8 q, slot k1 ^ ((m1 << 1) & 15) ^ (q & 1).  No twiddles; only max |X_k|^2 is kept.
    {
      const int k1c = wg * Cfg::F + (lane & 3);
#pragma unroll 1
      for (int bb = 0; bb < 2; ++bb) {
        const int q = (lane >> 2) + 8 * bb;
        const float2* rd = buf_a + (8 * q) * 16;
        const int sl = k1c ^ (q & 1);
        float2 u[8];
#pragma unroll
        for (int m1 = 0; m1 < 8; ++m1) u[m1] = rd[m1 * 16 + (sl ^ ((m1 << 1) & 15))];
        dft8(u);
#pragma unroll
        for (int m1 = 0; m1 < 8; ++m1) vmax = fmaxf(vmax, fmaf(u[m1].x, u[m1].x, u[m1].y * u[m1].y));
      }
    }
    vmax = warp_max(vmax);
    if (lane == 0) part_d(par, wg)[24] = static_cast<double>(vmax);
    __syncwarp();
    if (lane == 0) mbar_arrive(rbar);   // this warp is done with buf_a and with writing its partials of this frame
  }

  if (my_frames > 0) {
    __syncthreads();                                    // the last frame's partials are visible
    if (wg == 0) park_and_finalize(my_frames - 1);
  }
}

}  // namespace amc
#endif  // AMC_F16_AMP32
