// Fused sm_100a kernel for SHORT frames (N = 256): one WARP per frame, eight frames in flight per CTA.
// Same numerics as the other fused kernels; because a frame never leaves its warp there is no block
// barrier at all (only __syncwarp), the neighbour phase of lane 31 is lane 0's next sample (one
// rotating shuffle, no edge re-evaluation), totals are parked per warp and finalised eight frames
// at a time with one lane per frame.  FFT: radix 8 x 8 x 4 through the frame's own TMA slot.
// The 16 FP64 partials per lane are summed through a warp-private padded transposition buffer
// (16 STS.64 + 16 LDS.64 + 16 DADD instead of a 60-select / 32-shuffle butterfly).
#pragma once
#include "amc_fused.cuh"

namespace amc {

constexpr int kWBatch = 8;
constexpr int kWPendStride = 29;   // 28 parked values per frame (odd stride: conflict-free lane-per-frame reads)
constexpr int kWRow = 17;          // doubles per lane row of the reduction buffer (16 + 1 pad)

template <int N, typename CT>
struct FusedWCfg {
  static_assert(N == 256, "warp-per-frame kernel is instantiated for N = 256");
  static constexpr int SPT = 8;
  static constexpr int CTA = 256, G = 8, MIN_BLOCKS = 2;     // 8 warps = 8 frames in flight; smaller CTAs (1-4 warps,
                                                             // 12-16 warps per SM) were measured 1.5-5 % slower
  static constexpr int SLOT_BYTES = N * static_cast<int>(sizeof(CT));
  static constexpr bool C128 = sizeof(CT) == 16;
  static constexpr int FFTB_BYTES = C128 ? 0 : N * 8;        // c128: both FFT buffers live in the slot
  static constexpr int PEND_BYTES = kWBatch * kWPendStride * 8;
  static constexpr int RED_BYTES = 32 * kWRow * 8;
  static constexpr int GROUP_BYTES = 2 * SLOT_BYTES + FFTB_BYTES + PEND_BYTES + RED_BYTES + 64;
  static constexpr int SMEM_BYTES = G * GROUP_BYTES;
  static_assert(GROUP_BYTES % 16 == 0, "group region must keep 16-byte alignment");
};

template <int N, typename CT, int PROF = kProfAll>
__global__ void __launch_bounds__(FusedWCfg<N, CT>::CTA, FusedWCfg<N, CT>::MIN_BLOCKS)
fusedw_features_kernel(const CT* __restrict__ iq, int64_t n_frames, int64_t frame_stride,
                       double* __restrict__ out, int64_t out_stride, unsigned long long ticket) {
  pdl_launch_dependents();   // the careful-path kernel may be launched now; it waits for this grid to complete
  using Cfg = FusedWCfg<N, CT>;
  constexpr int SPT = Cfg::SPT;
  // feature groups of this instantiation (feature_mask profiles, see amc_device.cuh: kProf*)
  constexpr bool DO_FFT = (PROF & kProfFft) != 0, DO_PHASE = (PROF & kProfPhase) != 0, DO_AMP = (PROF & kProfAmp) != 0,
                 DO_MOM = (PROF & kProfMom) != 0;
  static_assert(DO_MOM, "every compiled profile keeps the monomial sums: finalize_features detects NaN input from them");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int g = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr unsigned FULL = 0xffffffffu;

  unsigned char* gbase = smem_raw + static_cast<size_t>(g) * Cfg::GROUP_BYTES;
  float2* fft_b_extra = reinterpret_cast<float2*>(gbase + 2 * Cfg::SLOT_BYTES);
  double* pend = reinterpret_cast<double*>(gbase + 2 * Cfg::SLOT_BYTES + Cfg::FFTB_BYTES);
  double* red = reinterpret_cast<double*>(gbase + 2 * Cfg::SLOT_BYTES + Cfg::FFTB_BYTES + Cfg::PEND_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(gbase + 2 * Cfg::SLOT_BYTES + Cfg::FFTB_BYTES + Cfg::PEND_BYTES +
                                               Cfg::RED_BYTES);

  // 32-bit group ids, L2 policy created at use, per-lane addresses recomputed from an opaque lane id: the loop
  // sits at the 128-register limit of 2 CTAs x 256 threads and spilled exactly these loop invariants
  const int gg = static_cast<int>(blockIdx.x) * Cfg::G + g;
  const int tg = static_cast<int>(gridDim.x) * Cfg::G;
  const int my_frames = (gg < n_frames) ? static_cast<int>((n_frames - gg + tg - 1) / tg) : 0;

  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait_primary();          // the predecessor in the stream has completed: global memory may be touched
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < 2; ++s)
      if (s < my_frames) {
        mbar_arrive_expect_tx(&bars[s], Cfg::SLOT_BYTES);
        bulk_copy_g2s(gbase + s * Cfg::SLOT_BYTES, iq + (gg + static_cast<int64_t>(s) * tg) * frame_stride,
                      Cfg::SLOT_BYTES, &bars[s], l2_evict_first_policy());
      }
  }

  for (int it = 0; it < my_frames; ++it) {
    const int slot = it & 1;
    unsigned char* slot_ptr = gbase + slot * Cfg::SLOT_BYTES;
    const CT* xs = reinterpret_cast<const CT*>(slot_ptr);
    mbar_wait(&bars[slot], static_cast<uint32_t>((it >> 1) & 1));

    // ---------------------------------------------------------------- pass 1 (sample lane + 32 j)
    Monomials mono;
    double sum_r;
    double r[SPT];
    float ph[SPT], xr[SPT], xi[SPT];
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      double a, b;
      load_sample<CT>(xs + lane + 32 * j, a, b, xr[j], xi[j]);
      double s = 0.0;
      if constexpr (DO_MOM) {
        s = (j == 0) ? mono.init(a, b) : mono.add(a, b);
      } else {
        if (j == 0) mono.clear();
        if constexpr (DO_AMP) s = __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b));   // as Monomials::add forms it
      }
      if constexpr (DO_AMP) {
        r[j] = sqrt_nr(s);
        sum_r = (j == 0) ? r[j] : sum_r + r[j];
      } else {
        r[j] = 0.0;
        sum_r = 0.0;
      }
      if constexpr (DO_PHASE) ph[j] = atan2_fast(xi[j], xr[j]);
      else ph[j] = 0.0f;
    }
    // ONE-PASS phase / frequency statistics (round 2, as in amc_fused16.cuh): raw float32 power sums of phi, of
    // t = |phi| - pi/2 and of the unwrapped phase steps f (radians), centred in float64 at finalisation
    float fq[SPT];
    float s_ph = 0.0f, s_aph = 0.0f, s_p2 = 0.0f, s_t2 = 0.0f;   // s_aph = sum t
    float s_f = 0.0f, s_f2 = 0.0f, s_f3 = 0.0f, s_f4 = 0.0f;
    const float last_keep = (lane == 31) ? 0.0f : 1.0f;       // sample N-1 has no successor
    if constexpr (!DO_PHASE) {
#pragma unroll
      for (int j = 0; j < SPT; ++j) fq[j] = 0.0f;
    } else {
    float tie_max = 0.0f;                                     // max |wrapped step| over this lane's steps
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      // successor of (lane, j) is (lane+1, j); for lane 31 it is (0, j+1): lane 0 offers its next sample
      const float offer = (lane == 0) ? ph[(j + 1) % SPT] : ph[j];
      const float nb = __shfl_sync(FULL, offer, (lane + 1) & 31);
      float dd = wrap_step_f32(nb - ph[j]);
      tie_max = fmaxf(tie_max, fabsf(dd));
      if (j == SPT - 1) dd *= last_keep;
      fq[j] = dd;
      s_ph += ph[j];
      const float tt = fabsf(ph[j]) - kPiO2F;
      s_aph += tt;
      s_p2 = fmaf(ph[j], ph[j], s_p2);
      s_t2 = fmaf(tt, tt, s_t2);
    }
    if (tie_max > kPiF - kTieEps) {   // rare (about once per 10^5 samples on noisy data): FP64 re-decision, see amc_fused16.cuh
      unsigned tie_mask = 0u;
#pragma unroll
      for (int j = 0; j < SPT; ++j)
        if (fabsf(kPiF - fabsf(fq[j])) < 2.0f * kTieEps) tie_mask |= 1u << j;
      while (tie_mask != 0u) {
        const int j = __ffs(tie_mask) - 1;
        tie_mask &= tie_mask - 1u;
        const float val = exact_phase_step<CT>(xs, lane + 32 * j);
#pragma unroll
        for (int q = 0; q < SPT; ++q) fq[q] = (q == j) ? val : fq[q];
      }
    }
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      const float f2 = fq[j] * fq[j];
      s_f += fq[j];
      s_f2 += f2;
      s_f3 = fmaf(f2, fq[j], s_f3);
      s_f4 = fmaf(f2, f2, s_f4);
    }
    }   // DO_PHASE

    // 16 FP64 partials per lane -> lane l (and l + 16) holds the warp total of value l
    __syncwarp();                                             // the previous frame's column reads are done
        const int lv = opaque_if<true>(lane);               // (without it: 32 B of spills at the 128-register cap, -2.5 %)
#pragma unroll
    for (int i = 0; i < 15; ++i) red[lv * kWRow + i] = mono.s[i];
    red[lv * kWRow + 15] = sum_r;
    float accf[8] = {s_p2, s_t2, s_f2, s_f4, s_f, s_ph, s_aph, s_f3};   // -> parked 19, 20, 21, 22, 23, 25, 26, 27
    warp_sum_multi<float, 8>(accf, lane);                     // lane l: total of value l >> 2
    __syncwarp();
    double tot16;
    {
      const double* col = red + (lv >> 4) * (16 * kWRow) + (lv & 15);
      double cs[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) cs[i] = col[i * kWRow];
#pragma unroll
      for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
        for (int i = 0; i < w; ++i) cs[i] += cs[i + w];
      tot16 = cs[0] + __shfl_xor_sync(FULL, cs[0], 16);
    }
    const double mu_r = __shfl_sync(FULL, tot16, 15) * (1.0 / N);

    // ---------------------------------------------------------------- pass 2 (registers only)
    double c2acc[4] = {0.0, 0.0, 0.0, 0.0};                   // sum |r-mu|, sum (r-mu)^2, sum (r-mu)^4, -
    if constexpr (DO_AMP) {
#pragma unroll
      for (int j = 0; j < SPT; ++j) {
        const double d = r[j] - mu_r;
        const double d2 = d * d;
        c2acc[0] += fabs(d);
        c2acc[1] += d2;
        c2acc[2] = fma(d2, d2, c2acc[2]);
      }
      warp_sum_multi<double, 4>(c2acc, lane);                 // lane l: value l >> 3
    }

    // ---------------------------------------------------------------- FFT 8 x 8 x 4 through the slot
    __syncwarp();                                             // every lane has finished reading x
    float2* buf_a = reinterpret_cast<float2*>(slot_ptr);
    float2* buf_b = Cfg::C128 ? reinterpret_cast<float2*>(slot_ptr + N * 8) : fft_b_extra;
    float vmax = 0.0f;
    if constexpr (DO_FFT) {
      // (the ~30 swizzled exchange indices are recomputed per frame from an opaque lane id: hoisted out of the
      // loop they were spilled and reloaded with LDL in front of every use)
      const int lf = opaque_if<true>(lane);
      float2 v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = make_float2(xr[q], xi[q]);
      dft8(v);
#pragma unroll
      for (int q = 0; q < 8; ++q) buf_a[swz(8 * lf + q)] = v[out8(q)];
      __syncwarp();
      const int k = lf & 7;
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = buf_a[swz(lf + 32 * q)];
#pragma unroll
      for (int q = 1; q < 8; ++q) v[q] = c_mul(v[q], g_tw8_s2[(q - 1) * 8 + k]);
      dft8(v);
      const int base = (lf >> 3) * 64 + k;
#pragma unroll
      for (int q = 0; q < 8; ++q) buf_b[swz(base + 8 * q)] = v[out8(q)];
      __syncwarp();
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        const int jj = lf + 32 * bb;                          // 0..63
        float2 u[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) u[q] = buf_b[swz(jj + 64 * q)];
#pragma unroll
        for (int q = 1; q < 4; ++q) u[q] = c_mul(u[q], g_tw8_s3a[(q - 1) * 64 + jj]);
        dft4(u[0], u[1], u[2], u[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) vmax = fmaxf(vmax, fmaf(u[q].x, u[q].x, u[q].y * u[q].y));
      }
    }
    vmax = warp_max(vmax);
    __syncwarp();                                             // slot (FFT scratch) no longer needed

    if (lane == 0 && it + 2 < my_frames) {
      fence_proxy_async_smem();
      mbar_arrive_expect_tx(&bars[slot], Cfg::SLOT_BYTES);
      bulk_copy_g2s(slot_ptr, iq + (gg + static_cast<int64_t>(it + 2) * tg) * frame_stride, Cfg::SLOT_BYTES,
                    &bars[slot], l2_evict_first_policy());
    }

    // ---------------------------------------------------------------- park this frame's 28 totals: 0..14 monomial
    // sums, 15 sum|x|, 16..18 centred amplitude sums, 19 sum phi^2, 20 sum t^2, 21 sum f^2, 22 sum f^4, 23 sum f, 24 max,
    // 25 sum phi, 26 sum t, 27 sum f^3
    const int bi = it % kWBatch;
    double* pe = pend + bi * kWPendStride;
    if (lane < 16) pe[lane] = tot16;                                   // 0..15
    if ((lane & 7) == 0 && lane < 24) pe[16 + (lane >> 3)] = c2acc[0];
    if ((lane & 3) == 0) {
      const int kk = lane >> 2;                                        // 0..7
      pe[19 + kk + (kk >= 5 ? 1 : 0)] = static_cast<double>(accf[0]);
    }
    if (lane == 3) pe[24] = static_cast<double>(vmax);
    if (bi == kWBatch - 1 || it == my_frames - 1) {
      __syncwarp();
      if (lane <= bi) {
        const double* pl = pend + lane * kWPendStride;
        FrameSums fs;
#pragma unroll
        for (int i = 0; i < 15; ++i) fs.mono[i] = pl[i];
        fs.sum_r = pl[15];
        fs.c_abs1 = pl[16];
        fs.c2 = pl[17];
        fs.c4 = pl[18];
        constexpr double dn = N, n1 = N - 1;
        constexpr double k1 = 0.15915494309189533577, k2 = k1 * k1;
        const double ph_m2 = pl[19] - pl[25] * pl[25] / dn;
        const double aph_m2 = pl[20] - pl[26] * pl[26] / dn;
        const double mu_f = pl[23] / n1;
        const double f_m2 = pl[21] - pl[23] * mu_f;
        const double f_m4 = pl[22] - 4.0 * mu_f * pl[27] + 6.0 * mu_f * mu_f * pl[21] - 3.0 * n1 * mu_f * mu_f * mu_f * mu_f;
        fs.ph_m2 = ph_m2;
        fs.aph_m2 = aph_m2;
        fs.f_m2 = f_m2 * k2;                                           // radians -> cycles
        fs.f_m4 = f_m4 * k2 * k2;
        fs.mean_f = mu_f * k1;
        fs.spec_max = pl[24];
        // cancellation factor of the float32 raw sums <= 4, frequency mean small against its spread; otherwise the
        // careful path recomputes the frame
        const bool unsafe = DO_PHASE && (!(pl[19] <= 4.0 * ph_m2) || !(pl[20] <= 4.0 * aph_m2) ||
                                         !(mu_f * mu_f * n1 <= 0.16 * f_m2));
        const int64_t fo = gg + static_cast<int64_t>(it - bi + lane) * tg;
        constexpr int kChecks = ((DO_FFT || DO_PHASE) ? kCheckRange : 0) | (DO_PHASE ? kCheckPhase : 0) |
                                (DO_AMP ? kCheckAmp : 0);
        if (!finalize_features(fs, N, out + fo * out_stride, kChecks | (unsafe ? kCheckForce : 0), ticket))
          blank_skipped_groups<PROF>(out + fo * out_stride);
      }
      __syncwarp();
    }
  }
}

}  // namespace amc
