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
constexpr int kWPendStride = 25;   // doubles per parked frame (odd: conflict-free lane-per-frame reads)
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
    float fq[SPT];                                            // unwrapped phase steps in RADIANS (scaled when parked)
    float s_ph = 0.0f, s_aph = 0.0f, s_f = 0.0f;
    const float last_keep = (lane == 31) ? 0.0f : 1.0f;       // sample N-1 has no successor
    if constexpr (!DO_PHASE) {
#pragma unroll
      for (int j = 0; j < SPT; ++j) fq[j] = 0.0f;
    } else {
    float tie_min = 1.0f;                                     // min | |dd| - pi | over this lane's steps
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      // successor of (lane, j) is (lane+1, j); for lane 31 it is (0, j+1): lane 0 offers its next sample
      const float offer = (lane == 0) ? ph[(j + 1) % SPT] : ph[j];
      const float nb = __shfl_sync(FULL, offer, (lane + 1) & 31);
      float dd = nb - ph[j];
      const float over = fabsf(dd) - kPiF;
      tie_min = fminf(tie_min, fabsf(over));
      if (over > 0.0f) dd -= copysignf(kTwoPiF, dd);
      if (j == SPT - 1) dd *= last_keep;
      fq[j] = dd;
      s_ph += ph[j];
      s_aph += fabsf(ph[j]);
    }
    if (tie_min < kTieEps) {   // rare (about once per 10^5 samples on noisy data): FP64 re-decision, see amc_fused16.cuh
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
    for (int j = 0; j < SPT; ++j) s_f += fq[j];
    }   // DO_PHASE

    // 16 FP64 partials per lane -> lane l (and l + 16) holds the warp total of value l
    __syncwarp();                                             // the previous frame's column reads are done
    const int lv = opaque_if<true>(lane);
#pragma unroll
    for (int i = 0; i < 15; ++i) red[lv * kWRow + i] = mono.s[i];
    red[lv * kWRow + 15] = sum_r;
    float accf[4] = {s_ph, s_aph, s_f, 0.0f};
    warp_sum_multi<float, 4>(accf, lane);                     // lane l: total of value l >> 3
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
    const float mu_ph = __shfl_sync(FULL, accf[0], 0) * (1.0f / N);
    const float mu_aph = __shfl_sync(FULL, accf[0], 8) * (1.0f / N);
    const float tot_f = __shfl_sync(FULL, accf[0], 16);
    const float mu_f = tot_f * (1.0f / (N - 1));              // radians

    // ---------------------------------------------------------------- pass 2 (registers only)
    double c2acc[4] = {0.0, 0.0, 0.0, 0.0};                   // sum |r-mu|, sum (r-mu)^2, sum (r-mu)^4, -
    float q2acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      if constexpr (DO_AMP) {
        const double d = r[j] - mu_r;
        const double d2 = d * d;
        c2acc[0] += fabs(d);
        c2acc[1] += d2;
        c2acc[2] = fma(d2, d2, c2acc[2]);
      }
      if constexpr (DO_PHASE) {
        const float e = ph[j] - mu_ph;
        q2acc[0] = fmaf(e, e, q2acc[0]);
        const float ea = fabsf(ph[j]) - mu_aph;
        q2acc[1] = fmaf(ea, ea, q2acc[1]);
        float ef = fq[j] - mu_f;
        if (j == SPT - 1) ef *= last_keep;
        const float ef2 = ef * ef;
        q2acc[2] += ef2;
        q2acc[3] = fmaf(ef2, ef2, q2acc[3]);
      }
    }
    if constexpr (DO_AMP) warp_sum_multi<double, 4>(c2acc, lane);   // lane l: value l >> 3
    if constexpr (DO_PHASE) warp_sum_multi<float, 4>(q2acc, lane);

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

    // ---------------------------------------------------------------- park this frame's 25 totals
    const int bi = it % kWBatch;
    double* pe = pend + bi * kWPendStride;
    constexpr double kInv2Pi = 0.15915494309189533577, kInv2Pi2 = kInv2Pi * kInv2Pi;
    if (lane < 16) pe[lane] = tot16;                                   // 0..15
    if ((lane & 7) == 0) {                                             // 16..18, 19..22; frequency sums: radians -> cycles
      if (lane < 24) pe[16 + (lane >> 3)] = c2acc[0];
      const double sc = (lane == 16) ? kInv2Pi2 : (lane == 24 ? kInv2Pi2 * kInv2Pi2 : 1.0);
      pe[19 + (lane >> 3)] = static_cast<double>(q2acc[0]) * sc;
    }
    if (lane == 1) pe[23] = static_cast<double>(tot_f) * kInv2Pi;
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
        fs.ph_m2 = pl[19];
        fs.aph_m2 = pl[20];
        fs.f_m2 = pl[21];
        fs.f_m4 = pl[22];
        fs.mean_f = pl[23] / (N - 1);
        fs.spec_max = pl[24];
        const int64_t fo = gg + static_cast<int64_t>(it - bi + lane) * tg;
        constexpr int kChecks = ((DO_FFT || DO_PHASE) ? kCheckRange : 0) | (DO_PHASE ? kCheckPhase : 0) |
                                (DO_AMP ? kCheckAmp : 0);
        if (!finalize_features(fs, N, out + fo * out_stride, kChecks, ticket)) blank_skipped_groups<PROF>(out + fo * out_stride);
      }
      __syncwarp();
    }
  }
}

}  // namespace amc
