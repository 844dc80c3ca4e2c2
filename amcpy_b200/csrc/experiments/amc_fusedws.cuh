// Warp-specialised variant of the 16-samples-per-thread kernel (N = 2048), behind AMC_FLAG_FUSED_WS for A/B runs.
//
// Same arithmetic as amc_fused16.cuh, different assignment of work to warps: a 256-thread CTA owns one
// frame at a time; warps 0-3 ("M") do the FP64 work of that frame (monomial sums, |x|, centred amplitude
// sums, parking / finalisation), warps 4-7 ("P") the FP32 work (atan2, wrapped differences, phase /
// frequency statistics, FFT).  Both groups read the same TMA slot, each has its own named barrier, and they
// only meet through three counters:
//   * full   (mbarrier)  the frame has landed in the x slot;
//   * x_cnt  (atomic)    every warp bumps it after its last read of x; the eighth one issues the TMA copy of
//                        the next frame - nobody ever blocks on the other group for it;
//   * p_done (counter)   P warps bump it at the end of a frame; M warp 0 checks it before it parks that
//                        frame's totals (partial rows are triple-buffered, so the groups may drift by a frame).
// The point of the experiment (profiles/r1_experiments.txt): two CTAs x 8 warps = 16 warps per SM at 128
// registers (the unified kernel needs 168: 12 warps), and every scheduler always holds FP64-heavy and
// FP32-heavy warps side by side.
#pragma once
#include "../amc_fused16.cuh"

namespace amc {

template <int N, typename CT>
struct FusedWsCfg {
  static_assert(N == 2048, "the warp-specialised variant is instantiated for N = 2048");
  static constexpr int SPT = 16, GROUP = N / SPT, W = GROUP / 32;       // 128 threads, 4 warps per role
  static constexpr int CTA = 2 * GROUP;
  static constexpr int M1 = N / 256, LOG_M1 = 3, F = 32 / M1;
  static constexpr int SLOT_BYTES = N * static_cast<int>(sizeof(CT));
  static constexpr int FFT_BYTES = N * 8;
  static constexpr int T_BYTES = 2 * W * 32 * kTRow * 8;                 // one scratch per warp (both roles)
  static constexpr int PART_D = 25, PART_F = 4;
  static constexpr int PART_BYTES = PART_D * 8 + PART_F * 4;             // per (buffer, role warp)
  static constexpr int N_PART_BUF = 3;
  static constexpr int EDGE_BYTES = W * 16 * 4;
  static constexpr int BATCH = 32;
  static constexpr int PEND_BYTES = BATCH * kPend16Stride * 8;
  static constexpr int RAW_BYTES = SLOT_BYTES + FFT_BYTES + T_BYTES + N_PART_BUF * 2 * W * PART_BYTES + EDGE_BYTES + PEND_BYTES + 64;
  static constexpr int SMEM_BYTES = (RAW_BYTES + 127) / 128 * 128;
};

template <int N, typename CT>
__global__ void __launch_bounds__(FusedWsCfg<N, CT>::CTA, 2)
fusedws_features_kernel(const CT* __restrict__ iq, int64_t n_frames, int64_t frame_stride,
                        double* __restrict__ out, int64_t out_stride) {
  using Cfg = FusedWsCfg<N, CT>;
  constexpr int GROUP = Cfg::GROUP, W = Cfg::W, SPT = Cfg::SPT, M1 = Cfg::M1, LOG_M1 = Cfg::LOG_M1;
  extern __shared__ __align__(128) unsigned char smem_raw[];

  const int tid = threadIdx.x;
  const int role = tid >> 7;                   // 0: M (FP64), 1: P (FP32)
  const int t = tid & (GROUP - 1);
  const int wg = t >> 5;
  const int lane = tid & 31;

  const CT* xs = reinterpret_cast<const CT*>(smem_raw);
  float2* buf_a = reinterpret_cast<float2*>(smem_raw + Cfg::SLOT_BYTES);
  float2* tbuf = reinterpret_cast<float2*>(smem_raw + Cfg::SLOT_BYTES + Cfg::FFT_BYTES) + (role * W + wg) * (32 * kTRow);
  unsigned char* part_base = smem_raw + Cfg::SLOT_BYTES + Cfg::FFT_BYTES + Cfg::T_BYTES;
  float* edge_s = reinterpret_cast<float*>(part_base + Cfg::N_PART_BUF * 2 * W * Cfg::PART_BYTES) + wg * 16;
  double* pend = reinterpret_cast<double*>(part_base + Cfg::N_PART_BUF * 2 * W * Cfg::PART_BYTES + Cfg::EDGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(pend) + Cfg::PEND_BYTES);
  uint64_t* rbar = full + 1;                   // P: every P warp has consumed the previous stage-A output
  int* x_cnt = reinterpret_cast<int*>(full + 2);
  volatile int* p_done = reinterpret_cast<volatile int*>(x_cnt + 1);

  const int gg = static_cast<int>(blockIdx.x);
  const int tg = static_cast<int>(gridDim.x);
  const int my_frames = (gg < n_frames) ? static_cast<int>((n_frames - gg + tg - 1) / tg) : 0;

  // partial rows: buffer b (frame % 3), role-warp w (0..3 M, 4..7 P)
  auto part_d = [&](int b, int w) { return reinterpret_cast<double*>(part_base + (b * 2 * W + w) * Cfg::PART_BYTES); };
  auto part_f = [&](int b, int w) {
    return reinterpret_cast<float*>(part_base + (b * 2 * W + w) * Cfg::PART_BYTES + Cfg::PART_D * 8);
  };

  if (tid == 0) {
    mbar_init(full, 1);
    mbar_init(rbar, W);
    *x_cnt = 0;
    *p_done = 0;
    fence_mbar_init();
  }
  __syncthreads();
  if (role == 1 && lane == 0) mbar_arrive(rbar);     // phase 0 = "nothing to wait for"
  if (tid == 0 && my_frames > 0) {
    mbar_arrive_expect_tx(full, Cfg::SLOT_BYTES);
    bulk_copy_g2s(smem_raw, iq + static_cast<int64_t>(gg) * frame_stride, Cfg::SLOT_BYTES, full, l2_evict_first_policy());
  }

  // called by lane 0 of every warp after the warp's last read of x for frame `it`
  auto release_x = [&](int it) {
    __threadfence_block();
    if (atomicAdd(x_cnt, 1) == 2 * W - 1) {
      atomicExch(x_cnt, 0);
      if (it + 1 < my_frames) {
        __threadfence_block();
        fence_proxy_async_smem();
        mbar_arrive_expect_tx(full, Cfg::SLOT_BYTES);
        bulk_copy_g2s(smem_raw, iq + (gg + static_cast<int64_t>(it + 1) * tg) * frame_stride, Cfg::SLOT_BYTES, full,
                      l2_evict_first_policy());
      }
    }
  };

  if (role == 0) {
    // ======================================================================================= M: FP64
    // M warp 0: collect frame k's totals (lane i owns value i: 0..18 from the M rows, 19..24 from the P rows)
    auto park_and_finalize = [&](int k) {
      const int b = k % Cfg::N_PART_BUF, bi = k % Cfg::BATCH;
      while (*p_done < W * (k + 1)) {
      }                                                // the P warps have finished frame k
      __threadfence_block();
      double* pe = pend + bi * kPend16Stride;
      if (lane < 25) {
        const int w0 = lane < 19 ? 0 : W;
        double s = part_d(b, w0)[lane];
        if (lane == 24) {
#pragma unroll
          for (int w = 1; w < W; ++w) s = fmax(s, part_d(b, w0 + w)[lane]);
        } else {
#pragma unroll
          for (int w = 1; w < W; ++w) s += part_d(b, w0 + w)[lane];
        }
        constexpr double k1 = 0.15915494309189533577, k2 = k1 * k1;
        if (lane >= 21 && lane <= 23) s *= (lane == 21) ? k2 : (lane == 22 ? k2 * k2 : k1);
        pe[lane] = s;
      }
      if (bi == Cfg::BATCH - 1 || k == my_frames - 1) {
        __syncwarp();
        if (lane <= bi) {
          const double* pl = pend + lane * kPend16Stride;
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
          const int64_t fo = gg + static_cast<int64_t>(k - bi + lane) * tg;
          finalize_features(fs, N, out + fo * out_stride, kCheckAll);
        }
        __syncwarp();
      }
    };

    for (int it = 0; it < my_frames; ++it) {
      const int b = it % Cfg::N_PART_BUF;
      mbar_wait(full, static_cast<uint32_t>(it & 1));
      Monomials mono;
      double sum_r;
      double r[SPT];
#pragma unroll
      for (int j = 0; j < SPT; ++j) {
        const CT v = xs[t + GROUP * j];
        const double a = static_cast<double>(v.x), bb = static_cast<double>(v.y);
        const double s = (j == 0) ? mono.init(a, bb) : mono.add(a, bb);
        r[j] = sqrt_nr(s);
        sum_r = (j == 0) ? r[j] : sum_r + r[j];
      }
      __syncwarp();
      if (lane == 0) release_x(it);
      {
        double* red = reinterpret_cast<double*>(tbuf);
#pragma unroll
        for (int i = 0; i < 15; ++i) red[lane * kTRow + i] = mono.s[i];
        red[lane * kTRow + 15] = sum_r;
        __syncwarp();
        const double* col = red + (lane >> 4) * (16 * kTRow) + (lane & 15);
        double cs[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) cs[i] = col[i * kTRow];
#pragma unroll
        for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
          for (int i = 0; i < w; ++i) cs[i] += cs[i + w];
        const double tot = cs[0] + __shfl_xor_sync(0xffffffffu, cs[0], 16);
        __syncwarp();                                  // column reads done before the next frame's row writes
        if (lane < 16) part_d(b, wg)[lane] = tot;
      }
      named_bar_sync(1, GROUP);                        // M-group barrier: sum|x| of the four M warps visible
      double tot_r = 0.0;
#pragma unroll
      for (int w = 0; w < W; ++w) tot_r += part_d(b, w)[15];
      const double mu_r = tot_r * (1.0 / N);
      {
        double c2acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
          const double d = r[j] - mu_r;
          const double d2 = d * d;
          c2acc[0] += fabs(d);
          c2acc[1] += d2;
          c2acc[2] = fma(d2, d2, c2acc[2]);
        }
        warp_sum_multi<double, 4>(c2acc, lane);
        if ((lane & 7) == 0 && lane < 24) part_d(b, wg)[16 + (lane >> 3)] = c2acc[0];
      }
      if (wg == 0 && it > 0) park_and_finalize(it - 1);
    }
    if (my_frames > 0) {
      named_bar_sync(1, GROUP);                        // the last frame's centred sums are visible
      if (wg == 0) park_and_finalize(my_frames - 1);
    }
  } else {
    // ======================================================================================= P: FP32
    for (int it = 0; it < my_frames; ++it) {
      const int b = it % Cfg::N_PART_BUF;
      const int par = it & 1;
      mbar_wait(full, static_cast<uint32_t>(par));
      float ph[SPT], xr[SPT], xi[SPT];
#pragma unroll
      for (int j = 0; j < SPT; ++j) {
        double a, bb;
        load_sample<CT>(xs + t + GROUP * j, a, bb, xr[j], xi[j]);
        ph[j] = atan2_fast(xi[j], xr[j]);
      }
      if (lane < SPT) {
        const int idx = 32 * (wg + 1) + GROUP * lane;
        float pe = 0.0f;
        if (idx < N) {
          double a, bb;
          float af, bf;
          load_sample<CT>(xs + idx, a, bb, af, bf);
          pe = atan2_fast(bf, af);
        }
        edge_s[lane] = pe;
      }
      __syncwarp();
      float fq[SPT];
      float s_ph = 0.0f, s_aph = 0.0f;
      float tie_min = 1.0f;
      const float last_keep = (t == GROUP - 1) ? 0.0f : 1.0f;
      float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < SPT; ++j) {
        float nb = __shfl_down_sync(0xffffffffu, ph[j], 1);
        if ((j & 3) == 0) e4 = reinterpret_cast<const float4*>(edge_s)[j >> 2];
        const float ej = (j & 3) == 0 ? e4.x : ((j & 3) == 1 ? e4.y : ((j & 3) == 2 ? e4.z : e4.w));
        if (lane == 31) nb = ej;
        float dd = nb - ph[j];
        const float over = fabsf(dd) - kPiF;
        tie_min = fminf(tie_min, fabsf(over));
        if (over > 0.0f) dd -= copysignf(kTwoPiF, dd);
        if (j == SPT - 1) dd *= last_keep;
        fq[j] = dd;
        s_ph += ph[j];
        s_aph += fabsf(ph[j]);
      }
      if (tie_min < kTieEps) {
        unsigned tie_mask = 0u;
#pragma unroll
        for (int j = 0; j < SPT; ++j)
          if (fabsf(kPiF - fabsf(fq[j])) < 2.0f * kTieEps) tie_mask |= 1u << j;
        while (tie_mask != 0u) {
          const int j = __ffs(tie_mask) - 1;
          tie_mask &= tie_mask - 1u;
          const float val = exact_phase_step<CT>(xs, t + GROUP * j);
#pragma unroll
          for (int q = 0; q < SPT; ++q) fq[q] = (q == j) ? val : fq[q];
        }
      }
      __syncwarp();
      if (lane == 0) release_x(it);
      float s_f = 0.0f;
#pragma unroll
      for (int j = 0; j < SPT; ++j) s_f += fq[j];
      {
        float accf[4] = {s_ph, s_aph, s_f, 0.0f};
        warp_sum_multi<float, 4>(accf, lane);
        mbar_wait(rbar, static_cast<uint32_t>(par));   // previous stage-A output consumed by every P warp
        if ((lane & 7) == 0 && lane < 24) part_f(b, W + wg)[lane >> 3] = accf[0];
        if (lane == 16) part_d(b, W + wg)[23] = static_cast<double>(accf[0]);
      }
      // ---------------------------------------------------------------- FFT stage A
      float2 v[16];
      {
        const int rot_t = (((t & 15) << (4 - LOG_M1)) | ((t & 15) >> LOG_M1)) & 15;
        const uint32_t row_a = (smem_u32(buf_a) + 128u * t) ^ (8u * rot_t);
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = make_float2(xr[q], xi[q]);
        dft16(v);
        float2 tw[15];
#pragma unroll
        for (int q = 1; q < 16; ++q) tw[q - 1] = g_tw_a[tw_a_offset(N) + (q - 1) * GROUP + t];
#pragma unroll
        for (int q = 1; q < 16; ++q) v[bitrev4(q)] = c_mul(v[bitrev4(q)], tw[q - 1]);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float2 o = v[bitrev4(q)];
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(row_a ^ (8u * q)), "f"(o.x), "f"(o.y) : "memory");
        }
      }
      named_bar_sync(2, GROUP);                        // P-group barrier: float sums + stage-A output visible

      float tot_ph = 0.0f, tot_aph = 0.0f, tot_f = 0.0f;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        tot_ph += part_f(b, W + w)[0];
        tot_aph += part_f(b, W + w)[1];
        tot_f += part_f(b, W + w)[2];
      }
      const float mu_ph = tot_ph * (1.0f / N), mu_aph = tot_aph * (1.0f / N);
      const float mu_f = tot_f * (1.0f / (N - 1));
      {
        float q2acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
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
        warp_sum_multi<float, 4>(q2acc, lane);
        if ((lane & 7) == 0) part_d(b, W + wg)[19 + (lane >> 3)] = static_cast<double>(q2acc[0]);
      }
      // ---------------------------------------------------------------- FFT stages B and C (warp-local)
      float vmax = 0.0f;
      {
        const int m1 = lane & (M1 - 1);
        const int k1b = wg * Cfg::F + (lane >> LOG_M1);
        const int kk_b = k1b ^ ((m1 << (4 - LOG_M1)) & 15);
        float2* wr_t = tbuf + lane * kTRow;
        float2 tw[15];
#pragma unroll
        for (int q = 1; q < 16; ++q) tw[q - 1] = g_tw_b[tw_b_offset(N) + (q - 1) * M1 + m1];
#pragma unroll
        for (int m2 = 0; m2 < 16; ++m2) v[m2] = buf_a[m1 * 16 + (kk_b ^ (m2 & (16 / M1 - 1))) + 16 * M1 * m2];
        dft16(v);
#pragma unroll
        for (int q = 1; q < 16; ++q) v[bitrev4(q)] = c_mul(v[bitrev4(q)], tw[q - 1]);
#pragma unroll
        for (int q = 0; q < 16; ++q) wr_t[q] = v[bitrev4(q)];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(rbar);
      {
        const float2* rd = tbuf + (lane >> 4) * (M1 * kTRow) + (lane & 15);
#pragma unroll 1
        for (int bb = 0; bb < 16 / M1; ++bb, rd += 2 * M1 * kTRow) {
          float2 u[M1];
#pragma unroll
          for (int q = 0; q < M1; ++q) u[q] = rd[q * kTRow];
          float2(&u8)[8] = reinterpret_cast<float2(&)[8]>(u);
          dft8(u8);
#pragma unroll
          for (int q = 0; q < M1; ++q) vmax = fmaxf(vmax, fmaf(u[q].x, u[q].x, u[q].y * u[q].y));
        }
      }
      vmax = warp_max(vmax);
      __syncwarp();                                    // the other lanes' partial stores precede lane 0's signal
      if (lane == 0) {
        part_d(b, W + wg)[24] = static_cast<double>(vmax);
        __threadfence_block();
        atomicAdd(const_cast<int*>(p_done), 1);        // this warp's partials of frame `it` are complete
      }
      __syncwarp();                                    // tbuf reads done before the next frame's use
    }
  }
}

}  // namespace amc
