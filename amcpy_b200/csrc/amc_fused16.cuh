// Fused sm_100a kernel, 16 samples per thread (frame sizes 512..4096).
//
// Same algorithm and numerics as amc_fused.cuh, re-shaped around what the B200 pipe measurements
// (profiles/r1_pipe_microbench.txt) say is scarce - instruction dispatch and the MIO pipe
// (shuffles, shared-memory wavefronts), not HBM:
//   * N/16 threads per frame (N=2048: 128 threads = one CTA, three CTAs per SM): every per-thread
//     fixed cost (cross-lane reductions, partial stores, barriers, edge handling) is paid once per
//     16 samples instead of once per 8;
//   * FFT as radix 16 x 16 x (N/256), decimation in frequency: stage A (radix 16 over the 16 samples a
//     thread already holds) straight from the registers of pass 1; ONE block-wide exchange hands
//     each WARP complete (N/16)-point sub-transforms (512 points per warp), whose two stages
//     exchange through a warp-private padded buffer with __syncwarp only;
//   * the edge phase (first sample after each warp's run of 32) travels through 64 bytes of shared
//     memory instead of one shuffle per sample;
//   * the 18-feature finalisation is batched: warp 0 parks each frame's 25 totals in shared memory
//     and finalises 32 frames at once, one lane per frame, instead of 32 redundant lanes per frame;
//   * only ONE block barrier per frame: stage A is written before the pass-1 barrier, the x slot is
//     refilled by TMA right after that barrier, everything after it is warp-local, and a frame's
//     totals are collected one frame later (double-buffered partials);
//   * (round 2) phase / |phase| / frequency statistics are ONE-PASS: pass 1 accumulates raw float32 power sums
//     (sum phi, sum phi^2, sum t, sum t^2 with t = |phi| - pi/2, sum f .. sum f^4 of the unwrapped phase steps), the
//     wrapped difference of a sample is formed in the iteration that computes its phase, and the sums are centred in
//     float64 when the frame is finalised - ph[16] and fq[16] no longer live across the barrier (165 registers, no
//     spills, -4 % instructions, 0.6287 -> 0.6079 ms).  Raw sums cancel when |mean| is large against the spread:
//     finalisation computes the cancellation factor from the same sums and hands the frame to the careful path
//     (amc_device.cuh) when it exceeds 4 (phases clustered away from 0 and +-pi/2, or a frequency mean above
//     0.4 sigma - a carrier offset of more than ~0.1 cycles/sample on noisy data): correct, not fast.  Only the
//     amplitude statistics (FP64, |r - mean r|) still need pass 2.
#pragma once
#include "amc_fused.cuh"

namespace amc {

__device__ __forceinline__ constexpr int bitrev4(int r) {
  return ((r & 1) << 3) | ((r & 2) << 1) | ((r & 4) >> 1) | ((r & 8) >> 3);
}

// 16-point forward DFT in registers (4 x 4 Cooley-Tukey); X_r is left in v[bitrev4(r)].
// The 1/sqrt(2) of the W16^2 / W16^6 twiddles is not applied to the values but rides on the butterflies that consume
// them as fused multiply-adds (dft4_hu / dft4_w8, amc_fused.cuh).
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  // now Y[k1][n2]: k1=0 -> v[n2], k1=2 -> v[4+n2], k1=1 -> v[8+n2], k1=3 -> v[12+n2]; twiddle W16^(n2*k1)
  v[9] = c_mul(v[9], make_float2(c1, -s1));                                        // W16^1
  v[11] = c_mul(v[11], make_float2(s1, -c1));                                      // W16^3
  v[13] = c_mul(v[13], make_float2(s1, -c1));                                      // W16^3
  v[15] = c_mul(v[15], make_float2(-c1, s1));                                      // W16^9
  dft4(v[0], v[1], v[2], v[3]);
  dft4_hu(v[8], v[9], v[10], v[11], v[10].x + v[10].y, v[10].y - v[10].x);         // W16^2 v10 = h (x + y, y - x)
  dft4_w8(v[4], v[5], v[6], v[7]);                                                 // W16^2, W16^4, W16^6
  dft4_hu(v[12], v[13], v[14], v[15], v[14].y - v[14].x, -(v[14].x + v[14].y));    // W16^6 v14 = h (y - x, -(x + y))
}

// Twiddle tables of the in-place 16 x 16 x 16 Stockham stages (now used by the long-frame kernel,
// amc_large.cuh: rows N = 4096 of g_tw_s3), laid out the way the warps read them (lane-contiguous), so
// every twiddle load touches one or two 128-byte lines.  (Indexing the generic W_4096 table made each load touch
// 16-28 lines and the L1TEX pipe - not HBM, not the FP pipes - became the kernel's bottleneck:
// 32 % of the run time, see profiles/r1_experiments.txt.)
//   g_tw_s2[q-1][tx]        = W_256^(tx q)          q = 1..15, tx = 0..15
//   g_tw_s3[off(N)][q-1][j] = W_N^(j q)             q = 1..N/256-1, j = 0..255
__device__ float2 g_tw_s2[15 * 16];
__device__ float2 g_tw_s3[(1 + 3 + 7 + 15) * 256];
__host__ __device__ constexpr int tw_s3_offset(int n) {       // rows before frame size n
  return (n == 512 ? 0 : (n == 1024 ? 1 : (n == 2048 ? 4 : 11))) * 256;
}
__global__ void init_twiddle16_kernel() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 15 * 16) {
    const int q = i / 16 + 1, tx = i % 16;
    double sn, cs;
    sincospi(-2.0 * static_cast<double>(tx * q) / 256.0, &sn, &cs);
    g_tw_s2[i] = make_float2(static_cast<float>(cs), static_cast<float>(sn));
  }
  if (i < 26 * 256) {
    const int row = i / 256, j = i % 256;
    int n, q;
    if (row < 1) { n = 512; q = row + 1; }
    else if (row < 4) { n = 1024; q = row - 1 + 1; }
    else if (row < 11) { n = 2048; q = row - 4 + 1; }
    else { n = 4096; q = row - 11 + 1; }
    double sn, cs;
    sincospi(-2.0 * static_cast<double>(j * q) / static_cast<double>(n), &sn, &cs);
    g_tw_s3[i] = make_float2(static_cast<float>(cs), static_cast<float>(sn));
  }
}

// Twiddles of the decimation-in-frequency form used by fused16_features_kernel (lane-contiguous):
//   g_tw_a[off_a(N)][k1-1][t]  = W_N^(t k1)        k1 = 1..15, t  = 0..N/16-1   (after stage A)
//   g_tw_b[off_b(N)][q-1][m1]  = W_(N/16)^(m1 q)   q  = 1..15, m1 = 0..N/256-1  (after stage B)
__device__ float2 g_tw_a[15 * (32 + 64 + 128 + 256)];
__device__ float2 g_tw_b[15 * (2 + 4 + 8 + 16)];
__host__ __device__ constexpr int tw_a_offset(int n) { return 15 * (n == 512 ? 0 : (n == 1024 ? 32 : (n == 2048 ? 96 : 224))); }
__host__ __device__ constexpr int tw_b_offset(int n) { return 15 * (n == 512 ? 0 : (n == 1024 ? 2 : (n == 2048 ? 6 : 14))); }
// The same twiddles in PAIRS, one 16-byte load for two of them (8 loads per stage instead of 15):
//   g_tw_a4[off_a4(N)][p][t]  = {W_N^(t (2p+1)), W_N^(t (2p+2))}            p = 0..7 (the second half of p = 7 is unused)
//   g_tw_b4[off_b4(N)][p][m1] = {W_(N/16)^(m1 (2p+1)), W_(N/16)^(m1 (2p+2))}
__device__ float4 g_tw_a4[8 * (32 + 64 + 128 + 256)];
__device__ float4 g_tw_b4[8 * (2 + 4 + 8 + 16)];
__host__ __device__ constexpr int tw_a4_offset(int n) { return 8 * (n == 512 ? 0 : (n == 1024 ? 32 : (n == 2048 ? 96 : 224))); }
__host__ __device__ constexpr int tw_b4_offset(int n) { return 8 * (n == 512 ? 0 : (n == 1024 ? 2 : (n == 2048 ? 6 : 14))); }
__global__ void init_twiddle16dif_kernel() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 8 * 480) {
    int n = 512, rel = i;
    while (rel >= 8 * (n / 16)) { rel -= 8 * (n / 16); n *= 2; }
    const int grp = n / 16, p = rel / grp, t = rel % grp;
    double s0, c0, s1, c1;
    sincospi(-2.0 * static_cast<double>(t * (2 * p + 1)) / static_cast<double>(n), &s0, &c0);
    sincospi(-2.0 * static_cast<double>(t * (2 * p + 2)) / static_cast<double>(n), &s1, &c1);
    g_tw_a4[i] = make_float4(static_cast<float>(c0), static_cast<float>(s0), static_cast<float>(c1), static_cast<float>(s1));
  }
  if (i < 8 * 30) {
    int n = 512, rel = i;
    while (rel >= 8 * (n / 256)) { rel -= 8 * (n / 256); n *= 2; }
    const int m1n = n / 256, p = rel / m1n, m1 = rel % m1n;
    double s0, c0, s1, c1;
    sincospi(-2.0 * static_cast<double>(m1 * (2 * p + 1)) / static_cast<double>(n / 16), &s0, &c0);
    sincospi(-2.0 * static_cast<double>(m1 * (2 * p + 2)) / static_cast<double>(n / 16), &s1, &c1);
    g_tw_b4[i] = make_float4(static_cast<float>(c0), static_cast<float>(s0), static_cast<float>(c1), static_cast<float>(s1));
  }
  if (i < 15 * 480) {
    int n = 512, rel = i;
    while (rel >= 15 * (n / 16)) { rel -= 15 * (n / 16); n *= 2; }
    const int grp = n / 16, k1 = rel / grp + 1, t = rel % grp;
    double sn, cs;
    sincospi(-2.0 * static_cast<double>(t * k1) / static_cast<double>(n), &sn, &cs);
    g_tw_a[i] = make_float2(static_cast<float>(cs), static_cast<float>(sn));
  }
  if (i < 15 * 30) {
    int n = 512, rel = i;
    while (rel >= 15 * (n / 256)) { rel -= 15 * (n / 256); n *= 2; }
    const int m1n = n / 256, q = rel / m1n + 1, m1 = rel % m1n;
    double sn, cs;
    sincospi(-2.0 * static_cast<double>(m1 * q) / static_cast<double>(n / 16), &sn, &cs);
    g_tw_b[i] = make_float2(static_cast<float>(cs), static_cast<float>(sn));
  }
}

// conflict-free exchange layout for the 16 x 16 x R Stockham passes: low 4 bits ^= bits 4..7
__device__ __forceinline__ int swz16(int e) { return e ^ ((e >> 4) & 15); }

constexpr int kTRow = 17;         // float2 per lane row of the warp-private exchange buffer (16 + 1 pad)
constexpr int kPend16Stride = 29; // doubles per parked frame (28 totals; odd stride: conflict-free lane-per-frame reads)
constexpr int kPartD16 = 28;

template <int N, typename CT>
struct Fused16Cfg {
  static constexpr int SPT = 16;
  static constexpr int GROUP = N / SPT;                  // threads per frame: 32..256
  // one frame group per CTA from N = 1024 up (N = 1024: 64 threads, four CTAs per SM - measured 3 % faster than
  // two 128-thread CTAs holding two groups each); N = 512 keeps four one-warp groups per 128-thread CTA
  static constexpr int CTA = GROUP < 64 ? 128 : GROUP;
  static constexpr int G = CTA / GROUP;
  static constexpr int W = GROUP / 32;
  static constexpr int SLOT_BYTES = N * static_cast<int>(sizeof(CT));   // one x slot (TMA target)
  static constexpr int FFT_BYTES = N * 8;                               // stage-A output, block-wide exchange
  static constexpr int T_BYTES = W * 32 * kTRow * 8;                    // warp-private stage-B -> C exchange
  // per (parity, warp): the frame totals as doubles (index = FrameSums order, see park_and_finalize; slot 24 unused)
  // + float slot 0: the warp's max |X_k|^2 (three more floats of padding)
  static constexpr int PART_D = kPartD16, PART_F = 4;
  static constexpr int PART_BYTES = PART_D * 8 + PART_F * 4;            // 216 per (parity, warp)
  static constexpr int EDGE_BYTES = W * 16 * 4;
  static constexpr int BATCH = 32;                                      // frames finalised together, one lane each
  static constexpr int PEND_BYTES = BATCH * kPend16Stride * 8;
  static constexpr int RAW_BYTES = SLOT_BYTES + FFT_BYTES + T_BYTES + 2 * W * PART_BYTES + EDGE_BYTES + PEND_BYTES + 16;
  static constexpr int GROUP_BYTES = (RAW_BYTES + 127) / 128 * 128;
  static_assert(SLOT_BYTES % 128 == 0 && GROUP_BYTES % 128 == 0, "stage-A rows must stay 128-byte aligned (XOR addressing)");
  static constexpr int SMEM_BYTES = G * GROUP_BYTES;
#if defined(AMC_EXP_2CTA)
  static constexpr int MIN_BLOCKS = (SMEM_BYTES <= 113 * 1024) ? 2 : 1;
#else
  static constexpr int MIN_BLOCKS = (SMEM_BYTES <= 55 * 1024 && CTA <= 128) ? 4 : (SMEM_BYTES <= 75 * 1024) ? 3 : ((SMEM_BYTES <= 113 * 1024) ? 2 : 1);
#endif
  // Round 1 hid the per-lane exchange geometry behind opaque copies of the thread index (opaque_if) so that it was
  // recomputed per frame instead of living in - or being spilled from - registers across pass 1.  With the one-pass
  // statistics (no ph[] / fq[] across the barrier) the kernel has the registers: letting the compiler hoist the
  // invariants is 1.7 % faster (0.6087 -> 0.5982 ms, 168 registers, no spills).
  static constexpr bool REG_BOUND = false;
  static constexpr int M1 = N / 256;                     // radix of the last FFT stage: 2, 4, 8, 16
  static constexpr int LOG_M1 = M1 == 2 ? 1 : (M1 == 4 ? 2 : (M1 == 8 ? 3 : 4));
  static constexpr int F = 32 / M1;                      // (N/16)-point sub-transforms per warp
  static_assert(N >= 512 && N <= 4096 && GROUP % 32 == 0, "frame size outside the 16-samples-per-thread kernel");
  static_assert(GROUP_BYTES % 16 == 0, "group region must keep 16-byte alignment");
  static_assert(F * W == 16, "16 sub-transforms per frame");
};

// Rare: some step of this thread came within kTieEps of +-pi.  Recompute the thread's 16 steps exactly as pass 1 did
// (same float32 function on the same samples -> the same values), re-decide the near-ties in float64 (np.unwrap's
// rules) and return the corrections to the four frequency power sums.
template <int N, typename CT>
__device__ __noinline__ void tie_corrections16(const CT* xs, int t, float (&corr)[4]) {
  constexpr int GROUP = N / 16;
  corr[0] = corr[1] = corr[2] = corr[3] = 0.0f;
  for (int j = 0; j < 16; ++j) {
    const int idx = t + GROUP * j;
    if (idx + 1 >= N) break;
    double a, b;
    float af, bf, cf, df;
    load_sample<CT>(xs + idx, a, b, af, bf);
    load_sample<CT>(xs + idx + 1, a, b, cf, df);
    const float dd = wrap_step_f32(atan2_fast(df, cf) - atan2_fast(bf, af));
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

template <int N, typename CT, int PROF = kProfAll>
__global__ void __launch_bounds__(Fused16Cfg<N, CT>::CTA, Fused16Cfg<N, CT>::MIN_BLOCKS)
fused16_features_kernel(const CT* __restrict__ iq, int64_t n_frames, int64_t frame_stride,
                        double* __restrict__ out, int64_t out_stride, unsigned long long ticket) {
  pdl_launch_dependents();   // the careful-path kernel may be launched now; it waits for this grid to complete
  using Cfg = Fused16Cfg<N, CT>;
  constexpr int GROUP = Cfg::GROUP, W = Cfg::W, SPT = Cfg::SPT, M1 = Cfg::M1, LOG_M1 = Cfg::LOG_M1;
  constexpr bool DO_FFT = (PROF & kProfFft) != 0, DO_PHASE = (PROF & kProfPhase) != 0, DO_AMP = (PROF & kProfAmp) != 0,
                 DO_MOM = (PROF & kProfMom) != 0;
  static_assert(DO_MOM, "every compiled profile keeps the monomial sums: finalize_features detects NaN input from them");
  static_assert(PROF > 0 && PROF <= kProfAll, "unknown feature profile");
  extern __shared__ __align__(128) unsigned char smem_raw[];

  const int tid = threadIdx.x;
  const int g = tid / GROUP;
  const int t = tid % GROUP;
  const int wg = t >> 5;
  const int lane = tid & 31;

  unsigned char* gbase = smem_raw + static_cast<size_t>(g) * Cfg::GROUP_BYTES;
  const CT* xs = reinterpret_cast<const CT*>(gbase);
  float2* buf_a = reinterpret_cast<float2*>(gbase + Cfg::SLOT_BYTES);
  float2* tbuf = reinterpret_cast<float2*>(gbase + Cfg::SLOT_BYTES + Cfg::FFT_BYTES) + wg * (32 * kTRow);
  unsigned char* part_base = gbase + Cfg::SLOT_BYTES + Cfg::FFT_BYTES + Cfg::T_BYTES;
  float* edge_s = reinterpret_cast<float*>(part_base + 2 * W * Cfg::PART_BYTES) + wg * 16;
  double* pend = reinterpret_cast<double*>(part_base + 2 * W * Cfg::PART_BYTES + Cfg::EDGE_BYTES);
  uint64_t* bar = reinterpret_cast<uint64_t*>(part_base + 2 * W * Cfg::PART_BYTES + Cfg::EDGE_BYTES + Cfg::PEND_BYTES);
  uint64_t* rbar = bar + 1;   // "every warp of the group has consumed the block-wide exchange + partials of a frame"

  // (32-bit group ids and an L2 policy created at each use: the hot loop is register-bound and these
  // values are only needed by one thread per frame)
  const int gg = static_cast<int>(blockIdx.x) * Cfg::G + g;
  const int tg = static_cast<int>(gridDim.x) * Cfg::G;
  // frames this group will process in total
  const int my_frames = (gg < n_frames) ? static_cast<int>((n_frames - gg + tg - 1) / tg) : 0;

  if (t == 0) {
    mbar_init(bar, 1);
    mbar_init(rbar, W);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait_primary();                 // the predecessor in the stream has completed: global memory may be touched
  if (lane == 0) mbar_arrive(rbar);   // phase 0 = "nothing to wait for" (keeps the wait in the loop unconditional)
  if (t == 0 && my_frames > 0) {
    mbar_arrive_expect_tx(bar, Cfg::SLOT_BYTES);
    bulk_copy_g2s(gbase, iq + static_cast<int64_t>(gg) * frame_stride, Cfg::SLOT_BYTES, bar, l2_evict_first_policy());
  }

  auto part_d = [&](int par, int w) { return reinterpret_cast<double*>(part_base + (par * W + w) * Cfg::PART_BYTES); };
  auto part_f = [&](int par, int w) {
    return reinterpret_cast<float*>(part_base + (par * W + w) * Cfg::PART_BYTES + Cfg::PART_D * 8);
  };
  // warp 0: collect frame k's totals (25 values, lane i owns value i) and finalise up to 32 frames at a time.
  // Values: 0..14 monomial sums, 15 sum|x|, 16 sum|r-mu|, 17/18 sum (r-mu)^2/^4, 19 sum (phi-mu)^2,
  // 20 sum (|phi|-mu)^2, 21/22 sum (f-mu)^2/^4, 23 sum f, 24 max|X|^2.  Frequency sums were accumulated in
  // radians: 21 <- /(2 pi)^2, 22 <- /(2 pi)^4, 23 <- /(2 pi).
  // ONE-PASS phase / frequency statistics: parked values 19 sum phi^2, 20 sum t^2 (t = |phi| - pi/2), 21 sum f^2, 22 sum f^4,
  // 23 sum f, 24 max|X|^2, 25 sum phi, 26 sum t, 27 sum f^3 (f = unwrapped phase step in radians), centred here in float64.
  auto park_and_finalize = [&](int k) {
    const int par = k & 1, bi = k % Cfg::BATCH;
    double* pe = pend + bi * kPend16Stride;
    // spectral max: float partials (uniform addresses), selected into lane 24 without a divergent FP64 max
    float smax = part_f(par, 0)[0];
#pragma unroll
    for (int w = 1; w < W; ++w) smax = fmaxf(smax, part_f(par, w)[0]);
    if (lane < kPartD16) {
      double s = part_d(par, 0)[lane];
#pragma unroll
      for (int w = 1; w < W; ++w) s += part_d(par, w)[lane];
      pe[lane] = (lane == 24) ? static_cast<double>(smax) : s;
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
        // cancellation factor of the float32 raw sums <= 4, frequency mean small against its spread; otherwise the
        // careful path recomputes the frame
        const bool unsafe = DO_PHASE && (!(pl[19] <= 4.0 * ph_m2) || !(pl[20] <= 4.0 * aph_m2) ||
                                         !(mu_f * mu_f * n1 <= 0.16 * f_m2));
        const int64_t fo = gg + static_cast<int64_t>(k - bi + lane) * tg;
        double* row = out + fo * out_stride;
        constexpr int kChecks = ((DO_FFT || DO_PHASE) ? kCheckRange : 0) | (DO_PHASE ? kCheckPhase : 0) |
                                (DO_AMP ? kCheckAmp : 0);
        if (!finalize_features(fs, N, row, kChecks | (unsafe ? kCheckForce : 0), ticket)) blank_skipped_groups<PROF>(row);
      }
      __syncwarp();
    }
  };

  // loop-invariant exchange geometry (float2 units)
  //   block-wide buffer: element (n1, k1) of stage A's output lives at n1*16 + (k1 ^ rot(n1 & 15)),
  //   rot = rotate-left by 4 - log2(M1) inside 4 bits: conflict-free for the writers (fixed k1, 16
  //   consecutive n1) and for the stage-B readers (fixed m2, lanes = (f, m1)).
  //   (The per-lane addresses below are cheap functions of the thread index; they are recomputed in
  //   every iteration from an opaque copy of it so that they do not occupy registers across pass 1,
  //   where the kernel sits at its 168-register limit.)

  for (int it = 0; it < my_frames; ++it) {
    const int par = it & 1;

    mbar_wait(bar, static_cast<uint32_t>(par));        // frame `it` has landed in the x slot

    // ---------------------------------------------------------------- pass 1 (the only pass over phase / frequency)
    // phase of the sample after this warp's run of 32, for every j: lane j evaluates it, lane 31 uses it
    if constexpr (DO_PHASE) {
      if (lane < SPT) {
        const int te = opaque_if<Cfg::REG_BOUND>(t);
        const int idx = 32 * ((te >> 5) + 1) + GROUP * (te & 31);
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
    }
    Monomials mono;
    double sum_r;
    double r[SPT];
    float xr[SPT], xi[SPT];
    float s_ph = 0.0f, s_aph = 0.0f, s_p2 = 0.0f, s_t2 = 0.0f;   // s_aph = sum t, t = |phi| - pi/2
    float s_f = 0.0f, s_f2 = 0.0f, s_f3 = 0.0f, s_f4 = 0.0f;
    float tie_max = 0.0f;                                      // max |wrapped step| over this thread's steps
    const float last_keep = (t == GROUP - 1) ? 0.0f : 1.0f;   // sample N-1 has no successor
    float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      double a, b;
      load_sample<CT>(xs + t + GROUP * j, a, b, xr[j], xi[j]);
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
      if constexpr (DO_PHASE) {
        const float ph = atan2_fast(xi[j], xr[j]);
        float nb = __shfl_down_sync(0xffffffffu, ph, 1);
        if ((j & 3) == 0) e4 = reinterpret_cast<const float4*>(edge_s)[j >> 2];   // uniform address: broadcast
        const float ej = (j & 3) == 0 ? e4.x : ((j & 3) == 1 ? e4.y : ((j & 3) == 2 ? e4.z : e4.w));
        if (lane == 31) nb = ej;
        float dd = wrap_step_f32(nb - ph);
        tie_max = fmaxf(tie_max, fabsf(dd));
        if (j == SPT - 1) dd *= last_keep;
        s_ph += ph;
        const float tt = fabsf(ph) - kPiO2F;
        s_aph += tt;
        s_p2 = fmaf(ph, ph, s_p2);
        s_t2 = fmaf(tt, tt, s_t2);
        const float d2 = dd * dd;
        s_f += dd;
        s_f2 += d2;
        s_f3 = fmaf(d2, dd, s_f3);
        s_f4 = fmaf(d2, d2, s_f4);
      }
    }
    if constexpr (DO_PHASE) {
      if (tie_max > kPiF - kTieEps) {   // rare (about once per 10^5 samples on noisy data)
        float corr[4];
        tie_corrections16<N, CT>(xs, t, corr);
        s_f += corr[0];
        s_f2 += corr[1];
        s_f3 += corr[2];
        s_f4 += corr[3];
      }
    }
    {
      // 16 FP64 partials per lane -> 16 warp totals: transposed through the warp-private buffer
      // (16 STS.64 + 16 LDS.64 + 16 DADD; the shuffle butterfly needed 60 selects + 32 shuffles).
      // (Interleaving this shared-memory round trip with the register-only radix-16 of stage A in one basic block was
      //  measured in round 2: 0.6321 vs 0.6279 ms - slower; profiles/r2_experiments.txt.)
      double* red = reinterpret_cast<double*>(tbuf);
      __syncwarp();                                     // stage C of the previous frame has read tbuf
#pragma unroll
      for (int i = 0; i < 15; ++i) red[lane * kTRow + i] = mono.s[i];
      red[lane * kTRow + 15] = sum_r;
      float accf[8] = {s_p2, s_t2, s_f2, s_f4, s_f, s_ph, s_aph, s_f3};   // -> parked 19, 20, 21, 22, 23, 25, 26, 27
      warp_sum_multi<float, 8>(accf, lane);
      __syncwarp();
      const double* col = red + (lane >> 4) * (16 * kTRow) + (lane & 15);
      double cs[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) cs[i] = col[i * kTRow];
#pragma unroll
      for (int w = 8; w >= 1; w >>= 1)                  // pairwise tree: same error growth as the butterfly
#pragma unroll
        for (int i = 0; i < w; ++i) cs[i] += cs[i + w];
      const double tot = cs[0] + __shfl_xor_sync(0xffffffffu, cs[0], 16);
      // every warp has finished reading the previous frame's stage-A output and (warp 0) the partials
      // that are about to be overwritten; in steady state this completed long ago
      mbar_wait(rbar, static_cast<uint32_t>(par));
      if (lane < 16) part_d(par, wg)[lane] = tot;
      if ((lane & 3) == 0) {
        const int kk = lane >> 2;                        // 0..7
        part_d(par, wg)[19 + kk + (kk >= 5 ? 1 : 0)] = static_cast<double>(accf[0]);
      }
    }

    // ---------------------------------------------------------------- FFT stage A: radix 16 over this
    // thread's own samples x[t + GROUP j] (registers), then the twiddle W_N^(t k1)
    float2 v[16];
    const int tv = opaque_if<Cfg::REG_BOUND>(t);          // keeps the geometry below out of pass 1
    const int lv = tv & 31, wv = tv >> 5;
    if constexpr (DO_FFT) {
      const int rot_t = (((tv & 15) << (4 - LOG_M1)) | ((tv & 15) >> LOG_M1)) & 15;
      const uint32_t row_a = (smem_u32(buf_a) + 128u * tv) ^ (8u * rot_t);   // row start is 128-byte aligned
#pragma unroll
#ifdef AMC_EXP_RELOAD_X
      for (int q = 0; q < 16; ++q) {                      // A/B: re-read x from the slot instead of keeping xr/xi live
        double a, b;
        load_sample<CT>(xs + tv + GROUP * q, a, b, v[q].x, v[q].y);
      }
#else
      for (int q = 0; q < 16; ++q) v[q] = make_float2(xr[q], xi[q]);
#endif
      dft16(v);
      float4 tw[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) tw[p] = g_tw_a4[tw_a4_offset(N) + p * GROUP + tv];            // 512 B per warp load
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

    group_sync<GROUP, Cfg::CTA>(g);   // THE barrier: pass-1 partials + stage-A output visible; x slot fully read

    if (t == 0 && it + 1 < my_frames) {                 // refill the x slot: frame it+1 streams in during pass 2 + FFT
      fence_proxy_async_smem();
      mbar_arrive_expect_tx(bar, Cfg::SLOT_BYTES);
      bulk_copy_g2s(gbase, iq + (gg + static_cast<int64_t>(it + 1) * tg) * frame_stride, Cfg::SLOT_BYTES, bar,
                    l2_evict_first_policy());
    }
    if (wg == 0 && it > 0) park_and_finalize(it - 1);   // the previous frame's totals are complete now

    double tot_r = 0.0;
#pragma unroll
    for (int w = 0; w < W; ++w) tot_r += part_d(par, w)[15];
    const double mu_r = tot_r * (1.0 / N);

    // ---------------------------------------------------------------- pass 2 (registers only)
    {
      // (sum (r-mu)^2 must be accumulated: deriving it as sum|x|^2 - (sum|x|)^2/N is one FP64 op cheaper and
      //  2 % faster, but it exposes the ~5e-14 one-sided bias of sqrt_nr multiplied by (mu/sigma)^2 - 1.3e-9
      //  on the amplitude kurtosis at 30 dB SNR; test_high_snr_amplitude_features_keep_the_1e9_class)
      double c2acc[4] = {0.0, 0.0, 0.0, 0.0};            // sum |r-mu|, sum (r-mu)^2, sum (r-mu)^4, -
#pragma unroll
      for (int j = 0; j < SPT; ++j) {
        if constexpr (DO_AMP) {
          const double d = r[j] - mu_r;
          const double d2 = d * d;
          c2acc[0] += fabs(d);
          c2acc[1] += d2;
          c2acc[2] = fma(d2, d2, c2acc[2]);
        }
      }
      if constexpr (DO_AMP) warp_sum_multi<double, 4>(c2acc, lane);
      if ((lane & 7) == 0 && lane < 24) part_d(par, wg)[16 + (lane >> 3)] = c2acc[0];
    }

    float vmax = 0.0f;
    if constexpr (!DO_FFT) {
      __syncwarp();
      if (lane == 0) mbar_arrive(rbar);   // this warp's reads of the parked partials are done
    } else {
    // ---------------------------------------------------------------- FFT stage B: this warp owns the
    // sub-transforms k1 = wg F .. wg F + F-1 (each GROUP points, n1 = m1 + M1 m2); lane (f, m1) does the
    // radix-16 over m2 and applies W_GROUP^(m1 q)
    float2* tb = reinterpret_cast<float2*>(gbase + Cfg::SLOT_BYTES + Cfg::FFT_BYTES) + wv * (32 * kTRow);
    {
      const int m1 = lv & (M1 - 1);
      const int k1b = wv * Cfg::F + (lv >> LOG_M1);                  // sub-transform of this lane in stage B
      const int kk_b = k1b ^ ((m1 << (4 - LOG_M1)) & 15);
      float2* wr_t = tb + lv * kTRow;
      float4 tw[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) tw[p] = g_tw_b4[tw_b4_offset(N) + p * M1 + m1];
#pragma unroll
      for (int m2 = 0; m2 < 16; ++m2) {
        // (n1 & 15) = m1 + M1 (m2 mod 16/M1): the rotated swizzle is base ^ (m2 mod 16/M1)
        v[m2] = buf_a[m1 * 16 + (kk_b ^ (m2 & (16 / M1 - 1))) + 16 * M1 * m2];
      }
      dft16(v);
#pragma unroll
      for (int q = 1; q < 16; ++q) {
        const float4 w = tw[(q - 1) >> 1];
        v[bitrev4(q)] = c_mul(v[bitrev4(q)], (q & 1) ? make_float2(w.x, w.y) : make_float2(w.z, w.w));
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) wr_t[q] = v[bitrev4(q)];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(rbar);   // this warp's reads of buf_a (and warp 0's parked partials) are done
    // ---------------------------------------------------------------- FFT stage C (last): radix M1 over m1,
    // 16 / M1 butterflies per lane, no twiddles; only max |X_k|^2 is kept
    {
      const float2* rd = tb + (lv >> 4) * (M1 * kTRow) + (lv & 15);
#pragma unroll 1   // (rolled: unrolled it is 1 % SLOWER - 0.6063 vs 0.6002 ms, round 2 - and 1.3 KB more code)
      for (int bb = 0; bb < 16 / M1; ++bb, rd += 2 * M1 * kTRow) {
        float2 u[M1];
#pragma unroll
        for (int q = 0; q < M1; ++q) u[q] = rd[q * kTRow];
        if constexpr (M1 == 2) {
          bfly2(u[0], u[1]);
        } else if constexpr (M1 == 4) {
          dft4(u[0], u[1], u[2], u[3]);
        } else if constexpr (M1 == 8) {
          float2(&u8)[8] = reinterpret_cast<float2(&)[8]>(u);
          dft8(u8);
        } else {
          float2(&u16)[16] = reinterpret_cast<float2(&)[16]>(u);
          dft16(u16);
        }
#pragma unroll
        for (int q = 0; q < M1; ++q) vmax = fmaxf(vmax, fmaf(u[q].x, u[q].x, u[q].y * u[q].y));
      }
    }
    vmax = warp_max(vmax);
    }   // DO_FFT
    if (lane == 0) part_f(par, wg)[0] = vmax;
    // no barrier here: buf_a and the partial arrays are protected by rbar (waited on in the next frame's
    // pass 1); the warp-private buffer is only touched by this warp
  }

  if (my_frames > 0) {                                  // (uniform per group)
    group_sync<GROUP, Cfg::CTA>(g);                     // the last frame's spectral max is visible
    if (wg == 0) park_and_finalize(my_frames - 1);
  }
}

}  // namespace amc
