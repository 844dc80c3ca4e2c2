// Device-side building blocks shared by the fused and the general feature kernels (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace amc {

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 6.28318530717958647692;
constexpr double kPiO4 = 0.78539816339744830962;
constexpr double k3PiO4 = 2.35619449019234492885;
constexpr float kPiF = 3.14159265358979323846f;
constexpr float kTwoPiF = 6.28318530717958647692f;
constexpr float kPiO2F = 1.57079632679489661923f;
constexpr float kInvTwoPiF = 0.15915494309189533577f;
// |dd| within this distance of pi (float32 path) is re-decided in float64 (np.unwrap tie rules)
constexpr float kTieEps = 4.0e-6f;

// ------------------------------------------------------------------ mbarrier / bulk async copy (TMA)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// One bulk (non-tensor) TMA copy global -> shared, completion signalled on `bar` (SASS: UBLKCP).
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                              uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Returns v, hidden from the optimiser when HIDE: values derived from it are recomputed where they are
// used instead of being kept in (or spilled from) registers across the whole frame loop.
template <bool HIDE>
__device__ __forceinline__ int opaque_if(int v) {
  if constexpr (HIDE) asm volatile("" : "+r"(v));
  return v;
}

// ------------------------------------------------------------------ warp reductions
// Sum V values (V a power of two <= 32) across the warp with ~V shuffles instead of 5*V:
// at every step half of the values travel to the partner lane.  Afterwards lane l holds the
// warp total of value index (l >> (5 - log2 V)) in v[0].  Fixed order => bitwise reproducible.
template <typename T, int V>
__device__ __forceinline__ void warp_sum_multi(T (&v)[V], int lane) {
  int width = 16;
#pragma unroll
  for (int cnt = V; cnt > 1; cnt >>= 1, width >>= 1) {
    const bool up = (lane & width) != 0;
#pragma unroll
    for (int i = 0; i < cnt / 2; ++i) {
      const T keep = up ? v[i + cnt / 2] : v[i];
      const T send = up ? v[i] : v[i + cnt / 2];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, width);
    }
  }
#pragma unroll
  for (; width >= 1; width >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], width);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, w));
  return v;
}

// ------------------------------------------------------------------ float64 reference-grade pieces
// atan2 with the exact octant constants a correctly rounded libm returns (np.angle on |re|==|im|).
__device__ __forceinline__ double atan2_exact(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  if (ax == ay && ax != 0.0 && ax <= 1.7976931348623157e308) {
    return copysign(x > 0.0 ? kPiO4 : k3PiO4, y);
  }
  return atan2(y, x);
}

// One step of np.unwrap (period 2*pi, discont pi) applied to a raw phase difference:
// returns diff(unwrapped)[i] = dd + ph_correct.  numpy: ddmod = mod(dd+pi, 2pi) - pi with
// floor-mod; (ddmod == -pi and dd > 0) -> +pi; |dd| < pi -> no correction.
__device__ __forceinline__ double unwrap_step(double dd) {
  if (fabs(dd) < kPi) return dd;
  double m = fmod(dd + kPi, kTwoPi);
  if (m != 0.0) {
    if (m < 0.0) m += kTwoPi;
  } else {
    m = 0.0;
  }
  double ddmod = m - kPi;
  if (ddmod == -kPi && dd > 0.0) ddmod = kPi;
  return ddmod;
}

// ------------------------------------------------------------------ float32 atan2 (1e-6 class)
// atan(q) on [0,1] = q + q*s*P(s), s = q*q, degree-7 P fitted minimax (abs err 7e-9 before rounding).  One term more
// than float32 rounding alone would ask for: the equi-oscillating error of the degree-6 fit (5e-8, period ~0.1 rad)
// has a local SLOPE of ~1e-6, which is the relative error it puts on the standard deviation of a narrow phase cluster
// (unmodulated carrier / strong DC line at high SNR: found by tests/soak.py, 1.2e-6 on features 2 and 3).
// Predicate-free octant / quadrant fix-ups (FSET.BF + sign-bit masks) so that many evaluations can
// be interleaved without spilling predicates.  atan2(+-0, +-0) follows IEEE / np.angle.
__device__ __forceinline__ float atan2_fast(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(fmaxf(ax, ay), 1.0e-37f), mn = fminf(ax, ay);
  float rc;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(mx));
  const float q = mn * rc;
  const float s = q * q;
  float p = 0.0026222064831683623f;
  p = fmaf(p, s, -0.015132382451695708f);
  p = fmaf(p, s, 0.0411216062546977f);
  p = fmaf(p, s, -0.07366684174003307f);
  p = fmaf(p, s, 0.10573921500635099f);
  p = fmaf(p, s, -0.14185972498001842f);
  p = fmaf(p, s, 0.19990396259243107f);
  p = fmaf(p, s, -0.33332987041851964f);
  const float r0 = fmaf(q * s, p, q);                       // atan(q) in [0, pi/4]
  // Octant and quadrant in five instructions: with r1 = (ay > ax ? pi/2 - r0 : r0) the answer is
  // copysign(x >= 0 ? r1 : pi - r1, y) = copysign(pi/2 - copysign(pi/2 - r1, x), y), and
  // u = pi/2 - r1 = | r0 - [ay <= ax] pi/2 | needs no separate r1 (round 1 spent six: the sign of x through an integer
  // shift + mask; 0.6023 -> 0.5951 ms at N = 2048).  Signs travel as bits, so x = -0 counts as negative like np.angle / atan2 do; pi/2 + pi/2 is exactly
  // the float32 pi that pi - r1 used.
  const float keep = (ay > ax) ? 0.0f : 1.0f;               // FSET.BF
  const float u = fmaf(keep, -kPiO2F, r0);                  // +-(pi/2 - r1); only the magnitude is used
  unsigned vb;                                              // copysign(|u|, x) as ONE LOP3: (u & ~m) | (x & m)  (LUT 0xD8)
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=r"(vb) : "r"(__float_as_uint(u)), "r"(__float_as_uint(x)));
  const float r = kPiO2F - __uint_as_float(vb);             // [0, pi]
  // copysign(r, y) as ONE LOP3: (r & ~m) | (y & m), m = sign mask  (LUT 0xD8)
  unsigned res;
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xD8;" : "=r"(res) : "r"(__float_as_uint(r)), "r"(__float_as_uint(y)));
  return __uint_as_float(res);
}

// One float32 step of np.unwrap away from the ties: dd - 2 pi rint(dd / 2 pi), |dd| < 2 pi + eps, with the rounding done by
// the 1.5 * 2^23 trick - three FP32-pipe instructions and no predicate (compare + copysign + predicated add: four).  For
// k = +-1 the fused multiply-add returns the same float as dd -+ kTwoPiF.  The decision boundary sits within 2e-7 of pi;
// callers treat every step that ends within kTieEps of +-pi as a tie and re-decide it in float64.
__device__ __forceinline__ float wrap_step_f32(float dd) {
  const float k = __fadd_rn(fmaf(dd, kInvTwoPiF, 12582912.0f), -12582912.0f);
  return fmaf(k, -kTwoPiF, dd);
}

// ------------------------------------------------------------------ |x| in float64 without DSQRT
// r = sqrt(s) from MUFU.RSQ64H (rsqrt.approx.f64, ~2^-22) + one coupled Newton step (rel err ~1e-13).
// The seed's input is clamped and halved with integer ops on the high word (s == 0 -> r == 0).
__device__ __forceinline__ double sqrt_nr(double s) {
  const int hi = max(__double2hiint(s), 0x00200000);        // >= 2^-1021: rsqrt stays finite, NaN stays NaN
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(__hiloint2double(hi, 0)));
  // note: for normal s the seed ignores the low word anyway (MUFU.RSQ64H reads the high word only)
  const double yh = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));   // y / 2
  const double r0 = s * y;
  const double e = fma(-r0, yh, 0.5);
  return fma(r0, e, r0);
}

// ------------------------------------------------------------------ per-sample moment monomials
// 15 running sums of a^p b^q, p+q in {2,4,6}; every M_pq of features.py:46-58 is a fixed linear
// combination of them (see finalize_features).  21 FP64 ops per sample.
struct Monomials {
  double s[15];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 15; ++i) s[i] = 0.0;
  }
  // first sample: plain stores instead of 15 accumulations onto zero; returns a*a + b*b
  __device__ __forceinline__ double init(double a, double b) {
    const double a2 = a * a, b2 = b * b, ab = a * b;
    s[0] = a2;
    s[1] = b2;
    s[2] = ab;
    const double a4 = a2 * a2, b4 = b2 * b2, a2b2 = a2 * b2;
    s[3] = a4;
    s[4] = a2 * ab;
    s[5] = a2b2;
    s[6] = ab * b2;
    s[7] = b4;
    s[8] = a4 * a2;
    s[9] = a4 * ab;
    s[10] = a4 * b2;
    s[11] = a2b2 * ab;
    s[12] = b4 * a2;
    s[13] = b4 * ab;
    s[14] = b4 * b2;
    return __dadd_rn(a2, b2);   // never contracted: the reduced feature profiles form |x|^2 the same way
  }
  // returns a*a + b*b
  __device__ __forceinline__ double add(double a, double b) {
    const double a2 = a * a, b2 = b * b, ab = a * b;
    s[0] += a2;
    s[1] += b2;
    s[2] += ab;
    const double a4 = a2 * a2, b4 = b2 * b2, a2b2 = a2 * b2;
    s[3] += a4;
    s[4] = fma(a2, ab, s[4]);
    s[5] += a2b2;
    s[6] = fma(ab, b2, s[6]);
    s[7] += b4;
    s[8] = fma(a4, a2, s[8]);
    s[9] = fma(a4, ab, s[9]);
    s[10] = fma(a4, b2, s[10]);
    s[11] = fma(a2b2, ab, s[11]);
    s[12] = fma(b4, a2, s[12]);
    s[13] = fma(b4, ab, s[13]);
    s[14] = fma(b4, b2, s[14]);
    return __dadd_rn(a2, b2);   // never contracted: the reduced feature profiles form |x|^2 the same way
  }
};

struct Cplx {
  double re, im;
};
__device__ __forceinline__ Cplx cmul(Cplx a, Cplx b) {
  return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__device__ __forceinline__ Cplx cscale(Cplx a, double k) { return {a.re * k, a.im * k}; }
__device__ __forceinline__ Cplx cadd(Cplx a, Cplx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ Cplx csub(Cplx a, Cplx b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ Cplx cconj(Cplx a) { return {a.re, -a.im}; }
__device__ __forceinline__ double cabs(Cplx a) { return hypot(a.re, a.im); }   // np.abs (sqrt(re^2+im^2): -0.2 %, not taken)

// Feature groups one instantiation of a fused kernel computes (amc_fused16.cuh, amc_fusedw.cuh).  The C ABI's feature_mask picks the cheapest compiled
// profile that covers the requested features (amc_api.cu: pick_profile); columns of groups that were not
// computed are written as NaN.  kProfAll is the drop-in default (and the benchmarked kernel).
constexpr int kProfFft = 1;     // feature 1            (spectral max: the three FFT stages)
constexpr int kProfPhase = 2;   // features 2, 3, 5, 9  (atan2, wrapped differences, phase / frequency statistics)
constexpr int kProfAmp = 4;     // features 4, 6, 7, 8  (|x| in FP64, centred amplitude sums)
constexpr int kProfMom = 8;     // features 10..18      (the 15 monomial sums)
constexpr int kProfAll = 15;

// The eleven mixed moments from the (already /N) monomial means (features.py:46-58).
struct Moments {
  Cplx m20, m22, m40, m41, m43, m60, m61, m63;
  double m21, m42, m62;
};
__device__ __forceinline__ Moments moments_from_monomials(const double (&S)[15], double inv_n) {
  const double A2 = S[0] * inv_n, B2 = S[1] * inv_n, AB = S[2] * inv_n;
  const double A4 = S[3] * inv_n, A3B = S[4] * inv_n, A2B2 = S[5] * inv_n, AB3 = S[6] * inv_n,
               B4 = S[7] * inv_n;
  const double A6 = S[8] * inv_n, A5B = S[9] * inv_n, A4B2 = S[10] * inv_n, A3B3 = S[11] * inv_n,
               A2B4 = S[12] * inv_n, AB5 = S[13] * inv_n, B6 = S[14] * inv_n;
  Moments m;
  m.m20 = {A2 - B2, 2.0 * AB};                                   // mean(x^2)
  m.m21 = A2 + B2;                                               // mean(|x|^2)
  m.m22 = cconj(m.m20);                                          // mean(conj(x)^2)
  m.m40 = {A4 - 6.0 * A2B2 + B4, 4.0 * (A3B - AB3)};             // mean(x^4)
  m.m41 = {A4 - B4, 2.0 * (A3B + AB3)};                          // mean(x^3 conj x)
  m.m42 = A4 + 2.0 * A2B2 + B4;                                  // mean(|x|^4)
  m.m43 = cconj(m.m41);                                          // mean(x conj(x)^3)
  m.m60 = {A6 - 15.0 * A4B2 + 15.0 * A2B4 - B6, 6.0 * A5B - 20.0 * A3B3 + 6.0 * AB5};  // mean(x^6)
  m.m61 = {A6 - 5.0 * A4B2 - 5.0 * A2B4 + B6, 4.0 * (A5B - AB5)};                      // mean(x^5 conj x)
  m.m62 = A6 + A4B2 - A2B4 - B6;                                 // Re mean(x^4 conj(x)^2) (features.py:57)
  m.m63 = {A6 + 3.0 * A4B2 + 3.0 * A2B4 + B6, 0.0};              // mean(|x|^6)
  return m;
}

// Everything a frame needs once its sums are complete.
struct FrameSums {
  double mono[15];   // sums of a^p b^q
  double sum_r;      // sum |x|
  double c_abs1;     // sum |r - mean r|
  double c_abs2;     // sum (|cn| - mean |cn|)^2, cn = r / mean r - 1   (careful path only: general kernel, third pass)
  double c2, c4;     // sum (r-mean)^2, ^4
  double ph_m2;      // sum (phi - mean)^2
  double aph_m2;     // sum (|phi| - mean)^2
  double f_m2, f_m4; // sum (f-mean)^2, ^4 over N-1 frequency samples
  double mean_f;     // mean of frequency (for scipy's NaN rule)
  double spec_max;   // max_k |X_k|^2
};

// ------------------------------------------------------------------ fast path -> careful path hand-over
// The fused kernels compute the spectrum and the phases in FLOAT32 and feature 4 from a one-pass formula.  That is
// inside the 1e-6 / 1e-9 classes for every frame except a few well-defined kinds, which finalize_features detects
// from the frame's sums and hands to the general (float64, scaled-FFT, three-pass) kernel by writing this tag into
// column 0 of the frame's row: the library launches `general_features_kernel` in redo mode right after every fused
// kernel; it scans column 0 and recomputes exactly the tagged rows (amc_api.cu: launch_redo).
//   kCheckRange : mean power outside [2^-100, 2^80 * 2048/N]: float32 squares would under/overflow (features 1,2,3,5,9)
//   kCheckPhase : phase / |phase| / frequency spread below kNarrowRad: the float32 phases carry ~1e-7 rad of rounding
//                 noise, i.e. ~3e-8/sigma relative on features 2, 3, 5, 9 (unmodulated carrier, dominant DC line)
//   kCheckAmp   : spread of |r - mean r| below 0.3 % of its mean: sum d^2 - (sum|d|)^2/N cancels (feature 4;
//                 two-level amplitudes at very high SNR, 2-sample frames)
// A quiet NaN with a payload no arithmetic produces; NaN input frames keep the ordinary NaN.
constexpr unsigned long long kRedoTagBits = 0x7ff8b200a3c10001ULL;
// "Did launch `ticket` tag anything?"  Every fused launch carries a unique, increasing ticket; tagging a row raises
// slot ticket % 256 to the ticket (atomicMax: the slots only ever grow, nothing is reset, launches on different
// streams cannot erase each other's mark).  The careful-path launch that follows returns at once while its slot is
// still below its ticket - the case for every batch of ordinary frames - and scans the rows otherwise (a slot shared
// with a later launch can only cause a superfluous scan, never a missed one).
__device__ unsigned long long g_redo_ring[256];
// Programmatic dependent launch: the careful-path kernel is launched while the fused kernel is still running and
// parks here until that grid has completed and its writes are visible (no-ops for ordinary launches).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_primary() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
constexpr int kCheckRange = 1, kCheckPhase = 2, kCheckAmp = 4;
constexpr int kCheckAll = 7;
constexpr int kCheckForce = 8;              // the kernel itself found the frame outside its fast path (amc_fused16x.cuh)
constexpr double kNarrowRad = 0.1;          // rad; below it the careful path takes over

// Writes the 18 features (column k = feature id k+1, features.py:192-211).
// `checks` (kCheck*): which fast-path validity tests apply to the kernel that produced `fs`; 0 = the sums are
// float64-grade already (general kernel).  Returns true when the row was tagged for the careful path instead.
__device__ __noinline__ bool finalize_features(const FrameSums& fs, int n, double* __restrict__ out, int checks,
                                               unsigned long long ticket = 0) {
  const double dn = static_cast<double>(n);
  const double inv_n = 1.0 / dn;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const double pw = (fs.mono[0] + fs.mono[1]) * inv_n;               // mean |x|^2
  if (checks != 0 && n >= 2 && pw == pw) {                           // (NaN frames stay on the NaN rule below)
    bool redo = (checks & kCheckForce) != 0;
    if (checks & kCheckRange) {
      const double hi = 1.2089258196146292e+24 * (2048.0 * inv_n);   // 2^80 * 2048/N: N^2 * sum|x|^2 stays < 2^126
      redo = redo || !(pw >= 7.888609052210118e-31 && pw <= hi);     // 2^-100; also catches +inf and an all-zero frame
    }
    if (checks & kCheckPhase) {
      const double thr = kNarrowRad * kNarrowRad;
      constexpr double k4pi2 = 39.47841760435743;                    // f_m2 is in cycles^2
      redo = redo || fs.ph_m2 < thr * (dn - 1.0) || fs.aph_m2 < thr * (dn - 1.0) ||
             (n >= 3 && fs.f_m2 * k4pi2 < thr * (dn - 2.0));
    }
    if (checks & kCheckAmp) {
      redo = redo || (fs.c2 - fs.c_abs1 * fs.c_abs1 * inv_n) < 1.0e-5 * fs.c2;
    }
    if (redo) {
      if (ticket != 0) atomicMax(&g_redo_ring[ticket & 255], ticket);
      out[0] = __longlong_as_double(static_cast<long long>(kRedoTagBits));
#pragma unroll
      for (int i = 1; i < 18; ++i) out[i] = nan;
      return true;
    }
  }
  out[0] = fs.spec_max / dn;                                         // features.py:68-69
  out[1] = sqrt(fs.aph_m2 / (dn - 1.0));                             // :74
  out[2] = sqrt(fs.ph_m2 / (dn - 1.0));                              // :79
  const double mu = fs.sum_r * inv_n;
  if (checks == 0) {
    out[3] = sqrt(fs.c_abs2 / (dn - 1.0));                           // careful path: np.std's two passes over |cn_amplitude|
  } else {
    // std(|r/mu - 1|, ddof=1) = sqrt((sum d^2 - (sum|d|)^2/N)/(N-1)) / mu, d = r - mu    (:82-85)
    const double v4 = (fs.c2 - fs.c_abs1 * fs.c_abs1 * inv_n) / (dn - 1.0);
    out[3] = sqrt(fmax(v4, 0.0)) / mu;
  }
  out[4] = sqrt(fs.f_m2 / (dn - 2.0));                               // :88-91 (N-1 values, ddof=1)
  out[5] = mu;                                                       // :96
  out[6] = sqrt(fs.sum_r) / dn;                                      // :101
  {                                                                  // :104-107, scipy: m4/m2^2 or NaN
    const double m2 = fs.c2 * inv_n, m4 = fs.c4 * inv_n;
    out[7] = (m2 > 0.0) ? m4 / (m2 * m2) : nan;  // kurtosis is invariant to the r/mu - 1 map
  }
  {                                                                  // :110-113
    const double m2 = fs.f_m2 / (dn - 1.0), m4 = fs.f_m4 / (dn - 1.0);
    const double thr = 2.220446049250313e-16 * fs.mean_f;
    out[8] = (m2 <= thr * thr) ? nan : m4 / (m2 * m2);
  }
  const Moments m = moments_from_monomials(fs.mono, inv_n);
  const Cplx m20sq = cmul(m.m20, m.m20);
  const double abs20sq = m.m20.re * m.m20.re + m.m20.im * m.m20.im;
  out[9] = cabs(m.m20);                                              // :116-118
  out[10] = fabs(m.m21);                                             // :121-123
  out[11] = cabs(csub(m.m40, cscale(m20sq, 3.0)));                   // :126-129
  out[12] = cabs(csub(m.m41, cscale(m.m20, 3.0 * m.m21)));           // :132-135
  {                                                                  // :138-141 (np.abs(m20)**2)
    const double a20 = cabs(m.m20);
    out[13] = fabs(m.m42 - a20 * a20 - 2.0 * m.m21 * m.m21);
  }
  {                                                                  // :144-147  (+3*m20^3, sic)
    Cplx c = csub(m.m60, cscale(cmul(m.m20, m.m40), 15.0));
    c = cadd(c, cscale(cmul(m20sq, m.m20), 3.0));
    out[14] = cabs(c);
  }
  {                                                                  // :150-155
    Cplx c = csub(m.m61, cscale(m.m40, 5.0 * m.m21));
    c = csub(c, cscale(cmul(m.m20, m.m41), 10.0));
    c = cadd(c, cscale(m20sq, 30.0 * m.m21));
    out[15] = cabs(c);
  }
  {                                                                  // :158-170 (real-only m62)
    Cplx c = {m.m62, 0.0};
    c = csub(c, cscale(m.m20, 6.0 * m.m42));
    c = csub(c, cscale(m.m41, 8.0 * m.m21));
    c = csub(c, cmul(m.m22, m.m40));
    c = cadd(c, cscale(cmul(m20sq, m.m22), 6.0));
    c = cadd(c, cscale(m.m20, 24.0 * m.m21 * m.m21));
    out[16] = cabs(c);
  }
  {                                                                  // :173-185
    Cplx c = m.m63;
    c.re -= 9.0 * m.m21 * m.m42;
    c.re += 12.0 * m.m21 * m.m21 * m.m21;
    c = csub(c, cscale(cmul(m.m20, m.m43), 3.0));
    c = csub(c, cscale(cmul(m.m22, m.m41), 3.0));
    c.re += 18.0 * m.m21 * abs20sq;
    out[17] = cabs(c);
  }
  // A NaN anywhere in the frame reaches all 18 features of the reference (through the FFT, the means and the
  // sums).  The GPU's min/max instructions drop NaNs (spectral max, tie detection), so the rule is applied here:
  // sum a^2 + sum b^2 is NaN if and only if some sample holds a NaN (infinities and overflow give +inf).
  if (isnan(fs.mono[0] + fs.mono[1])) {
#pragma unroll
    for (int i = 0; i < 18; ++i) out[i] = nan;
  }
  return false;
}

// columns of the feature groups a reduced profile did not compute
template <int PROF>
__device__ __forceinline__ void blank_skipped_groups(double* __restrict__ row) {
  if constexpr (PROF != kProfAll) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (!(PROF & kProfFft)) row[0] = nan;
    if (!(PROF & kProfPhase)) row[1] = row[2] = row[4] = row[8] = nan;
    if (!(PROF & kProfAmp)) row[3] = row[5] = row[6] = row[7] = nan;
    if (!(PROF & kProfMom)) {
#pragma unroll
      for (int i = 9; i < 18; ++i) row[i] = nan;
    }
  }
}

}  // namespace amc
