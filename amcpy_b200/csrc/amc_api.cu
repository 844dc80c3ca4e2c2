// C ABI of amcpy_b200 (see include/amcpy_b200.h): argument checking, kernel selection, launches,
// and the chunked host-buffer pipeline.  No exceptions cross the boundary; no CPU compute path.
#include "../../include/amcpy_b200.h"

#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <mutex>
#include <cstring>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include <atomic>
#include <condition_variable>

#include "amc_fused.cuh"
#include "amc_fused16.cuh"
#include "amc_fusedw.cuh"
#ifdef AMC_F16_V2
#include "experiments/amc_fused16x.cuh"   // N = 2048 at four CTAs per SM: correct, 13 % slower (profiles/r2_experiments.txt)
#endif
#ifdef AMC_EXPERIMENTS
#include "experiments/amc_fusedws.cuh"
#endif
#include "amc_large.cuh"
#include "amc_general.cuh"
#include "amc_generate.cuh"

namespace {

thread_local std::string t_err;
thread_local int64_t t_launches = 0;
thread_local bool t_allow_pdl = true;   // set per call by amc_extract_batch (see there)
std::atomic<unsigned long long> g_ticket{1};   // unique, increasing id of every fused launch (amc_device.cuh: g_redo_ring)

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_err = buf;
  return code;
}

#define AMC_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return fail(AMC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

constexpr int kMaxDevices = 64;
std::mutex g_mu;

// Per-device library state.  Tables are built on an internal stream and published through an event: callers'
// streams wait for the event on the device (cudaStreamWaitEvent) - no host-side synchronisation, nothing is ever
// enqueued on the caller's stream except its own kernels - until the event is known to have completed.
struct DeviceState {
  int sms = 0;
  bool tw_launched = false;
  bool tw_done = false;
  cudaStream_t init_stream = nullptr;
  cudaEvent_t tw_event = nullptr;
  cudaMemPool_t ws_pool = nullptr;   // the library's own stream-ordered pool (scratch of the long-frame / general kernels)
};
DeviceState g_dev[kMaxDevices];

struct DevInfo {
  int dev;
  int sms;
};

// Restores the caller's current device on every exit path of the host entries.
struct DeviceGuard {
  int prev = -1;
  bool armed = false;
  ~DeviceGuard() {
    if (armed) cudaSetDevice(prev);
  }
};

int device_info(DevInfo* di) {
  int dev = 0;
  AMC_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail(AMC_ERR_INVALID_ARG, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_dev[dev].sms == 0) {
    int sms = 0;
    AMC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    g_dev[dev].sms = sms;
  }
  di->dev = dev;
  di->sms = g_dev[dev].sms;
  return AMC_OK;
}

// Stream-ordered scratch from the library's OWN memory pool (never the process-wide default pool, whose release
// threshold belongs to the application): memory is kept across calls, so a steady stream of launches allocates once.
int ws_alloc(int dev, void** ptr, size_t bytes, cudaStream_t stream) {
  cudaMemPool_t pool;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceState& d = g_dev[dev];
    if (!d.ws_pool) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      AMC_CUDA(cudaMemPoolCreate(&d.ws_pool, &props));
      unsigned long long keep = ~0ull;
      AMC_CUDA(cudaMemPoolSetAttribute(d.ws_pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    pool = d.ws_pool;
  }
  AMC_CUDA(cudaMallocFromPoolAsync(ptr, bytes, pool, stream));
  return AMC_OK;
}

// internal stream of the current device (g_mu held)
int init_stream_locked(DeviceState& d) {
  if (!d.init_stream) AMC_CUDA(cudaStreamCreateWithFlags(&d.init_stream, cudaStreamNonBlocking));
  return AMC_OK;
}

// `stream` will see the tables: either they are known to be complete, or the stream waits for their event.
int wait_for_event(cudaStream_t stream, cudaEvent_t ev, bool* done_flag) {
  const cudaError_t q = cudaEventQuery(ev);
  if (q == cudaSuccess) {
    std::lock_guard<std::mutex> lk(g_mu);
    *done_flag = true;
    return AMC_OK;
  }
  if (q != cudaErrorNotReady) return fail(AMC_ERR_CUDA, "table initialisation failed: %s", cudaGetErrorString(q));
  AMC_CUDA(cudaStreamWaitEvent(stream, ev, 0));
  return AMC_OK;
}

int ensure_twiddles(int dev, cudaStream_t stream) {
  cudaEvent_t ev;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceState& d = g_dev[dev];
    if (d.tw_done) return AMC_OK;
    if (!d.tw_launched) {
      int rc = init_stream_locked(d);
      if (rc != AMC_OK) return rc;
      if (!d.tw_event) AMC_CUDA(cudaEventCreateWithFlags(&d.tw_event, cudaEventDisableTiming));
      amc::init_twiddle16_kernel<<<26, 256, 0, d.init_stream>>>();
      amc::init_twiddle16dif_kernel<<<29, 256, 0, d.init_stream>>>();
      amc::init_twiddle8_kernel<<<22, 256, 0, d.init_stream>>>();
      amc::init_twiddle_large_kernel<<<64, 256, 0, d.init_stream>>>();
      t_launches += 4;
      AMC_CUDA(cudaGetLastError());
      AMC_CUDA(cudaEventRecord(d.tw_event, d.init_stream));
      d.tw_launched = true;
    }
    ev = d.tw_event;
  }
  return wait_for_event(stream, ev, &g_dev[dev].tw_done);
}

bool is_pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }

// ---------------------------------------------------------------------------------------------
// Bluestein tables for frame sizes that are not powers of two (general kernel, fft_mode 2): built once per
// (device, N) in float64 on the host - chirp c_n = exp(-i pi n^2 / N) and the length-M FFT of the wrapped
// conjugate chirp, stored in the bit-reversed order the kernel's forward pass produces, scaled by 1/M -
// and kept on the device.  The reference's np.fft.fft accepts any length (features.py:68).
// ---------------------------------------------------------------------------------------------
struct BluesteinTables {
  float2* chirp = nullptr;   // N
  float2* bfft = nullptr;    // M
  int m = 0;
  cudaEvent_t ready = nullptr;
  bool done = false;
};
std::map<std::pair<int, int64_t>, BluesteinTables> g_bluestein;
constexpr size_t kBluesteinMaxEntries = 64;          // beyond that many distinct sizes: direct DFT
constexpr int64_t kBluesteinMinN = 128;              // below: the float64 direct DFT is cheap
constexpr int64_t kBluesteinSmemM = 16384;           // up to here the M float2 live in shared memory
constexpr int64_t kBluesteinMaxM = 1 << 21;          // beyond: AMC_ERR_UNSUPPORTED (frames of > 1 Mi samples)

void host_fft_pow2(std::vector<std::complex<double>>& v) {
  const size_t m = v.size();
  for (size_t i = 1, j = 0; i < m; ++i) {            // bit-reversal permutation
    size_t bit = m >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(v[i], v[j]);
  }
  const double pi = std::acos(-1.0);
  for (size_t len = 2; len <= m; len <<= 1) {
    for (size_t k = 0; k < len / 2; ++k) {
      const std::complex<double> w = std::polar(1.0, -2.0 * pi * static_cast<double>(k) / static_cast<double>(len));
      for (size_t i = k; i < m; i += len) {
        const std::complex<double> u = v[i], t = w * v[i + len / 2];
        v[i] = u + t;
        v[i + len / 2] = u - t;
      }
    }
  }
}

int64_t bluestein_m(int64_t n) {
  if (n < kBluesteinMinN) return 0;
  int64_t m = 1;
  while (m < 2 * n - 1) m <<= 1;
  return m > kBluesteinMaxM ? 0 : m;
}

// returns AMC_OK with tab->m == 0 when this size should use the direct DFT instead.  The tables are uploaded on the
// library's internal stream; `stream` waits for them on the device (no host synchronisation).
int ensure_bluestein(int dev, int64_t n, cudaStream_t stream, BluesteinTables* tab) {
  *tab = BluesteinTables();
  const int64_t m = bluestein_m(n);
  if (m == 0) return AMC_OK;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_bluestein.find({dev, n});
    if (it != g_bluestein.end()) {
      *tab = it->second;
    } else if (g_bluestein.size() >= kBluesteinMaxEntries) {
      return AMC_OK;
    }
  }
  if (tab->m == 0) {
    // host-side construction outside the lock (other threads / devices keep running); a racing thread building the
    // same table loses at insertion time and frees its copy
    const double pi = std::acos(-1.0);
    std::vector<std::complex<double>> b;
    std::vector<float2> chirp, bf;
    try {
      b.resize(static_cast<size_t>(m));
      chirp.resize(static_cast<size_t>(n));
      bf.resize(static_cast<size_t>(m));
    } catch (...) {
      return AMC_OK;   // no host memory for the tables: the direct DFT needs none
    }
    for (int64_t k = 0; k < n; ++k) {
      const double ang = pi * static_cast<double>((k * k) % (2 * n)) / static_cast<double>(n);   // k^2 mod 2N: exact
      const double cs = std::cos(ang), sn = std::sin(ang);
      chirp[static_cast<size_t>(k)] = make_float2(static_cast<float>(cs), static_cast<float>(-sn));
      b[static_cast<size_t>(k)] = {cs, sn};
      if (k > 0) b[static_cast<size_t>(m - k)] = {cs, sn};
    }
    host_fft_pow2(b);
    int bits = 0;
    while ((int64_t{1} << bits) < m) ++bits;
    for (int64_t j = 0; j < m; ++j) {
      int64_t r = 0;
      for (int q = 0; q < bits; ++q) r |= ((j >> q) & 1) << (bits - 1 - q);
      const std::complex<double> v = b[static_cast<size_t>(r)] / static_cast<double>(m);
      bf[static_cast<size_t>(j)] = make_float2(static_cast<float>(v.real()), static_cast<float>(v.imag()));
    }
    BluesteinTables t;
    t.m = static_cast<int>(m);
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_bluestein.find({dev, n});
    if (it != g_bluestein.end()) {
      *tab = it->second;                             // another thread was faster
    } else {
      DeviceState& d = g_dev[dev];
      int rc = init_stream_locked(d);
      if (rc != AMC_OK) return rc;
      cudaError_t e = cudaMalloc(&t.chirp, static_cast<size_t>(n) * sizeof(float2));
      if (e == cudaSuccess) e = cudaMalloc(&t.bfft, static_cast<size_t>(m) * sizeof(float2));
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t.ready, cudaEventDisableTiming);
      // (pageable sources: the runtime has staged the bytes by the time cudaMemcpyAsync returns)
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(t.chirp, chirp.data(), static_cast<size_t>(n) * sizeof(float2), cudaMemcpyHostToDevice, d.init_stream);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(t.bfft, bf.data(), static_cast<size_t>(m) * sizeof(float2), cudaMemcpyHostToDevice, d.init_stream);
      if (e == cudaSuccess) e = cudaEventRecord(t.ready, d.init_stream);
      bool cached = false;
      if (e == cudaSuccess) {
        try {
          g_bluestein[{dev, n}] = t;
          cached = true;
        } catch (...) {
        }
      }
      if (!cached) {                                 // nothing may leak: wait for the copies, then free everything
        cudaStreamSynchronize(d.init_stream);
        if (t.ready) cudaEventDestroy(t.ready);
        if (t.chirp) cudaFree(t.chirp);
        if (t.bfft) cudaFree(t.bfft);
        if (e != cudaSuccess) return fail(AMC_ERR_CUDA, "Bluestein table upload failed: %s", cudaGetErrorString(e));
        return AMC_OK;                               // not cacheable (out of host memory): direct DFT for this call
      }
      *tab = t;
    }
  }
  if (!tab->done) {
    bool done = false;
    const int rc = wait_for_event(stream, tab->ready, &done);
    if (rc != AMC_OK) return rc;
    if (done) {
      std::lock_guard<std::mutex> lk(g_mu);
      auto it = g_bluestein.find({dev, n});
      if (it != g_bluestein.end()) it->second.done = true;
    }
  }
  return AMC_OK;
}

// AMCPY_B200_NO_PDL=1: ordinary stream-ordered launches (A/B and debugging)
bool use_pdl() {
  static const bool on = [] {
    const char* env = std::getenv("AMCPY_B200_NO_PDL");
    return !(env && *env && std::atoi(env) != 0);
  }();
  return on;
}

// Launch with programmatic stream serialisation allowed: the grid may be scheduled while its predecessor in the
// stream is still running (launch latency, CTA scheduling and the kernel prologue overlap with the predecessor's
// tail).  Every kernel launched this way executes griddepcontrol.wait (amc_device.cuh: pdl_wait_primary) before its
// first global-memory access, which blocks until the predecessor has completed and flushed - ordinary stream
// semantics for everything the kernel reads or writes.
template <typename... Params, typename... Args>
cudaError_t launch_pdl(void (*kern)(Params...), int grid, int block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(static_cast<unsigned>(block));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (use_pdl() && t_allow_pdl) ? 1 : 0;
  if (cfg.numAttrs) {   // a stream being captured into a CUDA graph gets ordinary (fully serialised) kernel nodes
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) cudaGetLastError();
    if (cap != cudaStreamCaptureStatusNone) cfg.numAttrs = 0;
  }
  return cudaLaunchKernelEx(&cfg, kern, static_cast<Params>(args)...);
}

#ifdef AMC_EXPERIMENTS
template <int N, typename CT>
int launch_fused(const void* iq, int64_t n_frames, int64_t frame_stride, double* out, int64_t out_stride,
                 int sms, cudaStream_t stream) {
  using Cfg = amc::FusedCfg<N, CT>;
  auto kern = amc::fused_features_kernel<N, CT>;
  static thread_local int blocks_per_sm[kMaxDevices] = {};
  int dev = 0;
  AMC_CUDA(cudaGetDevice(&dev));
  if (blocks_per_sm[dev] == 0) {
    AMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    int occ = 0;
    AMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::CTA, Cfg::SMEM_BYTES));
    if (occ < 1) return fail(AMC_ERR_CUDA, "fused kernel N=%d does not fit on this device", N);
    blocks_per_sm[dev] = occ;
  }
  const int64_t want = (n_frames + Cfg::G - 1) / Cfg::G;
  const int64_t cap = static_cast<int64_t>(sms) * blocks_per_sm[dev];
  const int grid = static_cast<int>(want < cap ? want : cap);
  kern<<<grid, Cfg::CTA, Cfg::SMEM_BYTES, stream>>>(static_cast<const CT*>(iq), n_frames, frame_stride, out,
                                                   out_stride);
  ++t_launches;
  AMC_CUDA(cudaGetLastError());
  return AMC_OK;
}

#endif  // AMC_EXPERIMENTS

// feature_mask -> the cheapest compiled profile of the 16-samples-per-thread kernel that covers it
// (amc_fused16.cuh: kProf*).  Compiled: moments only, amplitude + moments, everything but the FFT, all.
int pick_profile(uint32_t feature_mask) {
  int need = 0;
  if (feature_mask & 0x00001u) need |= amc::kProfFft;
  if (feature_mask & 0x00116u) need |= amc::kProfPhase;   // features 2, 3, 5, 9
  if (feature_mask & 0x000e8u) need |= amc::kProfAmp;     // features 4, 6, 7, 8
  if (feature_mask & 0x3fe00u) need |= amc::kProfMom;     // features 10..18
  constexpr int compiled[] = {amc::kProfMom, amc::kProfAmp | amc::kProfMom,
                              amc::kProfPhase | amc::kProfAmp | amc::kProfMom, amc::kProfAll};
  for (int p : compiled)
    if ((p & need) == need) return p;
  return amc::kProfAll;
}

template <int N, typename CT, int PROF = amc::kProfAll>
int launch_fused16(const void* iq, int64_t n_frames, int64_t frame_stride, double* out, int64_t out_stride,
                   int sms, cudaStream_t stream, unsigned long long ticket) {
  using Cfg = amc::Fused16Cfg<N, CT>;
  auto kern = amc::fused16_features_kernel<N, CT, PROF>;
  static thread_local int blocks_per_sm[kMaxDevices] = {};
  int dev = 0;
  AMC_CUDA(cudaGetDevice(&dev));
  if (blocks_per_sm[dev] == 0) {
    AMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    int occ = 0;
    AMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::CTA, Cfg::SMEM_BYTES));
    if (occ < 1) return fail(AMC_ERR_CUDA, "fused16 kernel N=%d does not fit on this device", N);
    blocks_per_sm[dev] = occ;
  }
  const int64_t want = (n_frames + Cfg::G - 1) / Cfg::G;
  const int64_t cap = static_cast<int64_t>(sms) * blocks_per_sm[dev];
  const int grid = static_cast<int>(want < cap ? want : cap);
  AMC_CUDA(launch_pdl(kern, grid, Cfg::CTA, Cfg::SMEM_BYTES, stream, static_cast<const CT*>(iq), n_frames, frame_stride,
                      out, out_stride, ticket));
  ++t_launches;
  return AMC_OK;
}

#ifdef AMC_F16_V2
// N = 2048, four CTAs per SM (experiments/amc_fused16x.cuh)
template <typename CT>
int launch_fused16x(const void* iq, int64_t n_frames, int64_t frame_stride, double* out, int64_t out_stride,
                    int sms, cudaStream_t stream, unsigned long long ticket) {
  using Cfg = amc::Fused16xCfg<CT>;
  auto kern = amc::fused16x_features_kernel<CT>;
  static thread_local int blocks_per_sm[kMaxDevices] = {};
  int dev = 0;
  AMC_CUDA(cudaGetDevice(&dev));
  if (blocks_per_sm[dev] == 0) {
    AMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    int occ = 0;
    AMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::CTA, Cfg::SMEM_BYTES));
    if (occ < 1) return fail(AMC_ERR_CUDA, "fused16x kernel does not fit on this device");
    blocks_per_sm[dev] = occ;
  }
  const int64_t cap = static_cast<int64_t>(sms) * blocks_per_sm[dev];
  const int grid = static_cast<int>(n_frames < cap ? n_frames : cap);
  AMC_CUDA(launch_pdl(kern, grid, Cfg::CTA, Cfg::SMEM_BYTES, stream, static_cast<const CT*>(iq), n_frames, frame_stride,
                      out, out_stride, ticket));
  ++t_launches;
  return AMC_OK;
}
#endif  // AMC_F16_V2

#ifdef AMC_EXPERIMENTS
template <int N, typename CT>
int launch_fusedws(const void* iq, int64_t n_frames, int64_t frame_stride, double* out, int64_t out_stride,
                   int sms, cudaStream_t stream) {
  using Cfg = amc::FusedWsCfg<N, CT>;
  auto kern = amc::fusedws_features_kernel<N, CT>;
  static thread_local int blocks_per_sm[kMaxDevices] = {};
  int dev = 0;
  AMC_CUDA(cudaGetDevice(&dev));
  if (blocks_per_sm[dev] == 0) {
    AMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    int occ = 0;
    AMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::CTA, Cfg::SMEM_BYTES));
    if (occ < 1) return fail(AMC_ERR_CUDA, "warp-specialised kernel N=%d does not fit on this device", N);
    blocks_per_sm[dev] = occ;
  }
  const int64_t cap = static_cast<int64_t>(sms) * blocks_per_sm[dev];
  const int grid = static_cast<int>(n_frames < cap ? n_frames : cap);
  kern<<<grid, Cfg::CTA, Cfg::SMEM_BYTES, stream>>>(static_cast<const CT*>(iq), n_frames, frame_stride, out,
                                                   out_stride);
  ++t_launches;
  AMC_CUDA(cudaGetLastError());
  return AMC_OK;
}
#endif  // AMC_EXPERIMENTS

template <int N, typename CT, int PROF = amc::kProfAll>
int launch_fusedw(const void* iq, int64_t n_frames, int64_t frame_stride, double* out, int64_t out_stride,
                  int sms, cudaStream_t stream, unsigned long long ticket) {
  using Cfg = amc::FusedWCfg<N, CT>;
  auto kern = amc::fusedw_features_kernel<N, CT, PROF>;
  static thread_local int blocks_per_sm[kMaxDevices] = {};
  int dev = 0;
  AMC_CUDA(cudaGetDevice(&dev));
  if (blocks_per_sm[dev] == 0) {
    AMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    int occ = 0;
    AMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::CTA, Cfg::SMEM_BYTES));
    if (occ < 1) return fail(AMC_ERR_CUDA, "warp-per-frame kernel N=%d does not fit on this device", N);
    blocks_per_sm[dev] = occ;
  }
  const int64_t want = (n_frames + Cfg::G - 1) / Cfg::G;
  const int64_t cap = static_cast<int64_t>(sms) * blocks_per_sm[dev];
  const int grid = static_cast<int>(want < cap ? want : cap);
  AMC_CUDA(launch_pdl(kern, grid, Cfg::CTA, Cfg::SMEM_BYTES, stream, static_cast<const CT*>(iq), n_frames, frame_stride,
                      out, out_stride, ticket));
  ++t_launches;
  return AMC_OK;
}

template <int N, typename CT>
int launch_large(const void* iq, int64_t n_frames, int64_t frame_stride, double* out, int64_t out_stride, int sms,
                 cudaStream_t stream, unsigned long long ticket) {
  using Cfg = amc::LargeCfg<N>;
  auto kern = amc::large_features_kernel<N, CT, true>;
  auto kern_plain = amc::large_features_kernel<N, CT, false>;
  static thread_local bool attr_set[kMaxDevices] = {};
  int dev = 0;
  AMC_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev]) {   // once per (thread, device), not on every launch
    AMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    AMC_CUDA(cudaFuncSetAttribute(kern_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set[dev] = true;
  }
  const int64_t cap = static_cast<int64_t>(sms) * Cfg::MIN_BLOCKS;
  const int grid = static_cast<int>(n_frames < cap ? n_frames : cap);
  // |x| scratch: N doubles per CTA from the stream-ordered pool (no synchronisation); without it the kernel recomputes
  double* ws = nullptr;
  if (ws_alloc(dev, reinterpret_cast<void**>(&ws), static_cast<size_t>(grid) * N * sizeof(double), stream) != AMC_OK) {
    cudaGetLastError();
    ws = nullptr;
  }
  cudaError_t e;
  if (ws)
    e = launch_pdl(kern, grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream, static_cast<const CT*>(iq), n_frames, frame_stride,
                   out, out_stride, ticket, ws);
  else
    e = launch_pdl(kern_plain, grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream, static_cast<const CT*>(iq), n_frames,
                   frame_stride, out, out_stride, ticket, static_cast<double*>(nullptr));
  ++t_launches;
  if (ws) {
    const cudaError_t e2 = cudaFreeAsync(ws, stream);
    if (e == cudaSuccess) e = e2;
  }
  if (e != cudaSuccess) return fail(AMC_ERR_CUDA, "long-frame kernel launch failed: %s", cudaGetErrorString(e));
  return AMC_OK;
}

template <int N, typename CT>
int launch_fused16_profile(int prof, const void* iq, int64_t n_frames, int64_t frame_stride, double* out,
                           int64_t out_stride, int sms, cudaStream_t stream, unsigned long long ticket) {
  switch (prof) {
    case amc::kProfMom:
      return launch_fused16<N, CT, amc::kProfMom>(iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
    case amc::kProfAmp | amc::kProfMom:
      return launch_fused16<N, CT, amc::kProfAmp | amc::kProfMom>(iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
    case amc::kProfPhase | amc::kProfAmp | amc::kProfMom:
      return launch_fused16<N, CT, amc::kProfPhase | amc::kProfAmp | amc::kProfMom>(iq, n_frames, frame_stride, out,
                                                                                     out_stride, sms, stream, ticket);
    default:
      return launch_fused16<N, CT>(iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
  }
}

// *used_prof receives the feature profile the launched kernel computed (kProfAll unless a reduced profile exists)
template <typename CT>
int dispatch_fused(int64_t n, const void* iq, int64_t n_frames, int64_t frame_stride, double* out,
                   int64_t out_stride, int sms, cudaStream_t stream, int flags, int prof, int* used_prof,
                   unsigned long long ticket) {
  *used_prof = amc::kProfAll;
#ifdef AMC_EXPERIMENTS
  if (flags & (AMC_FLAG_FUSED_WS | AMC_FLAG_FUSED_SPT8)) *used_prof = -1;   // experiment kernels carry no ticket
  if ((flags & AMC_FLAG_FUSED_WS) && n == 2048)
    return launch_fusedws<2048, CT>(iq, n_frames, frame_stride, out, out_stride, sms, stream);
  if (flags & AMC_FLAG_FUSED_SPT8) {
    switch (n) {
      case 256: return launch_fused<256, CT>(iq, n_frames, frame_stride, out, out_stride, sms, stream);
      case 512: return launch_fused<512, CT>(iq, n_frames, frame_stride, out, out_stride, sms, stream);
      case 1024: return launch_fused<1024, CT>(iq, n_frames, frame_stride, out, out_stride, sms, stream);
      case 2048: return launch_fused<2048, CT>(iq, n_frames, frame_stride, out, out_stride, sms, stream);
      case 4096: return launch_fused<4096, CT>(iq, n_frames, frame_stride, out, out_stride, sms, stream);
      default: break;
    }
  }
#else
  (void)flags;
#endif
  *used_prof = amc::kProfAll;
  if (prof != amc::kProfAll) {   // reduced feature profiles: warp-per-frame and 16-samples-per-thread kernels
    constexpr int kAM = amc::kProfAmp | amc::kProfMom, kPAM = amc::kProfPhase | kAM;
    if (n == 256) {
      *used_prof = prof;
      if (prof == amc::kProfMom)
        return launch_fusedw<256, CT, amc::kProfMom>(iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
      if (prof == kAM) return launch_fusedw<256, CT, kAM>(iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
      if (prof == kPAM) return launch_fusedw<256, CT, kPAM>(iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
      *used_prof = amc::kProfAll;
    }
    if (n == 512 || n == 1024 || n == 2048 || n == 4096) *used_prof = prof;
    switch (n) {
      case 512: return launch_fused16_profile<512, CT>(prof, iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
      case 1024: return launch_fused16_profile<1024, CT>(prof, iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
      case 2048: return launch_fused16_profile<2048, CT>(prof, iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
      case 4096: return launch_fused16_profile<4096, CT>(prof, iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
      default: break;
    }
  }
#ifdef AMC_F16_V2
  if (n == 2048) return launch_fused16x<CT>(iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
#endif
  switch (n) {
    case 256: return launch_fusedw<256, CT>(iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
    case 512: return launch_fused16<512, CT>(iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
    case 1024: return launch_fused16<1024, CT>(iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
    case 2048: return launch_fused16<2048, CT>(iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
    case 4096: return launch_fused16<4096, CT>(iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
    case 8192: return launch_large<8192, CT>(iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
    case 16384: return launch_large<16384, CT>(iq, n_frames, frame_stride, out, out_stride, sms, stream, ticket);
    default: return fail(AMC_ERR_UNSUPPORTED, "no fused kernel for frame_size %lld", static_cast<long long>(n));
  }
}

bool fused_size(int64_t n) {
  return n == 256 || n == 512 || n == 1024 || n == 2048 || n == 4096 || n == 8192 || n == 16384;
}

constexpr int64_t kGeneralPow2Smem = 16384;  // up to here the N float2 of FFT scratch live in shared memory
constexpr int64_t kGeneralPow2Max = 1 << 20; // beyond: AMC_ERR_UNSUPPORTED
constexpr int64_t kGeneralDftMax = 12288;    // direct DFT: N double2 of twiddles must fit in shared memory
constexpr size_t kFftWorkspaceMax = size_t{1} << 30;   // global FFT workspace of one launch (long frames)
constexpr size_t kGeneralSmemMax = 200 * 1024;

template <typename CT>
int general_kernel_attr() {
  static thread_local bool attr_set[kMaxDevices] = {};
  int dev = 0;
  AMC_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev]) {
    AMC_CUDA(cudaFuncSetAttribute(amc::general_features_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(kGeneralSmemMax)));
    attr_set[dev] = true;
  }
  return AMC_OK;
}

// phase / amplitude cache of 2 N doubles between the passes: aliases the FFT buffer (modes 1, 2) or follows the
// DFT twiddle table (mode 0); skipped when it does not fit.  Returns the cache offset (-1: none), grows *dyn.
int place_cache(int64_t n, int fft_mode, size_t* dyn) {
  const size_t cache = static_cast<size_t>(n) * 2 * sizeof(double);
  const size_t off = fft_mode == 0 ? *dyn : 0;
  if (off + cache > kGeneralSmemMax) return -1;
  if (off + cache > *dyn) *dyn = off + cache;
  return static_cast<int>(off);
}

template <typename CT>
int launch_general(const void* iq, int64_t n_frames, int64_t n, int64_t frame_stride, int64_t sample_stride,
                   double* out, int64_t out_stride, int sms, cudaStream_t stream, int flags) {
  int fft_mode;
  size_t dyn;
  size_t ws_elems = 0;                       // > 0: transform buffer of that many float2 per CTA in global memory
  BluesteinTables bl;
  if (is_pow2(n) && n >= 2) {
    if (n > kGeneralPow2Max)
      return fail(AMC_ERR_UNSUPPORTED, "power-of-two frame_size %lld > %lld not supported", (long long)n,
                  (long long)kGeneralPow2Max);
    fft_mode = 1;
    dyn = static_cast<size_t>(n) * sizeof(float2);
    if (n > kGeneralPow2Smem) {
      ws_elems = static_cast<size_t>(n);
      dyn = 0;
    }
  } else {
    int dev = 0;
    AMC_CUDA(cudaGetDevice(&dev));
    const bool direct = (flags & AMC_FLAG_DIRECT_DFT) != 0;
    if (direct && n > kGeneralDftMax)
      return fail(AMC_ERR_UNSUPPORTED, "direct DFT: frame_size %lld > %lld not supported", (long long)n,
                  (long long)kGeneralDftMax);
    if (!direct && n >= kBluesteinMinN && bluestein_m(n) == 0)
      return fail(AMC_ERR_UNSUPPORTED, "non-power-of-two frame_size %lld too large (Bluestein length > %lld)",
                  (long long)n, (long long)kBluesteinMaxM);
    const int rc = direct ? AMC_OK : ensure_bluestein(dev, n, stream, &bl);
    if (rc != AMC_OK) return rc;
    if (bl.m > 0) {
      fft_mode = 2;
      dyn = static_cast<size_t>(bl.m) * sizeof(float2);
      if (bl.m > kBluesteinSmemM) {
        ws_elems = static_cast<size_t>(bl.m);
        dyn = 0;
      }
    } else {
      if (n > kGeneralDftMax)
        return fail(AMC_ERR_UNSUPPORTED, "frame_size %lld: no Bluestein tables and too long for the direct DFT",
                    (long long)n);
      fft_mode = 0;
      dyn = static_cast<size_t>(n) * sizeof(double2);
    }
  }
  const int cache_off = place_cache(n, fft_mode, &dyn);
  int rc = general_kernel_attr<CT>();
  if (rc != AMC_OK) return rc;
  int64_t cap = static_cast<int64_t>(sms) * 4;
  float2* ws = nullptr;
  if (ws_elems > 0) {
    // stream-ordered workspace (the library's own memory pool, allocation and free on the caller's stream: no synchronisation, safe
    // between concurrent calls); the grid is trimmed so that it stays below kFftWorkspaceMax
    const int64_t fit = static_cast<int64_t>(kFftWorkspaceMax / (ws_elems * sizeof(float2)));
    if (cap > fit) cap = fit < 1 ? 1 : fit;
    const int64_t ctas = n_frames < cap ? n_frames : cap;
    int dev_ws = 0;
    AMC_CUDA(cudaGetDevice(&dev_ws));
    rc = ws_alloc(dev_ws, reinterpret_cast<void**>(&ws), static_cast<size_t>(ctas) * ws_elems * sizeof(float2), stream);
    if (rc != AMC_OK) return rc;
  }
  const int grid = static_cast<int>(n_frames < cap ? n_frames : cap);
  amc::general_features_kernel<CT><<<grid, amc::kGenThreads, dyn, stream>>>(
      static_cast<const CT*>(iq), n_frames, static_cast<int>(n), frame_stride, sample_stride, out, out_stride, fft_mode,
      bl.chirp, bl.bfft, bl.m, cache_off, 0, 0ull, ws);
  ++t_launches;
  cudaError_t e = cudaGetLastError();
  if (ws) {
    const cudaError_t e2 = cudaFreeAsync(ws, stream);
    if (e == cudaSuccess) e = e2;
  }
  if (e != cudaSuccess) return fail(AMC_ERR_CUDA, "general kernel launch failed: %s", cudaGetErrorString(e));
  return AMC_OK;
}

// The careful-path pass behind every fused launch: rows the fused kernel tagged (amc_device.cuh: kRedoTagBits) are
// recomputed by the general kernel; when nothing is tagged - every frame of ordinary data - it only scans column 0
// of the output (one 8-byte load per frame).  Fused sizes are powers of two <= 16384: float32 radix-2 FFT.
template <typename CT>
int launch_redo(const void* iq, int64_t n_frames, int64_t n, int64_t frame_stride, double* out, int64_t out_stride,
                int sms, cudaStream_t stream, unsigned long long ticket) {
  size_t dyn = static_cast<size_t>(n) * sizeof(float2);
  const int cache_off = place_cache(n, 1, &dyn);
  int rc = general_kernel_attr<CT>();
  if (rc != AMC_OK) return rc;
  const int64_t want = (n_frames + 63) / 64;                 // >= 64 rows per CTA: one scan step of 256 threads
  // four CTAs per SM: when frames ARE tagged the recomputation runs at the general kernel's full occupancy; when none
  // is (the normal case) the grid exits at once - 148 vs 592 CTAs: 0.5996 vs 0.6003 ms per step, within the noise
  const int64_t cap = static_cast<int64_t>(sms) * 4;
  const int grid = static_cast<int>(want < cap ? want : cap);
  AMC_CUDA(launch_pdl(amc::general_features_kernel<CT>, grid, amc::kGenThreads, dyn, stream, static_cast<const CT*>(iq),
                      n_frames, static_cast<int>(n), frame_stride, 1, out, out_stride, 1, nullptr, nullptr, 0, cache_off, 1,
                      ticket, nullptr));
  ++t_launches;
  return AMC_OK;
}

#ifdef AMC_EXPERIMENTS
constexpr int kKnownFlags = AMC_FLAG_FORCE_GENERAL | AMC_FLAG_DIRECT_DFT | AMC_FLAG_FUSED_SPT8 | AMC_FLAG_FUSED_WS;
#else
constexpr int kKnownFlags = AMC_FLAG_FORCE_GENERAL | AMC_FLAG_DIRECT_DFT;
#endif

int check_common(const void* iq, int iq_dtype, int64_t n_frames, int64_t frame_size, int64_t frame_stride,
                 int64_t sample_stride) {
  if (iq_dtype != AMC_C64 && iq_dtype != AMC_C128) return fail(AMC_ERR_INVALID_ARG, "iq_dtype %d unknown", iq_dtype);
  if (n_frames < 0) return fail(AMC_ERR_INVALID_ARG, "n_frames %lld < 0", (long long)n_frames);
  if (frame_size < 1) return fail(AMC_ERR_INVALID_ARG, "frame_size %lld < 1", (long long)frame_size);
  if (frame_size > (1 << 30)) return fail(AMC_ERR_UNSUPPORTED, "frame_size %lld too large", (long long)frame_size);
  if (n_frames > 0 && iq == nullptr) return fail(AMC_ERR_INVALID_ARG, "iq is NULL");
  if (sample_stride < 1) return fail(AMC_ERR_INVALID_ARG, "sample_stride %lld < 1", (long long)sample_stride);
  if (frame_stride < 0) return fail(AMC_ERR_INVALID_ARG, "frame_stride %lld < 0", (long long)frame_stride);
  return AMC_OK;
}

// ---------------------------------------------------------------- host pipeline state (a small pool per device)
// One pipe = two streams + double-buffered device / pinned staging buffers.  Concurrent host calls on the same device
// each take their own pipe (up to kMaxPipes, then they queue), so one long call does not serialise the others.
struct HostPipe {
  bool busy = false;
  bool ready = false;
  cudaStream_t stream[2] = {nullptr, nullptr};
  void* d_in[2] = {nullptr, nullptr};
  void* d_tr[2] = {nullptr, nullptr};   // transposed copy for sample-major input
  double* d_out[2] = {nullptr, nullptr};
  size_t in_bytes[2] = {0, 0}, tr_bytes[2] = {0, 0}, out_bytes[2] = {0, 0};
  // pageable sources (numpy arrays, memory-mapped .mat planes): gathered by a few host threads into these pinned
  // buffers and copied from there - a cudaMemcpy2DAsync from pageable memory is staged by the driver on ONE thread
  void* h_in[2] = {nullptr, nullptr};
  size_t h_bytes[2] = {0, 0};
  cudaEvent_t h2d_done[2] = {nullptr, nullptr};
};
constexpr int kMaxPipes = 4;
HostPipe g_pipe[kMaxDevices][kMaxPipes];
std::mutex g_pipe_mu[kMaxDevices];
std::condition_variable g_pipe_cv[kMaxDevices];

struct PipeLease {
  int device = -1;
  HostPipe* pipe = nullptr;
  ~PipeLease() {
    if (pipe) {
      {
        std::lock_guard<std::mutex> lk(g_pipe_mu[device]);
        pipe->busy = false;
      }
      g_pipe_cv[device].notify_one();
    }
  }
};
void acquire_pipe(int device, PipeLease* lease) {
  std::unique_lock<std::mutex> lk(g_pipe_mu[device]);
  for (;;) {
    for (int i = 0; i < kMaxPipes; ++i) {
      if (!g_pipe[device][i].busy) {
        g_pipe[device][i].busy = true;
        lease->device = device;
        lease->pipe = &g_pipe[device][i];
        return;
      }
    }
    g_pipe_cv[device].wait(lk);
  }
}

// grow one buffer; its recorded size is valid only while the pointer is (a failed allocation leaves {nullptr, 0})
template <typename Alloc, typename Free>
int grow(void** ptr, size_t* have, size_t want, Alloc alloc, Free release, const char* what) {
  if (*have >= want) return AMC_OK;
  if (*ptr) {
    void* old = *ptr;
    *ptr = nullptr;
    *have = 0;
    const cudaError_t e = release(old);
    if (e != cudaSuccess) return fail(AMC_ERR_CUDA, "freeing the %s buffer failed: %s", what, cudaGetErrorString(e));
  }
  const cudaError_t e = alloc(ptr, want);
  if (e != cudaSuccess) {
    *ptr = nullptr;
    *have = 0;
    cudaGetLastError();
    return fail(AMC_ERR_CUDA, "allocating %zu bytes for the %s buffer failed: %s", want, what, cudaGetErrorString(e));
  }
  *have = want;
  return AMC_OK;
}

int ensure_pipe(HostPipe& p, size_t in_bytes, size_t tr_bytes, size_t out_bytes) {
  if (!p.ready) {
    for (int i = 0; i < 2; ++i)
      if (!p.stream[i]) AMC_CUDA(cudaStreamCreateWithFlags(&p.stream[i], cudaStreamNonBlocking));
    p.ready = true;
  }
  auto dev_alloc = [](void** q, size_t n) { return cudaMalloc(q, n); };
  auto dev_free = [](void* q) { return cudaFree(q); };
  for (int i = 0; i < 2; ++i) {
    int rc = grow(&p.d_in[i], &p.in_bytes[i], in_bytes, dev_alloc, dev_free, "device input");
    if (rc == AMC_OK) rc = grow(&p.d_tr[i], &p.tr_bytes[i], tr_bytes, dev_alloc, dev_free, "device re-layout");
    if (rc == AMC_OK)
      rc = grow(reinterpret_cast<void**>(&p.d_out[i]), &p.out_bytes[i], out_bytes, dev_alloc, dev_free, "device output");
    if (rc != AMC_OK) return rc;
  }
  return AMC_OK;
}

// AMCPY_B200_STAGING_WC=1: write-combined pinned staging (the host threads only ever write it)
bool staging_write_combined() {
  static const bool wc = [] {
    const char* env = std::getenv("AMCPY_B200_STAGING_WC");
    return env && *env && std::atoi(env) != 0;
  }();
  return wc;
}

int ensure_pinned_staging(HostPipe& p, size_t bytes) {
  const unsigned flags = staging_write_combined() ? cudaHostAllocWriteCombined : cudaHostAllocDefault;
  auto host_alloc = [flags](void** q, size_t n) { return cudaHostAlloc(q, n, flags); };
  auto host_free = [](void* q) { return cudaFreeHost(q); };
  for (int i = 0; i < 2; ++i) {
    if (!p.h2d_done[i]) AMC_CUDA(cudaEventCreateWithFlags(&p.h2d_done[i], cudaEventDisableTiming));
    const int rc = grow(&p.h_in[i], &p.h_bytes[i], bytes, host_alloc, host_free, "pinned staging");
    if (rc != AMC_OK) return rc;
  }
  return AMC_OK;
}

// true for ordinary (pageable, unregistered) host memory; pinned / registered / managed memory is copied directly
bool is_pageable_host(const void* ptr) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
    cudaGetLastError();   // older runtimes report unregistered host memory as an error: clear it
    return true;
  }
  return attr.type == cudaMemoryTypeUnregistered;
}

int host_copy_threads() {
  static const int n = [] {
    const char* env = std::getenv("AMCPY_B200_COPY_THREADS");
    if (env && *env) {
      const int v = std::atoi(env);
      if (v >= 1) return v > 64 ? 64 : v;
    }
    const unsigned hc = std::thread::hardware_concurrency();
    const int v = hc >= 16 ? 8 : (hc >= 4 ? static_cast<int>(hc / 2) : 1);
    return v;
  }();
  return n;
}

// dst[r * dst_pitch .. + row_bytes) = src[r * src_pitch .. + row_bytes) for r < n_rows, spread over host threads.
// Never throws (the C ABI must not): whatever could not be handed to a thread is copied by the caller.
void gather_rows(unsigned char* dst, size_t dst_pitch, const unsigned char* src, size_t src_pitch, size_t row_bytes,
                 size_t n_rows) {
  const bool linear = (dst_pitch == row_bytes && src_pitch == row_bytes) || n_rows == 1;
  // unit of work: bytes of one contiguous block, or rows
  const size_t units = linear ? row_bytes * n_rows : n_rows;
  auto work = [=](size_t u0, size_t u1) {
    if (linear) {
      std::memcpy(dst + u0, src + u0, u1 - u0);
    } else {
      for (size_t r = u0; r < u1; ++r) std::memcpy(dst + r * dst_pitch, src + r * src_pitch, row_bytes);
    }
  };
  size_t nt = static_cast<size_t>(host_copy_threads());
  if (row_bytes * n_rows < (4u << 20)) nt = 1;
  if (nt > units) nt = units;
  if (nt <= 1) {
    work(0, units);
    return;
  }
  const size_t per = (units + nt - 1) / nt;
  size_t handed = per;                      // the caller copies the first share itself
  {
    std::thread th[64];
    size_t started = 0;
    try {
      for (size_t t = 1; t < nt && handed < units; ++t) {
        const size_t u0 = handed, u1 = (u0 + per < units) ? u0 + per : units;
        th[started] = std::thread(work, u0, u1);
        ++started;
        handed = u1;
      }
    } catch (...) {
      // thread creation failed: the remaining shares are copied below
    }
    work(0, per < units ? per : units);
    for (size_t t = 0; t < started; ++t) th[t].join();
  }
  if (handed < units) work(handed, units);
}

}  // namespace

extern "C" {

int amc_version(void) { return 2000; }

const char* amc_last_error_string(void) { return t_err.c_str(); }

int64_t amc_launch_count(void) { return t_launches; }

int amc_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n < 1) {
    cudaGetLastError();
    return fail(AMC_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  }
  return n;
}

int amc_init(int device) {
  if (device < 0 || device >= kMaxDevices) return fail(AMC_ERR_INVALID_ARG, "device %d out of range", device);
  DeviceGuard guard;
  AMC_CUDA(cudaGetDevice(&guard.prev));
  AMC_CUDA(cudaSetDevice(device));
  guard.armed = true;
  DevInfo di;
  int rc = device_info(&di);
  if (rc != AMC_OK) return rc;
  cudaStream_t init_stream;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    rc = init_stream_locked(g_dev[device]);
    if (rc != AMC_OK) return rc;
    init_stream = g_dev[device].init_stream;
  }
  rc = ensure_twiddles(device, init_stream);
  if (rc != AMC_OK) return rc;
  AMC_CUDA(cudaStreamSynchronize(init_stream));
  std::lock_guard<std::mutex> lk(g_mu);
  g_dev[device].tw_done = true;
  return AMC_OK;
}

int64_t amc_workspace_bytes(int iq_dtype, int64_t n_frames, int64_t frame_size, int host_path) {
  if (iq_dtype != AMC_C64 && iq_dtype != AMC_C128) return fail(AMC_ERR_INVALID_ARG, "iq_dtype %d unknown", iq_dtype);
  if (n_frames < 0 || frame_size < 1) return fail(AMC_ERR_INVALID_ARG, "bad shape");
  int64_t bytes = 0;
  if (!is_pow2(frame_size)) {                       // Bluestein tables, cached per (device, frame_size)
    const int64_t m = bluestein_m(frame_size);
    if (m > 0) bytes += (frame_size + m) * static_cast<int64_t>(sizeof(float2));
    if (m > kBluesteinSmemM) {                      // + the stream-ordered transform workspace of one launch
      const int64_t per = m * static_cast<int64_t>(sizeof(float2));
      int64_t ctas = static_cast<int64_t>(kFftWorkspaceMax) / per;
      if (ctas < 1) ctas = 1;
      bytes += (n_frames < ctas ? n_frames : ctas) * per;
    }
  } else if (frame_size > kGeneralPow2Smem) {
    const int64_t per = frame_size * static_cast<int64_t>(sizeof(float2));
    int64_t ctas = static_cast<int64_t>(kFftWorkspaceMax) / per;
    if (ctas < 1) ctas = 1;
    bytes += (n_frames < ctas ? n_frames : ctas) * per;
  }
  if (frame_size == 8192 || frame_size == 16384) {  // |x| scratch of the long-frame kernel: N doubles per resident CTA
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      cudaGetLastError();
      sms = 148;                                    // no device visible: a B200's SM count
    }
    const int64_t ctas = static_cast<int64_t>(sms) * (frame_size == 8192 ? 2 : 1);
    bytes += (n_frames < ctas ? n_frames : ctas) * frame_size * static_cast<int64_t>(sizeof(double));
  }
  if (host_path) {                                  // double-buffered chunk buffers of one pipe
    const int64_t elt = iq_dtype == AMC_C128 ? 16 : 8;
    const int64_t frame_bytes = frame_size * elt;
    int64_t chunk = (64ll << 20) / frame_bytes;
    chunk = chunk < 32 ? 32 : (chunk / 32) * 32;
    if (chunk > n_frames) chunk = n_frames;
    bytes += 2 * (2 * chunk * frame_bytes + chunk * AMC_N_FEATURES * static_cast<int64_t>(sizeof(double)));
  }
  return bytes;
}

int amc_extract_batch(const void* iq, int iq_dtype, int64_t n_frames, int64_t frame_size, int64_t frame_stride,
                      int64_t sample_stride, double* out, int64_t out_stride, uint32_t feature_mask, int flags,
                      void* cuda_stream) {
  int rc = check_common(iq, iq_dtype, n_frames, frame_size, frame_stride, sample_stride);
  if (rc != AMC_OK) return rc;
  if (flags & ~kKnownFlags) return fail(AMC_ERR_INVALID_ARG, "unknown flag bits 0x%x", flags & ~kKnownFlags);
  if ((feature_mask & AMC_ALL_FEATURES) == 0) return fail(AMC_ERR_INVALID_ARG, "feature_mask selects nothing");
  if (out_stride < AMC_N_FEATURES) return fail(AMC_ERR_INVALID_ARG, "out_stride %lld < 18", (long long)out_stride);
  if (n_frames == 0) return AMC_OK;
  if (out == nullptr) return fail(AMC_ERR_INVALID_ARG, "out is NULL");
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  DevInfo di;
  rc = device_info(&di);
  if (rc != AMC_OK) return rc;

  const size_t elt = iq_dtype == AMC_C128 ? 16 : 8;
  const bool aligned = (reinterpret_cast<uintptr_t>(iq) % 16 == 0) && ((frame_stride * elt) % 16 == 0);
  const bool fused = !(flags & AMC_FLAG_FORCE_GENERAL) && sample_stride == 1 && fused_size(frame_size) && aligned &&
                     (frame_stride >= frame_size || n_frames == 1);
  if (fused) {
    rc = ensure_twiddles(di.dev, stream);
    if (rc != AMC_OK) return rc;
    const int prof = pick_profile(feature_mask);
    int used = amc::kProfAll;
    const unsigned long long ticket = g_ticket.fetch_add(1, std::memory_order_relaxed);
    // Programmatic dependent launch is worth 0.3-0.6 % of a step at every fused size except N = 1024, whose
    // 64-thread CTAs (four per SM) lose 10 % when the successor's CTAs become resident during the tail
    // (profiles/r2_experiments.txt): that size gets ordinary launches.
    t_allow_pdl = frame_size != 1024;
    if (iq_dtype == AMC_C128)
      rc = dispatch_fused<double2>(frame_size, iq, n_frames, frame_stride, out, out_stride, di.sms, stream, flags, prof,
                                   &used, ticket);
    else
      rc = dispatch_fused<float2>(frame_size, iq, n_frames, frame_stride, out, out_stride, di.sms, stream, flags, prof,
                                  &used, ticket);
    if (rc != AMC_OK) return rc;
    if (used == amc::kProfMom) return AMC_OK;        // float64 sums only: nothing for the careful path to redo
    const unsigned long long tk = used < 0 ? 0ull : ticket;   // 0: always scan (A/B experiment kernels)
    if (iq_dtype == AMC_C128)
      return launch_redo<double2>(iq, n_frames, frame_size, frame_stride, out, out_stride, di.sms, stream, tk);
    return launch_redo<float2>(iq, n_frames, frame_size, frame_stride, out, out_stride, di.sms, stream, tk);
  }
  if (iq_dtype == AMC_C128)
    return launch_general<double2>(iq, n_frames, frame_size, frame_stride, sample_stride, out, out_stride, di.sms,
                                   stream, flags);
  return launch_general<float2>(iq, n_frames, frame_size, frame_stride, sample_stride, out, out_stride, di.sms,
                                stream, flags);
}

int amc_frames_from_sample_major(const void* src, int iq_dtype, int64_t n_frames, int64_t frame_size,
                                 int64_t src_sample_stride, void* dst, void* cuda_stream) {
  int rc = check_common(src, iq_dtype, n_frames, frame_size, 1, src_sample_stride);
  if (rc != AMC_OK) return rc;
  if (n_frames == 0) return AMC_OK;
  if (dst == nullptr) return fail(AMC_ERR_INVALID_ARG, "dst is NULL");
  if (src_sample_stride < n_frames) return fail(AMC_ERR_INVALID_ARG, "src_sample_stride < n_frames");
  if (frame_size > 65535LL * 32) return fail(AMC_ERR_UNSUPPORTED, "frame_size too large for the re-layout grid");
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  dim3 grid(static_cast<unsigned>((n_frames + 31) / 32), static_cast<unsigned>((frame_size + 31) / 32));
  if (iq_dtype == AMC_C128)
    amc::frames_from_sample_major_kernel<double2><<<grid, 256, 0, stream>>>(
        static_cast<const double2*>(src), n_frames, static_cast<int>(frame_size), src_sample_stride,
        static_cast<double2*>(dst));
  else
    amc::frames_from_sample_major_kernel<float2><<<grid, 256, 0, stream>>>(
        static_cast<const float2*>(src), n_frames, static_cast<int>(frame_size), src_sample_stride,
        static_cast<float2*>(dst));
  ++t_launches;
  AMC_CUDA(cudaGetLastError());
  return AMC_OK;
}

int amc_instantaneous_batch(const void* iq, int iq_dtype, int64_t n_frames, int64_t frame_size, int64_t frame_stride,
                            int64_t sample_stride, double* abs_out, double* phase_out, double* unwrapped_out,
                            double* frequency_out, double* cn_amplitude_out, void* cuda_stream) {
  int rc = check_common(iq, iq_dtype, n_frames, frame_size, frame_stride, sample_stride);
  if (rc != AMC_OK) return rc;
  if (n_frames == 0) return AMC_OK;
  DevInfo di;
  rc = device_info(&di);
  if (rc != AMC_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  const int64_t cap = static_cast<int64_t>(di.sms) * 8;
  const int grid = static_cast<int>(n_frames < cap ? n_frames : cap);
  if (iq_dtype == AMC_C128)
    amc::instantaneous_kernel<double2><<<grid, amc::kGenThreads, 0, stream>>>(
        static_cast<const double2*>(iq), n_frames, static_cast<int>(frame_size), frame_stride, sample_stride, abs_out,
        phase_out, unwrapped_out, frequency_out, cn_amplitude_out);
  else
    amc::instantaneous_kernel<float2><<<grid, amc::kGenThreads, 0, stream>>>(
        static_cast<const float2*>(iq), n_frames, static_cast<int>(frame_size), frame_stride, sample_stride, abs_out,
        phase_out, unwrapped_out, frequency_out, cn_amplitude_out);
  ++t_launches;
  AMC_CUDA(cudaGetLastError());
  return AMC_OK;
}

int amc_moments_batch(const void* iq, int iq_dtype, int64_t n_frames, int64_t frame_size, int64_t frame_stride,
                      int64_t sample_stride, double* out, void* cuda_stream) {
  int rc = check_common(iq, iq_dtype, n_frames, frame_size, frame_stride, sample_stride);
  if (rc != AMC_OK) return rc;
  if (n_frames == 0) return AMC_OK;
  if (out == nullptr) return fail(AMC_ERR_INVALID_ARG, "out is NULL");
  DevInfo di;
  rc = device_info(&di);
  if (rc != AMC_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  const int64_t cap = static_cast<int64_t>(di.sms) * 8;
  const int grid = static_cast<int>(n_frames < cap ? n_frames : cap);
  if (iq_dtype == AMC_C128)
    amc::moments_kernel<double2><<<grid, amc::kGenThreads, 0, stream>>>(
        static_cast<const double2*>(iq), n_frames, static_cast<int>(frame_size), frame_stride, sample_stride, out);
  else
    amc::moments_kernel<float2><<<grid, amc::kGenThreads, 0, stream>>>(
        static_cast<const float2*>(iq), n_frames, static_cast<int>(frame_size), frame_stride, sample_stride, out);
  ++t_launches;
  AMC_CUDA(cudaGetLastError());
  return AMC_OK;
}

int amc_generate_frames(void* out, int iq_dtype, int n_cells, int64_t frames_per_cell, int64_t first_frame,
                        int64_t frame_size, const int* cell_mod, const int* cell_snr_idx, const double* cell_sigma,
                        uint64_t seed, void* cuda_stream) {
  if (iq_dtype != AMC_C64 && iq_dtype != AMC_C128) return fail(AMC_ERR_INVALID_ARG, "iq_dtype %d unknown", iq_dtype);
  if (n_cells < 0 || frames_per_cell < 0 || frame_size < 1 || frame_size > (1 << 30) || first_frame < 0)
    return fail(AMC_ERR_INVALID_ARG, "bad generator shape");
  if (n_cells == 0 || frames_per_cell == 0) return AMC_OK;
  if (!out || !cell_mod || !cell_snr_idx || !cell_sigma) return fail(AMC_ERR_INVALID_ARG, "NULL pointer");
  DevInfo di;
  int rc = device_info(&di);
  if (rc != AMC_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  const int grid = di.sms * 8;
  if (iq_dtype == AMC_C128)
    amc::generate_frames_kernel<double2><<<grid, 256, 0, stream>>>(static_cast<double2*>(out), n_cells, frames_per_cell,
                                                                   first_frame, static_cast<int>(frame_size), cell_mod,
                                                                   cell_snr_idx, cell_sigma, seed);
  else
    amc::generate_frames_kernel<float2><<<grid, 256, 0, stream>>>(static_cast<float2*>(out), n_cells, frames_per_cell,
                                                                  first_frame, static_cast<int>(frame_size), cell_mod,
                                                                  cell_snr_idx, cell_sigma, seed);
  ++t_launches;
  AMC_CUDA(cudaGetLastError());
  return AMC_OK;
}

int amc_extract_host(const void* iq, int iq_dtype, int64_t n_frames, int64_t frame_size, int64_t frame_stride,
                     int64_t sample_stride, double* out, int64_t out_stride, uint32_t feature_mask, int flags,
                     int device) {
  int rc = check_common(iq, iq_dtype, n_frames, frame_size, frame_stride, sample_stride);
  if (rc != AMC_OK) return rc;
  if ((feature_mask & AMC_ALL_FEATURES) == 0) return fail(AMC_ERR_INVALID_ARG, "feature_mask selects nothing");
  if (out_stride < AMC_N_FEATURES) return fail(AMC_ERR_INVALID_ARG, "out_stride %lld < 18", (long long)out_stride);
  if (n_frames == 0) return AMC_OK;
  if (out == nullptr) return fail(AMC_ERR_INVALID_ARG, "out is NULL");
  const bool row_major = sample_stride == 1 && (frame_stride >= frame_size || n_frames == 1);
  const bool sample_major = !row_major && frame_stride == 1 && sample_stride >= n_frames;
  if (!row_major && !sample_major)
    return fail(AMC_ERR_UNSUPPORTED,
                "host layout must be row-per-frame (sample_stride 1) or sample-major (frame_stride 1)");
  if (device < 0 || device >= kMaxDevices) return fail(AMC_ERR_INVALID_ARG, "device %d out of range", device);
  DeviceGuard guard;                                 // restores the caller's device on every exit path
  AMC_CUDA(cudaGetDevice(&guard.prev));
  AMC_CUDA(cudaSetDevice(device));
  guard.armed = true;

  const size_t elt = iq_dtype == AMC_C128 ? 16 : 8;
  const size_t frame_bytes = static_cast<size_t>(frame_size) * elt;
  // ~64 MiB of samples per chunk, a multiple of 32 frames, at least 32
  int64_t chunk = static_cast<int64_t>((64u << 20) / frame_bytes);
  chunk = chunk < 32 ? 32 : (chunk / 32) * 32;
  if (chunk > n_frames) chunk = n_frames;

  PipeLease lease;                                   // this call's own streams + staging buffers
  acquire_pipe(device, &lease);
  HostPipe& p = *lease.pipe;
  rc = ensure_pipe(p, static_cast<size_t>(chunk) * frame_bytes, sample_major ? static_cast<size_t>(chunk) * frame_bytes : 0,
                   static_cast<size_t>(chunk) * AMC_N_FEATURES * sizeof(double));
  if (rc != AMC_OK) return rc;
  const unsigned char* src = static_cast<const unsigned char*>(iq);
  const bool staged = is_pageable_host(iq) && static_cast<size_t>(n_frames) * frame_bytes >= (8u << 20);
  if (staged) {
    rc = ensure_pinned_staging(p, static_cast<size_t>(chunk) * frame_bytes);
    if (rc != AMC_OK) return rc;
  }
  int status = AMC_OK;
  int64_t c = 0;
  for (int64_t f0 = 0; f0 < n_frames && status == AMC_OK; f0 += chunk, ++c) {
    const int b = static_cast<int>(c & 1);
    const int64_t nf = (n_frames - f0) < chunk ? (n_frames - f0) : chunk;
    cudaStream_t st = p.stream[b];
    cudaError_t e;
    const void* dev_frames = p.d_in[b];
    if (staged) {   // host threads gather the chunk into pinned memory (same layout as the device buffer), one linear copy
      e = c >= 2 ? cudaEventSynchronize(p.h2d_done[b]) : cudaSuccess;   // the copy that last read this buffer is done
      if (e == cudaSuccess) {
        unsigned char* h = static_cast<unsigned char*>(p.h_in[b]);
        if (row_major)
          gather_rows(h, frame_bytes, src + static_cast<size_t>(f0) * frame_stride * elt,
                      (nf == 1 ? frame_bytes : static_cast<size_t>(frame_stride) * elt), frame_bytes, static_cast<size_t>(nf));
        else
          gather_rows(h, static_cast<size_t>(nf) * elt, src + static_cast<size_t>(f0) * elt,
                      static_cast<size_t>(sample_stride) * elt, static_cast<size_t>(nf) * elt,
                      static_cast<size_t>(frame_size));
        e = cudaMemcpyAsync(p.d_in[b], h, static_cast<size_t>(nf) * frame_bytes, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaEventRecord(p.h2d_done[b], st);
      }
    } else if (row_major && (frame_stride == frame_size || nf == 1)) {   // contiguous rows: one linear copy
      e = cudaMemcpyAsync(p.d_in[b], src + static_cast<size_t>(f0) * frame_stride * elt,
                          static_cast<size_t>(nf) * frame_bytes, cudaMemcpyHostToDevice, st);
    } else if (row_major) {
      e = cudaMemcpy2DAsync(p.d_in[b], frame_bytes, src + static_cast<size_t>(f0) * frame_stride * elt,
                            static_cast<size_t>(frame_stride) * elt, frame_bytes, static_cast<size_t>(nf),
                            cudaMemcpyHostToDevice, st);
    } else {
      e = cudaMemcpy2DAsync(p.d_in[b], static_cast<size_t>(nf) * elt, src + static_cast<size_t>(f0) * elt,
                            static_cast<size_t>(sample_stride) * elt, static_cast<size_t>(nf) * elt,
                            static_cast<size_t>(frame_size), cudaMemcpyHostToDevice, st);
    }
    if (e != cudaSuccess) {
      status = fail(AMC_ERR_CUDA, "host->device copy failed: %s", cudaGetErrorString(e));
      break;
    }
    if (sample_major) {
      status = amc_frames_from_sample_major(p.d_in[b], iq_dtype, nf, frame_size, nf, p.d_tr[b], st);
      if (status != AMC_OK) break;
      dev_frames = p.d_tr[b];
    }
    status = amc_extract_batch(dev_frames, iq_dtype, nf, frame_size, frame_size, 1, p.d_out[b], AMC_N_FEATURES,
                               feature_mask, flags, st);
    if (status != AMC_OK) break;
    if (out_stride == AMC_N_FEATURES)
      e = cudaMemcpyAsync(out + f0 * out_stride, p.d_out[b], static_cast<size_t>(nf) * AMC_N_FEATURES * sizeof(double),
                          cudaMemcpyDeviceToHost, st);
    else
      e = cudaMemcpy2DAsync(out + f0 * out_stride, static_cast<size_t>(out_stride) * sizeof(double), p.d_out[b],
                            AMC_N_FEATURES * sizeof(double), AMC_N_FEATURES * sizeof(double), static_cast<size_t>(nf),
                            cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) status = fail(AMC_ERR_CUDA, "device->host copy failed: %s", cudaGetErrorString(e));
  }
  for (int i = 0; i < 2; ++i) {
    cudaError_t e = cudaStreamSynchronize(p.stream[i]);
    if (e != cudaSuccess && status == AMC_OK)
      status = fail(AMC_ERR_CUDA, "stream sync failed: %s", cudaGetErrorString(e));
  }
  return status;
}

int amc_extract_host_planar(const void* re, const void* im, int iq_dtype, int64_t n_frames, int64_t frame_size,
                            int64_t sample_stride, double* out, int64_t out_stride, uint32_t feature_mask, int flags,
                            int device) {
  int rc = check_common(re, iq_dtype, n_frames, frame_size, 1, sample_stride);
  if (rc != AMC_OK) return rc;
  if ((feature_mask & AMC_ALL_FEATURES) == 0) return fail(AMC_ERR_INVALID_ARG, "feature_mask selects nothing");
  if (out_stride < AMC_N_FEATURES) return fail(AMC_ERR_INVALID_ARG, "out_stride %lld < 18", (long long)out_stride);
  if (n_frames == 0) return AMC_OK;
  if (out == nullptr) return fail(AMC_ERR_INVALID_ARG, "out is NULL");
  if (sample_stride < n_frames) return fail(AMC_ERR_INVALID_ARG, "sample_stride < n_frames");
  if (frame_size > 65535LL * 32) return fail(AMC_ERR_UNSUPPORTED, "frame_size too large for the re-layout grid");
  if (device < 0 || device >= kMaxDevices) return fail(AMC_ERR_INVALID_ARG, "device %d out of range", device);
  DeviceGuard guard;                                 // restores the caller's device on every exit path
  AMC_CUDA(cudaGetDevice(&guard.prev));
  AMC_CUDA(cudaSetDevice(device));
  guard.armed = true;

  const size_t relt = iq_dtype == AMC_C128 ? 8 : 4;            // bytes per real plane element
  const size_t frame_bytes = static_cast<size_t>(frame_size) * 2 * relt;
  int64_t chunk = static_cast<int64_t>((64u << 20) / frame_bytes);   // ~64 MiB of samples per chunk
  chunk = chunk < 32 ? 32 : (chunk / 32) * 32;
  if (chunk > n_frames) chunk = n_frames;

  PipeLease lease;                                   // this call's own streams + staging buffers
  acquire_pipe(device, &lease);
  HostPipe& p = *lease.pipe;
  rc = ensure_pipe(p, static_cast<size_t>(chunk) * frame_bytes, static_cast<size_t>(chunk) * frame_bytes,
                   static_cast<size_t>(chunk) * AMC_N_FEATURES * sizeof(double));
  if (rc != AMC_OK) return rc;
  const unsigned char* src_re = static_cast<const unsigned char*>(re);
  const unsigned char* src_im = static_cast<const unsigned char*>(im);
  const bool staged = is_pageable_host(re) && static_cast<size_t>(n_frames) * frame_bytes >= (8u << 20);
  if (staged) {
    rc = ensure_pinned_staging(p, static_cast<size_t>(chunk) * frame_bytes);
    if (rc != AMC_OK) return rc;
  }
  int status = AMC_OK;
  int64_t c = 0;
  for (int64_t f0 = 0; f0 < n_frames && status == AMC_OK; f0 += chunk, ++c) {
    const int b = static_cast<int>(c & 1);
    const int64_t nf = (n_frames - f0) < chunk ? (n_frames - f0) : chunk;
    cudaStream_t st = p.stream[b];
    unsigned char* d_re = static_cast<unsigned char*>(p.d_in[b]);
    unsigned char* d_im = d_re + static_cast<size_t>(nf) * frame_size * relt;   // second half of the staging buffer
    const size_t row = static_cast<size_t>(nf) * relt, pitch = static_cast<size_t>(sample_stride) * relt;
    cudaError_t e;
    if (staged) {   // host threads gather both planes of the chunk into pinned memory, one linear copy
      e = c >= 2 ? cudaEventSynchronize(p.h2d_done[b]) : cudaSuccess;
      if (e == cudaSuccess) {
        unsigned char* h = static_cast<unsigned char*>(p.h_in[b]);
        const size_t plane = row * static_cast<size_t>(frame_size);
        gather_rows(h, row, src_re + static_cast<size_t>(f0) * relt, pitch, row, static_cast<size_t>(frame_size));
        if (src_im)
          gather_rows(h + plane, row, src_im + static_cast<size_t>(f0) * relt, pitch, row, static_cast<size_t>(frame_size));
        e = cudaMemcpyAsync(d_re, h, src_im ? 2 * plane : plane, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaEventRecord(p.h2d_done[b], st);
      }
    } else {
      e = cudaMemcpy2DAsync(d_re, row, src_re + static_cast<size_t>(f0) * relt, pitch, row,
                            static_cast<size_t>(frame_size), cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess && src_im)
        e = cudaMemcpy2DAsync(d_im, row, src_im + static_cast<size_t>(f0) * relt, pitch, row,
                              static_cast<size_t>(frame_size), cudaMemcpyHostToDevice, st);
    }
    if (e != cudaSuccess) {
      status = fail(AMC_ERR_CUDA, "host->device copy failed: %s", cudaGetErrorString(e));
      break;
    }
    dim3 grid(static_cast<unsigned>((nf + 31) / 32), static_cast<unsigned>((frame_size + 31) / 32));
    if (iq_dtype == AMC_C128)
      amc::frames_from_planar_kernel<double, double2><<<grid, 256, 0, st>>>(
          reinterpret_cast<const double*>(d_re), src_im ? reinterpret_cast<const double*>(d_im) : nullptr, nf,
          static_cast<int>(frame_size), nf, static_cast<double2*>(p.d_tr[b]));
    else
      amc::frames_from_planar_kernel<float, float2><<<grid, 256, 0, st>>>(
          reinterpret_cast<const float*>(d_re), src_im ? reinterpret_cast<const float*>(d_im) : nullptr, nf,
          static_cast<int>(frame_size), nf, static_cast<float2*>(p.d_tr[b]));
    ++t_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) {
      status = fail(AMC_ERR_CUDA, "re-layout launch failed: %s", cudaGetErrorString(e));
      break;
    }
    status = amc_extract_batch(p.d_tr[b], iq_dtype, nf, frame_size, frame_size, 1, p.d_out[b], AMC_N_FEATURES,
                               feature_mask, flags, st);
    if (status != AMC_OK) break;
    e = cudaMemcpy2DAsync(out + f0 * out_stride, static_cast<size_t>(out_stride) * sizeof(double), p.d_out[b],
                          AMC_N_FEATURES * sizeof(double), AMC_N_FEATURES * sizeof(double), static_cast<size_t>(nf),
                          cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) status = fail(AMC_ERR_CUDA, "device->host copy failed: %s", cudaGetErrorString(e));
  }
  for (int i = 0; i < 2; ++i) {
    cudaError_t e = cudaStreamSynchronize(p.stream[i]);
    if (e != cudaSuccess && status == AMC_OK)
      status = fail(AMC_ERR_CUDA, "stream sync failed: %s", cudaGetErrorString(e));
  }
  return status;
}

}  // extern "C"
