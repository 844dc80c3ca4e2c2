// On-device synthetic IQ generator: BPSK / QPSK / 8PSK / 16QAM / 64QAM / WGN
// frames with AWGN, one sample per symbol, unit mean power, noise variance 10^(-SNR/10).
// Counter-based: every sample is a pure function of (seed, modulation, snr index, frame, sample),
// so any shard on any rank regenerates identical frames (Philox4x32-10, Box-Muller).
// The reference has no generator (its data came from GNU Radio captures); the recipe is the one
// amcpy_b200/synth.py documents (different random stream: numpy's Philox/ziggurat is not reproduced).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace amc {

struct Philox4 {
  uint32_t c[4];
};
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += W0;
    k1 += W1;
  }
  return {{c0, c1, c2, c3}};
}

// uniform in (0, 1] with 53 random bits
__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {
  const uint64_t v = (static_cast<uint64_t>(hi) << 32) | lo;
  return (static_cast<double>(v >> 11) + 1.0) * (1.0 / 9007199254740992.0);
}

// constellation point `sym` of modulation `mod` (0..4), unit mean power
__device__ __forceinline__ double2 constellation_point(int mod, uint32_t sym) {
  if (mod == 0) return make_double2((sym & 1) ? -1.0 : 1.0, 0.0);
  if (mod == 1 || mod == 2) {
    const int m = mod == 1 ? 4 : 8;
    const int k = sym % m;
    double s, c;
    sincospi(mod == 1 ? 0.25 + 0.5 * k : 0.25 * k, &s, &c);
    return make_double2(c, s);
  }
  const int side = mod == 3 ? 4 : 8;                     // 16QAM / 64QAM square grid
  const double norm = mod == 3 ? 0.31622776601683794 : 0.1543033499620919;   // 1/sqrt(10), 1/sqrt(42)
  const int i = sym % side, q = (sym / side) % side;
  return make_double2((2 * i - (side - 1)) * norm, (2 * q - (side - 1)) * norm);
}

// out[(cell * n_frames + frame) * frame_size + n]; cell = flattened (modulation, snr) index given by
// cell_mod[cell], cell_sigma[cell] (noise std per rail).  CT = double2 or float2.
template <typename CT>
__global__ void __launch_bounds__(256)
generate_frames_kernel(CT* __restrict__ out, int n_cells, int64_t frames_per_cell, int64_t first_frame,
                       int frame_size, const int* __restrict__ cell_mod, const int* __restrict__ cell_snr_idx,
                       const double* __restrict__ cell_sigma, uint64_t seed) {
  const int64_t total = static_cast<int64_t>(n_cells) * frames_per_cell * frame_size;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i % frame_size);
    const int64_t fr = (i / frame_size) % frames_per_cell + first_frame;
    const int cell = static_cast<int>(i / (static_cast<int64_t>(frame_size) * frames_per_cell));
    const int mod = cell_mod[cell];
    // counter = (sample, frame lo, frame hi | snr idx << 16, modulation); key = seed
    const Philox4 a = philox4x32_10(static_cast<uint32_t>(n), static_cast<uint32_t>(fr),
                                    static_cast<uint32_t>(fr >> 32) | (static_cast<uint32_t>(cell_snr_idx[cell]) << 16),
                                    static_cast<uint32_t>(mod), static_cast<uint32_t>(seed),
                                    static_cast<uint32_t>(seed >> 32));
    const Philox4 b = philox4x32_10(static_cast<uint32_t>(n), static_cast<uint32_t>(fr),
                                    static_cast<uint32_t>(fr >> 32) | (static_cast<uint32_t>(cell_snr_idx[cell]) << 16),
                                    static_cast<uint32_t>(mod) | 0x80000000u, static_cast<uint32_t>(seed),
                                    static_cast<uint32_t>(seed >> 32));
    // Box-Muller: two independent N(0,1)
    const double rad = sqrt(-2.0 * log(u01(a.c[0], a.c[1])));
    double sn, cs;
    sincospi(2.0 * u01(a.c[2], a.c[3]), &sn, &cs);
    const double sigma = cell_sigma[cell];
    double re = sigma * rad * cs, im = sigma * rad * sn;
    if (mod < 5) {
      const double2 p = constellation_point(mod, b.c[0]);
      re += p.x;
      im += p.y;
    }
    CT v;
    v.x = re;
    v.y = im;
    out[i] = v;
  }
}

}  // namespace amc
