/* Minimal C host for the amcpy_b200 C ABI (include/amcpy_b200.h): no Python, no torch.
 *
 *   gcc -O2 -Iinclude examples/extract_host.c -o extract_host -Lamcpy_b200/_lib -lamcpy_b200 -lm \
 *       -Wl,-rpath,$PWD/amcpy_b200/_lib
 *   ./extract_host [n_frames]
 *
 * Builds n_frames QPSK frames of 2048 complex128 samples (unit power, 10 dB SNR, a fixed LCG so the run is
 * reproducible), calls amc_extract_host - the batched replacement of the reference's per-frame
 * calculate_features(range(1,19), signal) fan-out (feature_extraction.py:30-39, :64-74) - and prints the 18 features
 * of frame 0 plus the batch means of |C20|, |C40|, |C42| (QPSK: ~0, ~1, ~1).  Without a CUDA device it reports the
 * library's error string and exits 0 after the ABI checks that need no GPU. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "amcpy_b200.h"

static uint64_t lcg_state = 0x9E3779B97F4A7C15ull;
static double uniform01(void) {
  lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
  return ((double)(lcg_state >> 11) + 0.5) / 9007199254740992.0;
}
static double gauss(void) { return sqrt(-2.0 * log(uniform01())) * cos(6.283185307179586 * uniform01()); }

int main(int argc, char** argv) {
  const int64_t n_frames = argc > 1 ? atoll(argv[1]) : 64, n = 2048;
  printf("amcpy_b200 ABI version %d\n", amc_version());
  /* argument checking needs no device: a bad dtype code is AMC_ERR_INVALID_ARG with a message */
  if (amc_extract_batch(NULL, 99, 1, n, n, 1, NULL, 18, AMC_ALL_FEATURES, 0, NULL) != AMC_ERR_INVALID_ARG) {
    fprintf(stderr, "argument check did not fire\n");
    return 1;
  }
  printf("argument check: %s\n", amc_last_error_string());
  /* what the library itself would allocate for this call shape (device path at a power-of-two size: nothing) */
  printf("workspace: device path %lld B, host path %lld B\n", (long long)amc_workspace_bytes(AMC_C128, n_frames, n, 0),
         (long long)amc_workspace_bytes(AMC_C128, n_frames, n, 1));
  const int devices = amc_device_count();
  if (devices < 1) {
    printf("no CUDA device (%d): %s - nothing computed, there is no CPU path\n", devices, amc_last_error_string());
    return 0;
  }
  if (amc_init(0) != AMC_OK) { /* optional: builds the tables now instead of on the first call */
    fprintf(stderr, "amc_init failed: %s\n", amc_last_error_string());
    return 1;
  }
  double* iq = (double*)malloc((size_t)n_frames * n * 2 * sizeof(double));
  double* out = (double*)malloc((size_t)n_frames * AMC_N_FEATURES * sizeof(double));
  if (!iq || !out) return 1;
  const double sigma = sqrt(pow(10.0, -10.0 / 10.0) / 2.0), h = 0.7071067811865476;
  for (int64_t i = 0; i < n_frames * n; ++i) {
    const int s = (int)(uniform01() * 4.0) & 3;
    iq[2 * i] = ((s & 1) ? -h : h) + sigma * gauss();
    iq[2 * i + 1] = ((s & 2) ? -h : h) + sigma * gauss();
  }
  const int rc = amc_extract_host(iq, AMC_C128, n_frames, n, /*frame_stride*/ n, /*sample_stride*/ 1, out,
                                  AMC_N_FEATURES, AMC_ALL_FEATURES, 0, /*device*/ 0);
  if (rc != AMC_OK) {
    fprintf(stderr, "amc_extract_host failed (%d): %s\n", rc, amc_last_error_string());
    return 1;
  }
  printf("frame 0:");
  for (int k = 0; k < AMC_N_FEATURES; ++k) printf(" %.6g", out[k]);
  double c20 = 0, c40 = 0, c42 = 0;
  for (int64_t f = 0; f < n_frames; ++f) {
    c20 += out[f * AMC_N_FEATURES + 9];
    c40 += out[f * AMC_N_FEATURES + 11];
    c42 += out[f * AMC_N_FEATURES + 13];
  }
  printf("\nmean |C20| %.4f  |C40| %.4f  |C42| %.4f over %lld frames (QPSK at 10 dB: ~0, ~1, ~1); %lld launches\n",
         c20 / n_frames, c40 / n_frames, c42 / n_frames, (long long)n_frames, (long long)amc_launch_count());
  const int ok = c20 / n_frames < 0.1 && fabs(c40 / n_frames - 1.0) < 0.15 && fabs(c42 / n_frames - 1.0) < 0.15;
  free(iq);
  free(out);
  return ok ? 0 : 2;
}
