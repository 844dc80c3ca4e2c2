/*
 * amcpy_b200 - C ABI of the B200-native feature-extraction hot path of amcpy.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types, no exceptions
 * across the ABI.  Every entry returns AMC_OK (0) or a negative AMC_ERR_* code and records a
 * message retrievable (per thread) with amc_last_error_string().
 *
 * What each entry replaces in the reference (paths relative to /root/reference/src/amcpy/):
 *
 *   amc_extract_batch / amc_extract_host
 *       the per-frame operator fan-out: `_Worker.run` calling
 *       `calculate_features(list(range(1, 19)), signal)` for every (snr, frame)
 *       (feature_extraction.py:30-39, :64-74) and, inside it, the 18 functions of
 *       features.py:66-185 dispatched through `_FEATURE_FUNCTIONS` (features.py:192-232).
 *       Output row k of a frame = feature id k+1, float64 (the reference's Python floats);
 *       the float32 cast of feature_extraction.py:56 is done by the caller that writes .mat.
 *   amc_instantaneous_batch
 *       `InstantaneousValues.__init__` (features.py:17-31): abs, phase, unwrapped_phase,
 *       frequency (N-1 values), cn_amplitude.
 *   amc_moments_batch
 *       `MomentValues.__init__` (features.py:39-58): m20 m21 m22 m40 m41 m42 m43 m60 m61 m62 m63.
 *   amc_frames_from_sample_major
 *       the strided `parsed[snr, frame, 0:frame_size]` views of a Fortran-ordered loadmat array
 *       (feature_extraction.py:48,68): turns sample-major storage into one contiguous row per frame.
 *
 * Layout contract: a "frame" is `frame_size` complex samples; frame f starts at element
 * f*frame_stride, sample n of it is at +n*sample_stride (strides in complex ELEMENTS).
 * The fused sm_100a kernels run when sample_stride == 1, frame_size is one of
 * {256, 512, 1024, 2048, 4096, 8192, 16384} and every frame start is 16-byte aligned; every other
 * shape (any length >= 1, any strides) runs the general kernel.  All run on the GPU; there is no CPU path.
 *
 * Precision classes vs the reference on complex128 input (tests/ hold the tolerances):
 *   relative 1e-9 : features 4, 6, 7, 8 and 10..18 (float64 accumulation)
 *   relative 1e-6 : features 1, 2, 3, 5, 9 (float32 FFT / atan2 in the fused kernel; unwrap branch
 *                   decisions within 4e-6 rad of +-pi are re-decided in float64 exactly as np.unwrap)
 *   AMC_FLAG_FORCE_GENERAL computes everything except the FFT in float64.
 * The fused kernels hold these classes for every frame because the few kinds of frame their float32 parts cannot
 * handle are detected from the frame's own sums and recomputed by the general (float64) kernel in a second, tiny
 * launch that follows every fused launch ("careful path": mean power outside [2^-100, 2^80*2048/N] - float32 would
 * under/overflow -, phase / |phase| / frequency spread below 0.1 rad - an unmodulated carrier or a dominant DC line,
 * where float32 phase noise would show -, and a degenerate |r - mean r| distribution for feature 4).  Supported
 * input range: any finite complex128 / complex64 values whose |x|^6 sums stay finite in float64 (the reference itself
 * overflows beyond that); a single sample below 1e-37 inside a frame of ordinary scale is treated as 0 by the
 * float32 phase path.
 * complex64 input is widened exactly and computed like complex128 (the reference's numpy computes complex64 frames in
 * float32 and differs from its own complex128 result by ~4e-7; this library matches the widened result at the
 * tolerances above and the float32-numpy result at 2e-4).
 * General-kernel spectrum: power-of-two sizes <= 16384 by a float32 radix-2 FFT; other sizes from 128 to 8192 by a
 * float32 Bluestein (chirp-z) FFT; smaller sizes and 8193..12288 by a float64 direct DFT; anything larger returns
 * AMC_ERR_UNSUPPORTED.  The float32 transforms run on the frame scaled by an exact power of two.
 */
#ifndef AMCPY_B200_H
#define AMCPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMC_N_FEATURES 18
#define AMC_N_MOMENTS 11
#define AMC_ALL_FEATURES 0x3FFFFu /* bit k = feature id k+1 */

/* iq_dtype */
#define AMC_C64 0  /* interleaved float32 re,im  */
#define AMC_C128 1 /* interleaved float64 re,im  */

/* flags */
#define AMC_FLAG_FORCE_GENERAL 1 /* never take the fused fixed-size kernel */
#define AMC_FLAG_DIRECT_DFT 8     /* non-power-of-two sizes: float64 direct DFT instead of the float32 Bluestein FFT (cross-check) */
/* Unknown flag bits are rejected with AMC_ERR_INVALID_ARG.  Bits 2 and 4 select A/B experiment kernels that exist
 * only in a library built with -DAMC_EXPERIMENTS (tools/exp/); they are not part of this ABI. */
#ifdef AMC_EXPERIMENTS
#define AMC_FLAG_FUSED_SPT8 2 /* round-1 kernel with 8 samples per thread */
#define AMC_FLAG_FUSED_WS 4   /* N = 2048: warp-specialised variant (FP64 warps / FP32 warps) */
#endif

/* return codes */
#define AMC_OK 0
#define AMC_ERR_INVALID_ARG (-1)
#define AMC_ERR_UNSUPPORTED (-2) /* frame_size outside what the general kernel's FFT supports */
#define AMC_ERR_CUDA (-3)
#define AMC_ERR_NO_DEVICE (-4)

/* Library / ABI version (major*1000 + minor). */
int amc_version(void);

/*
 * Optional eager initialisation of `device`: builds the twiddle tables of the fused kernels and waits for them.
 * Without it the first amc_extract_batch call on a device builds them on an internal stream and makes the caller's
 * stream wait for them on the device (cudaStreamWaitEvent; no host synchronisation) - which is not allowed while
 * that stream is being captured into a CUDA graph: call amc_init first in that case.
 */
int amc_init(int device);

/*
 * Device memory (bytes) the library itself allocates - once, cached or kept in its own stream-ordered memory pool - for
 * calls of this shape: Bluestein tables of a non-power-of-two frame_size, the FFT workspace of frames longer than
 * 16384 samples, the |x| scratch of the long-frame kernel (frame_size 8192 / 16384: N doubles per resident CTA) and,
 * when host_path != 0, the double-buffered chunk buffers of amc_extract_host / amc_extract_host_planar.  0 for the
 * device-pointer path at frame sizes 256..4096 (and any other power of two <= 4096): the caller owns every buffer
 * there.  Negative = AMC_ERR_*.
 */
int64_t amc_workspace_bytes(int iq_dtype, int64_t n_frames, int64_t frame_size, int host_path);

/* Message of the last failing call made by THIS thread ("" if none). Never NULL. */
const char* amc_last_error_string(void);

/* Number of CUDA devices visible, or AMC_ERR_NO_DEVICE. */
int amc_device_count(void);

/*
 * All 18 features of n_frames frames that already live in device memory (current device),
 * enqueued on `cuda_stream` (a cudaStream_t, may be NULL for the default stream); never synchronises the
 * host with the device.  The FIRST call on a device (twiddle tables, see amc_init) and the first call for a new
 * non-power-of-two frame size (Bluestein tables, <= 192 KB of device memory, cached per device and size: see
 * amc_workspace_bytes) build their tables on an internal stream; `cuda_stream` waits for them on the device.
 * A fused launch is followed by the careful-path launch described above (two kernels per call).
 *   iq           device pointer, complex64/complex128 interleaved
 *   out          device pointer, float64, row f at out + f*out_stride, out_stride >= 18
 *   feature_mask bit k = feature k+1 wanted; must be non-zero.  All 18 columns are always written: a wanted
 *                column holds the feature (bitwise the value an all-features call returns); an unwanted
 *                column holds either the feature or NaN - the library skips whole feature groups nobody
 *                asked for (FFT: 1; phase/frequency: 2,3,5,9; amplitude: 4,6,7,8; moments: 10..18) where a
 *                reduced kernel profile exists (frame sizes 256..4096).  AMC_ALL_FEATURES = the drop-in.
 * n_frames == 0 is a no-op.
 */
int amc_extract_batch(const void* iq, int iq_dtype, int64_t n_frames, int64_t frame_size,
                      int64_t frame_stride, int64_t sample_stride, double* out, int64_t out_stride,
                      uint32_t feature_mask, int flags, void* cuda_stream);

/*
 * Same computation for HOST buffers on device `device`: frames are copied to the GPU in
 * chunks on two streams (copy of chunk i+1 overlaps the kernel of chunk i), features are
 * copied back; returns when `out` is complete.  Pinned host memory gives full PCIe speed.
 * Accepts sample_stride == 1 (row per frame, any frame_stride >= frame_size) or the
 * sample-major layout frame_stride == 1 (what loadmat returns; sample_stride >= n_frames),
 * which is transposed on the device.
 */
int amc_extract_host(const void* iq, int iq_dtype, int64_t n_frames, int64_t frame_size,
                     int64_t frame_stride, int64_t sample_stride, double* out, int64_t out_stride,
                     uint32_t feature_mask, int flags, int device);

/*
 * Same for PLANAR sample-major host data: separate real and imaginary planes (float64 planes for
 * AMC_C128, float32 for AMC_C64; `im` may be NULL for real input), plane element (f, n) at
 * f + n*sample_stride (strides in REAL elements, sample_stride >= n_frames).  This is how a MATLAB
 * Level-5 .mat file stores the (n_snr, n_frames, n_samples) variables the reference slices
 * (feature_extraction.py:46-48, :68): the planes of a memory-mapped file go to the GPU as they are and
 * are interleaved + transposed there, instead of scipy.io.loadmat doing it on one host thread.
 */
int amc_extract_host_planar(const void* re, const void* im, int iq_dtype, int64_t n_frames, int64_t frame_size,
                            int64_t sample_stride, double* out, int64_t out_stride, uint32_t feature_mask,
                            int flags, int device);

/*
 * Device-side re-layout: src holds n_frames frames sample-major (element (f, n) at
 * f + n*src_sample_stride); dst receives one contiguous row of frame_size samples per frame.
 */
int amc_frames_from_sample_major(const void* src, int iq_dtype, int64_t n_frames, int64_t frame_size,
                                 int64_t src_sample_stride, void* dst, void* cuda_stream);

/*
 * InstantaneousValues for a batch (device pointers, float64 outputs, row-major):
 *   abs, phase, unwrapped, cn_amplitude : [n_frames, frame_size];  frequency : [n_frames, frame_size-1]
 * Any output pointer may be NULL to skip it.  float64 arithmetic, np.unwrap's rules.
 */
int amc_instantaneous_batch(const void* iq, int iq_dtype, int64_t n_frames, int64_t frame_size,
                            int64_t frame_stride, int64_t sample_stride, double* abs_out,
                            double* phase_out, double* unwrapped_out, double* frequency_out,
                            double* cn_amplitude_out, void* cuda_stream);

/*
 * MomentValues for a batch: out[f][2*i + {0,1}] = {re, im} of moment i in the order
 * m20 m21 m22 m40 m41 m42 m43 m60 m61 m62 m63 (m21, m42, m62 have im = 0 as in the reference).
 * out is device float64 [n_frames, 22].
 */
int amc_moments_batch(const void* iq, int iq_dtype, int64_t n_frames, int64_t frame_size,
                      int64_t frame_stride, int64_t sample_stride, double* out, void* cuda_stream);

/*
 * On-device synthetic IQ frames (no reference counterpart: the reference's data came from captures,
 * old/read_binary_stream.py:19-59; recipe of SURVEY.md section 8d).  Cell = one (modulation, SNR) pair:
 *   cell_mod[c]      0 BPSK, 1 QPSK, 2 8PSK, 3 16QAM, 4 64QAM, 5 WGN      (device int array, n_cells)
 *   cell_snr_idx[c]  SNR index used in the random-stream key               (device int array)
 *   cell_sigma[c]    noise standard deviation per rail                     (device double array)
 * out (device) receives [n_cells][frames_per_cell][frame_size] complex samples; frame numbering starts at
 * first_frame, so any shard of the frame axis regenerates exactly the frames of the full set
 * (Philox4x32-10 keyed by seed, counter = (sample, frame, snr index, modulation)).
 */
int amc_generate_frames(void* out, int iq_dtype, int n_cells, int64_t frames_per_cell, int64_t first_frame,
                        int64_t frame_size, const int* cell_mod, const int* cell_snr_idx, const double* cell_sigma,
                        uint64_t seed, void* cuda_stream);

/* How many kernels of this library the calling thread has launched so far (for bench accounting). */
int64_t amc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* AMCPY_B200_H */
