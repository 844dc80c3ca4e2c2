// Second pipe microbenchmark: dispatch cost of FP32 forms and what co-issues with FP64.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
enum { K_FADD, K_FMUL, K_FFMA3, K_FFMAIMM, K_FADD2, K_DFMA, K_DFMA_IADD, K_DFMA_LOP, K_DFMA_FADD, K_DFMA_LDS, K_DFMA_SHFL, K_FADD_IADD, K_COUNT };
const char* names[] = {"FADD r,r", "FMUL r,r", "FFMA r,r,r (distinct)", "FFMA r,imm,r", "FADD2", "DFMA", "DFMA+IADD3", "DFMA+LOP3", "DFMA+FADD", "DFMA+LDS.64", "DFMA+SHFL", "FADD+IADD3"};
template <int K>
__global__ void kern(float* out, unsigned long long* cyc, float seed) {
  __shared__ double sm[1024];
  float a[8], b[8], c[8]; double d[8]; int n[8]; unsigned long long p[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; b[i] = seed * 0.5f + i; c[i] = seed * 0.25f - i; d[i] = seed + i * 0.5 + threadIdx.x; n[i] = threadIdx.x + i; p[i] = 0x3f8000003f800000ull + i; }
  sm[threadIdx.x] = d[0];
  __syncthreads();
  const double dm = 1.0000001, dc = 0.25;
  unsigned long long pm = 0x3f8000013f800001ull;
  unsigned long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (K == K_FADD) a[i] = a[i] + b[i];
      if (K == K_FMUL) a[i] = a[i] * b[i];
      if (K == K_FFMA3) a[i] = fmaf(a[i], b[i], c[i]);
      if (K == K_FFMAIMM) a[i] = fmaf(a[i], 1.0001f, c[i]);
      if (K == K_FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pm));
      if (K == K_DFMA) d[i] = fma(d[i], dm, dc);
      if (K == K_DFMA_IADD) { d[i] = fma(d[i], dm, dc); n[i] = n[i] + n[(i + 1) & 7] + it; }
      if (K == K_DFMA_LOP) { d[i] = fma(d[i], dm, dc); n[i] = (n[i] ^ n[(i + 1) & 7]) | it; }
      if (K == K_DFMA_FADD) { d[i] = fma(d[i], dm, dc); a[i] = a[i] + b[i]; }
      if (K == K_DFMA_LDS) { d[i] = fma(d[i], dm, dc); b[i] += (float)sm[(n[i] + it) & 1023]; }
      if (K == K_DFMA_SHFL) { d[i] = fma(d[i], dm, dc); a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 15)); }
      if (K == K_FADD_IADD) { a[i] = a[i] + b[i]; n[i] = n[i] + n[(i + 1) & 7] + it; }
    }
  }
  unsigned long long t1 = clock64();
  float acc = 0; for (int i = 0; i < 8; ++i) acc += a[i] + b[i] + (float)d[i] + n[i] + (float)(p[i] & 0xffff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int K> void run(int warps) {
  float* out; unsigned long long* cyc; int blocks = 148;
  cudaMalloc(&out, blocks * 1024 * 4); cudaMalloc(&cyc, blocks * 8);
  for (int r = 0; r < 2; ++r) { kern<K><<<blocks, warps * 32>>>(out, cyc, 1.0f); cudaDeviceSynchronize(); }
  unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
  double groups = (double)ITERS * 8 * warps;      // warp-level groups per SM
  printf("%-24s warps/SM=%2d  cycles per group per SMSP = %5.2f\n", names[K], warps, avg / (groups / 4));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {12, 32}) {
    run<K_FADD>(w); run<K_FMUL>(w); run<K_FFMA3>(w); run<K_FFMAIMM>(w); run<K_FADD2>(w); run<K_DFMA>(w);
    run<K_DFMA_IADD>(w); run<K_DFMA_LOP>(w); run<K_DFMA_FADD>(w); run<K_DFMA_LDS>(w); run<K_DFMA_SHFL>(w); run<K_FADD_IADD>(w);
    printf("\n");
  }
  return 0;
}
