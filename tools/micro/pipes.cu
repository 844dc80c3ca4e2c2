// Pipe-throughput microbenchmarks for B200 (sm_100a): how many lane-ops per clock per SM each
// instruction class sustains, alone and mixed.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
enum { K_FFMA, K_FFMA2, K_DFMA, K_DADD, K_F2F, K_RSQ64, K_SHFL, K_MIX_DF, K_MIX_DF2, K_FADD2, K_LDS128, K_MUFU_RCP, K_MIX3, K_COUNT };
const char* names[] = {"FFMA", "FFMA2(x2 flop)", "DFMA", "DADD", "F2F.F32.F64", "MUFU.RSQ64H", "SHFL.BFLY", "DFMA+FFMA (1:1)", "DFMA+FFMA2 (1:1)", "FADD2", "LDS.128", "MUFU.RCP", "DFMA+2FFMA+1ALU"};

template <int K>
__global__ void kern(float* out, unsigned long long* cyc, float seed) {
  __shared__ float4 sm[1024];
  float a[8]; double d[8]; unsigned long long p[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; d[i] = seed + i * 0.5 + threadIdx.x; p[i] = (unsigned long long)(threadIdx.x + i) * 0x100000001ull + 0x3f8000003f800000ull; }
  sm[threadIdx.x] = make_float4(a[0], a[1], a[2], a[3]);
  __syncthreads();
  const float m = 1.0001f, c = 0.5f; const double dm = 1.0000001, dc = 0.25;
  unsigned long long pm = 0x3f8000013f800001ull;
  int ia = threadIdx.x;
  unsigned long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (K == K_FFMA) a[i] = fmaf(a[i], m, c);
      if (K == K_FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pm));
      if (K == K_FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pm));
      if (K == K_DFMA) d[i] = fma(d[i], dm, dc);
      if (K == K_DADD) d[i] = d[i] + dc;
      if (K == K_F2F) { float f; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(d[i])); a[i] += f; }
      if (K == K_RSQ64) { asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(d[i])); }
      if (K == K_MUFU_RCP) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
      if (K == K_SHFL) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 15));
      if (K == K_MIX_DF) { d[i] = fma(d[i], dm, dc); a[i] = fmaf(a[i], m, c); }
      if (K == K_MIX_DF2) { d[i] = fma(d[i], dm, dc); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pm)); }
      if (K == K_MIX3) { d[i] = fma(d[i], dm, dc); a[i] = fmaf(a[i], m, c); a[(i + 1) & 7] = fmaf(a[(i+1)&7], c, m); ia = (ia ^ (ia >> 3)) + i; }
      if (K == K_LDS128) { float4 v = sm[(ia + i * 32) & 1023]; a[i] += v.x; ia += (int)v.y & 1; }
    }
  }
  unsigned long long t1 = clock64();
  float acc = 0; for (int i = 0; i < 8; ++i) acc += a[i] + (float)d[i] + (float)(p[i] & 0xffff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + ia;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K>
void run(int warps) {
  float* out; unsigned long long* cyc;
  int blocks = 148;
  cudaMalloc(&out, blocks * 1024 * 4); cudaMalloc(&cyc, blocks * 8);
  kern<K><<<blocks, warps * 32>>>(out, cyc, 1.0f);
  cudaDeviceSynchronize();
  kern<K><<<blocks, warps * 32>>>(out, cyc, 1.0f);
  cudaDeviceSynchronize();
  unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
  double ops = (double)ITERS * 8 * warps * 32;   // lane-level "op groups" per SM
  printf("%-22s warps/SM=%2d  cycles=%9.0f  groups/clk/SM=%7.2f\n", names[K], warps, avg, ops / avg);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16, 32}) {
    run<K_FFMA>(w); run<K_FFMA2>(w); run<K_FADD2>(w); run<K_DFMA>(w); run<K_DADD>(w); run<K_F2F>(w); run<K_RSQ64>(w); run<K_MUFU_RCP>(w);
    run<K_SHFL>(w); run<K_LDS128>(w); run<K_MIX_DF>(w); run<K_MIX_DF2>(w); run<K_MIX3>(w);
    printf("\n");
  }
  return 0;
}
