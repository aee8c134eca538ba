// Micro-benchmark of the edge kernel's gate math in isolation (no TMEM, no memory): how many cycles does one "step"
// (COLS columns x r, z, n gates of one row per thread) take with W warps per scheduler?  nvcc -arch=sm_100a -O3 gate_math.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ f32x2 ex2_2(f32x2 v) { float a, b; up2(v, a, b); return pk2(ex2a(a), ex2a(b)); }
__device__ __forceinline__ f32x2 rcp_2(f32x2 v) { float a, b; up2(v, a, b); return pk2(rcpa(a), rcpa(b)); }

template <int PAIRS>   // PAIRS packed pairs per step: 2 = 4 columns (the shipped kernel), 4 = 8 columns, 8 = 16 columns
__global__ void k(float* out, long long* cyc, int iters, float seed) {
  const f32x2 NL = pk2(-1.4426950f, -1.4426950f), TL = pk2(2.885390f, 2.885390f), ONE = pk2(1.f, 1.f), NTWO = pk2(-2.f, -2.f), NONE = pk2(-1.f, -1.f);
  f32x2 ar[PAIRS], az[PAIRS], an[PAIRS], ah[PAIRS], hp[PAIRS];
  for (int i = 0; i < PAIRS; ++i) {
    const float b = seed * (threadIdx.x + 1) * 1e-3f + i * 0.01f;
    ar[i] = pk2(b, -b); az[i] = pk2(0.3f + b, 0.1f - b); an[i] = pk2(b * 0.5f, b * 0.25f); ah[i] = pk2(-b, b * 2.f); hp[i] = pk2(b, b);
  }
  f32x2 dot = 0ull;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int e = 0; e < PAIRS; ++e) {
      const f32x2 rg = rcp_2(add2(ex2_2(fma2(ar[e], NL, hp[e])), ONE));
      const f32x2 zg = rcp_2(add2(ex2_2(fma2(az[e], NL, hp[e])), ONE));
      const f32x2 u = fma2(rg, add2(ah[e], hp[e]), add2(an[e], hp[e]));
      const f32x2 ng = fma2(rcp_2(add2(ex2_2(mul2(u, TL)), ONE)), NTWO, ONE);
      const f32x2 ov = fma2(zg, fma2(ng, NONE, hp[e]), ng);
      dot = fma2(ov, ONE, dot);
      hp[e] = ov;               // next iteration depends on this one like the next tile does not: worst case for latency
      ar[e] = add2(ar[e], ov);  // keep the inputs moving
    }
  }
  const long long t1 = clock64();
  float a, b; up2(dot, a, b);
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int PAIRS> void run(int warps_per_cta, const char* name) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<PAIRS><<<148, 32 * warps_per_cta>>>(out, cyc, iters, 1.0f);
  k<PAIRS><<<148, 32 * warps_per_cta>>>(out, cyc, iters, 1.1f);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_step = (double)c / iters;
  printf("%s: %2d warps/CTA (%.1f per scheduler), %d columns/step: %.0f cycles per step, %.1f cycles per column-step, MUFU pipe %.0f %%\n", name,
         warps_per_cta, warps_per_cta / 4.0, 2 * PAIRS, per_step, per_step / (2 * PAIRS), 100.0 * (warps_per_cta / 4.0) * (2 * PAIRS * 6 * 8) / per_step);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 16, 32}) { run<2>(w, "4col"); run<4>(w, "8col"); run<8>(w, "16col"); }
  return 0;
}
