// Variants of the in-place sequential fold loop: cycles per element (512 elements, one warp).
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ void lds_v2f64(uint32_t a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts_v2f64(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
#define LOAD8(v, o) { lds_v2f64((o), v[0], v[1]); lds_v2f64((o) + 16u, v[2], v[3]); lds_v2f64((o) + 32u, v[4], v[5]); lds_v2f64((o) + 48u, v[6], v[7]); }
#define CHAIN8(v, o) { \
    const double p0 = acc + v[0]; const double p1 = p0 + v[1]; if (writer) sts_v2f64((o), p0, p1); \
    const double p2 = p1 + v[2]; const double p3 = p2 + v[3]; if (writer) sts_v2f64((o) + 16u, p2, p3); \
    const double p4 = p3 + v[4]; const double p5 = p4 + v[5]; if (writer) sts_v2f64((o) + 32u, p4, p5); \
    const double p6 = p5 + v[6]; acc = p6 + v[7]; if (writer) sts_v2f64((o) + 48u, p6, acc); }

// V2: groups of 8, double buffer, clamped unconditional prefetch
__device__ __forceinline__ double fold_v2(uint32_t sa, uint32_t count, double acc, bool writer) {
  const uint32_t ngrp = count >> 3;
  if (ngrp) {
    const uint32_t last_g = sa + (ngrp - 1u) * 64u;
    uint32_t o = sa;
    double a[8], b[8];
    LOAD8(a, o);
    for (uint32_t g = 0;;) {
      { uint32_t on = min(o + 64u, last_g); LOAD8(b, on); }
      CHAIN8(a, o);
      if (++g == ngrp) break;
      o += 64u;
      { uint32_t on = min(o + 64u, last_g); LOAD8(a, on); }
      CHAIN8(b, o);
      if (++g == ngrp) break;
      o += 64u;
    }
  }
  return acc;
}
// V3: distance 2 (three buffers)
__device__ __forceinline__ double fold_v3(uint32_t sa, uint32_t count, double acc, bool writer) {
  const uint32_t ngrp = count >> 3;
  if (ngrp) {
    const uint32_t last_g = sa + (ngrp - 1u) * 64u;
    uint32_t o = sa;
    double a[8], b[8], c[8];
    LOAD8(a, o);
    { uint32_t on = min(o + 64u, last_g); LOAD8(b, on); }
    for (uint32_t g = 0;;) {
      { uint32_t on = min(o + 128u, last_g); LOAD8(c, on); }
      CHAIN8(a, o);
      if (++g == ngrp) break;
      o += 64u;
      { uint32_t on = min(o + 128u, last_g); LOAD8(a, on); }
      CHAIN8(b, o);
      if (++g == ngrp) break;
      o += 64u;
      { uint32_t on = min(o + 128u, last_g); LOAD8(b, on); }
      CHAIN8(c, o);
      if (++g == ngrp) break;
      o += 64u;
    }
  }
  return acc;
}
// V4: operands in registers (lane l holds element k*32+l), broadcast by shuffles, stores by the owner lane
__device__ __forceinline__ double fold_v4(double* B, uint32_t count, double acc, int lane) {
  for (uint32_t base = 0; base < count; base += 32) {
    const double w = B[base + lane];
    double mine = 0.0;
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      acc = acc + __shfl_sync(0xffffffffu, w, l);
      if (lane == l) mine = acc;
    }
    B[base + lane] = mine;
  }
  return acc;
}
__global__ void k(double* out, long long* cyc, double seed, uint32_t S) {
  __shared__ __align__(16) double B[1024];
  const int lane = threadIdx.x & 31;
  const uint32_t sB = (uint32_t)__cvta_generic_to_shared(B);
  for (int v = 0; v < 3; ++v) {
    for (int rep = 0; rep < 2; ++rep) {
      for (uint32_t i = lane; i < 1024; i += 32) B[i] = seed * (1.0 + (i % 37) * 0.37);
      __syncwarp();
      long long t0 = clock64();
      double acc = 0.0;
      if (v == 0) acc = fold_v2(sB, S, 0.0, lane == 0);
      else if (v == 1) acc = fold_v3(sB, S, 0.0, lane == 0);
      else acc = fold_v4(B, S, 0.0, lane);
      long long t1 = clock64();
      __syncwarp();
      cyc[v * 2 + rep] = t1 - t0;
      out[v * 64 + rep * 32 + lane] = acc + B[lane] + B[S - 1];
      __syncwarp();
    }
  }
}
int main() {
  double* d; long long* c;
  cudaMalloc(&d, 1024 * 8); cudaMalloc(&c, 8 * 8);
  k<<<1, 32>>>(d, c, 1.5, 512);
  long long h[6];
  double o[192];
  cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
  cudaMemcpy(o, d, sizeof o, cudaMemcpyDeviceToHost);
  printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
  const char* nm[] = {"V2 8-group double buffer", "V3 8-group distance 2", "V4 shuffle broadcast"};
  for (int v = 0; v < 3; ++v) printf("%-28s %.2f %.2f cycles/element  (check %.17g)\n", nm[v], h[v * 2] / 512.0, h[v * 2 + 1] / 512.0, o[v * 64]);
  return 0;
}
