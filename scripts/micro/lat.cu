// Dependent-chain latencies on one warp (debug aid for DESIGN.md §2): cycles per op.
#include <cstdio>
#include <cstdint>
#define N 2048
__global__ void k(double* out, long long* cyc, double seed, unsigned long long iseed) {
  double a = seed, b = seed * 1e-9;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = a + b;
  long long t1 = clock64();
  cyc[0] = t1 - t0;
  unsigned long long x = iseed, y = iseed >> 3;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { x = x + y; x ^= (x >> 7); }
  t1 = clock64();
  cyc[1] = t1 - t0;
  unsigned long long z = iseed;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) z = z + y;
  t1 = clock64();
  cyc[2] = t1 - t0;
  unsigned int s = (unsigned)iseed;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) s = __shfl_sync(0xffffffffu, s, (s + 1) & 31) + 1;
  t1 = clock64();
  cyc[3] = t1 - t0;
  double m = seed;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) m = m * 1.0000001;
  t1 = clock64();
  cyc[4] = t1 - t0;
  // 4 independent DADD chains
  double c0 = seed, c1 = seed + 1, c2 = seed + 2, c3 = seed + 3;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) { c0 += b; c1 += b; c2 += b; c3 += b; }
  t1 = clock64();
  cyc[5] = t1 - t0;
  float f = (float)seed, g = 1e-7f;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) f = f + g;
  t1 = clock64();
  cyc[6] = t1 - t0;
  out[threadIdx.x] = a + (double)x + (double)z + s + m + c0 + c1 + c2 + c3 + f;
}
int main() {
  double* d; long long* c;
  cudaMalloc(&d, 32 * 8); cudaMalloc(&c, 8 * 8);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(d, c, 1.5, 0x123456789abcdefull);
  long long h[8];
  cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
  const char* nm[] = {"DADD dep", "IADD64+xor-shift dep", "IADD64 dep", "SHFL+IADD dep", "DMUL dep", "4x DADD indep (per 4)", "FADD dep"};
  for (int i = 0; i < 7; ++i) printf("%-24s %.2f cycles/op\n", nm[i], (double)h[i] / N);
  return 0;
}
