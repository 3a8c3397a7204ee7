// Isolated copy of the refold phase-B loop: cycles per element for a 512-element in-place fold.
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ void lds_v2f64(uint32_t a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts_v2f64(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__global__ void k(double* out, long long* cyc, double seed, uint32_t S, int mode) {
  __shared__ __align__(16) double B[1024];
  const int lane = threadIdx.x & 31;
  for (int rep = 0; rep < 3; ++rep) {
    for (uint32_t i = lane; i < 1024; i += 32) B[i] = seed * (1.0 + (i % (37 + mode * 1000)) * 0.37);
    __syncwarp();
    const uint32_t sB = (uint32_t)__cvta_generic_to_shared(B);
    const bool writer = lane == 0;
    double acc = 0.0;
    long long t0 = clock64();
    uint32_t i = 0;
    double w0, w1, w2, w3;
    lds_v2f64(sB, w0, w1);
    lds_v2f64(sB + 16u, w2, w3);
    for (; i + 4u <= S; i += 4u) {
      double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
      if (i + 8u <= S) {
        lds_v2f64(sB + i * 8u + 32u, n0, n1);
        lds_v2f64(sB + i * 8u + 48u, n2, n3);
      }
      const double p0 = acc + w0;
      const double p1 = p0 + w1;
      const double p2 = p1 + w2;
      acc = p2 + w3;
      if (writer) {
        sts_v2f64(sB + i * 8u, p0, p1);
        sts_v2f64(sB + i * 8u + 16u, p2, acc);
      }
      w0 = n0; w1 = n1; w2 = n2; w3 = n3;
    }
    long long t1 = clock64();
    cyc[rep] = t1 - t0;
    out[lane] = acc + B[lane];
    __syncwarp();
  }
}
int main() {
  double* d; long long* c;
  cudaMalloc(&d, 32 * 8); cudaMalloc(&c, 8 * 8);
  for (int mode = 0; mode < 2; ++mode)
    for (double seed : {1.5, 1e-310, 123.456}) {
      k<<<1, 32>>>(d, c, seed, 512, mode);
      long long h[3];
      cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
      printf("mode %d seed %g: %.2f %.2f %.2f cycles/element\n", mode, seed, h[0] / 512.0, h[1] / 512.0, h[2] / 512.0);
    }
  return 0;
}
