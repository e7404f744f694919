// Shared-memory wavefront counts of 64- / 128-bit loads with broadcast patterns (ncu: one kernel per pattern).
#include <cstdio>
#include <cuda_runtime.h>
__device__ double sink[64];
template <int P>
__global__ void probe(int iters, int stride_in) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double acc = 0.0;
  int active = 32, idx = 0;
  bool wide = false;
  const int stride = stride_in;   // runtime value: keeps the compiler from folding addresses
  if (P == 0) { active = 15; idx = (lane / 3) * (stride - 1); }               // LDS.64, 5 blocks stride 9
  if (P == 1) { active = 15; idx = (lane / 3) * stride; wide = true; }        // LDS.128, 5 blocks stride 10
  if (P == 2) { active = 32; idx = (lane / 3) * stride; wide = true; }        // LDS.128, 11 blocks
  if (P == 3) { active = 32; idx = lane * 2; wide = true; }                   // LDS.128, 32 distinct consecutive
  if (P == 4) { active = 16; idx = lane * 2; wide = true; }                   // LDS.128, 16 distinct consecutive
  if (P == 5) { active = 32; idx = (lane / 3) * (stride - 1); }               // LDS.64, 11 blocks stride 9
  if (P == 6) { active = 15; idx = lane * stride; wide = true; }              // LDS.128, 15 distinct stride 10
  if (P == 7) { active = 15; idx = lane * (stride - 1); }                     // LDS.64, 15 distinct stride 9
  if (P == 8) { active = 30; idx = (lane / 3) * stride; wide = true; }        // LDS.128, 10 blocks, 30 lanes
  if (P == 9) { active = 8; idx = lane * 2; wide = true; }                    // LDS.128, 8 distinct
  if (lane < active) {
    if (wide) {
      const unsigned base = (unsigned)__cvta_generic_to_shared(sm + idx);
      for (int i = 0; i < iters; ++i) {
        double x, y;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(base + 16u * (i & 3)));
        acc += x + y;
      }
    } else {
      const unsigned base = (unsigned)__cvta_generic_to_shared(sm + idx);
      for (int i = 0; i < iters; ++i) {
        double x;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(base + 8u * (i & 7)));
        acc += x;
      }
    }
  }
  sink[threadIdx.x & 63] = acc;
}
int main() {
  const int iters = 4096;
#define RUN(P) probe<P><<<1, 32, 4096 * 8>>>(iters, 10);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
  cudaError_t e = cudaDeviceSynchronize();
  printf("done %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
