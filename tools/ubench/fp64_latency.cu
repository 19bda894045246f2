// Dependent-chain latencies (cycles) of the FP64 operations that sit on the critical path of potf2 / trsm:
// DFMA, DMUL, SHFL of a double, rsqrt(), 1.0/x, LDS.  One warp, clock64 around an unrolled dependent chain.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/fp64_latency tools/ubench/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 512;
template <int OP>
__global__ void chain(double* out, long long* cyc, double seed, double mulc) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = seed;
  __syncthreads();
  double x = seed + threadIdx.x * 1e-9, y = mulc;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, y, 1e-30);
    if (OP == 1) x = x * y;
    if (OP == 2) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    if (OP == 3) x = rsqrt(x) + 1.5;
    if (OP == 4) x = 1.0 / x + 0.5;
    if (OP == 5) x = sm[(__double2loint(x) & 7)] ;
    if (OP == 6) x = sqrt(x) + 2.0;
    if (OP == 7) { x = __shfl_sync(0xffffffffu, x * y, (threadIdx.x + 1) & 31); x = fma(-x, y, 1.0); }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name, double sub) {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
  for (int r = 0; r < 3; ++r) chain<OP><<<1, 32>>>(out, cyc, 1.25, 0.999999);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s %7.1f cycles per step (minus %.0f for the helper op)\n", name, (double)h / N, sub);
}
int main() {
  run<0>("DFMA dependent", 0);
  run<1>("DMUL dependent", 0);
  run<2>("SHFL.IDX of a double (2 x 32 bit)", 0);
  run<3>("rsqrt(double) + DADD", 0);
  run<4>("1.0 / x + DADD", 0);
  run<5>("LDS.64 dependent (+ address)", 0);
  run<6>("sqrt(double) + DADD", 0);
  run<7>("DMUL -> SHFL -> DFMA", 0);
  return 0;
}
