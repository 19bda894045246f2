// FP64 micro-peaks used as roofline denominators for the Cholesky (DMMA) and for the ALU-bound
// kernels (DFMA).  MEASURED_PEAKS.json only carries HBM and bf16 numbers; SURVEY.md section 8d asks the
// build to measure these itself.
#include "tgp_common.cuh"

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (double)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 123.456) out[0] = s;  // never true; keeps the chain alive
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  const double av = a + threadIdx.x, bv = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(av), "d"(bv));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

__global__ void empty_kernel(double* out) { if (out == nullptr) return; }
__global__ void empty_pdl_kernel(double* out) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (out == nullptr) return;
}

// kind 2 / 3: average microseconds per launch of a chain of `iters` dependent empty kernels on one stream,
// plain (2) or with programmatic dependent launch (3).  Result returned through *tflops (as microseconds).
static int launch_chain_probe(int kind, int iters, double* out_us) {
  double* d = nullptr;
  TGP_CUDA(cudaMalloc(&d, 8));
  cudaStream_t st;
  TGP_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaEvent_t e0, e1;
  TGP_CUDA(cudaEventCreate(&e0));
  TGP_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    TGP_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) {
      if (kind == 2) {
        empty_kernel<<<8, 256, 0, st>>>(d);
      } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(8); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        TGP_CUDA(cudaLaunchKernelEx(&cfg, empty_pdl_kernel, d));
      }
    }
    TGP_CUDA(cudaEventRecord(e1, st));
    TGP_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    TGP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  TGP_LAUNCH_CHECK();
  *out_us = (double)best * 1e3 / iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(st); cudaFree(d);
  return TGP_OK;
}

extern "C" int tgp_microbench_fp64(int kind, int iters, double* tflops) {
  if (kind == 2 || kind == 3) {
    TGP_CHECK_ARG(tflops && iters > 0, "iters");
    return launch_chain_probe(kind, iters, tflops);
  }
  TGP_CHECK_ARG(tflops && iters > 0 && (kind == 0 || kind == 1), "kind/iters");
  double* d = nullptr;
  TGP_CUDA(cudaMalloc(&d, 8));
  const int grid = tgp_num_sms() * 8, block = 256;
  cudaEvent_t e0, e1;
  TGP_CUDA(cudaEventCreate(&e0));
  TGP_CUDA(cudaEventCreate(&e1));
  float best_ms = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    TGP_CUDA(cudaEventRecord(e0));
    if (kind == 0) dfma_peak_kernel<<<grid, block>>>(d, iters, 1.0000001, 1e-9);
    else dmma_peak_kernel<<<grid, block>>>(d, iters, 1e-3, 1e-3);
    TGP_CUDA(cudaEventRecord(e1));
    TGP_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    TGP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best_ms) best_ms = ms;
  }
  TGP_LAUNCH_CHECK();
  // flops per thread-iteration: DFMA 16 FMAs = 32 flop; DMMA: 16 mma x (8*8*4*2 flop / 32 lanes) = 16*16 flop
  const double per_thread = (kind == 0) ? 32.0 : 256.0;
  *tflops = per_thread * (double)iters * (double)grid * (double)block / ((double)best_ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return TGP_OK;
}
