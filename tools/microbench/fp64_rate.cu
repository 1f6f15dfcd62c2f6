// FP64 issue-rate probe for sm_100a: DFMA (scalar) vs DMMA (mma.sync.m8n8k4.f64) FMA throughput, with the operand
// conversion the FLAME decode needs (fp32 shared-memory operand -> fp64).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

template <int MODE>  // 0: DFMA x24 per (3 cvt), 1: DMMA, operands in registers, 2: DMMA with LDS.32 + cvt per B fragment
__global__ void __launch_bounds__(256) probe(double* out, int iters, const float* src) {
  __shared__ float sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = src[i];
  __syncthreads();
  double acc[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) acc[i] = 0.0;
  const int lane = threadIdx.x & 31;
  double a = 1.0 + 1e-9 * lane, b = 1.0 - 1e-9 * lane;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      const double s0 = static_cast<double>(sm[(it * 3 + lane) & 2047]), s1 = static_cast<double>(sm[(it * 3 + 1 + lane) & 2047]),
                   s2 = static_cast<double>(sm[(it * 3 + 2 + lane) & 2047]);
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        acc[3 * h] = fma(s0, a, acc[3 * h]);
        acc[3 * h + 1] = fma(s1, a, acc[3 * h + 1]);
        acc[3 * h + 2] = fma(s2, b, acc[3 * h + 2]);
      }
    } else if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        double d[2] = {acc[2 * j], acc[2 * j + 1]};
        dmma884(d, a, b);
        acc[2 * j] = d[0]; acc[2 * j + 1] = d[1];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const double bb = static_cast<double>(sm[(it * 12 + j * 32 + lane) & 2047]);
        double d[2] = {acc[2 * j], acc[2 * j + 1]};
        dmma884(d, a, bb);
        acc[2 * j] = d[0]; acc[2 * j + 1] = d[1];
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, double fma_per_thread_iter, int ctas_per_sm) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * ctas_per_sm, iters = 20000;
  double* out; float* src;
  cudaMalloc(&out, sizeof(double) * grid * 256);
  cudaMalloc(&src, sizeof(float) * 2048);
  cudaMemset(src, 0, sizeof(float) * 2048);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<grid, 256>>>(out, 100, src);
  cudaEventRecord(e0);
  probe<MODE><<<grid, 256>>>(out, iters, src);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fmas = fma_per_thread_iter * iters * 256.0 * grid;
  printf("%-44s ctas/sm %d: %.3f ms, %.2f TFMA/s (%.1f TFLOP/s), %.1f FMA/clk/SM at 1.9 GHz, err=%s\n", name, ctas_per_sm, ms, fmas / ms / 1e9,
         2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / sms / 1.9e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(src);
}

int main() {
  for (int c : {1, 2, 3}) {
    run<0>("DFMA 24/iter + 3 LDS + 3 cvt", 24.0, c);
    run<1>("DMMA m8n8k4 x12/iter (register operands)", 12.0 * 8.0, c);   // 256 FMA per warp-instr = 8 per thread
    run<2>("DMMA m8n8k4 x12/iter + LDS.32 + cvt each", 12.0 * 8.0, c);
  }
  return 0;
}
