// Microbenchmark: issue throughput of scalar vs packed (f32x2) FP32 instructions on sm_100a, and
// shared-memory exchange bandwidth for 8-byte / 16-byte accesses.  Decides how the FFT butterflies
// are written (fft_core.cuh).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipe fp32_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
#define ITER 4096
#define NCH 8

template <int MODE>
__global__ void __launch_bounds__(256) pipe_kernel(float* out, float seed) {
  // NCH independent dependency chains per thread
  float a[NCH], b[NCH];
  u64 p[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    a[i] = seed + i + threadIdx.x;
    b[i] = seed * 0.5f + i;
    float2 t = make_float2(a[i], b[i]);
    p[i] = *reinterpret_cast<u64*>(&t);
  }
  float c0 = seed * 1.0001f, c1 = seed * 0.9999f;
  float2 ct = make_float2(c0, c1);
  u64 pc = *reinterpret_cast<u64*>(&ct);
  u64 pd = pc ^ 0x0000100000001000ull;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (MODE == 0) { a[i] = fmaf(a[i], c0, c1); b[i] = fmaf(b[i], c1, c0); }            // 2 FFMA (3 distinct regs)
      if (MODE == 1) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc), "l"(pd)); }  // 1 FFMA2
      if (MODE == 2) { a[i] = a[i] + c0; b[i] = b[i] + c1; }                              // 2 FADD
      if (MODE == 3) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc)); }  // 1 FADD2
      if (MODE == 4) { a[i] = a[i] * c0; b[i] = b[i] * c1; }                              // 2 FMUL
      if (MODE == 5) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc)); }  // 1 FMUL2
      if (MODE == 6) { a[i] = fmaf(a[i], c0, b[i]); b[i] = fmaf(b[i], c1, a[i]); }         // dependent FFMA mix
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    float2 t = *reinterpret_cast<float2*>(&p[i]);
    s += a[i] + b[i] + t.x + t.y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory exchange: every thread writes W words then reads W words per iteration
template <int BYTES>
__global__ void __launch_bounds__(256) smem_kernel(float* out, int iters) {
  extern __shared__ __align__(16) unsigned char sm[];
  float acc = 0.f;
  const int t = threadIdx.x;
  if (BYTES == 8) {
    float2* s = reinterpret_cast<float2*>(sm);
    float2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = make_float2(t + k, t - k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) s[t + 256 * k] = v[k];
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = s[((t + 32 * k + it) & 255) + 256 * k];
      __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].y;
  } else {
    float4* s = reinterpret_cast<float4*>(sm);
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = make_float4(t + k, t - k, t, k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) s[t + 256 * k] = v[k];
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = s[((t + 32 * k + it) & 255) + 256 * k];
      __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
  }
  out[blockIdx.x * blockDim.x + t] = acc;
}

template <int MODE>
void run_pipe(const char* name, int flops_per_inner, float* d, int ctas_per_sm) {
  int sms = 148;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  pipe_kernel<MODE><<<sms * ctas_per_sm, 256>>>(d, 1.0f);
  cudaEventRecord(e0);
  pipe_kernel<MODE><<<sms * ctas_per_sm, 256>>>(d, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double elems = (double)sms * ctas_per_sm * 256 * (double)ITER * NCH * 2;  // fp32 element-ops
  printf("%-22s ctas/sm=%d  %.3f ms  %.2f T elem-op/s  (%.2f TFLOP/s)\n", name, ctas_per_sm, ms, elems / ms * 1e-9,
         elems * flops_per_inner / ms * 1e-9);
}

int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4 * 4);
  for (int c : {2, 4, 8}) {
    run_pipe<0>("FFMA  (scalar)", 2, d, c);
    run_pipe<1>("FFMA2 (packed)", 2, d, c);
    run_pipe<2>("FADD  (scalar)", 1, d, c);
    run_pipe<3>("FADD2 (packed)", 1, d, c);
    run_pipe<4>("FMUL  (scalar)", 1, d, c);
    run_pipe<5>("FMUL2 (packed)", 1, d, c);
    run_pipe<6>("FFMA dependent", 2, d, c);
  }
  for (int ctas : {1, 2, 4}) {
    for (int bytes : {8, 16}) {
      int iters = 2000;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      size_t sm = 256 * 8 * 8;
      if (bytes == 8) smem_kernel<8><<<148 * ctas, 256, sm>>>(d, 10); else smem_kernel<16><<<148 * ctas, 256, sm>>>(d, 10);
      cudaEventRecord(e0);
      if (bytes == 8) smem_kernel<8><<<148 * ctas, 256, sm>>>(d, iters); else smem_kernel<16><<<148 * ctas, 256, sm>>>(d, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double bytes_moved = (double)148 * ctas * iters * 2.0 * sm;  // write + read
      printf("smem exchange %2dB ctas/sm=%d: %.3f ms  %.1f B/clk/SM @1.9GHz  (%.2f TB/s chip)\n", bytes, ctas, ms,
             bytes_moved / (ms * 1e-3) / 148 / 1.9e9, bytes_moved / ms * 1e-9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
