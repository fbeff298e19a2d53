// Microbenchmark of the register/shared-memory Stockham FFT core (torch-cfd_b200/csrc/fft_core.cuh):
// cycles per 512-point complex FFT per SM for scalar vs packed (f32x2) lanes, CTA-wide vs
// per-group named barriers, and different occupancies.  Data stay in registers between
// transforms, so this is the ceiling of the FFT part of the NS kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I torch-cfd_b200/csrc -o fft_bench fft_bench.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "fft_core.cuh"
using namespace tcfd;

struct CtaSync2 { TCFD_D void operator()() const { __syncthreads(); } };
template <int NTHREADS>
struct GroupSync {
  int id;
  TCFD_D void operator()() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NTHREADS) : "memory"); }
};

template <class T, int N, int G, int V, bool PP, bool NAMED, int MINB>
__global__ void __launch_bounds__(G*(N / 8), MINB)
fft_loop(const cx<float>* __restrict__ in, cx<float>* __restrict__ out, const cx<float>* __restrict__ table, int reps) {
  constexpr int NT = N / 8;
  constexpr int W = lane_traits<T>::width;
  constexpr int HALF = V * N;
  constexpr int BUF = HALF * (PP ? 2 : 1);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cx<T>* smem = reinterpret_cast<cx<T>*>(smem_raw);
  const int g = threadIdx.x / NT, t = threadIdx.x % NT;
  cx<T>* buf = smem + g * BUF;
  FftTwiddles<float, N> tw;
  tw.load(table, t);
  cx<T> v[V][8];
  const cx<float>* src = in + ((size_t)(blockIdx.x * G + g) * V * W) * N;
#pragma unroll
  for (int j = 0; j < V; ++j)
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      if constexpr (W == 2) {
        cx<float> a = src[(2 * j) * N + t + m * NT], b = src[(2 * j + 1) * N + t + m * NT];
        v[j][m] = cx<f2>{f2(a.x, b.x), f2(a.y, b.y)};
      } else {
        v[j][m] = src[j * N + t + m * NT];
      }
    }
  int parity = 0;
  const float sc = 1.0f / N;
  for (int r = 0; r < reps; ++r) {
    if constexpr (NAMED) {
      GroupSync<NT> sync{g + 1};
      fft_run<T, N, -1, V, PP, HALF>(v, tw, buf, parity, t, sync);
      fft_run<T, N, +1, V, PP, HALF>(v, tw, buf, parity, t, sync);
    } else {
      CtaSync2 sync;
      fft_run<T, N, -1, V, PP, HALF>(v, tw, buf, parity, t, sync);
      fft_run<T, N, +1, V, PP, HALF>(v, tw, buf, parity, t, sync);
    }
#pragma unroll
    for (int j = 0; j < V; ++j)
#pragma unroll
      for (int m = 0; m < 8; ++m) v[j][m] = cx<T>{v[j][m].x * sc, v[j][m].y * sc};
  }
  cx<float>* dst = out + ((size_t)(blockIdx.x * G + g) * V * W) * N;
#pragma unroll
  for (int j = 0; j < V; ++j)
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      if constexpr (W == 2) {
        dst[(2 * j) * N + t + m * NT] = cx<float>{v[j][m].x.lo, v[j][m].y.lo};
        dst[(2 * j + 1) * N + t + m * NT] = cx<float>{v[j][m].x.hi, v[j][m].y.hi};
      } else {
        dst[j * N + t + m * NT] = v[j][m];
      }
    }
}

template <class T, int N, int G, int V, bool PP, bool NAMED, int MINB>
void run(const char* name, const cx<float>* in, cx<float>* out, const cx<float>* table, int ctas_per_sm) {
  constexpr int W = lane_traits<T>::width;
  auto k = fft_loop<T, N, G, V, PP, NAMED, MINB>;
  size_t smem = (size_t)G * V * N * (PP ? 2 : 1) * sizeof(cx<T>);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, G * N / 8, smem);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
  int grid = 148 * (ctas_per_sm > 0 ? ctas_per_sm : occ);
  int reps = 200;
  k<<<grid, G * N / 8, smem>>>(in, out, table, 2);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<<<grid, G * N / 8, smem>>>(in, out, table, reps);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ffts = (double)grid * G * V * W * 2.0 * reps;
  double cyc_per_fft_sm = ms * 1e-3 * 1.965e9 / (ffts / 148.0);
  // check: in -> out should be identity (fwd then inv scaled)
  std::vector<cx<float>> hi(N), ho(N);
  cudaMemcpy(hi.data(), in, N * sizeof(cx<float>), cudaMemcpyDeviceToHost);
  cudaMemcpy(ho.data(), out, N * sizeof(cx<float>), cudaMemcpyDeviceToHost);
  double err = 0; for (int i = 0; i < N; ++i) err = fmax(err, fabs(hi[i].x - ho[i].x) + fabs(hi[i].y - ho[i].y));
  printf("%-28s regs=%3d occ=%d ctas/sm=%d smem=%6zu  %.3f ms  %.1f cyc/FFT/SM  (%.2f GFFT/s) err=%.2e %s\n", name, fa.numRegs, occ,
         grid / 148, smem, ms, cyc_per_fft_sm, ffts / ms * 1e-6, err, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  constexpr int N = 512;
  size_t nfft = (size_t)148 * 8 * 4 * 4 * 2;  // upper bound of transforms touched
  cx<float>* in; cx<float>* out; cx<float>* table;
  cudaMalloc(&in, nfft * N * sizeof(cx<float>)); cudaMalloc(&out, nfft * N * sizeof(cx<float>));
  std::vector<cx<float>> h(nfft * N);
  for (size_t i = 0; i < h.size(); ++i) h[i] = cx<float>{(float)((i * 2654435761u) % 1000) / 1000.f - 0.5f, (float)((i * 40503u) % 1000) / 1000.f - 0.5f};
  cudaMemcpy(in, h.data(), h.size() * sizeof(cx<float>), cudaMemcpyHostToDevice);
  std::vector<cx<float>> tw(N);
  for (int j = 0; j < N; ++j) tw[j] = cx<float>{(float)cos(-2 * M_PI * j / N), (float)sin(-2 * M_PI * j / N)};
  cudaMalloc(&table, N * sizeof(cx<float>)); cudaMemcpy(table, tw.data(), N * sizeof(cx<float>), cudaMemcpyHostToDevice);
  //            T     N  G  V  PP    NAMED MINB
  run<float, N, 4, 1, true, false, 1>("scalar V1 cta-sync", in, out, table, 0);
  run<float, N, 4, 1, true, true, 1>("scalar V1 named", in, out, table, 0);
  run<float, N, 4, 2, true, false, 1>("scalar V2 cta-sync", in, out, table, 0);
  run<float, N, 4, 2, true, true, 1>("scalar V2 named", in, out, table, 0);
  run<float, N, 4, 2, false, true, 1>("scalar V2 named noPP", in, out, table, 0);
  run<f2, N, 4, 1, true, false, 1>("packed V1 cta-sync", in, out, table, 0);
  run<f2, N, 4, 1, true, true, 1>("packed V1 named", in, out, table, 0);
  run<f2, N, 4, 1, false, true, 1>("packed V1 named noPP", in, out, table, 0);
  run<f2, N, 2, 1, true, true, 1>("packed V1 named G2", in, out, table, 0);
  run<f2, N, 4, 1, true, true, 3>("packed V1 named minb3", in, out, table, 0);
  run<f2, N, 4, 1, true, true, 4>("packed V1 named minb4", in, out, table, 0);
  run<f2, N, 4, 2, true, true, 1>("packed V2 named", in, out, table, 0);
  run<float, N, 4, 1, true, true, 4>("scalar V1 named minb4", in, out, table, 0);
  run<float, N, 4, 1, true, true, 6>("scalar V1 named minb6", in, out, table, 0);
  run<float, N, 4, 2, true, true, 3>("scalar V2 named minb3", in, out, table, 0);
  for (int c : {1, 2}) run<f2, N, 4, 1, true, true, 1>("packed V1 named (limited)", in, out, table, c);
  return 0;
}
