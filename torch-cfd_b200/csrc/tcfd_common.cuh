// Common definitions for the torch-cfd B200 kernels.
#pragma once
#ifdef TCFD_EMU
#include "tcfd_emu.h"
#else
#include <cuda_runtime.h>
#include <cstdint>
#define TCFD_HD __host__ __device__ __forceinline__
#define TCFD_D __device__ __forceinline__
#define TCFD_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define TCFD_LAUNCH3(kernel, gx, gy, gz, block, smem, stream, ...) \
  kernel<<<dim3((gx), (gy), (gz)), (block), (smem), (stream)>>>(__VA_ARGS__)
#define TCFD_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

namespace tcfd {

template <class T>
struct alignas(2 * sizeof(T)) cx {
  T x, y;
};
template <class T> TCFD_D cx<T> operator+(cx<T> a, cx<T> b) { return cx<T>{a.x + b.x, a.y + b.y}; }
template <class T> TCFD_D cx<T> operator-(cx<T> a, cx<T> b) { return cx<T>{a.x - b.x, a.y - b.y}; }
template <class T> TCFD_D cx<T> operator*(T s, cx<T> a) { return cx<T>{s * a.x, s * a.y}; }
template <class T> TCFD_D cx<T> conj(cx<T> a) { return cx<T>{a.x, -a.y}; }

// explicit fused multiply-add (the build disables implicit contraction, -fmad=false)
TCFD_HD float fma_rn(float a, float b, float c) { return fmaf(a, b, c); }
TCFD_HD double fma_rn(double a, double b, double c) { return fma(a, b, c); }

// reciprocal of the Crank-Nicolson denominator 1 - mu L (>= 1).  fp32: hardware approximation refined by one
// Newton step (3 instructions, < 1 ulp -- far inside the 1e-3 / 5e-6 parity bars) instead of the ~10
// instructions of the correctly rounded __frcp_rn, 20 of which a rows unit needs; fp64: correctly rounded.
#ifndef TCFD_EMU
TCFD_D float rcp_cn(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.0f), r);
}
TCFD_D double rcp_cn(double x) { return __drcp_rn(x); }
#else
TCFD_D float rcp_cn(float x) { return 1.0f / x; }
TCFD_D double rcp_cn(double x) { return 1.0 / x; }
#endif
// correctly rounded reciprocal (== 1/x in IEEE arithmetic, without the generic division's slow path)
#ifndef TCFD_EMU
TCFD_D float rcp_rn(float x) { return __frcp_rn(x); }
TCFD_D double rcp_rn(double x) { return __drcp_rn(x); }
#else
TCFD_D float rcp_rn(float x) { return 1.0f / x; }
TCFD_D double rcp_rn(double x) { return 1.0 / x; }
#endif

}  // namespace tcfd
