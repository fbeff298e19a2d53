// Pointwise channel mixes of the FNO3d layer (reference: fno/fno3d.py:119-130 MLP, :223-235 layer loop,
// SURVEY 8a row B5): the 1x1x1 Conv3d's and GELUs that surround the spectral convolution, fused so that
// an activation tensor is read once and written once per layer.
//
//   layer glue   y = act( mlp2( gelu( mlp1(c) ) ) + w(x) )          c = spectral_conv(x)
//   linear       y = W x + b                                        (lifting Conv3d `p`)
//   project      y = mlp2( act( mlp1(x) ) )                         (projection MLP `q`, Co = 1)
//
// Tensors are (b, C, X, Y, T) fp32, channel planes of npts = X*Y*T contiguous points.  A thread owns P
// consecutive points (one vector load per channel plane: a warp reads 32*P contiguous floats) and keeps
// the per-point channel vectors in registers.  The weights travel as a KERNEL PARAMETER (host pointers
// in the ABI, packed and zero-padded on the host): every multiply-add takes its weight straight from the
// constant bank (FFMA R, R, c[0][..], R), so the products cost no load instructions and no shared memory.
// HBM bound: (Ci + Co) * 4 bytes per point.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>

#include "../../include/tcfd.h"
#include "tcfd_common.cuh"
#ifndef TCFD_EMU
#include <cstring>
#include "fno_glue_tc.cuh"
#endif

extern "C" void tcfd_set_last_error(const char* msg);

namespace tcfd {
namespace {

// GELU, erf form (torch's nn.GELU() default).  (An Abramowitz-Stegun 7.1.26 erf with one reciprocal and one
// exponential was tried: same parity, 10 % SLOWER in the fused layer -- the kernel is not bound by the GELU
// instruction count -- so the library erff stays.)
TCFD_D float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

template <int P>
struct vecf {
  float v[P];
};
template <int P>
TCFD_D vecf<P> ld_vec(const float* p) {
  vecf<P> r;
#ifndef TCFD_EMU
  if constexpr (P == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
  } else if constexpr (P == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    r.v[0] = t.x; r.v[1] = t.y;
    return r;
  }
#endif
#pragma unroll
  for (int j = 0; j < P; ++j) r.v[j] = p[j];
  return r;
}
template <int P>
TCFD_D void st_vec(float* p, const vecf<P>& r) {
#ifndef TCFD_EMU
  if constexpr (P == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    return;
  } else if constexpr (P == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(r.v[0], r.v[1]);
    return;
  }
#endif
#pragma unroll
  for (int j = 0; j < P; ++j) p[j] = r.v[j];
}

// shared-memory image of a weight matrix W[o][i] (torch Conv3d layout) TRANSPOSED to [i][OP] with the
// output index padded to a multiple of 4, so that for a fixed input channel the outputs are read by
// 16-byte loads that every thread of the warp shares (one wavefront)
TCFD_D void stage_wT(float* dst, const float* w, int Co, int Ci, int OP) {
  for (int e = threadIdx.x; e < Ci * OP; e += blockDim.x) {
    const int i = e / OP, o = e % OP;
    dst[e] = o < Co ? w[o * Ci + i] : 0.f;
  }
}
TCFD_D void stage_vec(float* dst, const float* b, int n, int np) {
  for (int e = threadIdx.x; e < np; e += blockDim.x) dst[e] = (b && e < n) ? b[e] : 0.f;
}

// ------------------------------------------------------------------------------------------
// y[b][o][q] = b[o] + sum_i W[o][i] x[b][i][q]        CO = padded number of outputs (multiple of 4)
template <int CO, int P>
__global__ void __launch_bounds__(256)
fno_linear_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ w,
                  const float* __restrict__ bias, int Ci, int Co, size_t npts, size_t groups_per_sample, size_t total) {
  TCFD_DYN_SMEM(smem_raw);
  float* wT = reinterpret_cast<float*>(smem_raw);  // [Ci][CO]
  float* bs = wT + Ci * CO;                        // [CO]
  stage_wT(wT, w, Co, Ci, CO);
  stage_vec(bs, bias, Co, CO);
  __syncthreads();
  for (size_t gidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < total; gidx += (size_t)gridDim.x * blockDim.x) {
    const size_t b = gidx / groups_per_sample, q = (gidx % groups_per_sample) * P;
    const float* xp = x + (size_t)b * Ci * npts + q;
    float acc[CO][P];
#pragma unroll
    for (int o = 0; o < CO; ++o)
#pragma unroll
      for (int j = 0; j < P; ++j) acc[o][j] = bs[o];
    // LDN channel planes in flight per round: the thread's time is load latency, not arithmetic (C5 lifting: 13 planes
    // in two rounds instead of four)
    constexpr int LDN = (CO * P <= 40) ? 8 : 4;
    for (int i0 = 0; i0 < Ci; i0 += LDN) {
      vecf<P> xv[LDN];
#pragma unroll
      for (int u = 0; u < LDN; ++u)
        if (i0 + u < Ci) xv[u] = ld_vec<P>(xp + (size_t)(i0 + u) * npts);
#pragma unroll
      for (int u = 0; u < LDN; ++u)
        if (i0 + u < Ci) {
          const float* wr = wT + (i0 + u) * CO;
#pragma unroll
          for (int o = 0; o < CO; ++o)
#pragma unroll
            for (int j = 0; j < P; ++j) acc[o][j] = fmaf(wr[o], xv[u].v[j], acc[o][j]);
        }
    }
    float* yp = y + (size_t)b * Co * npts + q;
#pragma unroll
    for (int o = 0; o < CO; ++o)
      if (o < Co) {
        vecf<P> r;
#pragma unroll
        for (int j = 0; j < P; ++j) r.v[j] = acc[o][j];
        st_vec<P>(yp + (size_t)o * npts, r);
      }
  }
}

// ------------------------------------------------------------------------------------------
// y = act( W2 gelu(W1 c + b1) + b2 + Ww x + bw )       C channels in and out, CP = C padded to 4
template <int CP>
struct GlueWeights {  // transposed ([input][output]) and zero padded; b2 holds b2 + bw
  float w1T[CP][CP], w2T[CP][CP], wwT[CP][CP], b1[CP], b2[CP];
};

template <int CP, int P>
__global__ void __launch_bounds__(128, (CP * P <= 40 ? 2 : 1))
fno_layer_glue_kernel(const float* __restrict__ c, const float* __restrict__ x, float* __restrict__ y,
                      const
#ifndef TCFD_EMU
                      __grid_constant__
#endif
                      GlueWeights<CP> W, int act, int C, size_t npts, size_t groups_per_sample, size_t total) {
  for (size_t gidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < total; gidx += (size_t)gridDim.x * blockDim.x) {
    const size_t b = gidx / groups_per_sample, q = (gidx % groups_per_sample) * P;
    const size_t base = (size_t)b * C * npts + q;
    // the C channel planes of c are requested up front (C independent vector loads in flight), the planes
    // of x as soon as c has been consumed: they land under the GELUs and the second product
    float cv[CP][P];
#pragma unroll
    for (int i = 0; i < CP; ++i) {
      if (i < C) {
        const vecf<P> a = ld_vec<P>(c + base + (size_t)i * npts);
#pragma unroll
        for (int j = 0; j < P; ++j) cv[i][j] = a.v[j];
      } else {
#pragma unroll
        for (int j = 0; j < P; ++j) cv[i][j] = 0.f;
      }
    }
    float h[CP][P];
#pragma unroll
    for (int o = 0; o < CP; ++o)
#pragma unroll
      for (int j = 0; j < P; ++j) h[o][j] = W.b1[o];
#pragma unroll
    for (int i = 0; i < CP; ++i)
#pragma unroll
      for (int o = 0; o < CP; ++o)
#pragma unroll
        for (int j = 0; j < P; ++j) h[o][j] = fmaf(W.w1T[i][o], cv[i][j], h[o][j]);
    float xv[CP][P];
#pragma unroll
    for (int i = 0; i < CP; ++i) {
      if (i < C) {
        const vecf<P> b_ = ld_vec<P>(x + base + (size_t)i * npts);
#pragma unroll
        for (int j = 0; j < P; ++j) xv[i][j] = b_.v[j];
      } else {
#pragma unroll
        for (int j = 0; j < P; ++j) xv[i][j] = 0.f;
      }
    }
#pragma unroll
    for (int o = 0; o < CP; ++o)
#pragma unroll
      for (int j = 0; j < P; ++j) h[o][j] = gelu_erf(h[o][j]);
    float acc[CP][P];
#pragma unroll
    for (int o = 0; o < CP; ++o)
#pragma unroll
      for (int j = 0; j < P; ++j) acc[o][j] = W.b2[o];
#pragma unroll
    for (int i = 0; i < CP; ++i)  // hidden -> output (rows i >= C of w2T are zero, gelu(pad) is harmless)
#pragma unroll
      for (int o = 0; o < CP; ++o)
#pragma unroll
        for (int j = 0; j < P; ++j) acc[o][j] = fmaf(W.w2T[i][o], h[i][j], acc[o][j]);
#pragma unroll
    for (int i = 0; i < CP; ++i)  // skip path w(x)
#pragma unroll
      for (int o = 0; o < CP; ++o)
#pragma unroll
        for (int j = 0; j < P; ++j) acc[o][j] = fmaf(W.wwT[i][o], xv[i][j], acc[o][j]);
#pragma unroll
    for (int o = 0; o < CP; ++o)
      if (o < C) {
        vecf<P> r;
#pragma unroll
        for (int j = 0; j < P; ++j) r.v[j] = act ? gelu_erf(acc[o][j]) : acc[o][j];
        st_vec<P>(y + base + (size_t)o * npts, r);
      }
  }
}

// ------------------------------------------------------------------------------------------
// y[b][q] = b2 + sum_m W2[m] act( b1[m] + sum_i W1[m][i] x[b][i][q] )      (Co = 1)
template <int CP, int P>
__global__ void __launch_bounds__(256)
fno_project_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ w1,
                   const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2, int act,
                   int C, int M, size_t npts, size_t groups_per_sample, size_t total) {
  TCFD_DYN_SMEM(smem_raw);
  float* w1s = reinterpret_cast<float*>(smem_raw);  // [M][CP]: row m = the C input weights of hidden unit m
  float* b1s = w1s + M * CP;                        // [M]
  float* w2s = b1s + M;                             // [M]
  for (int e = threadIdx.x; e < M * CP; e += blockDim.x) {
    const int m = e / CP, i = e % CP;
    w1s[e] = i < C ? w1[m * C + i] : 0.f;
  }
  for (int e = threadIdx.x; e < M; e += blockDim.x) {
    b1s[e] = b1 ? b1[e] : 0.f;
    w2s[e] = w2[e];
  }
  __syncthreads();
  const float bias2 = b2 ? b2[0] : 0.f;
  for (size_t gidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < total; gidx += (size_t)gridDim.x * blockDim.x) {
    const size_t b = gidx / groups_per_sample, q = (gidx % groups_per_sample) * P;
    const float* xp = x + (size_t)b * C * npts + q;
    float xv[CP][P];
#pragma unroll
    for (int i = 0; i < CP; ++i) {
      if (i < C) {
        const vecf<P> v = ld_vec<P>(xp + (size_t)i * npts);
#pragma unroll
        for (int j = 0; j < P; ++j) xv[i][j] = v.v[j];
      } else {
#pragma unroll
        for (int j = 0; j < P; ++j) xv[i][j] = 0.f;
      }
    }
    float out[P];
#pragma unroll
    for (int j = 0; j < P; ++j) out[j] = bias2;
    for (int m = 0; m < M; ++m) {
      const float* wr = w1s + m * CP;
      float hm[P];
#pragma unroll
      for (int j = 0; j < P; ++j) hm[j] = b1s[m];
#pragma unroll
      for (int i = 0; i < CP; ++i)
#pragma unroll
        for (int j = 0; j < P; ++j) hm[j] = fmaf(wr[i], xv[i][j], hm[j]);
      const float w2m = w2s[m];
#pragma unroll
      for (int j = 0; j < P; ++j) out[j] = fmaf(w2m, act ? gelu_erf(hm[j]) : hm[j], out[j]);
    }
    vecf<P> r;
#pragma unroll
    for (int j = 0; j < P; ++j) r.v[j] = out[j];
    st_vec<P>(y + (size_t)b * npts + q, r);
  }
}

int fail(int code, const std::string& m) {
  tcfd_set_last_error(m.c_str());
  return code;
}

int grid_for(size_t total, int threads) {
#ifndef TCFD_EMU
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const size_t want = (total + threads - 1) / threads, cap = (size_t)sms * (2048 / threads);
  return (int)(want < cap ? want : cap);
#else
  (void)threads;
  return total > 0 ? 2 : 0;
#endif
}
#ifdef TCFD_EMU
constexpr int THREADS = 32, GLUE_THREADS = 32;
#else
constexpr int THREADS = 256, GLUE_THREADS = 128;
#endif

template <class K>
int smem_attr(K k, size_t smem) {
#ifndef TCFD_EMU
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
#else
  (void)k; (void)smem;
#endif
  return 0;
}
int after_launch(const char* what) {
#ifndef TCFD_EMU
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TCFD_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
#else
  (void)what;
#endif
  return TCFD_OK;
}
// points per thread: 4 when every plane offset stays 16-byte aligned, else 2, else 1
int points_per_thread(size_t npts, const void* a, const void* b, const void* c) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c);
  int cap = 4;
  if (const char* e = getenv("TCFD_FNO_P")) cap = atoi(e) > 0 ? atoi(e) : cap;  // experiment knob
  if (cap >= 4 && npts % 4 == 0 && (al & 15u) == 0) return 4;
  if (cap >= 2 && npts % 2 == 0 && (al & 7u) == 0) return 2;
  return 1;
}
}  // namespace
}  // namespace tcfd

using namespace tcfd;

extern "C" int tcfd_fno_pointwise_linear(const float* x, float* y, const float* w, const float* bias, int batch, int Ci,
                                         int Co, size_t npts, void* stream_) {
  if (!x || !y || !w) return fail(TCFD_ERR_INVALID, "null argument");
  if (batch < 1 || Ci < 1 || Co < 1 || npts < 1) return fail(TCFD_ERR_INVALID, "bad sizes");
  if (Co > 32) return fail(TCFD_ERR_INVALID, "pointwise linear: at most 32 output channels");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int CO = (Co + 3) / 4 * 4;
  int P = points_per_thread(npts, x, y, y);
  if (CO > 16 && P > 2) P = 2;  // accumulators: CO * P registers
  const size_t gps = npts / P, total = gps * batch;
  const size_t smem = ((size_t)Ci * CO + CO) * 4;
#define TCFD_LIN(co, p)                                                                                          \
  if (CO == co && P == p) {                                                                                      \
    auto k = fno_linear_kernel<co, p>;                                                                           \
    if (int rc = smem_attr(k, smem)) return fail(TCFD_ERR_CUDA, "cudaFuncSetAttribute failed (" + std::to_string(rc) + ")"); \
    TCFD_LAUNCH(k, grid_for(total, THREADS), THREADS, smem, st, x, y, w, bias, Ci, Co, npts, gps, total);         \
    return after_launch("fno_linear_kernel");                                                                    \
  }
#define TCFD_LIN_P(co) TCFD_LIN(co, 1) TCFD_LIN(co, 2)
  TCFD_LIN_P(4) TCFD_LIN_P(8) TCFD_LIN_P(12) TCFD_LIN_P(16) TCFD_LIN_P(20) TCFD_LIN_P(24) TCFD_LIN_P(28) TCFD_LIN_P(32)
  TCFD_LIN(4, 4) TCFD_LIN(8, 4) TCFD_LIN(12, 4) TCFD_LIN(16, 4)
#undef TCFD_LIN_P
#undef TCFD_LIN
  return fail(TCFD_ERR_INVALID, "pointwise linear: unsupported channel count");
}

template <int CP>
GlueWeights<CP> pack_glue(const float* w1, const float* b1, const float* w2, const float* b2, const float* ww,
                          const float* bw, int C) {
  GlueWeights<CP> W;
  for (int i = 0; i < CP; ++i)
    for (int o = 0; o < CP; ++o) {
      const bool in = i < C && o < C;
      W.w1T[i][o] = in ? w1[o * C + i] : 0.f;
      W.w2T[i][o] = in ? w2[o * C + i] : 0.f;
      W.wwT[i][o] = in ? ww[o * C + i] : 0.f;
    }
  for (int o = 0; o < CP; ++o) {
    W.b1[o] = (b1 && o < C) ? b1[o] : 0.f;
    W.b2[o] = o < C ? (b2 ? b2[o] : 0.f) + (bw ? bw[o] : 0.f) : 0.f;
  }
  return W;
}

namespace { int g_last_glue_path = 0; }
extern "C" int tcfd_fno_layer_glue_path(void) { return g_last_glue_path; }

extern "C" int tcfd_fno_layer_glue(const float* conv_out, const float* x, float* y, const float* w1, const float* b1,
                                   const float* w2, const float* b2, const float* ww, const float* bw, int act, int batch,
                                   int C, size_t npts, void* stream_) {
  if (!conv_out || !x || !y || !w1 || !w2 || !ww) return fail(TCFD_ERR_INVALID, "null argument");
  if (batch < 1 || C < 1 || npts < 1) return fail(TCFD_ERR_INVALID, "bad sizes");
  if (C > 32) return fail(TCFD_ERR_INVALID, "layer glue: at most 32 channels");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
#ifndef TCFD_EMU
  {
    // tensor-core path (fno_glue_tc.cuh): the three C x C products as 3xTF32 UMMAs with the accumulators in TMEM.
    // TCFD_GLUE_TC=0 selects the CUDA-core kernel below (cross-check / A-B timing); TCFD_GLUE_SWAP is a bring-up knob.
    const char* e = getenv("TCFD_GLUE_TC");
    const bool use_tc = !(e && atoi(e) == 0);
    const char* sw = getenv("TCFD_GLUE_SWAP");
    const bool swap = sw && atoi(sw) != 0;
    if (use_tc && C % 2 == 0 && npts < ((size_t)1 << 27)) {
      static int sms = 0;
      if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      }
      const size_t tps = (npts + 127) / 128, ntiles = tps * batch;
      const bool tail = npts % 128 != 0;
      // dynamic shared memory only pads the footprint so that the resident CTAs of an SM never ask for more than
      // the 512 TMEM columns: 4 CTAs (128 columns each) for C <= 24, 2 CTAs (256 columns) above
      const int per_sm = C <= 24 ? TCFD_GLUE_MINB : 2;
      const size_t pad = C <= 24 ? 30 * 1024 : 64 * 1024;
      const size_t grid = ntiles < (size_t)sms * per_sm ? ntiles : (size_t)sms * per_sm;
#define TCFD_GLUE_TC_RUN(cc, tl, sw_)                                                                                  \
  {                                                                                                                    \
    auto k = gluetc::fno_layer_glue_tc_kernel<cc, tl, sw_>;                                                            \
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad) != cudaSuccess)                 \
      return fail(TCFD_ERR_CUDA, "cudaFuncSetAttribute failed");                                                       \
    const auto W = gluetc::pack_tc<(cc + 7) / 8 * 8>(w1, b1, w2, b2, ww, bw, C);                                       \
    g_last_glue_path = 1;                                                                                              \
    k<<<(unsigned)grid, 256, pad, st>>>(conv_out, x, y, W, act, npts, tps, ntiles);                                    \
    return after_launch("fno_layer_glue_tc_kernel");                                                                   \
  }
#define TCFD_GLUE_TC_CASE(cc)                                  \
  if (C == cc && !swap) {                                      \
    if (tail) TCFD_GLUE_TC_RUN(cc, true, false)                \
    else TCFD_GLUE_TC_RUN(cc, false, false)                    \
  }
      TCFD_GLUE_TC_CASE(2) TCFD_GLUE_TC_CASE(4) TCFD_GLUE_TC_CASE(6) TCFD_GLUE_TC_CASE(8) TCFD_GLUE_TC_CASE(10)
      TCFD_GLUE_TC_CASE(12) TCFD_GLUE_TC_CASE(14) TCFD_GLUE_TC_CASE(16) TCFD_GLUE_TC_CASE(18) TCFD_GLUE_TC_CASE(20)
      TCFD_GLUE_TC_CASE(22) TCFD_GLUE_TC_CASE(24) TCFD_GLUE_TC_CASE(26) TCFD_GLUE_TC_CASE(28) TCFD_GLUE_TC_CASE(30)
      TCFD_GLUE_TC_CASE(32)
      if (C == 20 && swap) TCFD_GLUE_TC_RUN(20, true, true)
#undef TCFD_GLUE_TC_CASE
#undef TCFD_GLUE_TC_RUN
    }
  }
#endif
  g_last_glue_path = 0;
  const int CP = (C + 3) / 4 * 4;
  int P = points_per_thread(npts, conv_out, x, y);
  // (a 4-point streaming variant -- each weight serving 4 multiply-adds -- was measured slower: 5.6 vs 3.6 ms)
  if (P > 2) P = 2;  // hidden + accumulators + inputs: 3 * CP * P registers
  if (CP > 20) P = 1;
  const size_t gps = npts / P, total = gps * batch;
#define TCFD_GLUE(cp, p)                                                                                          \
  if (CP == cp && P == p) {                                                                                       \
    auto k = fno_layer_glue_kernel<cp, p>;                                                                        \
    const GlueWeights<cp> W = pack_glue<cp>(w1, b1, w2, b2, ww, bw, C);                                           \
    TCFD_LAUNCH(k, grid_for(total, GLUE_THREADS), GLUE_THREADS, 0, st, conv_out, x, y, W, act, C, npts, gps, total); \
    return after_launch("fno_layer_glue_kernel");                                                                 \
  }
  TCFD_GLUE(4, 1) TCFD_GLUE(4, 2) TCFD_GLUE(8, 1) TCFD_GLUE(8, 2) TCFD_GLUE(12, 1) TCFD_GLUE(12, 2) TCFD_GLUE(16, 1)
  TCFD_GLUE(16, 2) TCFD_GLUE(20, 1) TCFD_GLUE(20, 2) TCFD_GLUE(24, 1) TCFD_GLUE(28, 1) TCFD_GLUE(32, 1)
#undef TCFD_GLUE
  return fail(TCFD_ERR_INVALID, "layer glue: unsupported channel count");
}

extern "C" int tcfd_fno_project(const float* x, float* y, const float* w1, const float* b1, const float* w2,
                                const float* b2, int act, int batch, int C, int M, size_t npts, void* stream_) {
  if (!x || !y || !w1 || !w2) return fail(TCFD_ERR_INVALID, "null argument");
  if (batch < 1 || C < 1 || M < 1 || npts < 1) return fail(TCFD_ERR_INVALID, "bad sizes");
  if (C > 32) return fail(TCFD_ERR_INVALID, "project: at most 32 input channels");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int CP = (C + 3) / 4 * 4;
  int P = points_per_thread(npts, x, y, y);
  if (CP > 16 && P > 2) P = 2;
  const size_t gps = npts / P, total = gps * batch;
  const size_t smem = ((size_t)M * CP + 2 * (size_t)M) * 4;
  if (smem > 200 * 1024) return fail(TCFD_ERR_INVALID, "project: hidden width too large for shared memory");
#define TCFD_PROJ(cp, p)                                                                                          \
  if (CP == cp && P == p) {                                                                                       \
    auto k = fno_project_kernel<cp, p>;                                                                           \
    if (int rc = smem_attr(k, smem)) return fail(TCFD_ERR_CUDA, "cudaFuncSetAttribute failed (" + std::to_string(rc) + ")"); \
    TCFD_LAUNCH(k, grid_for(total, THREADS), THREADS, smem, st, x, y, w1, b1, w2, b2, act, C, M, npts, gps, total); \
    return after_launch("fno_project_kernel");                                                                    \
  }
#define TCFD_PROJ_P(cp) TCFD_PROJ(cp, 1) TCFD_PROJ(cp, 2)
  TCFD_PROJ_P(4) TCFD_PROJ_P(8) TCFD_PROJ_P(12) TCFD_PROJ_P(16) TCFD_PROJ_P(20) TCFD_PROJ_P(24) TCFD_PROJ_P(28) TCFD_PROJ_P(32)
  TCFD_PROJ(4, 4) TCFD_PROJ(8, 4) TCFD_PROJ(12, 4) TCFD_PROJ(16, 4)
#undef TCFD_PROJ_P
#undef TCFD_PROJ
  return fail(TCFD_ERR_INVALID, "project: unsupported channel count");
}
