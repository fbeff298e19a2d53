// Stand-alone batched 2-D real transforms on the register/shared-memory Stockham core (fft_core.cuh) and
// the bilinear resampling that follows them in the reference's data-generation scripts:
//
//   tcfd_fft2_irfft2          == torch.fft.irfft2   (fno/data_gen/data_gen_Kolmogorov2d.py:178-179,
//                                data_gen_McWilliams2d.py:157; torch_cfd/equations.py:415-422 call sites)
//   tcfd_fft2_rfft2           == torch.fft.rfft2    (torch_cfd/equations.py:432-436 forcing spectra,
//                                torch_cfd/initial_conditions.py IC generators)
//   tcfd_resample_bilinear    == F.interpolate(value, size=(ns, ns), mode="bilinear")   (data_gen_*.py:186)
//
// Same pass order as torch: irfft2 = complex inverse along kx (axis -2) for every kept ky, then C2R along
// ky (axis -1), which drops the imaginary part of the ky = 0 and ky = n/2 bins; rfft2 = R2C along y, then a
// complex transform along x.  Layouts are the reference's: spectrum [count][n][n/2+1] interleaved complex,
// field [count][n][n] real; "backward" normalisation (1/n^2 on the inverse).
//
//   x pass    CTA = G adjacent ky columns x (n/8 threads each); thread index = t * G + g, so a warp reads
//             G * 8 (16) contiguous bytes of 32 / G rows: full sectors on the strided side
//   y passes  one group per pair of rows: two real rows travel as ONE complex transform (z = a + i b)
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/tcfd.h"
#include "fft_core.cuh"

extern "C" void tcfd_set_last_error(const char* msg);

namespace tcfd {
namespace {

struct CtaSyncF {
  TCFD_D void operator()() const { __syncthreads(); }
};

template <int N>
struct Fft2Geom {
  static constexpr int NT = N / 8;
  static constexpr int GX = (512 / NT) < 8 ? (512 / NT) : 8;       // ky columns per CTA of the x pass (<= 512 threads)
  static constexpr int GY0 = (128 / NT) > 0 ? (128 / NT) : 1;
  static constexpr int GY = GY0 < N / 2 ? GY0 : N / 2;             // row pairs per CTA of the y passes (divides N / 2)
};

// complex transform along kx (axis -2) of the half spectrum, in -> out (may alias), scaled
template <class T, int N, int DIR>
__global__ void __launch_bounds__(Fft2Geom<N>::GX * (N / 8))
fft2_xaxis_kernel(const cx<T>* in, cx<T>* out, const cx<T>* __restrict__ twtab, int nh, T scale) {
  constexpr int NT = N / 8, G = Fft2Geom<N>::GX;
  TCFD_DYN_SMEM(smem_raw);
  const int g = threadIdx.x % G, t = threadIdx.x / G;
  cx<T>* buf = reinterpret_cast<cx<T>*>(smem_raw) + (size_t)g * N;
  FftTwiddles<T, N> tw;
  tw.load(twtab, t);
  CtaSyncF sync;
  int parity = 0;
  const int ky = blockIdx.x * G + g;
  const bool valid = ky < nh;
  const size_t base = (size_t)blockIdx.y * N * nh + (valid ? ky : 0);
  cx<T> v[1][8];
#pragma unroll
  for (int m = 0; m < 8; ++m) v[0][m] = valid ? in[base + (size_t)(t + m * NT) * nh] : cx<T>{T(0), T(0)};
  fft_run<T, N, DIR, 1, false, N>(v, tw, buf, parity, t, sync);
  if (valid) {
#pragma unroll
    for (int m = 0; m < 8; ++m) out[base + (size_t)(t + m * NT) * nh] = cx<T>{scale * v[0][m].x, scale * v[0][m].y};
  }
}

// C2R along ky (axis -1): rows a = 2p, b = 2p + 1 of image blockIdx.y as one complex inverse transform
template <class T, int N>
__global__ void __launch_bounds__(Fft2Geom<N>::GY * (N / 8))
fft2_c2r_rows_kernel(const cx<T>* __restrict__ in, T* __restrict__ out, const cx<T>* __restrict__ twtab, int nh) {
  constexpr int NT = N / 8, G = Fft2Geom<N>::GY;
  TCFD_DYN_SMEM(smem_raw);
  const int g = threadIdx.x / NT, t = threadIdx.x % NT;
  cx<T>* buf = reinterpret_cast<cx<T>*>(smem_raw) + (size_t)g * N;
  FftTwiddles<T, N> tw;
  tw.load(twtab, t);
  CtaSyncF sync;
  int parity = 0;
  const int pair = blockIdx.x * G + g;  // N / 2 pairs, G divides N / 2
  const cx<T>* ra = in + ((size_t)blockIdx.y * N + 2 * pair) * nh;
  const cx<T>* rb = ra + nh;
  cx<T> v[1][8];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int k = t + m * NT;
    const bool lo = k <= N / 2;
    cx<T> a = ra[lo ? k : N - k], b = rb[lo ? k : N - k];
    if (!lo) { a.y = -a.y; b.y = -b.y; }
    if (k == 0 || k == N / 2) { a.y = T(0); b.y = T(0); }  // C2R drops the imaginary part of the self-conjugate bins
    v[0][m] = cx<T>{a.x - b.y, a.y + b.x};
  }
  fft_run<T, N, +1, 1, false, N>(v, tw, buf, parity, t, sync);
  T* oa = out + ((size_t)blockIdx.y * N + 2 * pair) * N;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    oa[t + m * NT] = v[0][m].x;
    oa[N + t + m * NT] = v[0][m].y;
  }
}

// R2C along y (axis -1): rows a, b as one complex forward transform, separated through shared memory
template <class T, int N>
__global__ void __launch_bounds__(Fft2Geom<N>::GY * (N / 8))
fft2_r2c_rows_kernel(const T* __restrict__ in, cx<T>* __restrict__ out, const cx<T>* __restrict__ twtab, int nh) {
  constexpr int NT = N / 8, G = Fft2Geom<N>::GY;
  TCFD_DYN_SMEM(smem_raw);
  const int g = threadIdx.x / NT, t = threadIdx.x % NT;
  cx<T>* buf = reinterpret_cast<cx<T>*>(smem_raw) + (size_t)g * N;
  FftTwiddles<T, N> tw;
  tw.load(twtab, t);
  CtaSyncF sync;
  int parity = 0;
  const int pair = blockIdx.x * G + g;
  const T* ia = in + ((size_t)blockIdx.y * N + 2 * pair) * N;
  cx<T> v[1][8];
#pragma unroll
  for (int m = 0; m < 8; ++m) v[0][m] = cx<T>{ia[t + m * NT], ia[N + t + m * NT]};
  fft_run<T, N, -1, 1, false, N>(v, tw, buf, parity, t, sync);
#pragma unroll
  for (int m = 0; m < 8; ++m) buf[t + m * NT] = v[0][m];
  __syncthreads();
  cx<T>* oa = out + ((size_t)blockIdx.y * N + 2 * pair) * nh;
  cx<T>* ob = oa + nh;
  for (int k = t; k <= N / 2; k += NT) {
    const cx<T> z = buf[k], zn = buf[(N - k) % N];
    // X_a = (Z(k) + conj Z(-k)) / 2 ;  X_b = (Z(k) - conj Z(-k)) / (2 i)
    oa[k] = cx<T>{T(0.5) * (z.x + zn.x), T(0.5) * (z.y - zn.y)};
    ob[k] = cx<T>{T(0.5) * (z.y + zn.y), T(0.5) * (zn.x - z.x)};
  }
}

// F.interpolate(mode="bilinear", align_corners=False) of [count][n_in][n_in] -> [count][n_out][n_out]
// (aten upsample_bilinear2d: source index = scale * (dst + 0.5) - 0.5 clamped at 0, weights in the compute type)
template <class TI, class TO>
__global__ void resample_bilinear_kernel(const TI* __restrict__ in, TO* __restrict__ out, int n_in, int n_out, size_t total) {
  typedef TO A;  // accumulation type = output type (float for fp32 results, double for fp64)
  const A scale = (A)n_in / (A)n_out;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % n_out), i = (int)((idx / n_out) % n_out);
    const size_t img = idx / ((size_t)n_out * n_out);
    A sh = scale * ((A)i + (A)0.5) - (A)0.5, sw = scale * ((A)j + (A)0.5) - (A)0.5;
    sh = sh < (A)0 ? (A)0 : sh;
    sw = sw < (A)0 ? (A)0 : sw;
    const int h0 = (int)sh, w0 = (int)sw;
    const int h1 = h0 + (h0 < n_in - 1 ? 1 : 0), w1 = w0 + (w0 < n_in - 1 ? 1 : 0);
    const A lh1 = sh - (A)h0, lh0 = (A)1 - lh1, lw1 = sw - (A)w0, lw0 = (A)1 - lw1;
    const TI* p = in + img * (size_t)n_in * n_in;
    const A v00 = (A)p[(size_t)h0 * n_in + w0], v01 = (A)p[(size_t)h0 * n_in + w1];
    const A v10 = (A)p[(size_t)h1 * n_in + w0], v11 = (A)p[(size_t)h1 * n_in + w1];
    out[idx] = (TO)(lh0 * (lw0 * v00 + lw1 * v01) + lh1 * (lw0 * v10 + lw1 * v11));
  }
}

template <class K>
int set_smem(K kernel, size_t smem) {
#ifndef TCFD_EMU
  if (smem > 48 * 1024) return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#else
  (void)kernel;
  (void)smem;
#endif
  return 0;
}

template <class T, int N>
int run_irfft2(const void* in_, void* scratch_, void* out_, const void* tw_, int count, cudaStream_t stream) {
  typedef Fft2Geom<N> Gm;
  constexpr int NT = N / 8, NH = N / 2 + 1;
  const cx<T>* in = static_cast<const cx<T>*>(in_);
  cx<T>* scratch = static_cast<cx<T>*>(scratch_);
  const cx<T>* tw = static_cast<const cx<T>*>(tw_);
  auto kx = fft2_xaxis_kernel<T, N, +1>;
  auto ky = fft2_c2r_rows_kernel<T, N>;
  const size_t sx = (size_t)Gm::GX * N * sizeof(cx<T>), sy = (size_t)Gm::GY * N * sizeof(cx<T>);
  int rc;
  if ((rc = set_smem(kx, sx)) || (rc = set_smem(ky, sy))) return rc;
  const T scale = (T)(1.0 / ((double)N * (double)N));
  TCFD_LAUNCH3(kx, (NH + Gm::GX - 1) / Gm::GX, count, 1, Gm::GX * NT, sx, stream, in, scratch, tw, NH, scale);
  TCFD_LAUNCH3(ky, (N / 2) / Gm::GY, count, 1, Gm::GY * NT, sy, stream, scratch, static_cast<T*>(out_), tw, NH);
  return 0;
}

template <class T, int N>
int run_rfft2(const void* in_, void* out_, const void* tw_, int count, cudaStream_t stream) {
  typedef Fft2Geom<N> Gm;
  constexpr int NT = N / 8, NH = N / 2 + 1;
  cx<T>* out = static_cast<cx<T>*>(out_);
  const cx<T>* tw = static_cast<const cx<T>*>(tw_);
  auto ky = fft2_r2c_rows_kernel<T, N>;
  auto kx = fft2_xaxis_kernel<T, N, -1>;
  const size_t sx = (size_t)Gm::GX * N * sizeof(cx<T>), sy = (size_t)Gm::GY * N * sizeof(cx<T>);
  int rc;
  if ((rc = set_smem(kx, sx)) || (rc = set_smem(ky, sy))) return rc;
  TCFD_LAUNCH3(ky, (N / 2) / Gm::GY, count, 1, Gm::GY * NT, sy, stream, static_cast<const T*>(in_), out, tw, NH);
  TCFD_LAUNCH3(kx, (NH + Gm::GX - 1) / Gm::GX, count, 1, Gm::GX * NT, sx, stream, out, out, tw, NH, T(1));
  return 0;
}

#define TCFD_FFT2_SIZES(X) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)

template <class T>
int dispatch_irfft2(int n, const void* in, void* scratch, void* out, const void* tw, int count, cudaStream_t s) {
#define TCFD_CASE(NN) if (n == NN) return run_irfft2<T, NN>(in, scratch, out, tw, count, s);
  TCFD_FFT2_SIZES(TCFD_CASE)
#undef TCFD_CASE
  return -1;
}
template <class T>
int dispatch_rfft2(int n, const void* in, void* out, const void* tw, int count, cudaStream_t s) {
#define TCFD_CASE(NN) if (n == NN) return run_rfft2<T, NN>(in, out, tw, count, s);
  TCFD_FFT2_SIZES(TCFD_CASE)
#undef TCFD_CASE
  return -1;
}

int fail2(int code, const std::string& msg) {
  tcfd_set_last_error(msg.c_str());
  return code;
}
}  // namespace
}  // namespace tcfd

struct tcfd_fft2 {
  int n = 0, nh = 0, prec = 0;
  size_t es = 0;
  void* tw = nullptr;
  void* scratch = nullptr;   // x-pass output of the inverse transform, `cap` images
  int cap = 0;
  int launches = 0;
};

extern "C" int tcfd_fft2_create(tcfd_fft2_t** out, int n, int prec) {
  using namespace tcfd;
  if (!out) return fail2(TCFD_ERR_INVALID, "null argument");
  *out = nullptr;
  if (prec != 32 && prec != 64) return fail2(TCFD_ERR_INVALID, "prec must be 32 or 64");
  bool ok = false;
#define TCFD_CASE(NN) ok = ok || n == NN;
  TCFD_FFT2_SIZES(TCFD_CASE)
#undef TCFD_CASE
  if (!ok) return fail2(TCFD_ERR_INVALID, "unsupported grid size n=" + std::to_string(n) + " (supported: powers of two 32..2048)");
  tcfd_fft2* h = new tcfd_fft2();
  h->n = n;
  h->nh = n / 2 + 1;
  h->prec = prec;
  h->es = prec / 8;
  const double PI = 3.14159265358979323846264338327950288;
  std::vector<unsigned char> tw(2 * (size_t)n * h->es);
  for (int j = 0; j < n; ++j) {
    const double a = -2.0 * PI * (double)j / (double)n;
    if (prec == 32) {
      reinterpret_cast<float*>(tw.data())[2 * j] = (float)std::cos(a);
      reinterpret_cast<float*>(tw.data())[2 * j + 1] = (float)std::sin(a);
    } else {
      reinterpret_cast<double*>(tw.data())[2 * j] = std::cos(a);
      reinterpret_cast<double*>(tw.data())[2 * j + 1] = std::sin(a);
    }
  }
  if (cudaMalloc(&h->tw, tw.size()) != cudaSuccess || cudaMemcpy(h->tw, tw.data(), tw.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    delete h;
    return fail2(TCFD_ERR_NOMEM, "twiddle table allocation failed");
  }
  *out = h;
  return TCFD_OK;
}

extern "C" int tcfd_fft2_destroy(tcfd_fft2_t* h) {
  if (!h) return TCFD_OK;
  cudaFree(h->tw);
  cudaFree(h->scratch);
  delete h;
  return TCFD_OK;
}

extern "C" int tcfd_fft2_last_launch_count(const tcfd_fft2_t* h) { return h ? h->launches : 0; }

extern "C" int tcfd_fft2_irfft2(tcfd_fft2_t* h, const void* in_hat, void* out, int count, void* stream) {
  using namespace tcfd;
  if (!h || !in_hat || !out) return fail2(TCFD_ERR_INVALID, "null argument");
  if (count < 1) return fail2(TCFD_ERR_INVALID, "count must be >= 1");
  const size_t img_hat = (size_t)h->n * h->nh * 2 * h->es, img = (size_t)h->n * h->n * h->es;
  // scratch: at most 256 MB (and at least one image); larger batches run in chunks
  int want = (int)((size_t)(256u << 20) / img_hat);
  if (want < 1) want = 1;
  if (want > count) want = count;
  if (want > h->cap) {
    cudaFree(h->scratch);  // (synchronises: earlier launches that use the old block are complete)
    h->scratch = nullptr;
    h->cap = 0;
    if (cudaMalloc(&h->scratch, (size_t)want * img_hat) != cudaSuccess) return fail2(TCFD_ERR_NOMEM, "irfft2 scratch allocation failed");
    h->cap = want;
  }
  h->launches = 0;
  for (int c0 = 0; c0 < count; c0 += h->cap) {
    const int cb = count - c0 < h->cap ? count - c0 : h->cap;
    const void* src = static_cast<const unsigned char*>(in_hat) + (size_t)c0 * img_hat;
    void* dst = static_cast<unsigned char*>(out) + (size_t)c0 * img;
    const int rc = h->prec == 32 ? dispatch_irfft2<float>(h->n, src, h->scratch, dst, h->tw, cb, static_cast<cudaStream_t>(stream))
                                 : dispatch_irfft2<double>(h->n, src, h->scratch, dst, h->tw, cb, static_cast<cudaStream_t>(stream));
    h->launches += 2;
    if (rc != 0) return fail2(TCFD_ERR_CUDA, "irfft2 launch failed");
  }
#ifndef TCFD_EMU
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail2(TCFD_ERR_CUDA, std::string("irfft2: ") + cudaGetErrorString(e));
#endif
  return TCFD_OK;
}

extern "C" int tcfd_fft2_rfft2(tcfd_fft2_t* h, const void* in, void* out_hat, int count, void* stream) {
  using namespace tcfd;
  if (!h || !in || !out_hat) return fail2(TCFD_ERR_INVALID, "null argument");
  if (count < 1) return fail2(TCFD_ERR_INVALID, "count must be >= 1");
  const int rc = h->prec == 32 ? dispatch_rfft2<float>(h->n, in, out_hat, h->tw, count, static_cast<cudaStream_t>(stream))
                               : dispatch_rfft2<double>(h->n, in, out_hat, h->tw, count, static_cast<cudaStream_t>(stream));
  h->launches = 2;
  if (rc != 0) return fail2(TCFD_ERR_CUDA, "rfft2 launch failed");
#ifndef TCFD_EMU
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail2(TCFD_ERR_CUDA, std::string("rfft2: ") + cudaGetErrorString(e));
#endif
  return TCFD_OK;
}

extern "C" int tcfd_resample_bilinear(const void* in, void* out, int prec_in, int prec_out, int count, int n_in, int n_out,
                                      void* stream) {
  using namespace tcfd;
  if (!in || !out) return fail2(TCFD_ERR_INVALID, "null argument");
  if ((prec_in != 32 && prec_in != 64) || (prec_out != 32 && prec_out != 64)) return fail2(TCFD_ERR_INVALID, "prec must be 32 or 64");
  if (count < 1 || n_in < 1 || n_out < 1) return fail2(TCFD_ERR_INVALID, "bad sizes");
  const size_t total = (size_t)count * n_out * n_out;
  const int threads = 256;
  size_t blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (prec_in == 32 && prec_out == 32) {
    auto k = resample_bilinear_kernel<float, float>;
    TCFD_LAUNCH(k, (unsigned)blocks, threads, 0, s, static_cast<const float*>(in), static_cast<float*>(out), n_in, n_out, total);
  } else if (prec_in == 64 && prec_out == 64) {
    auto k = resample_bilinear_kernel<double, double>;
    TCFD_LAUNCH(k, (unsigned)blocks, threads, 0, s, static_cast<const double*>(in), static_cast<double*>(out), n_in, n_out, total);
  } else if (prec_in == 64 && prec_out == 32) {
    auto k = resample_bilinear_kernel<double, float>;
    TCFD_LAUNCH(k, (unsigned)blocks, threads, 0, s, static_cast<const double*>(in), static_cast<float*>(out), n_in, n_out, total);
  } else {
    auto k = resample_bilinear_kernel<float, double>;
    TCFD_LAUNCH(k, (unsigned)blocks, threads, 0, s, static_cast<const float*>(in), static_cast<double*>(out), n_in, n_out, total);
  }
#ifndef TCFD_EMU
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail2(TCFD_ERR_CUDA, std::string("resample_bilinear: ") + cudaGetErrorString(e));
#endif
  return TCFD_OK;
}
