// C ABI of hot path B (see include/tcfd.h): host driver of the pruned spectral convolution.
// Owns the t-axis tables, the FFT twiddles and the (small) spectral workspaces; the arithmetic
// lives in sconv_kernels.cuh.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tcfd.h"
#include "sconv_kernels.cuh"

extern "C" void tcfd_set_last_error(const char* msg);

namespace {
thread_local std::string g_serr;
int sfail(int code, const std::string& msg) {
  g_serr = msg;
  tcfd_set_last_error(msg.c_str());
  return code;
}
#define SCUDA_TRY(expr)                                                                         \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return sfail(TCFD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));          \
  } while (0)

typedef tcfd::cx<float> cplx;
const double PI = 3.14159265358979323846264338327950288;

template <class K>
int set_smem(K kernel, size_t smem) {
#ifndef TCFD_EMU
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
#else
  (void)kernel;
  (void)smem;
#endif
  return 0;
}
}  // namespace

struct tcfd_sconv3d {
  tcfd_sconv3d_desc_t d{};
  int Tn_in = 0, Tn_out = 0, K = 0;
  void *twx = nullptr, *twy = nullptr;
  void *A_f = nullptr, *S_f = nullptr, *A_b = nullptr, *S_b = nullptr;
  void *Z = nullptr, *H1 = nullptr, *H2 = nullptr;
  size_t ws_bytes = 0;
  int launches = 0;
};

namespace {
int upload_c(void** dst, const std::vector<cplx>& v) {
  SCUDA_TRY(cudaMalloc(dst, v.size() * sizeof(cplx)));
  SCUDA_TRY(cudaMemcpy(*dst, v.data(), v.size() * sizeof(cplx), cudaMemcpyHostToDevice));
  return 0;
}
std::vector<cplx> twiddles(int n) {
  std::vector<cplx> tw(n);
  for (int j = 0; j < n; ++j) {
    const double a = -2.0 * PI * (double)j / (double)n;
    tw[j] = cplx{(float)std::cos(a), (float)std::sin(a)};
  }
  return tw;
}
bool pow2_ok(int n) { return n >= 32 && n <= 512 && (n & (n - 1)) == 0; }

using namespace tcfd;

template <int Y>
int launch_planes_fwd(const float* x, cplx* Z1, const cplx* A, const cplx* tw, SconvDims d, int nplanes, cudaStream_t st) {
  typedef PlanesSmem<Y> S;
  const size_t smem = S::GP * S::group_bytes(d.Tin, d.my);
  if (smem > 227 * 1024) return -100;
  auto k = sconv_planes_fwd_kernel<Y>;
  if (int rc = set_smem(k, smem)) return rc;
  TCFD_LAUNCH(k, (nplanes + S::GP - 1) / S::GP, S::GP * S::NT, smem, st, x, Z1, A, tw, d, nplanes);
  return 0;
}
template <int Y>
int launch_planes_inv(const cplx* Z2, float* y, const cplx* Sy, const cplx* tw, SconvDims d, int nplanes, cudaStream_t st) {
  typedef PlanesSmem<Y> S;
  const size_t smem = S::GP * S::group_bytes(d.Tout, d.my);
  if (smem > 227 * 1024) return -100;
  auto k = sconv_planes_inv_kernel<Y>;
  if (int rc = set_smem(k, smem)) return rc;
  TCFD_LAUNCH(k, (nplanes + S::GP - 1) / S::GP, S::GP * S::NT, smem, st, Z2, y, Sy, tw, d, nplanes);
  return 0;
}
// second-generation plane kernels: persistent grid (SM count x resident CTAs), pipelined staging
int planes_grid(const void* kernel, int threads, size_t smem, int want) {
#ifndef TCFD_EMU
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) occ = 1;
  const int cap = sms * occ;
  return want < cap ? want : cap;
#else
  (void)kernel; (void)threads; (void)smem;
  return want < 3 ? want : 3;  // several planes per group: exercises the double buffering
#endif
}
// TCFD_SCONV_PLANES = 1 / 2: force the first / second generation plane kernels (bring-up and A/B timing)
int planes_gen() {  // read at every call: tests switch generations inside one process
  const char* e = getenv("TCFD_SCONV_PLANES");
  return e ? atoi(e) : 0;
}
bool planes_v1() { return planes_gen() == 1; }
template <int Y>
int launch_planes_fwd2(const float* x, cplx* Z1, const cplx* A, const cplx* tw, SconvDims d, int nplanes, cudaStream_t st) {
  typedef Planes2Smem<Y> S;
  const size_t smem = S::table_bytes(d.Tin, d.mt) + S::GP * S::group_bytes(d.Tin, d.my, d.mt, false);
  // bulk copies need 16-byte aligned planes (always true for whole tensors; a sliced view may not be)
  if (smem > 227 * 1024 || (reinterpret_cast<uintptr_t>(x) & 15u)) return launch_planes_fwd<Y>(x, Z1, A, tw, d, nplanes, st);
  auto k = sconv_planes_fwd2_kernel<Y>;
  if (int rc = set_smem(k, smem)) return rc;
  const int grid = planes_grid(reinterpret_cast<const void*>(k), S::GP * S::NT, smem, (nplanes + S::GP - 1) / S::GP);
  TCFD_LAUNCH(k, grid, S::GP * S::NT, smem, st, x, Z1, A, tw, d, nplanes);
  return 0;
}
template <int Y>
int launch_planes_inv2(const cplx* Z2, float* y, const cplx* Sy, const cplx* tw, SconvDims d, int nplanes, cudaStream_t st) {
  typedef Planes2Smem<Y> S;
  const size_t smem = S::table_bytes(d.Tout, d.mt) + S::GP * S::group_bytes(d.Tout, d.my, d.mt, true);
  if (smem > 227 * 1024 || (reinterpret_cast<uintptr_t>(y) & 15u)) return launch_planes_inv<Y>(Z2, y, Sy, tw, d, nplanes, st);
  auto k = sconv_planes_inv2_kernel<Y>;
  if (int rc = set_smem(k, smem)) return rc;
  const int grid = planes_grid(reinterpret_cast<const void*>(k), S::GP * S::NT, smem, (nplanes + S::GP - 1) / S::GP);
  TCFD_LAUNCH(k, grid, S::GP * S::NT, smem, st, Z2, y, Sy, tw, d, nplanes);
  return 0;
}
// third generation (inverse only): pruned y transform without exchanges; MYT = compile-time bound of my
template <int Y, int MYT>
int launch_planes_inv3(const cplx* Z2, float* y, const cplx* Sy, const cplx* tw, SconvDims d, int nplanes, cudaStream_t st) {
  typedef Planes3Smem<Y> S;
  const size_t smem = S::table_bytes(d.Tout, d.mt) + S::GP * S::group_bytes(d.Tout, d.my, d.mt);
  if (smem > 227 * 1024 || (reinterpret_cast<uintptr_t>(y) & 15u) || (reinterpret_cast<uintptr_t>(Z2) & 15u))
    return launch_planes_inv2<Y>(Z2, y, Sy, tw, d, nplanes, st);
  if (d.my == MYT) {  // every twiddle of the instantiation is used: the variant without bound checks
    auto k = sconv_planes_inv3_kernel<Y, MYT, true>;
    if (int rc = set_smem(k, smem)) return rc;
    const int grid = planes_grid(reinterpret_cast<const void*>(k), S::GP * S::NT, smem, (nplanes + S::GP - 1) / S::GP);
    TCFD_LAUNCH(k, grid, S::GP * S::NT, smem, st, Z2, y, Sy, tw, d, nplanes);
    return 0;
  }
  auto k = sconv_planes_inv3_kernel<Y, MYT, false>;
  if (int rc = set_smem(k, smem)) return rc;
  const int grid = planes_grid(reinterpret_cast<const void*>(k), S::GP * S::NT, smem, (nplanes + S::GP - 1) / S::GP);
  TCFD_LAUNCH(k, grid, S::GP * S::NT, smem, st, Z2, y, Sy, tw, d, nplanes);
  return 0;
}
template <int Y>
int launch_planes_inv_best(const cplx* Z2, float* y, const cplx* Sy, const cplx* tw, SconvDims d, int nplanes, cudaStream_t st) {
  if (planes_gen() == 2 || d.my > 32 || 2 * d.my + 1 > Y) return launch_planes_inv2<Y>(Z2, y, Sy, tw, d, nplanes, st);
  if (d.my <= 8) return launch_planes_inv3<Y, 8>(Z2, y, Sy, tw, d, nplanes, st);
  if (d.my <= 16) return launch_planes_inv3<Y, 16>(Z2, y, Sy, tw, d, nplanes, st);
  if (d.my <= 20) return launch_planes_inv3<Y, 20>(Z2, y, Sy, tw, d, nplanes, st);
  return launch_planes_inv3<Y, 32>(Z2, y, Sy, tw, d, nplanes, st);
}
template <int X, bool FWD>
int launch_xaxis(const cplx* in, cplx* out, const cplx* tw, SconvDims d, int ncol, int nslabs, cudaStream_t st) {
  typedef XaxisSmem<X> S;
  auto k = sconv_xaxis_kernel<X, FWD>;
  if (int rc = set_smem(k, S::BYTES)) return rc;
  const int npairs = (ncol + 1) / 2;
  TCFD_LAUNCH3(k, (npairs + S::GP - 1) / S::GP, nslabs, 1, S::GP * S::NT, S::BYTES, st, in, out, tw, d, ncol);
  return 0;
}

// second-generation x-axis kernels (persistent, prefetching); TCFD_SCONV_XAXIS = 1 forces the first generation
int xaxis_gen() {  // read at every call: tests switch generations inside one process
  const char* e = getenv("TCFD_SCONV_XAXIS");
  return e ? atoi(e) : 0;
}
template <int X>
int launch_xaxis_fwd2(const cplx* in, cplx* out, const cplx* tw, SconvDims d, int ncol, int nslabs, cudaStream_t st) {
  typedef Xaxis2Smem<X> S;
  auto k = sconv_xaxis_fwd2_kernel<X>;
  if (int rc = set_smem(k, S::FWD_BYTES)) return rc;
  const int ntx = (ncol / 2 + S::GP - 1) / S::GP, ntiles = ntx * nslabs;
  const int grid = planes_grid(reinterpret_cast<const void*>(k), S::GP * S::NT, S::FWD_BYTES, ntiles);
  TCFD_LAUNCH(k, grid, S::GP * S::NT, S::FWD_BYTES, st, in, out, tw, d, ncol, ntx, ntiles);
  return 0;
}
template <int X, int MXT>
int launch_xaxis_inv2(const cplx* in, cplx* out, const cplx* tw, SconvDims d, int ncol, int nslabs, cudaStream_t st) {
  typedef Xaxis2Smem<X> S;
  auto k = sconv_xaxis_inv2_kernel<X, MXT>;
  const size_t smem = S::inv_bytes(d.mx);
  if (int rc = set_smem(k, smem)) return rc;
  const int ntx = (ncol / 2 + S::GP - 1) / S::GP, ntiles = ntx * nslabs;
  const int grid = planes_grid(reinterpret_cast<const void*>(k), S::GP * S::NT, smem, ntiles);
  TCFD_LAUNCH(k, grid, S::GP * S::NT, smem, st, in, out, tw, d, ncol, ntx, ntiles);
  return 0;
}
template <int X>
int launch_xaxis_best(bool fwd, const cplx* in, cplx* out, const cplx* tw, SconvDims d, int ncol, int nslabs, cudaStream_t st) {
  const bool aligned = !((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) && (ncol % 2) == 0;
  if (xaxis_gen() == 1 || !aligned || (long long)((ncol / 2 + 7) / 8) * nslabs > 0x3fffffffLL)
    return fwd ? launch_xaxis<X, true>(in, out, tw, d, ncol, nslabs, st) : launch_xaxis<X, false>(in, out, tw, d, ncol, nslabs, st);
  if (fwd) return launch_xaxis_fwd2<X>(in, out, tw, d, ncol, nslabs, st);
  if (d.mx > 32 || 2 * d.mx > X) return launch_xaxis<X, false>(in, out, tw, d, ncol, nslabs, st);
  if (d.mx <= 8) return launch_xaxis_inv2<X, 8>(in, out, tw, d, ncol, nslabs, st);
  if (d.mx <= 16) return launch_xaxis_inv2<X, 16>(in, out, tw, d, ncol, nslabs, st);
  if (d.mx <= 20) return launch_xaxis_inv2<X, 20>(in, out, tw, d, ncol, nslabs, st);
  return launch_xaxis_inv2<X, 32>(in, out, tw, d, ncol, nslabs, st);
}

#define SCONV_SIZES(F) F(32) F(64) F(128) F(256) F(512)

int planes_fwd(int Y, const float* x, cplx* Z1, const cplx* A, const cplx* tw, SconvDims d, int np, cudaStream_t st) {
#define CASE(n) if (Y == n) return planes_v1() ? launch_planes_fwd<n>(x, Z1, A, tw, d, np, st) : launch_planes_fwd2<n>(x, Z1, A, tw, d, np, st);
  SCONV_SIZES(CASE)
#undef CASE
  return -1;
}
int planes_inv(int Y, const cplx* Z2, float* y, const cplx* Sy, const cplx* tw, SconvDims d, int np, cudaStream_t st) {
#define CASE(n) if (Y == n) return planes_v1() ? launch_planes_inv<n>(Z2, y, Sy, tw, d, np, st) : launch_planes_inv_best<n>(Z2, y, Sy, tw, d, np, st);
  SCONV_SIZES(CASE)
#undef CASE
  return -1;
}
int xaxis(int X, bool fwd, const cplx* in, cplx* out, const cplx* tw, SconvDims d, int ncol, int nslabs, cudaStream_t st) {
#define CASE(n) if (X == n) return launch_xaxis_best<n>(fwd, in, out, tw, d, ncol, nslabs, st);
  SCONV_SIZES(CASE)
#undef CASE
  return -1;
}

int check_launch(tcfd_sconv3d* h, int rc, const char* what) {
  h->launches++;
  if (rc == -100) return sfail(TCFD_ERR_INVALID, std::string(what) + ": shared-memory footprint exceeds 227 KB (T too large)");
  if (rc != 0) return sfail(TCFD_ERR_CUDA, std::string(what) + ": launch failed (" + std::to_string(rc) + ")");
#ifndef TCFD_EMU
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return sfail(TCFD_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
#endif
  return 0;
}

SconvDims dims_of(const tcfd_sconv3d* h, int Tin, int Tout, int nslabs) {
  SconvDims d;
  d.X = h->d.X; d.Y = h->d.Y; d.mx = h->d.mx; d.my = h->d.my; d.mt = h->d.mt;
  d.Tin = Tin; d.Tout = Tout; d.nplanes_c = nslabs;
  return d;
}
}  // namespace

extern "C" int tcfd_sconv3d_create(tcfd_sconv3d_t** out, const tcfd_sconv3d_desc_t* d) {
  if (!out || !d) return sfail(TCFD_ERR_INVALID, "null argument");
  *out = nullptr;
  if (!pow2_ok(d->X) || !pow2_ok(d->Y)) return sfail(TCFD_ERR_INVALID, "X and Y must be powers of two in [32, 512]");
  if (d->T_in < 1 || d->T_out < 1 || d->t_pad < 0) return sfail(TCFD_ERR_INVALID, "bad T_in / T_out / t_pad");
  if (d->Ci < 1 || d->Co < 1 || d->max_batch < 1) return sfail(TCFD_ERR_INVALID, "bad channel count or batch");
  const int Tn_in = d->T_in + d->t_pad, Tn_out = d->T_out + d->t_pad;
  if (d->mx < 1 || d->my < 1 || d->mt < 1 || 2 * d->mx > d->X || 2 * d->my > d->Y)
    return sfail(TCFD_ERR_INVALID, "modes must satisfy 1 <= mx <= X/2, 1 <= my <= Y/2 (overlapping corner blocks are not supported)");
  if (d->mt > Tn_in / 2 + 1)
    return sfail(TCFD_ERR_INVALID, "modes_t exceeds the " + std::to_string(Tn_in / 2 + 1) +
                                       " retained t-frequencies of the input (the reference's einsum raises here too)");
  if (d->norm < 0 || d->norm > 2) return sfail(TCFD_ERR_INVALID, "norm must be 0 (backward), 1 (ortho) or 2 (forward)");
  tcfd_sconv3d* h = new tcfd_sconv3d();
  h->d = *d;
  h->Tn_in = Tn_in;
  h->Tn_out = Tn_out;
  h->K = 4 * d->mx * d->my * d->mt;
  const double n_in = (double)d->X * d->Y * Tn_in, n_out = (double)d->X * d->Y * Tn_out;
  const double fs = d->norm == 0 ? 1.0 : (d->norm == 1 ? 1.0 / std::sqrt(n_in) : 1.0 / n_in);
  const double is = d->norm == 0 ? 1.0 / n_out : (d->norm == 1 ? 1.0 / std::sqrt(n_out) : 1.0);
  const int mt = d->mt;
  std::vector<cplx> A_f((size_t)mt * d->T_in), S_f((size_t)d->T_out * mt), A_b((size_t)mt * d->T_out), S_b((size_t)d->T_in * mt);
  for (int kt = 0; kt < mt; ++kt)
    for (int t = 0; t < d->T_in; ++t) {
      const double th = -2.0 * PI * (double)((long long)kt * (t + d->t_pad) % Tn_in) / (double)Tn_in;
      A_f[(size_t)kt * d->T_in + t] = cplx{(float)(fs * std::cos(th)), (float)(fs * std::sin(th))};
      S_b[(size_t)t * mt + kt] = cplx{(float)(fs * std::cos(th)), (float)(-fs * std::sin(th))};
    }
  for (int t = 0; t < d->T_out; ++t)
    for (int kt = 0; kt < mt; ++kt) {
      double s = 2.0;
      if (kt == 0 || (Tn_out % 2 == 0 && kt == Tn_out / 2)) s = 1.0;
      if (kt > Tn_out / 2) s = 0.0;  // beyond the output's Nyquist: dropped by irfftn(s=...)
      const double th = 2.0 * PI * (double)((long long)kt * (t + d->t_pad) % Tn_out) / (double)Tn_out;
      const double re = is * s * std::cos(th), im = is * s * std::sin(th);
      S_f[(size_t)t * mt + kt] = cplx{(float)re, (float)im};
      A_b[(size_t)kt * d->T_out + t] = cplx{(float)re, (float)(-im)};
    }
  int rc = 0;
  if (!rc) rc = upload_c(&h->twx, twiddles(d->X));
  if (!rc) rc = upload_c(&h->twy, twiddles(d->Y));
  if (!rc) rc = upload_c(&h->A_f, A_f);
  if (!rc) rc = upload_c(&h->S_f, S_f);
  if (!rc) rc = upload_c(&h->A_b, A_b);
  if (!rc) rc = upload_c(&h->S_b, S_b);
  const int cmax = d->Ci > d->Co ? d->Ci : d->Co;
  const size_t zb = (size_t)d->max_batch * cmax * d->X * 2 * d->my * mt * sizeof(cplx);
  const size_t hb = (size_t)d->max_batch * cmax * h->K * sizeof(cplx);
  if (!rc && cudaMalloc(&h->Z, zb) != cudaSuccess) rc = sfail(TCFD_ERR_NOMEM, "workspace allocation failed");
  if (!rc && cudaMalloc(&h->H1, hb) != cudaSuccess) rc = sfail(TCFD_ERR_NOMEM, "workspace allocation failed");
  if (!rc && cudaMalloc(&h->H2, hb) != cudaSuccess) rc = sfail(TCFD_ERR_NOMEM, "workspace allocation failed");
  h->ws_bytes = zb + 2 * hb;
  if (rc) {
    std::string keep = g_serr;
    tcfd_sconv3d_destroy(h);
    tcfd_set_last_error(keep.c_str());
    return rc;
  }
  *out = h;
  return TCFD_OK;
}

extern "C" int tcfd_sconv3d_destroy(tcfd_sconv3d_t* h) {
  if (!h) return TCFD_OK;
  void* all[] = {h->twx, h->twy, h->A_f, h->S_f, h->A_b, h->S_b, h->Z, h->H1, h->H2};
  for (void* p : all)
    if (p) cudaFree(p);
  delete h;
  return TCFD_OK;
}

extern "C" size_t tcfd_sconv3d_workspace_bytes(const tcfd_sconv3d_t* h) { return h ? h->ws_bytes : 0; }
extern "C" size_t tcfd_sconv3d_xhat_elems(const tcfd_sconv3d_t* h, int batch) {
  return h ? (size_t)batch * h->d.Ci * h->K : 0;
}
extern "C" int tcfd_sconv3d_last_launch_count(const tcfd_sconv3d_t* h) { return h ? h->launches : 0; }

namespace {
// TCFD_SCONV_MIX = 1: first-generation mode-mixing kernels everywhere, 2: second generation everywhere (A/B timing)
int mix_gen() {  // read at every call: tests switch generations inside one process
  const char* e = getenv("TCFD_SCONV_MIX");
  return e ? atoi(e) : 0;
}
// Yh = Xh x W (BWD = false) or gXh = gYh x conj(W)^T (BWD = true)
template <bool BWD>
int launch_mix2(tcfd_sconv3d* h, const cplx* in, cplx* out, const MixArgs& a, const SconvDims& dm, int batch, cudaStream_t st) {
  const int P = BWD ? a.Co : a.Ci, Q = BWD ? a.Ci : a.Co;
  const int bt = (batch + MIX_BT - 1) / MIX_BT;
  const int warps = (Q + MIX_OT - 1) / MIX_OT;
  const size_t smem = (size_t)MIX_BT * P * 32 * sizeof(cplx);
  // measured at C4 (profiles/r2i_*): the shared-tile kernel wins for the adjoint product (233 -> 102 us) but not for the
  // forward one (113 vs 126 us), which keeps the first-generation kernel unless TCFD_SCONV_MIX = 2
  const bool gen1 = mix_gen() == 1 || (!BWD && mix_gen() != 2);
  if (gen1 || (h->K & 1) || warps > 8 || smem > 160 * 1024) {
    // enough CTAs for every SM: split the batch tiles over grid.z when modes x channel tiles alone are few
    const int gx = (h->K + 127) / 128, gy = warps;
    int gz = 1184 / (gx * gy);
    gz = gz < 1 ? 1 : (gz > bt ? bt : gz);
    if (BWD) TCFD_LAUNCH3(sconv_mix_bwd_x_kernel, gx, gy, gz, 128, 0, st, in, out, a, dm);
    else TCFD_LAUNCH3(sconv_mix_fwd_kernel, gx, gy, gz, 128, 0, st, in, out, a, dm);
    return 0;
  }
  auto k = sconv_mix2_kernel<BWD>;
  if (int rc = set_smem(k, smem)) return rc;
  // one wave of co-resident CTAs: the batch tiles are split over grid.y only as far as the SMs have room
  const int gx = (h->K + 31) / 32;
  const int cap = planes_grid(reinterpret_cast<const void*>(k), warps * 32, smem, 1 << 30);
  int gy = cap / gx;
  gy = gy < 1 ? 1 : (gy > bt ? bt : gy);
  TCFD_LAUNCH3(k, gx, gy, 1, warps * 32, smem, st, in, out, a, dm);
  return 0;
}
// analysis half: x -> truncated spectrum Xh (kept for backward when xhat_save is given) -> per-mode channel mix -> Yh
int analysis_impl(tcfd_sconv3d* h, const void* x, const void* const* w, const void* const* bias, float delta, cplx* Yh,
                  void* xhat_save, int batch, cudaStream_t st) {
  const tcfd_sconv3d_desc_t& d = h->d;
  const int ncol = 2 * d.my * d.mt;
  cplx* Xh = xhat_save ? static_cast<cplx*>(xhat_save) : static_cast<cplx*>(h->H1);
  cplx* Z = static_cast<cplx*>(h->Z);
  int rc;
  SconvDims dm = dims_of(h, d.T_in, d.T_out, batch * d.Ci);
  rc = check_launch(h, planes_fwd(d.Y, static_cast<const float*>(x), Z, static_cast<const cplx*>(h->A_f),
                                  static_cast<const cplx*>(h->twy), dm, batch * d.Ci * d.X, st), "planes_fwd");
  if (rc) return rc;
  rc = check_launch(h, xaxis(d.X, true, Z, Xh, static_cast<const cplx*>(h->twx), dm, ncol, batch * d.Ci, st), "xaxis_fwd");
  if (rc) return rc;
  MixArgs a{};
  for (int c = 0; c < 4; ++c) {
    a.w[c] = static_cast<const cplx*>(w[c]);
    a.bias[c] = bias ? static_cast<const cplx*>(bias[c]) : nullptr;
  }
  a.B = batch; a.Ci = d.Ci; a.Co = d.Co; a.delta = delta;
  if (int rc2 = launch_mix2<false>(h, Xh, Yh, a, dm, batch, st)) return rc2;
  return check_launch(h, 0, "mix_fwd");
}
// synthesis half: truncated spectrum Yh (batch, Co, 2mx, 2my, mt) -> y
int synthesis_impl(tcfd_sconv3d* h, const cplx* Yh, void* y, int batch, cudaStream_t st) {
  const tcfd_sconv3d_desc_t& d = h->d;
  const int ncol = 2 * d.my * d.mt;
  cplx* Z = static_cast<cplx*>(h->Z);
  SconvDims dm = dims_of(h, d.T_in, d.T_out, batch * d.Co);
  int rc = check_launch(h, xaxis(d.X, false, Yh, Z, static_cast<const cplx*>(h->twx), dm, ncol, batch * d.Co, st), "xaxis_inv");
  if (rc) return rc;
  return check_launch(h, planes_inv(d.Y, Z, static_cast<float*>(y), static_cast<const cplx*>(h->S_f),
                                    static_cast<const cplx*>(h->twy), dm, batch * d.Co * d.X, st), "planes_inv");
}
// adjoint of the synthesis half: grad_y -> gradient with respect to Yh (torch convention for complex tensors)
int synthesis_bwd_impl(tcfd_sconv3d* h, const void* grad_y, cplx* gYh, int batch, cudaStream_t st) {
  const tcfd_sconv3d_desc_t& d = h->d;
  const int ncol = 2 * d.my * d.mt;
  cplx* Z = static_cast<cplx*>(h->Z);
  SconvDims dm = dims_of(h, d.T_out, d.T_in, batch * d.Co);
  int rc = check_launch(h, planes_fwd(d.Y, static_cast<const float*>(grad_y), Z, static_cast<const cplx*>(h->A_b),
                                      static_cast<const cplx*>(h->twy), dm, batch * d.Co * d.X, st), "planes_fwd(bwd)");
  if (rc) return rc;
  return check_launch(h, xaxis(d.X, true, Z, gYh, static_cast<const cplx*>(h->twx), dm, ncol, batch * d.Co, st), "xaxis_fwd(bwd)");
}
// adjoint of the analysis half: gYh -> grad_w / grad_bias (needs the saved Xh) and grad_x
int analysis_bwd_impl(tcfd_sconv3d* h, const cplx* gYh, const void* xhat, const void* const* w, void* grad_x,
                      void* const* grad_w, void* const* grad_bias, float delta, int batch, cudaStream_t st) {
  const tcfd_sconv3d_desc_t& d = h->d;
  const int ncol = 2 * d.my * d.mt;
  cplx* gXh = static_cast<cplx*>(h->H2);
  cplx* Z = static_cast<cplx*>(h->Z);
  int rc;
  SconvDims dm = dims_of(h, d.T_out, d.T_in, batch * d.Co);
  MixArgs a{};
  for (int c = 0; c < 4; ++c) {
    if (!w[c]) return sfail(TCFD_ERR_INVALID, "null weight pointer");
    a.w[c] = static_cast<const cplx*>(w[c]);
    a.gw[c] = grad_w ? static_cast<cplx*>(grad_w[c]) : nullptr;
    a.gbias[c] = grad_bias ? static_cast<cplx*>(grad_bias[c]) : nullptr;
  }
  a.B = batch; a.Ci = d.Ci; a.Co = d.Co; a.delta = delta;
  if (grad_w) {
    for (int c = 0; c < 4; ++c)
      if (!grad_w[c]) return sfail(TCFD_ERR_INVALID, "null grad_w pointer");
    if (mix_gen() == 1)
      TCFD_LAUNCH3(sconv_mix_bwd_w_kernel, (h->K + 127) / 128, d.Co, d.Ci, 128, 0, st, static_cast<const cplx*>(xhat), gYh, a, dm);
    else
      TCFD_LAUNCH3(sconv_mix_bwd_w2_kernel, (h->K + 127) / 128, (d.Co + MIX_WT - 1) / MIX_WT, (d.Ci + MIX_WT - 1) / MIX_WT, 128, 0, st,
                   static_cast<const cplx*>(xhat), gYh, a, dm);
    if ((rc = check_launch(h, 0, "mix_bwd_w"))) return rc;
  }
  if (grad_x) {
    if ((rc = launch_mix2<true>(h, gYh, gXh, a, dm, batch, st))) return rc;
    if ((rc = check_launch(h, 0, "mix_bwd_x"))) return rc;
    dm = dims_of(h, d.T_out, d.T_in, batch * d.Ci);
    rc = check_launch(h, xaxis(d.X, false, gXh, Z, static_cast<const cplx*>(h->twx), dm, ncol, batch * d.Ci, st), "xaxis_inv(bwd)");
    if (rc) return rc;
    rc = check_launch(h, planes_inv(d.Y, Z, static_cast<float*>(grad_x), static_cast<const cplx*>(h->S_b),
                                    static_cast<const cplx*>(h->twy), dm, batch * d.Ci * d.X, st), "planes_inv(bwd)");
    if (rc) return rc;
  }
  return TCFD_OK;
}
int check_call(const tcfd_sconv3d* h, int batch) {
  if (!h) return sfail(TCFD_ERR_INVALID, "null argument");
  if (batch < 1 || batch > h->d.max_batch) return sfail(TCFD_ERR_INVALID, "batch outside [1, max_batch]");
  return 0;
}
}  // namespace

extern "C" size_t tcfd_sconv3d_yhat_elems(const tcfd_sconv3d_t* h, int batch) {
  return h ? (size_t)batch * h->d.Co * h->K : 0;
}

extern "C" int tcfd_sconv3d_forward(tcfd_sconv3d_t* h, const void* x, const void* const* w, const void* const* bias,
                                    float delta, void* y, void* xhat_save, int batch, void* stream_) {
  if (!h || !x || !w || !y) return sfail(TCFD_ERR_INVALID, "null argument");
  if (int rc = check_call(h, batch)) return rc;
  for (int c = 0; c < 4; ++c)
    if (!w[c]) return sfail(TCFD_ERR_INVALID, "null weight pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  h->launches = 0;
  cplx* Yh = static_cast<cplx*>(h->H2);
  if (int rc = analysis_impl(h, x, w, bias, delta, Yh, xhat_save, batch, st)) return rc;
  return synthesis_impl(h, Yh, y, batch, st);
}

extern "C" int tcfd_sconv3d_backward(tcfd_sconv3d_t* h, const void* grad_y, const void* xhat, const void* const* w,
                                     void* grad_x, void* const* grad_w, void* const* grad_bias, float delta, int batch,
                                     void* stream_) {
  if (!h || !grad_y || !xhat || !w) return sfail(TCFD_ERR_INVALID, "null argument");
  if (int rc = check_call(h, batch)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  h->launches = 0;
  cplx* gYh = static_cast<cplx*>(h->H1);
  if (int rc = synthesis_bwd_impl(h, grad_y, gYh, batch, st)) return rc;
  return analysis_bwd_impl(h, gYh, xhat, w, grad_x, grad_w, grad_bias, delta, batch, st);
}

// ---- the two halves as separate calls (a spectral post-process or a change of mesh sits between them)
extern "C" int tcfd_sconv3d_analysis(tcfd_sconv3d_t* h, const void* x, const void* const* w, const void* const* bias,
                                     float delta, void* yhat, void* xhat_save, int batch, void* stream_) {
  if (!h || !x || !w || !yhat) return sfail(TCFD_ERR_INVALID, "null argument");
  if (int rc = check_call(h, batch)) return rc;
  for (int c = 0; c < 4; ++c)
    if (!w[c]) return sfail(TCFD_ERR_INVALID, "null weight pointer");
  h->launches = 0;
  return analysis_impl(h, x, w, bias, delta, static_cast<cplx*>(yhat), xhat_save, batch, static_cast<cudaStream_t>(stream_));
}
extern "C" int tcfd_sconv3d_synthesis(tcfd_sconv3d_t* h, const void* yhat, void* y, int batch, void* stream_) {
  if (!h || !yhat || !y) return sfail(TCFD_ERR_INVALID, "null argument");
  if (int rc = check_call(h, batch)) return rc;
  h->launches = 0;
  return synthesis_impl(h, static_cast<const cplx*>(yhat), y, batch, static_cast<cudaStream_t>(stream_));
}
extern "C" int tcfd_sconv3d_synthesis_backward(tcfd_sconv3d_t* h, const void* grad_y, void* grad_yhat, int batch, void* stream_) {
  if (!h || !grad_y || !grad_yhat) return sfail(TCFD_ERR_INVALID, "null argument");
  if (int rc = check_call(h, batch)) return rc;
  h->launches = 0;
  return synthesis_bwd_impl(h, grad_y, static_cast<cplx*>(grad_yhat), batch, static_cast<cudaStream_t>(stream_));
}
extern "C" int tcfd_sconv3d_analysis_backward(tcfd_sconv3d_t* h, const void* grad_yhat, const void* xhat, const void* const* w,
                                              void* grad_x, void* const* grad_w, void* const* grad_bias, float delta, int batch,
                                              void* stream_) {
  if (!h || !grad_yhat || !xhat || !w) return sfail(TCFD_ERR_INVALID, "null argument");
  if (int rc = check_call(h, batch)) return rc;
  h->launches = 0;
  return analysis_bwd_impl(h, static_cast<const cplx*>(grad_yhat), xhat, w, grad_x, grad_w, grad_bias, delta, batch,
                           static_cast<cudaStream_t>(stream_));
}
