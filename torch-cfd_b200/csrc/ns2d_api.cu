// C ABI of hot path A (see include/tcfd.h).  Host-side driver: owns tables, workspaces and the
// launch sequence of one RK4+CN step; the arithmetic lives in ns2d_kernels.cuh.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tcfd.h"
#include "ns2d_flow.cuh"
#include "ns2d_plan.h"
#include "tma.cuh"
#ifndef TCFD_EMU
#include <cudaTypedefs.h>
#endif

#define TCFD_DECL(prec, n) extern "C" void tcfd_ns2d_entry_##prec##_##n(tcfd_ns2d_entry_t*);
#define TCFD_SIZES(X, prec) X(prec, 32) X(prec, 64) X(prec, 128) X(prec, 256) X(prec, 512) X(prec, 1024) X(prec, 2048)
TCFD_SIZES(TCFD_DECL, 32)
TCFD_SIZES(TCFD_DECL, 64)

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return fail(TCFD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));          \
  } while (0)

bool find_entry(int prec, int n, tcfd_ns2d_entry_t* e) {
#define TCFD_TRY(p, nn)                                      \
  if (prec == p && n == nn) {                                \
    tcfd_ns2d_entry_##p##_##nn(e);                           \
    return true;                                             \
  }
  TCFD_SIZES(TCFD_TRY, 32)
  TCFD_SIZES(TCFD_TRY, 64)
  return false;
}
}  // namespace

struct tcfd_ns2d {
  int n = 0, nh = 0, prec = 0, max_batch = 0, KF = 0, num_sms = 1;
  int chunk = 1;  // samples per L2-resident chunk
  size_t es = 0;  // sizeof(real)
  tcfd_ns2d_entry_t entry{};
  void *tw = nullptr, *kappa_x = nullptr, *kappa_y = nullptr, *nil = nullptr, *lin = nullptr,
       *filt = nullptr, *fhat = nullptr, *tab = nullptr, *frow = nullptr, *tabU = nullptr, *maskU = nullptr;
  void *hA = nullptr, *hB = nullptr, *wS = nullptr, *wT = nullptr, *H = nullptr, *advt = nullptr;
  void *stage_in = nullptr, *stage_out = nullptr, *stage_dw = nullptr;  // step_host staging
  tcfd::TileMaps maps{};  // TMA descriptors of the H tiles (second-generation cols kernel)
  // step_host pipeline: copy-in / copy-out streams and per-chunk events
  cudaStream_t s_in = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_done, ev_free;
  cudaEvent_t ev_start = nullptr, ev_end = nullptr;
  size_t ws_bytes = 0;
  int launches = 0;
  // whole-call resident kernel for n <= 64 (ns2d_small.cuh)
  bool small = false;
  // third-generation persistent dataflow schedule (ns2d_flow.cuh)
  bool flow = false;
  int* sync_dev = nullptr;   // ticket + per-sample counters
  int* err_host = nullptr;   // mapped pinned word the kernel raises on a dependency time-out
  int* err_dev = nullptr;
  unsigned long long* prof_dev = nullptr;  // region cycle counters of the profiling kernel variant (TCFD_FLOW_PROF=1)
  void* slab = nullptr;      // one allocation behind H, advt, wS, hA, wT, hB (dataflow schedule)
  tcfd_flow_window_t win{};  // L2 access-policy window over the hot part of the slab
  // measurement mode (tcfd_ns2d_step_timed): every launch is bracketed by events
  bool timed = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> ev_kind;
  size_t state_bytes(int b) const { return (size_t)b * n * nh * 2 * es; }
};

extern "C" const char* tcfd_last_error(void) { return g_err.c_str(); }
// internal: lets the other translation units of the library report through tcfd_last_error()
extern "C" void tcfd_set_last_error(const char* msg) { g_err = msg ? msg : ""; }
extern "C" const char* tcfd_version(void) {
#ifdef TCFD_EMU
  return "tcfd 0.1 host-emulation (tests only)";
#else
  return "tcfd 0.1 sm_100a";
#endif
}

namespace {
template <class T>
int upload(void** dst, const void* src, size_t count) {
  CUDA_TRY(cudaMalloc(dst, count * sizeof(T)));
  CUDA_TRY(cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

template <class T>
int create_tables(tcfd_ns2d* h, const tcfd_ns2d_desc_t* d) {
  const int n = h->n, nh = h->nh;
  // twiddles exp(-2 pi i j / n), evaluated in double
  std::vector<T> tw(2 * (size_t)n);
  const double PI = 3.14159265358979323846264338327950288;
  for (int j = 0; j < n; ++j) {
    // exact octant symmetries are not needed at this accuracy: cos/sin of double arguments
    const double a = -2.0 * PI * (double)j / (double)n;
    tw[2 * j] = (T)std::cos(a);
    tw[2 * j + 1] = (T)std::sin(a);
  }
  int rc;
  if ((rc = upload<T>(&h->tw, tw.data(), tw.size()))) return rc;
  // kappa / n^2 : exact (n is a power of two) -- carries the 1/n^2 of the inverse transforms
  const T scale = (T)(1.0 / ((double)n * (double)n));
  std::vector<T> kx(n), ky(nh);
  for (int i = 0; i < n; ++i) kx[i] = static_cast<const T*>(d->kappa_x)[i] * scale;
  for (int i = 0; i < nh; ++i) ky[i] = static_cast<const T*>(d->kappa_y)[i] * scale;
  if ((rc = upload<T>(&h->kappa_x, kx.data(), n))) return rc;
  if ((rc = upload<T>(&h->kappa_y, ky.data(), nh))) return rc;
  if ((rc = upload<T>(&h->nil, d->neg_inv_lap, (size_t)n * nh))) return rc;
  if ((rc = upload<T>(&h->lin, d->linear_term, (size_t)n * nh))) return rc;
  h->KF = nh;
  if (d->filter) {
    if ((rc = upload<T>(&h->filt, d->filter, (size_t)n * nh))) return rc;
    // rows |kx| that carry at least one unmasked mode
    const T* f = static_cast<const T*>(d->filter);
    int kf = 0;
    for (int p = 0; p < nh; ++p) {
      const int r1 = p, r2 = (n - p) % n;
      bool any = false;
      for (int c = 0; c < nh && !any; ++c) any = (f[(size_t)r1 * nh + c] != T(0)) || (f[(size_t)r2 * nh + c] != T(0));
      if (any) kf = p + 1;
    }
    h->KF = kf > 0 ? kf : 1;
  }
  // interleaved copy {linear_term, mask, -1/laplace', 0} for the second-generation kernels
  {
    std::vector<tcfd::tab4<T>> tab((size_t)n * nh);
    const T* lin = static_cast<const T*>(d->linear_term);
    const T* nil = static_cast<const T*>(d->neg_inv_lap);
    const T* f = static_cast<const T*>(d->filter);
    for (size_t i = 0; i < tab.size(); ++i) tab[i] = tcfd::tab4<T>{lin[i], f ? f[i] : T(1), nil[i], T(0)};
    if ((rc = upload<tcfd::tab4<T>>(&h->tab, tab.data(), tab.size()))) return rc;
    std::vector<unsigned char> zero(n, 0);
    if ((rc = upload<unsigned char>(&h->frow, zero.data(), n))) return rc;
    if (h->entry.v2) {
      // per-d blocks for the substage kernel: {lin_a, lin_b, nil_a, nil_b} per column (one copy serves
      // rows r and n-r: the tables must be even in kx) and the 0/1 mask as bits per half
      const int nd = n / 4 + 1;
      const size_t tab_row = ((size_t)nh * 4 * sizeof(T) + 15) / 16 * 16, mask_row = ((size_t)nh + 15) / 16 * 16;
      std::vector<unsigned char> tabU((size_t)nd * tab_row, 0), maskU((size_t)nd * 2 * mask_row, 0);
      for (int r = 1; r < n; ++r)
        for (int c = 0; c < nh; ++c)
          if (lin[(size_t)r * nh + c] != lin[(size_t)(n - r) * nh + c] || nil[(size_t)r * nh + c] != nil[(size_t)(n - r) * nh + c])
            return fail(TCFD_ERR_INVALID, "linear_term and neg_inv_lap must be even in kx (row r == row n-r)");
      for (size_t i = 0; f && i < (size_t)n * nh; ++i)
        if (f[i] != T(0) && f[i] != T(1)) return fail(TCFD_ERR_INVALID, "filter must be a 0/1 mask");
      for (int dd = 0; dd < nd; ++dd) {
        const int ra = 2 * dd, rb = (2 * dd + 1 <= n / 2) ? 2 * dd + 1 : ra;
        const int rows[2][2] = {{ra, rb}, {(n - ra) % n, (n - rb) % n}};
        T* tb = reinterpret_cast<T*>(tabU.data() + (size_t)dd * tab_row);
        for (int c = 0; c < nh; ++c) {
          tb[2 * c + 0] = lin[(size_t)ra * nh + c];  // plane 0: lin, lanes interleaved
          tb[2 * c + 1] = lin[(size_t)rb * nh + c];
          tb[2 * nh + c] = nil[(size_t)ra * nh + c];  // plane 1: nil of lane a
          tb[3 * nh + c] = nil[(size_t)rb * nh + c];  // plane 2: nil of lane b
          for (int hf = 0; hf < 2; ++hf) {
            const bool ka = !f || f[(size_t)rows[hf][0] * nh + c] != T(0), kb = !f || f[(size_t)rows[hf][1] * nh + c] != T(0);
            maskU[((size_t)dd * 2 + hf) * mask_row + c] = (unsigned char)((ka ? 1 : 0) | (kb ? 2 : 0));
          }
        }
      }
      if ((rc = upload<unsigned char>(&h->tabU, tabU.data(), tabU.size()))) return rc;
      if ((rc = upload<unsigned char>(&h->maskU, maskU.data(), maskU.size()))) return rc;
    }
  }
  return 0;
}

template <class T>
void fill_params(const tcfd_ns2d* h, tcfd::NsParams<T>& p, int batch) {
  std::memset(&p, 0, sizeof(p));
  p.B = batch;
  p.KF = h->KF;
  p.H = static_cast<tcfd::cx<T>*>(h->H);
  p.advt = static_cast<tcfd::cx<T>*>(h->advt);
  p.H2 = static_cast<T*>(h->H);
  p.advt2 = static_cast<T*>(h->advt);
  p.NDF = (h->KF + 1) / 2;
  p.tab = static_cast<const tcfd::tab4<T>*>(h->tab);
  p.frow = static_cast<const unsigned char*>(h->frow);
  p.tabU = h->tabU;
  if (const char* e = getenv("TCFD_DBG")) p.dbg = atoi(e);
  p.maskU = static_cast<const unsigned char*>(h->maskU);
  p.tw = static_cast<const tcfd::cx<T>*>(h->tw);
  p.kappa_x = static_cast<const T*>(h->kappa_x);
  p.kappa_y = static_cast<const T*>(h->kappa_y);
  p.nil = static_cast<const T*>(h->nil);
  p.lin = static_cast<const T*>(h->lin);
  p.filt = static_cast<const T*>(h->filt);
  p.fhat = static_cast<const tcfd::cx<T>*>(h->fhat);
}

// bound of a dependency wait of the dataflow kernel in SM cycles (about 2 GHz): TCFD_FLOW_TIMEOUT_S seconds, default 3,
// 0 = wait for ever (cuda-gdb, compute-sanitizer, time-sliced GPUs)
long long flow_wait_cycles() {
  double s = 3.0;
  if (const char* e = getenv("TCFD_FLOW_TIMEOUT_S")) s = atof(e);
  return s <= 0.0 ? 0ll : (long long)(s * 2.0e9);
}

// TMA descriptor over H2 = [chunk][N rows kx][N] packed-complex entries (4 reals each): a tile is 4
// consecutive entries of every row of one sample.
int make_tile_maps(tcfd_ns2d* h) {
  const size_t ent = 4 * h->es;
  const size_t row_bytes = (size_t)h->n * ent;
#ifdef TCFD_EMU
  h->maps.base = static_cast<const unsigned char*>(h->H);
  h->maps.row_bytes = row_bytes;
  h->maps.sample_bytes = row_bytes * h->n;
  h->maps.adv_base = static_cast<const unsigned char*>(h->advt);
  return 0;
#else
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) return fail(TCFD_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
  auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  const CUtensorMapDataType dt = h->prec == 32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
  const CUtensorMapSwizzle sw = h->prec == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  const cuuint64_t gdim[3] = {(cuuint64_t)h->n * 4, (cuuint64_t)h->n, (cuuint64_t)h->chunk};
  const cuuint64_t gstr[2] = {(cuuint64_t)row_bytes, (cuuint64_t)row_bytes * h->n};
  const cuuint32_t estr[3] = {1, 1, 1};
  const cuuint32_t rows = (cuuint32_t)(h->n < 256 ? h->n : 256);
  const cuuint32_t box_main[3] = {16, rows, 1};  // 16 reals = 4 entries = 64 B (fp32) / 128 B (fp64)
  CUresult r = encode(&h->maps.main, dt, 3, h->H, gdim, gstr, box_main, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TCFD_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  if (h->flow) {
    // advection rows of the dataflow schedule: [chunk][JB blocks][N columns][8 rows][4 reals] (tma.cuh)
    const cuuint64_t jb = (cuuint64_t)((h->n / 4 + 1 + tcfd::ADV_BLOCK - 1) / tcfd::ADV_BLOCK);
    const cuuint64_t adim[4] = {4, (cuuint64_t)tcfd::ADV_BLOCK, (cuuint64_t)h->n, (cuuint64_t)h->chunk * jb};
    const cuuint64_t astr[3] = {(cuuint64_t)ent, (cuuint64_t)ent * tcfd::ADV_BLOCK, (cuuint64_t)ent * tcfd::ADV_BLOCK * h->n};
    const cuuint32_t aest[4] = {1, 1, 1, 1};
    const cuuint32_t abox[4] = {4, 1, rows, 1};
    r = encode(&h->maps.adv, dt, 4, h->advt, adim, astr, abox, aest, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(TCFD_ERR_CUDA, "cuTensorMapEncodeTiled (advection rows) failed with CUresult " + std::to_string((int)r));
  }
  return 0;
#endif
}

int launch(tcfd_ns2d* h, int which, const void* params, void* stream) {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->timed) {
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventRecord(e0, static_cast<cudaStream_t>(stream)));
  }
  int rc = h->entry.launch(which, params, h->entry.v2 ? &h->maps : nullptr, h->num_sms, stream);
  if (h->timed) {
    CUDA_TRY(cudaEventRecord(e1, static_cast<cudaStream_t>(stream)));
    h->ev.push_back(e0);
    h->ev.push_back(e1);
    h->ev_kind.push_back(which);
  }
  h->launches++;
  if (rc != 0) return fail(TCFD_ERR_CUDA, std::string("kernel launch failed: ") + cudaGetErrorString((cudaError_t)rc));
  return 0;
}

template <class T>
int step_impl(tcfd_ns2d* h, const void* w_in_, void* w_out_, void* dwdt_, int batch, int steps, int nstages,
              const double* beta, const double* gdt, const double* mu, double inv_total_dt, void* stream) {
  typedef tcfd::cx<T> C;
  const int total = steps * nstages;
  const size_t per = (size_t)h->n * h->nh;  // spectrum entries per sample
  // Chunk-major schedule: all substages of all steps run on `chunk` samples before the next chunk
  // starts, so a chunk's state (w, h), H and advt stay resident in the 126 MB L2 from launch to
  // launch and HBM sees each state entry once per call instead of 18 S per step.
  for (int c0 = 0; c0 < batch; c0 += h->chunk) {
    const int cb = (batch - c0 < h->chunk) ? batch - c0 : h->chunk;
    const C* w_in = static_cast<const C*>(w_in_) + (size_t)c0 * per;
    C* w_out = static_cast<C*>(w_out_) + (size_t)c0 * per;
    C* dwdt = dwdt_ ? static_cast<C*>(dwdt_) + (size_t)c0 * per : nullptr;
    tcfd::NsParams<T> p;
    fill_params<T>(h, p, cb);
    // prologue: H[4] from the initial state
    p.w_in = w_in;
    int rc;
    if ((rc = launch(h, TCFD_K_ROWS_INV, &p, stream))) return rc;
    const C* src = w_in;
    for (int j = 0; j < total; ++j) {
      const int k = j % nstages;
      if ((rc = launch(h, TCFD_K_COLS, &p, stream))) return rc;
      const bool last_sub = (j == total - 1);
      C* dst = ((total - 1 - j) % 2 == 0) ? w_out : static_cast<C*>(h->wS);
      p.mode = tcfd::UPD_RK;
      p.w_in = src;
      p.w_out = dst;
      p.h_in = static_cast<const C*>((j % 2) ? h->hB : h->hA);
      p.h_out = static_cast<C*>((j % 2) ? h->hA : h->hB);
      if (h->entry.v2) {
        // the state travels in the unit layout between substages (wS / wT), only the first
        // substage reads and the last one writes the caller's reference layout
        typedef tcfd::cx<typename tcfd::pack2<T>::type> CU;
        p.in_user = (j == 0) ? 1 : 0;
        p.out_user = last_sub ? 1 : 0;
        p.w_in = w_in;    // reference-layout input (used when in_user; also dwdt's w_old)
        p.w_out = w_out;  // reference-layout output (used when out_user)
        p.wU_in = static_cast<const CU*>((j % 2) ? h->wS : h->wT);
        p.wU_out = static_cast<CU*>((j % 2) ? h->wT : h->wS);
        p.hU_in = static_cast<const CU*>((j % 2) ? h->hB : h->hA);
        p.hU_out = static_cast<CU*>((j % 2) ? h->hA : h->hB);
      }
      p.read_h = (k > 0 && beta[k] != 0.0) ? 1 : 0;
      p.write_h = (k + 1 < nstages && beta[k + 1] != 0.0) ? 1 : 0;
      p.beta = (T)beta[k];
      p.gdt = (T)gdt[k];
      p.mu = (T)mu[k];
      const bool last = (j == total - 1);
      p.w_old = nullptr;
      p.dwdt = nullptr;
      if (last && dwdt) {
        p.w_old = w_in;
        p.dwdt = dwdt;
        p.inv_tdt = (T)inv_total_dt;
      }
      if ((rc = launch(h, last ? TCFD_K_ROWS_FWD : TCFD_K_ROWS_FULL, &p, stream))) return rc;
      src = dst;
    }
  }
  return 0;
}

// One persistent launch for the whole call (ns2d_flow.cuh): chunk-major dataflow schedule.
template <class T>
int flow_step_impl(tcfd_ns2d* h, const void* w_in, void* w_out, void* dwdt, int batch, int steps, int nstages,
                   const double* beta, const double* gdt, const double* mu, double inv_total_dt, void* stream) {
  typedef tcfd::cx<T> C;
  typedef tcfd::cx<typename tcfd::pack2<T>::type> CU;
  tcfd::FlowParams<T> fp;
  std::memset(&fp, 0, sizeof(fp));
  fill_params<T>(h, fp.p, batch);
  fp.p.mode = tcfd::UPD_RK;
  fp.p.w_in = static_cast<const C*>(w_in);
  fp.p.w_out = static_cast<C*>(w_out);
  if (dwdt) {
    fp.p.w_old = static_cast<const C*>(w_in);
    fp.p.dwdt = static_cast<C*>(dwdt);
    fp.p.inv_tdt = (T)inv_total_dt;
    fp.w0U = static_cast<CU*>(h->wT);
  }
  fp.nsub = steps * nstages;
  fp.nstages = nstages;
  fp.W = h->chunk;
  for (int k = 0; k < nstages; ++k) {
    fp.beta[k] = (T)beta[k];
    fp.gdt[k] = (T)gdt[k];
    fp.mu[k] = (T)mu[k];
    fp.rd_h[k] = (k > 0 && beta[k] != 0.0) ? 1 : 0;
    fp.wr_h[k] = (k + 1 < nstages && beta[k + 1] != 0.0) ? 1 : 0;
  }
  fp.wU = static_cast<CU*>(h->wS);
  fp.hU = static_cast<CU*>(h->hA);
  fp.sync = h->sync_dev;
  fp.err = h->err_dev;
  fp.wait_cycles = flow_wait_cycles();
  fp.prof = h->prof_dev;
  CUDA_TRY(cudaMemsetAsync(h->sync_dev, 0, sizeof(int) * (2 + 2 * (size_t)h->max_batch), static_cast<cudaStream_t>(stream)));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->timed) {
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventRecord(e0, static_cast<cudaStream_t>(stream)));
  }
  // grouped items (3 double rows / 4 quads per ticket) only when this call brings enough samples to
  // keep every SM busy with them; small calls (e.g. the sub-batches of tcfd_ns2d_step_host) use single units
  tcfd_flow_window_t win = h->win;
  win.grouped = (h->n >= 512 && (batch < h->chunk ? batch : h->chunk) >= 24) ? 1 : 0;
  const int rc = h->entry.launch_flow(&fp, &h->maps, h->num_sms, stream, &win);
  if (h->timed) {
    CUDA_TRY(cudaEventRecord(e1, static_cast<cudaStream_t>(stream)));
    h->ev.push_back(e0);
    h->ev.push_back(e1);
    h->ev_kind.push_back(TCFD_K_ROWS_FULL);
  }
  h->launches++;
  if (rc != 0) return fail(TCFD_ERR_CUDA, std::string("flow kernel launch failed: ") + cudaGetErrorString((cudaError_t)rc));
  return 0;
}

// Small grids (n <= 64): the whole call as ONE launch, one CTA per sample, state resident in shared memory
// (ns2d_small.cuh).  TCFD_SMALL=0 selects the two-launches-per-sub-stage kernels instead (cross-check).
template <class T>
int small_step_impl(tcfd_ns2d* h, const void* w_in, void* w_out, void* dwdt, int batch, int steps, int nstages,
                    const double* beta, const double* gdt, const double* mu, double inv_total_dt, void* stream) {
  typedef tcfd::cx<T> C;
  tcfd::FlowParams<T> fp;
  std::memset(&fp, 0, sizeof(fp));
  fill_params<T>(h, fp.p, batch);
  fp.p.mode = tcfd::UPD_RK;
  fp.p.w_in = static_cast<const C*>(w_in);
  fp.p.w_out = static_cast<C*>(w_out);
  if (dwdt) {
    fp.p.dwdt = static_cast<C*>(dwdt);
    fp.p.inv_tdt = (T)inv_total_dt;
  }
  fp.nsub = steps * nstages;
  fp.nstages = nstages;
  for (int k = 0; k < nstages; ++k) {
    fp.beta[k] = (T)beta[k];
    fp.gdt[k] = (T)gdt[k];
    fp.mu[k] = (T)mu[k];
    fp.rd_h[k] = (k > 0 && beta[k] != 0.0) ? 1 : 0;
    fp.wr_h[k] = (k + 1 < nstages && beta[k + 1] != 0.0) ? 1 : 0;
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->timed) {
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventRecord(e0, static_cast<cudaStream_t>(stream)));
  }
  const int rc = h->entry.launch_small(&fp, batch, stream);
  if (h->timed) {
    CUDA_TRY(cudaEventRecord(e1, static_cast<cudaStream_t>(stream)));
    h->ev.push_back(e0);
    h->ev.push_back(e1);
    h->ev_kind.push_back(TCFD_K_ROWS_FULL);
  }
  h->launches++;
  if (rc != 0) return fail(TCFD_ERR_CUDA, std::string("resident small-grid kernel launch failed: ") + cudaGetErrorString((cudaError_t)rc));
  return 0;
}

template <class T>
int eval_impl(tcfd_ns2d* h, int mode, const void* w_in, const void* wt_in, void* out, int batch, void* stream) {
  typedef tcfd::cx<T> C;
  const size_t per = (size_t)h->n * h->nh;
  for (int c0 = 0; c0 < batch; c0 += h->chunk) {
    const int cb = (batch - c0 < h->chunk) ? batch - c0 : h->chunk;
    tcfd::NsParams<T> p;
    fill_params<T>(h, p, cb);
    p.w_in = static_cast<const C*>(w_in) + (size_t)c0 * per;
    int rc;
    if ((rc = launch(h, TCFD_K_ROWS_INV, &p, stream))) return rc;
    if ((rc = launch(h, TCFD_K_COLS, &p, stream))) return rc;
    p.mode = mode;
    p.w_old = wt_in ? static_cast<const C*>(wt_in) + (size_t)c0 * per : nullptr;
    p.h_out = static_cast<C*>(out) + (size_t)c0 * per;
    if ((rc = launch(h, TCFD_K_ROWS_EVAL, &p, stream))) return rc;
  }
  return 0;
}

int check_batch(const tcfd_ns2d* h, int batch) {
  if (!h) return fail(TCFD_ERR_INVALID, "null handle");
  if (batch < 1 || batch > h->max_batch)
    return fail(TCFD_ERR_INVALID, "batch " + std::to_string(batch) + " outside [1, max_batch=" + std::to_string(h->max_batch) + "]");
  return 0;
}
}  // namespace

extern "C" int tcfd_ns2d_create(tcfd_ns2d_t** out, const tcfd_ns2d_desc_t* d) {
  if (!out || !d) return fail(TCFD_ERR_INVALID, "null argument");
  *out = nullptr;
  if (d->prec != 32 && d->prec != 64) return fail(TCFD_ERR_INVALID, "prec must be 32 or 64");
  if (d->max_batch < 1) return fail(TCFD_ERR_INVALID, "max_batch must be >= 1");
  if (!d->kappa_x || !d->kappa_y || !d->neg_inv_lap || !d->linear_term)
    return fail(TCFD_ERR_INVALID, "kappa_x, kappa_y, neg_inv_lap and linear_term are required");
  tcfd_ns2d_entry_t e{};
  if (!find_entry(d->prec, d->n, &e))
    return fail(TCFD_ERR_INVALID, "unsupported grid size n=" + std::to_string(d->n) +
                                      " (supported: powers of two 32..2048)");
  if (d->prec == 64 && d->n > 1024)
    return fail(TCFD_ERR_INVALID, "n = " + std::to_string(d->n) + " is served in fp32 only (the fp64 column tile of " +
                                      std::to_string((d->n / 2 + 1) * 256 / 1024) + " KB does not fit in shared memory)");
  tcfd_ns2d* h = new tcfd_ns2d();
  h->n = d->n;
  h->nh = d->n / 2 + 1;
  h->prec = d->prec;
  h->es = d->prec / 8;
  h->max_batch = d->max_batch;
  h->entry = e;
#ifndef TCFD_EMU
  {
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (ce != cudaSuccess) {
      delete h;
      return fail(TCFD_ERR_CUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(ce));
    }
  }
#endif
  int rc = (d->prec == 32) ? create_tables<float>(h, d) : create_tables<double>(h, d);
  if (rc == 0 && d->f_hat) rc = tcfd_ns2d_set_forcing(h, d->f_hat);
  if (rc == 0) {
    // chunk: samples per launch.  Default = the whole batch: with two launches per substage a
    // chunk small enough to keep H L2-resident (~8 samples at 512^2) under-fills the 148 SMs and
    // loses more than the saved DRAM traffic (measured: profiles/r01 chunk sweep).  TCFD_CHUNK_MB
    // caps the per-chunk working set for experiments.
    double budget_mb = 1e9;
    if (const char* e = getenv("TCFD_CHUNK_MB")) budget_mb = atof(e) > 0 ? atof(e) : budget_mb;
    const double per_sample_mb = (double)h->nh * h->n * 2 * h->es * 8.0 / 1e6;  // H(4) + advt + w, h, wS
    const double want = budget_mb / per_sample_mb;
    int chunk = want >= (double)h->max_batch ? h->max_batch : (int)want;
    if (chunk < 1) chunk = 1;
    // Persistent dataflow schedule (default for N >= 256; TCFD_FLOW=0 selects the two-launch
    // schedule): the workspaces hold W samples (TCFD_FLOW_W).  Measured on B200 (profiles/r04_flow_sweep.md):
    // the step is bound by per-warp instruction latency at 8-10 resident warps per SM, not by DRAM -- a
    // window small enough to stay L2-resident (W ~ 8 at 512^2, with or without a persisting-L2 access
    // window) cuts DRAM traffic 3x and gains nothing, while its shorter dependency distance costs
    // 10-40 %; default W = 64.
    h->flow = h->entry.launch_flow != nullptr;
    if (const char* e = getenv("TCFD_FLOW")) h->flow = h->flow && atoi(e) != 0;
    h->small = h->entry.launch_small != nullptr;
    if (const char* e = getenv("TCFD_SMALL")) h->small = h->small && atoi(e) != 0;
    if (h->flow) {
      int W = 64;
      if (const char* e = getenv("TCFD_FLOW_W")) W = atoi(e) > 0 ? atoi(e) : W;
      chunk = W < h->max_batch ? W : h->max_batch;
      if (cudaMalloc(reinterpret_cast<void**>(&h->sync_dev), sizeof(int) * (2 + 2 * (size_t)h->max_batch)) != cudaSuccess ||
          cudaHostAlloc(reinterpret_cast<void**>(&h->err_host), sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
          cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->err_dev), h->err_host, 0) != cudaSuccess)
        rc = fail(TCFD_ERR_NOMEM, "flow schedule: synchronisation block allocation failed");
      else
        *h->err_host = 0;
      if (rc == 0 && getenv("TCFD_FLOW_PROF") && atoi(getenv("TCFD_FLOW_PROF"))) {
        if (cudaMalloc(reinterpret_cast<void**>(&h->prof_dev), 16 * sizeof(unsigned long long)) == cudaSuccess)
          cudaMemset(h->prof_dev, 0, 16 * sizeof(unsigned long long));
      }
    }
    h->chunk = chunk;
    const size_t sb = h->state_bytes(h->chunk);
    // advt: v1 [B][nh][n] complex; v2 [B][n/4+1][n][4 reals] (slightly larger)
    // (dataflow schedule: rows padded to whole blocks of ADV_BLOCK, tma.cuh)
    const size_t adv_rows = h->flow ? (size_t)((h->n / 4 + 1 + tcfd::ADV_BLOCK - 1) / tcfd::ADV_BLOCK) * tcfd::ADV_BLOCK : (size_t)(h->n / 4 + 1);
    const size_t ab = (size_t)h->chunk * adv_rows * h->n * 4 * h->es;
    // unit-layout state (v2): [B][n/4+1][2][nh] entries of 4 reals
    const size_t ub = (size_t)h->chunk * (h->n / 4 + 1) * 2 * h->nh * 4 * h->es;
    // H: first generation [B][nh][n/yt][4][yt] = 4 * nh * n complex per sample; second generation
    // [B][n][n] packed-complex entries of 4 reals (z layout, ns2d_v2.cuh)
    const size_t hb = h->entry.v2 ? (size_t)h->chunk * h->n * h->n * 4 * h->es
                                  : (size_t)h->chunk * h->nh * h->n * 4 * 2 * h->es;
    void** bufs[] = {&h->H, &h->advt, &h->wS, &h->hA, &h->wT, &h->hB};  // hot buffers first
    size_t nbs[6];
    for (int i = 0; i < 6; ++i) {
      void** b = bufs[i];
      size_t nb = (b == &h->advt && ab > sb) ? ab : sb;
      if (h->entry.v2 && b != &h->advt) nb = ub;
      if (b == &h->H) nb = hb;
      if (!h->entry.v2 && b == &h->wT) nb = 0;
      nbs[i] = (nb + 1023) / 1024 * 1024;
    }
    if (rc == 0 && h->flow) {
      // one slab, so that a single access-policy window covers the buffers the schedule keeps in L2
      size_t total = 0;
      for (size_t nb : nbs) total += nb;
      if (cudaMalloc(&h->slab, total) != cudaSuccess) {
        rc = fail(TCFD_ERR_NOMEM, "workspace allocation failed");
      } else {
        h->ws_bytes += total;
        size_t off = 0;
        for (int i = 0; i < 6; ++i) {
          *bufs[i] = static_cast<unsigned char*>(h->slab) + off;
          off += nbs[i];
        }
        // grouped items (3 double rows / 4 quads per ticket) need a wide window to keep every SM busy
        // (>= ~1000 grouped rows items per phase: 512^2 from 24 samples, 1024^2 from 12 -- measured 809 vs 734 steps/s at
        // 1024^2 x 16; at 256^2 the one-warp CTAs prefer single units: 3430 vs 2330 steps/s, profiles/r2o_flow_sweep.jsonl)
        h->win.grouped = (h->n >= 512 && (long long)h->chunk * ((h->n / 4 + 1 + 2) / 3) >= 1000) ? 1 : 0;
#ifndef TCFD_EMU
        const size_t hot = total - nbs[5];  // everything but hB (used by the two-launch schedule only)
        int dev = 0, max_persist = 0, max_win = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        // persisting-L2 window over the hot workspace: experiment knob, off by default (no gain measured,
        // and cudaLimitPersistingL2CacheSize is a device-wide setting)
        int want_persist = 0;
        if (const char* e = getenv("TCFD_FLOW_PERSIST")) want_persist = atoi(e);
        if (want_persist && hot <= (size_t)max_persist && hot <= (size_t)max_win &&
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, hot) == cudaSuccess) {
          h->win.base = h->slab;
          h->win.bytes = hot;
          h->win.hit_ratio = 1.0f;
        }
        if (getenv("TCFD_FLOW_VERBOSE"))
          fprintf(stderr, "tcfd: flow window W=%d hot=%.1f MB max_persist=%.1f MB max_window=%.1f MB persisting=%d\n",
                  h->chunk, hot / 1e6, max_persist / 1e6, max_win / 1e6, h->win.bytes ? 1 : 0);
#endif
      }
    } else if (rc == 0) {
      for (int i = 0; i < 6; ++i) {
        if (!nbs[i]) continue;
        if (cudaMalloc(bufs[i], nbs[i]) != cudaSuccess) { rc = fail(TCFD_ERR_NOMEM, "workspace allocation failed"); break; }
        h->ws_bytes += nbs[i];
      }
    }
    if (rc == 0 && h->entry.v2) rc = make_tile_maps(h);
  }
  if (rc != 0) {
    std::string keep = g_err;
    tcfd_ns2d_destroy(h);
    g_err = keep;
    return rc;
  }
  *out = h;
  return TCFD_OK;
}

extern "C" int tcfd_ns2d_destroy(tcfd_ns2d_t* h) {
  if (!h) return TCFD_OK;
  void* all[] = {h->tw, h->kappa_x, h->kappa_y, h->nil, h->lin, h->filt, h->fhat, h->tab, h->frow, h->tabU, h->maskU,
                 h->stage_in, h->stage_out, h->stage_dw};
  for (void* p : all)
    if (p) cudaFree(p);
  if (h->slab) {
    cudaFree(h->slab);
  } else {
    void* ws[] = {h->wT, h->hA, h->hB, h->wS, h->H, h->advt};
    for (void* p : ws)
      if (p) cudaFree(p);
  }
  if (h->sync_dev) cudaFree(h->sync_dev);
  if (h->prof_dev) cudaFree(h->prof_dev);
  if (h->err_host) cudaFreeHost(h->err_host);
  for (cudaEvent_t e : h->ev_in) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_done) cudaEventDestroy(e);
  if (h->ev_start) cudaEventDestroy(h->ev_start);
  if (h->ev_end) cudaEventDestroy(h->ev_end);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  delete h;
  return TCFD_OK;
}

extern "C" int tcfd_ns2d_set_forcing(tcfd_ns2d_t* h, const void* f_hat) {
  if (!h) return fail(TCFD_ERR_INVALID, "null handle");
  const size_t bytes = (size_t)h->n * h->nh * 2 * h->es;
  if (!f_hat) {
    if (h->fhat) cudaFree(h->fhat);
    h->fhat = nullptr;
    return TCFD_OK;
  }
  if (!h->fhat) CUDA_TRY(cudaMalloc(&h->fhat, bytes));
  CUDA_TRY(cudaMemcpy(h->fhat, f_hat, bytes, cudaMemcpyHostToDevice));
  // rows that carry a non-zero forcing entry (the kernels skip the f_hat loads of all other rows)
  std::vector<unsigned char> frow(h->n, 0);
  const size_t rb = (size_t)h->nh * 2 * h->es;
  for (int r = 0; r < h->n; ++r) {
    const unsigned char* row = static_cast<const unsigned char*>(f_hat) + (size_t)r * rb;
    bool any = false;
    if (h->prec == 32) {
      const float* v = reinterpret_cast<const float*>(row);
      for (int i = 0; i < 2 * h->nh && !any; ++i) any = v[i] != 0.0f;
    } else {
      const double* v = reinterpret_cast<const double*>(row);
      for (int i = 0; i < 2 * h->nh && !any; ++i) any = v[i] != 0.0;
    }
    frow[r] = any ? 1 : 0;
  }
  CUDA_TRY(cudaMemcpy(h->frow, frow.data(), h->n, cudaMemcpyHostToDevice));
  return TCFD_OK;
}

extern "C" int tcfd_ns2d_check(const tcfd_ns2d_t* h) {
  if (!h) return fail(TCFD_ERR_INVALID, "null handle");
  if (h->err_host && *reinterpret_cast<volatile int*>(h->err_host) != 0)
    return fail(TCFD_ERR_CUDA, "dataflow schedule: a dependency wait timed out in an earlier call; results are invalid");
  return TCFD_OK;
}

// experiments only (not declared in tcfd.h): region cycle counters of the profiling kernel variant
// (-DTCFD_FLOW_VARIANTS build, TCFD_FLOW_PROF=1, TCFD_FLOW_G="3,4,-64"); synchronises the device, reads and clears
extern "C" int tcfd_ns2d_flow_profile(tcfd_ns2d_t* h, unsigned long long* out16) {
  if (!h || !out16 || !h->prof_dev) return fail(TCFD_ERR_INVALID, "no profile counters on this handle");
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(out16, h->prof_dev, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemset(h->prof_dev, 0, 16 * sizeof(unsigned long long)));
  return TCFD_OK;
}

extern "C" int tcfd_ns2d_schedule(const tcfd_ns2d_t* h) { return (h && (h->flow || h->small)) ? 1 : 0; }

extern "C" size_t tcfd_ns2d_workspace_bytes(const tcfd_ns2d_t* h) { return h ? h->ws_bytes : 0; }
extern "C" int tcfd_ns2d_last_launch_count(const tcfd_ns2d_t* h) { return h ? h->launches : 0; }

extern "C" int tcfd_ns2d_step(tcfd_ns2d_t* h, const void* w_in, void* w_out, void* dwdt, int batch, int steps,
                              int nstages, const double* beta, const double* gdt, const double* mu,
                              double inv_total_dt, void* stream) {
  int rc = check_batch(h, batch);
  if (rc) return rc;
  if (!w_in || !w_out || !beta || !gdt || !mu) return fail(TCFD_ERR_INVALID, "null argument");
  if (w_in == w_out) return fail(TCFD_ERR_INVALID, "w_out must not alias w_in");
  if (steps < 1 || nstages < 1) return fail(TCFD_ERR_INVALID, "steps and nstages must be >= 1");
  h->launches = 0;
  if ((rc = tcfd_ns2d_check(h))) return rc;
  if (h->small && nstages <= tcfd::FLOW_MAX_STAGES)
    return h->prec == 32
               ? small_step_impl<float>(h, w_in, w_out, dwdt, batch, steps, nstages, beta, gdt, mu, inv_total_dt, stream)
               : small_step_impl<double>(h, w_in, w_out, dwdt, batch, steps, nstages, beta, gdt, mu, inv_total_dt, stream);
  if (h->flow && nstages <= tcfd::FLOW_MAX_STAGES) {
    const int nd = h->n / 4 + 1, nq = h->n / 4;
    const double items = (double)batch * ((double)nd + (double)steps * nstages * (nd + nq));
    if (items < 1.0e9)
      return h->prec == 32
                 ? flow_step_impl<float>(h, w_in, w_out, dwdt, batch, steps, nstages, beta, gdt, mu, inv_total_dt, stream)
                 : flow_step_impl<double>(h, w_in, w_out, dwdt, batch, steps, nstages, beta, gdt, mu, inv_total_dt, stream);
  }
  return h->prec == 32
             ? step_impl<float>(h, w_in, w_out, dwdt, batch, steps, nstages, beta, gdt, mu, inv_total_dt, stream)
             : step_impl<double>(h, w_in, w_out, dwdt, batch, steps, nstages, beta, gdt, mu, inv_total_dt, stream);
}

extern "C" int tcfd_ns2d_step_timed(tcfd_ns2d_t* h, const void* w_in, void* w_out, int batch, int steps, int nstages,
                                    const double* beta, const double* gdt, const double* mu, void* stream,
                                    float* ms, int* count) {
  if (!h || !ms || !count) return fail(TCFD_ERR_INVALID, "null argument");
  h->timed = true;
  int rc = tcfd_ns2d_step(h, w_in, w_out, nullptr, batch, steps, nstages, beta, gdt, mu, 0.0, stream);
  h->timed = false;
  for (int i = 0; i < 4; ++i) { ms[i] = 0.f; count[i] = 0; }
  cudaError_t ce = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  for (size_t i = 0; i < h->ev_kind.size(); ++i) {
    float t = 0.f;
    if (ce == cudaSuccess && rc == 0) ce = cudaEventElapsedTime(&t, h->ev[2 * i], h->ev[2 * i + 1]);
    ms[h->ev_kind[i]] += t;
    count[h->ev_kind[i]]++;
    cudaEventDestroy(h->ev[2 * i]);
    cudaEventDestroy(h->ev[2 * i + 1]);
  }
  h->ev.clear();
  h->ev_kind.clear();
  if (rc) return rc;
  if (ce != cudaSuccess) return fail(TCFD_ERR_CUDA, std::string("timed step: ") + cudaGetErrorString(ce));
  return TCFD_OK;
}

extern "C" int tcfd_ns2d_explicit_terms(tcfd_ns2d_t* h, const void* w_in, void* f_out, int batch, void* stream) {
  int rc = check_batch(h, batch);
  if (rc) return rc;
  if (!w_in || !f_out) return fail(TCFD_ERR_INVALID, "null argument");
  h->launches = 0;
  return h->prec == 32 ? eval_impl<float>(h, tcfd::UPD_F, w_in, nullptr, f_out, batch, stream)
                       : eval_impl<double>(h, tcfd::UPD_F, w_in, nullptr, f_out, batch, stream);
}

extern "C" int tcfd_ns2d_residual(tcfd_ns2d_t* h, const void* w_in, const void* wt_in, void* r_out, int batch,
                                  void* stream) {
  int rc = check_batch(h, batch);
  if (rc) return rc;
  if (!w_in || !wt_in || !r_out) return fail(TCFD_ERR_INVALID, "null argument");
  h->launches = 0;
  return h->prec == 32 ? eval_impl<float>(h, tcfd::UPD_RESID, w_in, wt_in, r_out, batch, stream)
                       : eval_impl<double>(h, tcfd::UPD_RESID, w_in, wt_in, r_out, batch, stream);
}

namespace {
template <class T, class O>
int record_impl(tcfd_ns2d* h, const void* w, const void* dwdt, const void* res, void* sw, void* spsi, void* sdw,
                void* sres, int batch, int n_t, int it, void* stream) {
  const size_t per = (size_t)h->n * h->nh, total = per * batch;
  const int threads = 256;
  size_t blocks = (total + threads - 1) / threads;
  const size_t cap = (size_t)h->num_sms * 8;
#ifndef TCFD_EMU
  if (blocks > cap) blocks = cap;
#else
  (void)cap;
  if (blocks > 4) blocks = 4;
#endif
  TCFD_LAUNCH((tcfd::ns2d_record_kernel<T, O>), (unsigned)blocks, threads, 0, static_cast<cudaStream_t>(stream),
              static_cast<const tcfd::cx<T>*>(w), static_cast<const tcfd::cx<T>*>(dwdt),
              static_cast<const tcfd::cx<T>*>(res), static_cast<const T*>(h->nil), static_cast<tcfd::cx<O>*>(sw),
              static_cast<tcfd::cx<O>*>(spsi), static_cast<tcfd::cx<O>*>(sdw), static_cast<tcfd::cx<O>*>(sres), per, n_t, it,
              total);
  h->launches++;
#ifndef TCFD_EMU
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TCFD_ERR_CUDA, std::string("record kernel: ") + cudaGetErrorString(e));
#endif
  return 0;
}
}  // namespace

extern "C" int tcfd_ns2d_record(tcfd_ns2d_t* h, const void* w, const void* dwdt, const void* res, void* snap_w,
                                void* snap_psi, void* snap_dwdt, void* snap_res, int batch, int n_t, int it,
                                int out_prec, void* stream) {
  if (!h || !w) return fail(TCFD_ERR_INVALID, "null argument");
  if (batch < 1 || n_t < 1 || it < 0 || it >= n_t) return fail(TCFD_ERR_INVALID, "bad batch / n_t / it");
  if ((snap_dwdt && !dwdt) || (snap_res && !res)) return fail(TCFD_ERR_INVALID, "snapshot requested without its source");
  if (out_prec != 32 && out_prec != 64) return fail(TCFD_ERR_INVALID, "out_prec must be 32 or 64");
  if (h->prec == 32)
    return out_prec == 32 ? record_impl<float, float>(h, w, dwdt, res, snap_w, snap_psi, snap_dwdt, snap_res, batch, n_t, it, stream)
                          : record_impl<float, double>(h, w, dwdt, res, snap_w, snap_psi, snap_dwdt, snap_res, batch, n_t, it, stream);
  return out_prec == 32 ? record_impl<double, float>(h, w, dwdt, res, snap_w, snap_psi, snap_dwdt, snap_res, batch, n_t, it, stream)
                        : record_impl<double, double>(h, w, dwdt, res, snap_w, snap_psi, snap_dwdt, snap_res, batch, n_t, it, stream);
}

extern "C" int tcfd_ns2d_step_host(tcfd_ns2d_t* h, const void* w_in_host, void* w_out_host, void* dwdt_host,
                                   int batch, int steps, int nstages, const double* beta, const double* gdt,
                                   const double* mu, double inv_total_dt, void* stream_) {
  int rc = check_batch(h, batch);
  if (rc) return rc;
  if (!w_in_host || !w_out_host) return fail(TCFD_ERR_INVALID, "null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t sb = h->state_bytes(h->max_batch);
  if (!h->stage_in) { CUDA_TRY(cudaMalloc(&h->stage_in, sb)); h->ws_bytes += sb; }
  if (!h->stage_out) { CUDA_TRY(cudaMalloc(&h->stage_out, sb)); h->ws_bytes += sb; }
  if (dwdt_host && !h->stage_dw) { CUDA_TRY(cudaMalloc(&h->stage_dw, sb)); h->ws_bytes += sb; }
  // Three-stage pipeline over sub-batches: upload (copy-in stream) | step (caller's stream) | download
  // (copy-out stream).  PCIe is full duplex, so the uploads of chunk c+1 and the downloads of chunk
  // c-1 run under the kernels of chunk c; the caller's stream finally waits for the last download.
  int nchunks = 8;
  if (const char* e = getenv("TCFD_HOST_CHUNKS")) nchunks = atoi(e) > 0 ? atoi(e) : nchunks;
  if (nchunks > batch) nchunks = batch;
  if (!h->s_in) {
    CUDA_TRY(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_end, cudaEventDisableTiming));
  }
  while ((int)h->ev_in.size() < nchunks) {
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    h->ev_in.push_back(a);
    h->ev_done.push_back(b);
  }
  // the staging buffers may still be in use by work queued earlier on the caller's stream
  CUDA_TRY(cudaEventRecord(h->ev_start, stream));
  CUDA_TRY(cudaStreamWaitEvent(h->s_in, h->ev_start, 0));
  CUDA_TRY(cudaStreamWaitEvent(h->s_out, h->ev_start, 0));
  const size_t per = (size_t)h->n * h->nh * 2 * h->es;  // bytes per sample
  const int base = batch / nchunks, extra = batch % nchunks;
  int b0 = 0;
  for (int c = 0; c < nchunks; ++c) {
    const int cb = base + (c < extra ? 1 : 0);
    const size_t off = (size_t)b0 * per, nb = (size_t)cb * per;
    unsigned char* din = static_cast<unsigned char*>(h->stage_in) + off;
    unsigned char* dout = static_cast<unsigned char*>(h->stage_out) + off;
    unsigned char* ddw = dwdt_host ? static_cast<unsigned char*>(h->stage_dw) + off : nullptr;
    CUDA_TRY(cudaMemcpyAsync(din, static_cast<const unsigned char*>(w_in_host) + off, nb, cudaMemcpyHostToDevice, h->s_in));
    CUDA_TRY(cudaEventRecord(h->ev_in[c], h->s_in));
    CUDA_TRY(cudaStreamWaitEvent(stream, h->ev_in[c], 0));
    rc = tcfd_ns2d_step(h, din, dout, ddw, cb, steps, nstages, beta, gdt, mu, inv_total_dt, stream_);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev_done[c], stream));
    CUDA_TRY(cudaStreamWaitEvent(h->s_out, h->ev_done[c], 0));
    CUDA_TRY(cudaMemcpyAsync(static_cast<unsigned char*>(w_out_host) + off, dout, nb, cudaMemcpyDeviceToHost, h->s_out));
    if (dwdt_host)
      CUDA_TRY(cudaMemcpyAsync(static_cast<unsigned char*>(dwdt_host) + off, ddw, nb, cudaMemcpyDeviceToHost, h->s_out));
    b0 += cb;
  }
  CUDA_TRY(cudaEventRecord(h->ev_end, h->s_out));
  CUDA_TRY(cudaStreamWaitEvent(stream, h->ev_end, 0));
  return TCFD_OK;
}
