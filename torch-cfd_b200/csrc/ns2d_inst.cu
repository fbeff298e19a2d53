// One translation unit per (precision, N): compiled with -DTCFD_PREC=32|64 -DTCFD_N=<n>.
// Exports a C launcher table entry used by ns2d_api.cu.  N >= 256 gets the second-generation
// kernels (ns2d_v2.cuh), smaller grids the CTA-tiled ones (ns2d_kernels.cuh).
#include "ns2d_flow.cuh"
#include "ns2d_small.cuh"
#include "ns2d_plan.h"
#include <cstdio>
#include <cstdlib>

#if TCFD_PREC == 32
typedef float real_t;
#else
typedef double real_t;
#endif

namespace {
using namespace tcfd;
constexpr int N = TCFD_N;
constexpr int NT = N / 8;
constexpr bool V2 = N >= 256;
typedef cx<real_t> cplx;

template <class K>
int prep(K kernel, size_t smem, int threads, int* occ) {
#ifndef TCFD_EMU
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kernel, threads, smem);
  if (e != cudaSuccess) return (int)e;
  if (*occ < 1) *occ = 1;
#else
  *occ = 1;
#endif
  return 0;
}

// occupancy / shared-memory attributes are PER DEVICE (cudaFuncSetAttribute applies to the current device
// only): the launchers cache them per device so that one process can drive several GPUs
constexpr int MAX_DEV = 64;
inline int cur_dev() {
#ifndef TCFD_EMU
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEV) d = 0;
  return d;
#else
  return 0;
#endif
}

inline int grid_for(int work, int num_sms, int o) {
#ifdef TCFD_EMU
  (void)num_sms;
  (void)o;
  return work;
#else
  const int cap = num_sms * (o > 0 ? o : 1);
  return work < cap ? work : cap;
#endif
}

#if TCFD_N < 256
// ------------------------------------------------------------------ first generation (N <= 128)
namespace v1 {
constexpr int CTA = 256;
constexpr int G = (CTA / NT) > 0 ? (CTA / NT) : 1;
constexpr int GC = G < N / 2 ? G : N / 2;
constexpr int YT = 2 * GC;
constexpr bool PP = true;
constexpr size_t smem_rows(bool inv) { return (size_t)G * ((inv ? 2 : 1) * N) * (PP ? 2 : 1) * sizeof(cplx); }
constexpr size_t smem_cols() {
  return ((size_t)(N / 2 + 1) * (4 * YT + 1) + (size_t)GC * N * (PP ? 2 : 1)) * sizeof(cplx);
}

int launch(int which, const NsParams<real_t>& p, int num_sms, cudaStream_t stream) {
  static int occ_all[MAX_DEV][4];
  int* occ = occ_all[cur_dev()];
  const int nblk_rows = (p.B * (N / 2 + 1) + G - 1) / G;
  const int ntiles = p.B * (N / YT);
  int rc = 0;
  switch (which) {
    case TCFD_K_ROWS_INV: {
      auto k = ns2d_rows_kernel<real_t, N, G, YT, false, true, PP>;
      if (!occ[0] && (rc = prep(k, smem_rows(true), G * NT, &occ[0]))) return rc;
      TCFD_LAUNCH(k, grid_for(nblk_rows, num_sms, occ[0]), G * NT, smem_rows(true), stream, p);
      break;
    }
    case TCFD_K_ROWS_FULL: {
      auto k = ns2d_rows_kernel<real_t, N, G, YT, true, true, PP>;
      if (!occ[1] && (rc = prep(k, smem_rows(true), G * NT, &occ[1]))) return rc;
      TCFD_LAUNCH(k, grid_for(nblk_rows, num_sms, occ[1]), G * NT, smem_rows(true), stream, p);
      break;
    }
    case TCFD_K_ROWS_EVAL:
    case TCFD_K_ROWS_FWD: {
      auto k = ns2d_rows_kernel<real_t, N, G, YT, true, false, PP>;
      if (!occ[2] && (rc = prep(k, smem_rows(false), G * NT, &occ[2]))) return rc;
      TCFD_LAUNCH(k, grid_for(nblk_rows, num_sms, occ[2]), G * NT, smem_rows(false), stream, p);
      break;
    }
    case TCFD_K_COLS: {
      auto k = ns2d_cols_kernel<real_t, N, GC, PP>;
      if (!occ[3] && (rc = prep(k, smem_cols(), GC * NT, &occ[3]))) return rc;
      TCFD_LAUNCH(k, grid_for(ntiles, num_sms, occ[3]), GC * NT, smem_cols(), stream, p);
      break;
    }
    default:
      return -1;
  }
  return 0;
}
}  // namespace v1
#else
// ------------------------------------------------------------------ second generation (N >= 256)
namespace v2 {
constexpr int NTC = NT >= 32 ? NT : 32;  // (only instantiated for N >= 256)
// resident threads per SM the register budget is tuned for: 512 (fp32: <= 128 regs) / 256 (fp64)
constexpr int TARGET_THREADS = sizeof(real_t) == 4 ? 512 : 256;
typedef pack2<real_t>::type lane_t;
constexpr int MINB_BY_THREADS = (TARGET_THREADS / NTC) > 0 ? (TARGET_THREADS / NTC) : 1;
// CTAs per SM the shared-memory footprint allows (227 KB usable, 1 KB reserved per CTA)
constexpr int ROWS_BY_SMEM = (int)(232448 / (RowsSmem<real_t, N>::BYTES + 1024));
#ifdef TCFD_ROWS3_MINB
constexpr int MINB_ROWS3 = TCFD_ROWS3_MINB;
#else
// one less than shared memory would allow: 168 instead of 128 registers per thread (no spills)
constexpr int ROWS3_CAP = ROWS_BY_SMEM > 2 ? ROWS_BY_SMEM - 1 : ROWS_BY_SMEM;
constexpr int MINB_ROWS3 = ROWS3_CAP < 1 ? 1 : (ROWS3_CAP < MINB_BY_THREADS ? ROWS3_CAP : MINB_BY_THREADS);
#endif
constexpr int MINB_ROWS = MINB_BY_THREADS;
constexpr int MINB_COLS = 5;
constexpr size_t smem_rows() { return (size_t)N * sizeof(cx<lane_t>); }
constexpr size_t smem_rows3() { return (size_t)RowsSmem<real_t, N>::BYTES; }
constexpr size_t smem_cols() {
  return 1024 + (size_t)TileGeom<N, 4 * (int)sizeof(cx<lane_t>)>::BYTES + (size_t)N * sizeof(cx<lane_t>) + 16;
}

int launch(int which, const NsParams<real_t>& p, const TileMaps* maps, int num_sms, cudaStream_t stream) {
  static int occ_all[MAX_DEV][7];
  int* occ = occ_all[cur_dev()];
  const int nunits = p.B * (N / 4 + 1);
  const int nquads = p.B * (N / 4);
  int rc = 0;
  switch (which) {
    case TCFD_K_ROWS_INV: {
      auto k = ns2d_rows2_kernel<real_t, N, false, true, ROWS_RK, MINB_ROWS>;
      if (!occ[0] && (rc = prep(k, smem_rows(), NT, &occ[0]))) return rc;
      TCFD_LAUNCH(k, grid_for(nunits, num_sms, occ[0]), NT, smem_rows(), stream, p);
      break;
    }
    case TCFD_K_ROWS_FULL: {  // substage: forward rows + update + inverse rows
      if (p.in_user) {
        auto k = ns2d_rows3_kernel<real_t, N, true, true, false, MINB_ROWS3>;
        if (!occ[1] && (rc = prep(k, smem_rows3(), NT, &occ[1]))) return rc;
        TCFD_LAUNCH(k, grid_for(nunits, num_sms, occ[1]), NT, smem_rows3(), stream, p);
      } else {
        auto k = ns2d_rows3_kernel<real_t, N, true, false, false, MINB_ROWS3>;
        if (!occ[5] && (rc = prep(k, smem_rows3(), NT, &occ[5]))) return rc;
        TCFD_LAUNCH(k, grid_for(nunits, num_sms, occ[5]), NT, smem_rows3(), stream, p);
      }
      break;
    }
    case TCFD_K_ROWS_FWD: {  // last substage: forward rows + update, reference layout out
      if (p.in_user) {
        auto k = ns2d_rows3_kernel<real_t, N, false, true, true, MINB_ROWS3>;
        if (!occ[2] && (rc = prep(k, smem_rows3(), NT, &occ[2]))) return rc;
        TCFD_LAUNCH(k, grid_for(nunits, num_sms, occ[2]), NT, smem_rows3(), stream, p);
      } else {
        auto k = ns2d_rows3_kernel<real_t, N, false, false, true, MINB_ROWS3>;
        if (!occ[6] && (rc = prep(k, smem_rows3(), NT, &occ[6]))) return rc;
        TCFD_LAUNCH(k, grid_for(nunits, num_sms, occ[6]), NT, smem_rows3(), stream, p);
      }
      break;
    }
    case TCFD_K_ROWS_EVAL: {
      auto k = ns2d_rows2_kernel<real_t, N, true, false, ROWS_EVAL, MINB_ROWS>;
      if (!occ[4] && (rc = prep(k, smem_rows(), NT, &occ[4]))) return rc;
      TCFD_LAUNCH(k, grid_for(nunits, num_sms, occ[4]), NT, smem_rows(), stream, p);
      break;
    }
    case TCFD_K_COLS: {
      if (!maps) return -2;
      auto k = ns2d_cols2_kernel<real_t, N, MINB_COLS>;
      if (!occ[3] && (rc = prep(k, smem_cols(), NT, &occ[3]))) return rc;
      TCFD_LAUNCH(k, grid_for(nquads, num_sms, occ[3]), NT, smem_cols(), stream, p, *maps);
      break;
    }
    default:
      return -1;
  }
  return 0;
}

#ifndef TCFD_FLOW_MINB
#define TCFD_FLOW_MINB MINB_BY_THREADS
#endif
constexpr size_t FLOW_SMEM = (size_t)FlowSmem<real_t, N>::BYTES;       // single-transform exchange buffer
constexpr int FLOW_BY_SMEM = (int)(232448 / (FLOW_SMEM + 1024));
constexpr int FLOW_MINB = FLOW_BY_SMEM < 1 ? 1 : (FLOW_BY_SMEM < TCFD_FLOW_MINB ? FLOW_BY_SMEM : TCFD_FLOW_MINB);

// MAXR: CTAs per SM the register budget is tuned for (0 = FLOW_MINB)
template <int GR, int GC, int MAXR, int MODE = 0>
int launch_flow_g(const FlowParams<real_t>& fp, const TileMaps* maps, int num_sms, cudaStream_t stream,
                  const tcfd_flow_window_t* win) {
  static int occ_all[MAX_DEV];
  int& occ = occ_all[cur_dev()];
  auto k = ns2d_flow_kernel<real_t, N, (MAXR > 0 ? MAXR : FLOW_MINB), GR, GC, MODE>;
  constexpr size_t smem = (MODE & 2) ? (size_t)FlowSmem<real_t, N, 2>::BYTES : FLOW_SMEM;
  int rc = 0;
  if (!occ && (rc = prep(k, smem, NT, &occ))) return rc;
#ifdef TCFD_EMU
  (void)win;
  const int grid = 2;
  TCFD_LAUNCH(k, grid, NT, smem, stream, fp, *maps);
  return 0;
#else
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(num_sms * occ));
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (win && win->bytes) {
    attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[0].val.accessPolicyWindow.base_ptr = win->base;
    attr[0].val.accessPolicyWindow.num_bytes = win->bytes;
    attr[0].val.accessPolicyWindow.hitRatio = win->hit_ratio;
    attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return (int)cudaLaunchKernelEx(&cfg, k, fp, *maps);
#endif
}

// group sizes (double rows per rows item, column quads per cols item) and CTAs per SM of the register
// budget (0: FLOW_MINB); TCFD_FLOW_G="<GR>,<GC>,<MINB>" selects one of the other compiled variants
// (schedule experiments, -DTCFD_FLOW_VARIANTS builds only)
int launch_flow(const FlowParams<real_t>& fp, const TileMaps* maps, int num_sms, cudaStream_t stream,
                const tcfd_flow_window_t* win) {
  if (!maps) return -2;
  // Items: grouped (3 double rows / 4 column quads per ticket, register budget for 256 resident threads
  // per SM -> no spills) when the window is wide, single units when it is narrow (a narrow window needs
  // every item it can get to keep the SMs busy).  The caller chooses (win->grouped); TCFD_FLOW_G=
  // "<GR>,<GC>,<MINB>" selects any compiled variant (-DTCFD_FLOW_VARIANTS builds: schedule experiments).
  constexpr int MR_RAW = (sizeof(real_t) == 4 ? 256 : 128) / NT;
  constexpr int MR = MR_RAW < 1 ? 1 : MR_RAW;
  int gr = (win && win->grouped) ? 3 : 1, gc = (win && win->grouped) ? 4 : 1, mr = MR;
  int egr = 0, egc = 0, emr = 0;  // read at every call: the variable is an experiment / test knob
  if (const char* e = getenv("TCFD_FLOW_G")) {
    int a = 0, b = 0, c = 0;
    const int nf = sscanf(e, "%d,%d,%d", &a, &b, &c);
    if (nf >= 2) { egr = a; egc = b; emr = nf == 3 ? c : 0; }
  }
  if (egr > 0) { gr = egr; gc = egc; mr = emr; }
#define TCFD_FLOW_CASE(A, B, C) if (gr == A && gc == B && mr == C) return launch_flow_g<A, B, C>(fp, maps, num_sms, stream, win);
  TCFD_FLOW_CASE(3, 4, MR)
  TCFD_FLOW_CASE(1, 1, MR)
#ifdef TCFD_FLOW_VARIANTS
  TCFD_FLOW_CASE(1, 1, 0)
  TCFD_FLOW_CASE(3, 4, 0)
  TCFD_FLOW_CASE(2, 2, MR)
  TCFD_FLOW_CASE(4, 4, MR)
  TCFD_FLOW_CASE(3, 8, MR)
  TCFD_FLOW_CASE(2, 4, MR)
  TCFD_FLOW_CASE(5, 4, MR)
  TCFD_FLOW_CASE(1, 1, 5)
  TCFD_FLOW_CASE(3, 4, 5)
  TCFD_FLOW_CASE(2, 2, 5)
  // rolled column loop of the cols items
  if (gr == 3 && gc == 4 && mr == 8) return launch_flow_g<3, 4, MR, 8>(fp, maps, num_sms, stream, win);
  // ping-pong exchange buffers
  if (gr == 3 && gc == 4 && mr == 2) return launch_flow_g<3, 4, MR, 2>(fp, maps, num_sms, stream, win);
  if (gr == 1 && gc == 1 && mr == 2) return launch_flow_g<1, 1, MR, 2>(fp, maps, num_sms, stream, win);
  if (gr == 3 && gc == 4 && mr == -4) return launch_flow_g<3, 4, MR, 1>(fp, maps, num_sms, stream, win);  // no-FFT timing
  if (gr == 1 && gc == 1 && mr == -4) return launch_flow_g<1, 1, MR, 1>(fp, maps, num_sms, stream, win);
  // cycle attribution per region (fp.prof)
  if (gr == 3 && gc == 4 && mr == -64) return launch_flow_g<3, 4, MR, 64>(fp, maps, num_sms, stream, win);
  if (gr == 1 && gc == 1 && mr == -64) return launch_flow_g<1, 1, MR, 64>(fp, maps, num_sms, stream, win);
  // timing experiments on top of the no-FFT twin: -5 no fields, -9 no tile unpack, -17 no update math, -29 all
  if (gr == 3 && gc == 4 && mr == -5) return launch_flow_g<3, 4, MR, 5>(fp, maps, num_sms, stream, win);
  if (gr == 3 && gc == 4 && mr == -9) return launch_flow_g<3, 4, MR, 9>(fp, maps, num_sms, stream, win);
  if (gr == 3 && gc == 4 && mr == -17) return launch_flow_g<3, 4, MR, 17>(fp, maps, num_sms, stream, win);
  if (gr == 3 && gc == 4 && mr == -29) return launch_flow_g<3, 4, MR, 29>(fp, maps, num_sms, stream, win);
#endif
#undef TCFD_FLOW_CASE
  return -3;
}
}  // namespace v2
#endif

int launch(int which, const void* params, const void* maps, int num_sms, void* stream_) {
  const NsParams<real_t>& p = *static_cast<const NsParams<real_t>*>(params);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
#if TCFD_N >= 256
  rc = v2::launch(which, p, static_cast<const TileMaps*>(maps), num_sms, stream);
#else
  (void)maps;
  rc = v1::launch(which, p, num_sms, stream);
#endif
  if (rc) return rc;
#ifndef TCFD_EMU
  return (int)cudaGetLastError();
#else
  return 0;
#endif
}

#if TCFD_N <= 64
// groups per CTA of the resident kernel: as many as 1024 threads / the 227 KB of shared memory allow
constexpr int SMALL_G_THREADS = 512 / NT;
constexpr int small_groups() {
  int g = SMALL_G_THREADS;
  while (g > 1 && SmallSmem<real_t, N, 1>::OFF_BUF + (size_t)g * N * sizeof(cx<real_t>) > 232448 - 1024) --g;
  return g;
}
int launch_small(const void* params, int batch, void* stream_) {
  constexpr int G = small_groups();
  static int ready[MAX_DEV];
  auto k = ns2d_small_kernel<real_t, N, G>;
  constexpr size_t smem = SmallSmem<real_t, N, G>::BYTES;
#ifndef TCFD_EMU
  int& r = ready[cur_dev()];
  if (!r) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    r = 1;
  }
#else
  (void)ready;
#endif
  TCFD_LAUNCH(k, batch, G * NT, smem, static_cast<cudaStream_t>(stream_), *static_cast<const FlowParams<real_t>*>(params));
#ifndef TCFD_EMU
  return (int)cudaGetLastError();
#else
  return 0;
#endif
}
#endif

#if TCFD_N >= 256
int launch_flow(const void* params, const void* maps, int num_sms, void* stream_, const tcfd_flow_window_t* win) {
  int rc = v2::launch_flow(*static_cast<const FlowParams<real_t>*>(params), static_cast<const TileMaps*>(maps), num_sms,
                           static_cast<cudaStream_t>(stream_), win);
  if (rc) return rc;
#ifndef TCFD_EMU
  return (int)cudaGetLastError();
#else
  return 0;
#endif
}
#endif
}  // namespace

#define TCFD_CAT3(a, b, c) a##b##_##c
#define TCFD_ENTRY(prec, n) TCFD_CAT3(tcfd_ns2d_entry_, prec, n)
extern "C" void TCFD_ENTRY(TCFD_PREC, TCFD_N)(tcfd_ns2d_entry_t* e) {
  e->n = N;
  e->prec = TCFD_PREC;
#if TCFD_N < 256
  e->yt = v1::YT;
#else
  e->yt = 4;
#endif
  e->v2 = V2 ? 1 : 0;
  e->launch = &launch;
#if TCFD_N >= 256
  e->flow_ctas_per_sm = (int)(232448 / ((size_t)FlowSmem<real_t, N>::BYTES + 1024));
  e->launch_flow = e->flow_ctas_per_sm > 0 ? &launch_flow : nullptr;
#else
  e->flow_ctas_per_sm = 0;
  e->launch_flow = nullptr;
#endif
#if TCFD_N <= 64
  e->launch_small = &launch_small;
#else
  e->launch_small = nullptr;
#endif
}
