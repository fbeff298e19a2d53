// One translation unit per (precision, N): compiled with -DTCFD_PREC=32|64 -DTCFD_N=<n>.
// Exports a C launcher table entry used by ns2d_api.cu.
#include "ns2d_kernels.cuh"
#include "ns2d_plan.h"

#if TCFD_PREC == 32
typedef float real_t;
#else
typedef double real_t;
#endif

namespace {
using namespace tcfd;
constexpr int N = TCFD_N;
constexpr int NT = N / 8;
constexpr int CTA = 256;
constexpr int G = (CTA / NT) > 0 ? (CTA / NT) : 1;
constexpr int GC0 = G;
constexpr int GC = GC0 < N / 2 ? GC0 : N / 2;
constexpr int YT = 2 * GC;
constexpr bool PP = true;
typedef cx<real_t> cplx;

constexpr size_t smem_rows(bool inv) { return (size_t)G * ((inv ? 2 : 1) * N) * (PP ? 2 : 1) * sizeof(cplx); }
constexpr size_t smem_cols() {
  return ((size_t)(N / 2 + 1) * (4 * YT + 1) + (size_t)GC * N * (PP ? 2 : 1)) * sizeof(cplx);
}

template <class K>
int prep(K kernel, size_t smem, int threads, int* occ) {
#ifndef TCFD_EMU
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kernel, threads, smem);
  if (e != cudaSuccess) return (int)e;
#else
  *occ = 1;
#endif
  return 0;
}

int launch(int which, const void* params, int num_sms, void* stream_) {
  const NsParams<real_t>& p = *static_cast<const NsParams<real_t>*>(params);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  static int occ[4] = {0, 0, 0, 0};
  const int nblk_rows = (p.B * (N / 2 + 1) + G - 1) / G;
  const int ntiles = p.B * (N / YT);
  int rc = 0;
  auto grid_for = [&](int work, int o) {
#ifdef TCFD_EMU
    (void)o;
    return work;
#else
    int cap = num_sms * (o > 0 ? o : 1);
    return work < cap ? work : cap;
#endif
  };
  switch (which) {
    case TCFD_K_ROWS_INV: {
      auto k = ns2d_rows_kernel<real_t, N, G, YT, false, true, PP>;
      if (!occ[0] && (rc = prep(k, smem_rows(true), G * NT, &occ[0]))) return rc;
      TCFD_LAUNCH(k, grid_for(nblk_rows, occ[0]), G * NT, smem_rows(true), stream, p);
      break;
    }
    case TCFD_K_ROWS_FULL: {
      auto k = ns2d_rows_kernel<real_t, N, G, YT, true, true, PP>;
      if (!occ[1] && (rc = prep(k, smem_rows(true), G * NT, &occ[1]))) return rc;
      TCFD_LAUNCH(k, grid_for(nblk_rows, occ[1]), G * NT, smem_rows(true), stream, p);
      break;
    }
    case TCFD_K_ROWS_FWD: {
      auto k = ns2d_rows_kernel<real_t, N, G, YT, true, false, PP>;
      if (!occ[2] && (rc = prep(k, smem_rows(false), G * NT, &occ[2]))) return rc;
      TCFD_LAUNCH(k, grid_for(nblk_rows, occ[2]), G * NT, smem_rows(false), stream, p);
      break;
    }
    case TCFD_K_COLS: {
      auto k = ns2d_cols_kernel<real_t, N, GC, PP>;
      if (!occ[3] && (rc = prep(k, smem_cols(), GC * NT, &occ[3]))) return rc;
      TCFD_LAUNCH(k, grid_for(ntiles, occ[3]), GC * NT, smem_cols(), stream, p);
      break;
    }
    default:
      return -1;
  }
#ifndef TCFD_EMU
  return (int)cudaGetLastError();
#else
  return 0;
#endif
}
}  // namespace

#define TCFD_CAT3(a, b, c) a##b##_##c
#define TCFD_ENTRY(prec, n) TCFD_CAT3(tcfd_ns2d_entry_, prec, n)
extern "C" void TCFD_ENTRY(TCFD_PREC, TCFD_N)(tcfd_ns2d_entry_t* e) {
  e->n = N;
  e->prec = TCFD_PREC;
  e->yt = YT;
  e->launch = &launch;
}
