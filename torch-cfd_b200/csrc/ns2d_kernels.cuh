// Kernels of hot path A: one Carpenter-Kennedy substage of the pseudo-spectral 2-D vorticity
// Navier-Stokes step (reference: torch_cfd/equations.py:328-358 stepper, :413-447 equation,
// torch_cfd/spectral.py:87-115 velocity from vorticity).
//
// A substage is two launches that never materialise a physical-space field in HBM:
//
//   rows kernel  (R): one work item = the pair of spectrum rows (kx, N-kx) of one sample.
//        [FWD]  y-axis forward FFT of the x-transformed advection row  ->  adv_hat rows kx, N-kx
//               -> 2/3 mask, + f_hat, h = F + beta h, CN solve  ->  new w_hat rows (HBM, float2)
//        [INV]  from the NEW rows (still in registers): the four spectra u, v, dw/dx, dw/dy,
//               Hermitian-completed along ky, y-axis inverse FFT  ->  H[4](kx, y)  (L2-resident)
//   cols kernel  (C): one work item = two physical columns (y, y+1) of one sample.
//        x-axis C2R of (u,v) and (dw/dx,dw/dy) packed as two complex FFTs per column,
//        adv = -(dw/dx u + dw/dy v) in registers, the two columns packed into ONE complex forward
//        x-FFT, separated, and only the rows |kx| that survive the 2/3 mask are written (advt).
//
// Layouts (all complex interleaved, T = float or double):
//   state w, h        [B][N][NH]              NH = N/2+1, reference layout (rfft2 half spectrum)
//   H                 [B][NH][N/YT][4][YT]    kx in 0..N/2; YT = columns per C tile
//   advt              [B][NH][N]              rows 0..KF-1 used
#pragma once
#include "fft_core.cuh"

namespace tcfd {

enum : int { UPD_RK = 0, UPD_F = 1, UPD_RESID = 2 };

// batch-shared tables of one spectrum entry, interleaved (second-generation kernels)
template <class T>
struct alignas(4 * sizeof(T)) tab4 {
  T lin, filt, nil, pad;  // linear_term, 2/3 mask (1 when smooth=False), -1/laplace'
};

template <class T>
struct NsParams {
  int B;
  int KF;          // rows kx in [0, KF) of advt carry modes that survive the mask
  int mode;        // UPD_*
  int read_h, write_h;
  const cx<T>* w_in;
  cx<T>* w_out;
  const cx<T>* h_in;
  cx<T>* h_out;     // RK accumulator; or F / residual output in UPD_F / UPD_RESID
  const cx<T>* w_old;  // dwdt reference state (UPD_RK with dwdt) or w_t (UPD_RESID)
  cx<T>* dwdt;
  cx<T>* H;
  cx<T>* advt;
  // second-generation kernels (ns2d_v2.cuh): packed layouts and the number of advt double rows
  T* H2;
  T* advt2;
  int NDF;
  const tab4<T>* tab;         // [N][NH]
  const unsigned char* frow;  // [N] row carries a non-zero forcing entry
  // unit-layout state between substages and per-d table blocks (ns2d_rows3_kernel)
  const cx<typename pack2<T>::type>* wU_in;
  const cx<typename pack2<T>::type>* hU_in;
  cx<typename pack2<T>::type>* wU_out;
  cx<typename pack2<T>::type>* hU_out;
  const void* tabU;            // [ND]{lin[NH][2], nil_a[NH], nil_b[NH]}
  const unsigned char* maskU;  // [ND][2][MASK_ROW]
  int in_user, out_user;       // substage reads / writes the reference layout
  int dbg;                     // timing experiments (only with -DTCFD_DEBUG_KNOBS)
  const cx<T>* tw;
  const T* kappa_x;  // [N]   2 pi kx / N^2
  const T* kappa_y;  // [NH]  2 pi ky / N^2
  const T* nil;      // [N][NH]  -1/laplace'
  const T* lin;      // [N][NH]  linear_term
  const T* filt;     // [N][NH] or nullptr
  const cx<T>* fhat; // [N][NH] or nullptr
  T beta, gdt, mu, inv_tdt;
};

// trajectory recording (fno/data_gen/solvers.py:245-256): cast w, psi = nil * w, dw/dt and the residual
// of one recorded step into slot `it` of the (B, n_t, n, nh) snapshot buffers
template <class T, class O>
__global__ void ns2d_record_kernel(const cx<T>* __restrict__ w, const cx<T>* __restrict__ dwdt,
                                   const cx<T>* __restrict__ res, const T* __restrict__ nil, cx<O>* sw, cx<O>* spsi,
                                   cx<O>* sdw, cx<O>* sres, size_t per, int n_t, int it, size_t total) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t b = idx / per, e = idx % per;
    const size_t o = (b * n_t + it) * per + e;
    const cx<T> wv = w[idx];
    if (sw) sw[o] = cx<O>{(O)wv.x, (O)wv.y};
    if (spsi) {
      const T f = nil[e];
      spsi[o] = cx<O>{(O)(f * wv.x), (O)(f * wv.y)};
    }
    if (sdw) sdw[o] = cx<O>{(O)dwdt[idx].x, (O)dwdt[idx].y};
    if (sres) sres[o] = cx<O>{(O)res[idx].x, (O)res[idx].y};
  }
}

struct CtaSync {
  TCFD_D void operator()() const { __syncthreads(); }
};

// ------------------------------------------------------------------------------------------
// state update of one spectrum entry; returns the new w_hat (or w_in when not stepping)
template <class T>
TCFD_D cx<T> ns_update(const NsParams<T>& p, size_t gidx, int tidx, cx<T> A, bool store) {
  cx<T> F = A;
  if (p.filt) {
    const T f = p.filt[tidx];
    F = cx<T>{f * A.x, f * A.y};
  }
  if (p.fhat) F = F + p.fhat[tidx];
  if (p.mode == UPD_F) {
    if (store) p.h_out[gidx] = F;
    return F;
  }
  const cx<T> w = p.w_in[gidx];
  const T L = p.lin[tidx];
  if (p.mode == UPD_RESID) {
    const cx<T> wt = p.w_old[gidx];
    const cx<T> r = (wt - F) - L * w;
    if (store) p.h_out[gidx] = r;
    return r;
  }
  cx<T> h = F;
  if (p.read_h) h = F + p.beta * p.h_in[gidx];
  if (p.write_h && store) p.h_out[gidx] = h;
  const T inv = T(1) / (T(1) - p.mu * L);
  const cx<T> x = (w + p.gdt * h) + p.mu * (L * w);
  const cx<T> wn = inv * x;
  if (store) {
    p.w_out[gidx] = wn;
    if (p.dwdt) p.dwdt[gidx] = p.inv_tdt * (wn - p.w_old[gidx]);
  }
  return wn;
}

// the four transformed spectra at one entry: u = i ky psi, v = -i kx psi, gx = i kx w, gy = i ky w
// with psi = -w/lap' (nil = -1/lap'), kappa = 2 pi k / N^2.  PAIR 0 -> (u, v), PAIR 1 -> (gx, gy)
template <int PAIR, class T>
TCFD_D void ns_fields(cx<T> w, T nil, T kx, T ky, cx<T>& a, cx<T>& b) {
  if (PAIR == 0) {
    const cx<T> q{nil * w.x, nil * w.y};
    a = cx<T>{-(ky * q.y), ky * q.x};
    b = cx<T>{kx * q.y, -(kx * q.x)};
  } else {
    a = cx<T>{-(kx * w.y), kx * w.x};
    b = cx<T>{-(ky * w.y), ky * w.x};
  }
}

// ------------------------------------------------------------------------------------------
// rows kernel.  G groups of NT = N/8 threads per CTA; PP = ping-pong exchange buffers.
template <class T, int N, int G, int YT, bool FWD, bool INV, bool PP>
__global__ void __launch_bounds__(G * (N / 8))
ns2d_rows_kernel(const NsParams<T> p) {
  constexpr int NT = N / 8, NH = N / 2 + 1;
  constexpr int HALF = (INV ? 2 : 1) * N;
  constexpr int BUF = HALF * (PP ? 2 : 1);  // elements per group
  TCFD_DYN_SMEM(smem_raw);
  cx<T>* smem = reinterpret_cast<cx<T>*>(smem_raw);
  const int g = threadIdx.x / NT, t = threadIdx.x % NT;
  cx<T>* buf = smem + g * BUF;
  FftTwiddles<T, N> tw;
  tw.load(p.tw, t);
  CtaSync sync;
  int parity = 0;
  const int nitems = p.B * NH;

  for (int blk = blockIdx.x; blk * G < nitems; blk += gridDim.x) {
    const int item = blk * G + g;
    const bool valid = item < nitems;
    const int it = valid ? item : nitems - 1;
    const int s = it / NH, pr = it % NH;
    const int r1 = pr, r2 = (N - pr) % N;
    const bool self = (r1 == r2);
    cx<T> wv[8], e0, e1;  // e0 = entry (r2, 0), e1 = entry (r1, N/2): owned by thread 0

    if constexpr (FWD) {
      // CTA-uniform: does any item of this block carry unmasked rows?
      const int pr_first = (blk * G) % NH;
      const bool do_fft = (pr_first < p.KF) || (pr_first + G > NH);
      cx<T> a[1][8];
#pragma unroll
      for (int m = 0; m < 8; ++m) a[0][m] = cx<T>{T(0), T(0)};
      if (do_fft) {
        if (pr < p.KF) {
          const cx<T>* src = p.advt + ((size_t)s * NH + pr) * N;
#pragma unroll
          for (int m = 0; m < 8; ++m) a[0][m] = src[t + m * NT];
        }
        fft_run<T, N, -1, 1, PP, HALF>(a, tw, buf, parity, t, sync);
      }
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int ky = t + m * NT;
        const bool lo = m < 4;
        const int row = lo ? r1 : r2, col = lo ? ky : N - ky;
        const cx<T> A = lo ? a[0][m] : conj(a[0][m]);
        const bool own = valid && (lo || !self || (m == 4 && t == 0));
        wv[m] = ns_update(p, ((size_t)s * N + row) * NH + col, row * NH + col, A, own);
      }
      if (t == 0) {
        e0 = ns_update(p, ((size_t)s * N + r2) * NH + 0, r2 * NH + 0, conj(a[0][0]), valid && !self);
        e1 = ns_update(p, ((size_t)s * N + r1) * NH + N / 2, r1 * NH + N / 2, a[0][4], valid && !self);
      }
    } else {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int ky = t + m * NT;
        const bool lo = m < 4;
        const int row = lo ? r1 : r2, col = lo ? ky : N - ky;
        wv[m] = p.w_in[((size_t)s * N + row) * NH + col];
      }
      if (t == 0) {
        e0 = p.w_in[((size_t)s * N + r2) * NH + 0];
        e1 = p.w_in[((size_t)s * N + r1) * NH + N / 2];
      }
    }

    if constexpr (INV) {
      const T kx1 = p.kappa_x[r1], kx2 = p.kappa_x[r2];
      cx<T>* Hrow = p.H + ((size_t)s * NH + pr) * (size_t)(N / YT) * (4 * YT);
#pragma unroll
      for (int pair = 0; pair < 2; ++pair) {
        cx<T> z[2][8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int ky = t + m * NT;
          const bool lo = m < 4;
          const int row = lo ? r1 : r2, col = lo ? ky : N - ky;
          const T nil = (pair == 0) ? p.nil[row * NH + col] : T(0);
          const T kyv = p.kappa_y[col];
          cx<T> a, b;
          if (pair == 0) ns_fields<0>(wv[m], nil, lo ? kx1 : kx2, kyv, a, b);
          else ns_fields<1>(wv[m], nil, lo ? kx1 : kx2, kyv, a, b);
          z[0][m] = lo ? a : conj(a);
          z[1][m] = lo ? b : conj(b);
        }
        if (t == 0) {
          // self-conjugate columns ky = 0 and ky = N/2: average the two rows (C2R semantics)
          cx<T> a1, b1, a2, b2;
          const T ky0 = p.kappa_y[0], kyh = p.kappa_y[N / 2];
          const T n10 = pair == 0 ? p.nil[r1 * NH] : T(0), n20 = pair == 0 ? p.nil[r2 * NH] : T(0);
          const T n1h = pair == 0 ? p.nil[r1 * NH + N / 2] : T(0);
          const T n2h = pair == 0 ? p.nil[r2 * NH + N / 2] : T(0);
          if (pair == 0) { ns_fields<0>(wv[0], n10, kx1, ky0, a1, b1); ns_fields<0>(e0, n20, kx2, ky0, a2, b2); }
          else { ns_fields<1>(wv[0], n10, kx1, ky0, a1, b1); ns_fields<1>(e0, n20, kx2, ky0, a2, b2); }
          z[0][0] = T(0.5) * (a1 + conj(a2));
          z[1][0] = T(0.5) * (b1 + conj(b2));
          if (pair == 0) { ns_fields<0>(e1, n1h, kx1, kyh, a1, b1); ns_fields<0>(wv[4], n2h, kx2, kyh, a2, b2); }
          else { ns_fields<1>(e1, n1h, kx1, kyh, a1, b1); ns_fields<1>(wv[4], n2h, kx2, kyh, a2, b2); }
          z[0][4] = T(0.5) * (a1 + conj(a2));
          z[1][4] = T(0.5) * (b1 + conj(b2));
        }
        fft_run<T, N, +1, 2, PP, HALF>(z, tw, buf, parity, t, sync);
        if (valid) {
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            const int y = t + m * NT;
            cx<T>* dst = Hrow + (size_t)(y / YT) * (4 * YT) + (y % YT);
            dst[(2 * pair) * YT] = z[0][m];
            dst[(2 * pair + 1) * YT] = z[1][m];
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// cols kernel.  GC groups per CTA, tile of YT = 2*GC columns.
template <class T, int N, int GC, bool PP>
__global__ void __launch_bounds__(GC * (N / 8))
ns2d_cols_kernel(const NsParams<T> p) {
  constexpr int NT = N / 8, NH = N / 2 + 1, YT = 2 * GC;
  constexpr int RS = 4 * YT + 1;  // padded tile row (odd -> conflict-free strided reads)
  constexpr int OS = YT + 1;      // padded out-tile row
  constexpr int HALF = N;
  constexpr int BUF = HALF * (PP ? 2 : 1);
  constexpr int NTHREADS = GC * NT;
  TCFD_DYN_SMEM(smem_raw);
  cx<T>* tile = reinterpret_cast<cx<T>*>(smem_raw);  // [NH][RS], reused as out tile [KF][OS]
  cx<T>* bufs = tile + NH * RS;
  const int g = threadIdx.x / NT, t = threadIdx.x % NT;
  cx<T>* buf = bufs + g * BUF;
  FftTwiddles<T, N> tw;
  tw.load(p.tw, t);
  CtaSync sync;
  int parity = 0;
  const int ntiles = p.B * (N / YT);

  for (int tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
    const int s = tl / (N / YT), yt = tl % (N / YT);
    __syncthreads();
    {  // tile fill: row kx <- 4*YT contiguous complex values
      const cx<T>* src = p.H + ((size_t)s * NH * (N / YT) + yt) * (4 * YT);
      for (int i = threadIdx.x; i < NH * 4 * YT; i += NTHREADS) {
        const int k = i / (4 * YT), j = i % (4 * YT);
        tile[k * RS + j] = src[(size_t)k * (N / YT) * (4 * YT) + j];
      }
    }
    __syncthreads();

    cx<T> c[1][8];
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int col = 2 * g + cc;
      cx<T> uv[8];
#pragma unroll
      for (int pair = 0; pair < 2; ++pair) {
        cx<T> z[1][8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int k = t + m * NT;
          const bool lo = (m < 4) || (m == 4 && t == 0);  // k <= N/2
          const int row = lo ? k : N - k;
          const cx<T> a = tile[row * RS + (2 * pair) * YT + col];
          const cx<T> b = tile[row * RS + (2 * pair + 1) * YT + col];
          z[0][m] = lo ? cx<T>{a.x - b.y, a.y + b.x} : cx<T>{a.x + b.y, b.x - a.y};
          if ((m == 0 || m == 4) && t == 0) z[0][m] = cx<T>{a.x, b.x};  // kx = 0, N/2: real rows
        }
        fft_run<T, N, +1, 1, PP, HALF>(z, tw, buf, parity, t, sync);
        if (pair == 0) {
#pragma unroll
          for (int m = 0; m < 8; ++m) uv[m] = z[0][m];
        } else {
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            const T adv = -(z[0][m].x * uv[m].x + z[0][m].y * uv[m].y);
            if (cc == 0) c[0][m].x = adv; else c[0][m].y = adv;
          }
        }
      }
    }
    fft_run<T, N, -1, 1, PP, HALF>(c, tw, buf, parity, t, sync);
    // separate the two real columns: needs C(N-k) from another thread
    {
      cx<T>* xb = buf + (PP ? parity * HALF : 0);
      if (PP) parity ^= 1;
#pragma unroll
      for (int m = 0; m < 8; ++m) xb[t + m * NT] = c[0][m];
      __syncthreads();  // also: every group is done reading the input tile
      cx<T>* outt = tile;
#pragma unroll
      for (int m = 0; m < 5; ++m) {
        const int k = t + m * NT;
        if (k < p.KF && (m < 4 || t == 0)) {
          const cx<T> ck = c[0][m];
          const cx<T> cn = xb[(N - k) % N];
          outt[k * OS + 2 * g] = cx<T>{T(0.5) * (ck.x + cn.x), T(0.5) * (ck.y - cn.y)};
          outt[k * OS + 2 * g + 1] = cx<T>{T(0.5) * (ck.y + cn.y), T(0.5) * (cn.x - ck.x)};
        }
      }
      __syncthreads();
      cx<T>* dst = p.advt + (size_t)s * NH * N + yt * YT;
      for (int i = threadIdx.x; i < p.KF * YT; i += NTHREADS) {
        const int k = i / YT, j = i % YT;
        dst[(size_t)k * N + j] = outt[k * OS + j];
      }
    }
  }
}

}  // namespace tcfd
