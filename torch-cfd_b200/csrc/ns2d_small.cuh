// Whole-call resident kernel of hot path A for small grids (N <= 64): ONE launch runs every sub-stage of every step of
// a call with the sample's state and all intermediate fields in the shared memory of one CTA (SURVEY 2.1 "K3", 7 step 5).
//
// The generic kernels need two grid-wide launches per sub-stage; at 64 x 64 a launch does a few microseconds of work
// behind ~10 us of launch / dependent-global-load latency (BASELINE configs[0], 64^2 x 1 x fp64 x 100 steps: 1001
// launches, 11.7 ms).  Here a CTA owns one sample:
//
//   shared memory   wS, hS   [N][NH] complex     state and RK accumulator, reference layout
//                   Z1, Z2   [N][N]  complex     z1 = u + i v, z2 = dw/dx + i dw/dy after the y-inverse (all N rows kx:
//                                                rows above N/2 are written as conjugates, cf. ns2d_v2.cuh), then re-used:
//                                                Z1[x][y].x <- advection(x, y),  Z2 <- advt[kx <= N/2][y] (x-transformed)
//                   buf      [G][N]  complex     exchange buffer of each FFT group (N/8 threads)
//   per sub-stage   R2  y-inverse of the P / Q combinations of every row pair (scalar lanes: 4 transforms per pair)
//                   C1  per physical column: the two packed x-inverses, advection product
//                   C2  per column pair: one forward x-transform of two real columns, separated -> advt
//                   R1  per row pair: forward y-transform, 2/3 mask, forcing, RK / CN update of wS, hS
//                   with a CTA barrier between the phases; the groups of a warp run in lock step, so the exchanges
//                   inside a transform only need __syncwarp().
// HBM sees the state once per CALL (read w_in, write w_out [+ dw/dt]) -- the "2 S" bound of SURVEY 8d -- and the
// batch-shared tables through L1/L2.  Arithmetic and operation order: ns2d_kernels.cuh / ns2d_v2.cuh (same update
// formula, same field construction), so results agree with the other schedules to rounding.
#pragma once
#include "ns2d_flow.cuh"

namespace tcfd {

struct WarpLockstepSync {
  TCFD_D void operator()() const {
#ifndef TCFD_EMU
    __syncwarp();
#else
    __syncthreads();  // emulation: every thread of the CTA follows the same control flow
#endif
  }
};

template <class T, int N, int G>
struct SmallSmem {
  static constexpr int NH = N / 2 + 1;
  static constexpr int ZS = N + 1;  // padded row of Z1 / Z2: the column walks of C1 / C2 hit distinct banks
  static constexpr size_t OFF_W = 0;
  static constexpr size_t OFF_H = OFF_W + (size_t)N * NH * sizeof(cx<T>);
  static constexpr size_t OFF_Z1 = OFF_H + (size_t)N * NH * sizeof(cx<T>);
  static constexpr size_t OFF_Z2 = OFF_Z1 + (size_t)N * ZS * sizeof(cx<T>);
  static constexpr size_t OFF_BUF = OFF_Z2 + (size_t)N * ZS * sizeof(cx<T>);
  static constexpr size_t BYTES = OFF_BUF + (size_t)G * N * sizeof(cx<T>);
};

// one lane of ns_fields_z: LANE 0 -> (sg kx + i ky) psi, LANE 1 -> (sg kx + i ky) (sg i w)
template <int LANE, class T>
TCFD_D cx<T> small_field(cx<T> w, T nil, T kx, T ky, T sg) {
  const cx<T> q = LANE == 0 ? cx<T>{nil * w.x, nil * w.y} : cx<T>{-(sg * w.y), sg * w.x};
  const T a = sg * kx, nky = -ky;
  return cx<T>{fma_rn(q.x, a, q.y * nky), fma_rn(q.y, a, q.x * ky)};
}

template <class T, int N, int G>
__global__ void __launch_bounds__(G * (N / 8))
ns2d_small_kernel(const FlowParams<T> fp) {
  typedef SmallSmem<T, N, G> S;
  constexpr int NT = N / 8, NH = N / 2 + 1, ZS = S::ZS;
  const NsParams<T>& p = fp.p;
  TCFD_DYN_SMEM(smem_raw);
  cx<T>* wS = reinterpret_cast<cx<T>*>(smem_raw + S::OFF_W);
  cx<T>* hS = reinterpret_cast<cx<T>*>(smem_raw + S::OFF_H);
  cx<T>* Z1 = reinterpret_cast<cx<T>*>(smem_raw + S::OFF_Z1);
  cx<T>* Z2 = reinterpret_cast<cx<T>*>(smem_raw + S::OFF_Z2);
  cx<T>* AT = Z2;  // advt[pr][y], pr <= N/2 (after C1 the z2 field is dead)
  const int g = threadIdx.x / NT, t = threadIdx.x % NT;
  cx<T>* buf = reinterpret_cast<cx<T>*>(smem_raw + S::OFF_BUF) + (size_t)g * N;
  FftTwiddles<T, N> tw;
  tw.load(p.tw, t);
  WarpLockstepSync sync;
  int parity = 0;
  const int s = blockIdx.x;
  const size_t sb = (size_t)s * N * NH;
  const bool want_dwdt = p.dwdt != nullptr;
  const T ky0 = p.kappa_y[0], kyh = p.kappa_y[N / 2];

  for (int i = threadIdx.x; i < N * NH; i += G * NT) wS[i] = p.w_in[sb + i];
  __syncthreads();

  for (int j = 0; j < fp.nsub; ++j) {
    const int k = j % fp.nstages;
    const bool last_sub = j == fp.nsub - 1;
    const bool rd_h = fp.rd_h[k] != 0, wr_h = fp.wr_h[k] != 0;
    const T beta = fp.beta[k], gdt = fp.gdt[k], mu = fp.mu[k];

    // ---------------------------------------------------------------- R2: y-inverse of P / Q, both lanes
    {
      constexpr int NTASK = NH * 4;  // (row pair, type, lane)
      for (int it = 0; it < (NTASK + G - 1) / G; ++it) {
        const int task = it * G + g;
        const bool in = task < NTASK;
        const int tk = in ? task : NTASK - 1;
        const int pr = tk >> 2, type = (tk >> 1) & 1, lane = tk & 1;
        const int r1 = pr, r2 = (N - pr) % N;
        const bool valid = in && !(type && r1 == r2);  // self-paired rows need P only
        const T kx1 = p.kappa_x[r1], kx2 = p.kappa_x[r2];
        const T sg = type ? T(-1) : T(1);
        auto fld = [&](cx<T> w, T nil, T kx, T ky, T sgn) {
          return lane ? small_field<1, T>(w, nil, kx, ky, sgn) : small_field<0, T>(w, nil, kx, ky, sgn);
        };
        cx<T> z[1][8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int ky = t + m * NT;
          const bool lo = m < 4;
          const int row = lo ? r1 : r2, col = lo ? ky : N - ky;
          const cx<T> f = fld(wS[row * NH + col], p.tab[row * NH + col].nil, lo ? kx1 : kx2, p.kappa_y[col], lo ? sg : -sg);
          z[0][m] = lo ? f : conj(f);
        }
        if (t == 0) {
          // self-conjugate columns ky = 0 and ky = N/2: Hermitian part of the two rows (C2R semantics)
          const cx<T> f1 = fld(wS[r1 * NH], p.tab[r1 * NH].nil, kx1, ky0, sg);
          const cx<T> f2 = fld(wS[r2 * NH], p.tab[r2 * NH].nil, kx2, ky0, -sg);
          const cx<T> g1 = fld(wS[r1 * NH + N / 2], p.tab[r1 * NH + N / 2].nil, kx1, kyh, sg);
          const cx<T> g2 = fld(wS[r2 * NH + N / 2], p.tab[r2 * NH + N / 2].nil, kx2, kyh, -sg);
          z[0][0] = T(0.5) * (f1 + conj(f2));
          z[0][4] = T(0.5) * (g1 + conj(g2));
        }
        fft_run<T, N, +1, 1, false, N>(z, tw, buf, parity, t, sync);
        if (valid) {
          cx<T>* dst = (lane ? Z2 : Z1) + (size_t)(type ? r2 : r1) * ZS;
#pragma unroll
          for (int m = 0; m < 8; ++m) dst[t + m * NT] = type ? conj(z[0][m]) : z[0][m];
        }
      }
    }
    __syncthreads();

    // ---------------------------------------------------------------- C1: x-inverses of column y, advection product
    for (int it = 0; it < (N + G - 1) / G; ++it) {
      const int col = it * G + g;
      const bool valid = col < N;
      const int y = valid ? col : N - 1;
      cx<T> a[1][8], b[1][8];
#pragma unroll
      // (groups past the last column run the transforms for the group barriers only: they must not read a column another
      // group is about to overwrite below)
      for (int m = 0; m < 8; ++m) a[0][m] = valid ? Z1[(size_t)(t + m * NT) * ZS + y] : cx<T>{T(0), T(0)};
      fft_run<T, N, +1, 1, false, N>(a, tw, buf, parity, t, sync);
#pragma unroll
      for (int m = 0; m < 8; ++m) b[0][m] = valid ? Z2[(size_t)(t + m * NT) * ZS + y] : cx<T>{T(0), T(0)};
      fft_run<T, N, +1, 1, false, N>(b, tw, buf, parity, t, sync);
      if (valid) {
#pragma unroll
        for (int m = 0; m < 8; ++m)  // a = u + i v, b = dw/dx + i dw/dy
          Z1[(size_t)(t + m * NT) * ZS + y].x = -(b[0][m].x * a[0][m].x + b[0][m].y * a[0][m].y);
      }
    }
    __syncthreads();

    // ---------------------------------------------------------------- C2: forward x-transform of column pairs -> advt
    for (int it = 0; it < (N / 2 + G - 1) / G; ++it) {
      const int pair = it * G + g;
      const bool valid = pair < N / 2;
      const int y = 2 * (valid ? pair : N / 2 - 1);
      cx<T> c[1][8];
#pragma unroll
      for (int m = 0; m < 8; ++m) c[0][m] = cx<T>{Z1[(size_t)(t + m * NT) * ZS + y].x, Z1[(size_t)(t + m * NT) * ZS + y + 1].x};
      fft_run<T, N, -1, 1, false, N>(c, tw, buf, parity, t, sync);
      sync();  // (the transform's last exchange may still be read by the slower lanes of the group)
#pragma unroll
      for (int m = 0; m < 8; ++m) buf[t + m * NT] = c[0][m];
      sync();
      if (valid) {
        for (int kx = t; kx <= N / 2; kx += NT) {
          const cx<T> zc = buf[kx], zn = buf[(N - kx) % N];
          // X_a = (Z(k) + conj Z(-k)) / 2 ;  X_b = (Z(k) - conj Z(-k)) / (2 i)
          AT[(size_t)kx * ZS + y] = cx<T>{T(0.5) * (zc.x + zn.x), T(0.5) * (zc.y - zn.y)};
          AT[(size_t)kx * ZS + y + 1] = cx<T>{T(0.5) * (zc.y + zn.y), T(0.5) * (zn.x - zc.x)};
        }
      }
      sync();
    }
    __syncthreads();

    // ---------------------------------------------------------------- R1: forward y-transform, mask, forcing, RK / CN
    for (int it = 0; it < (NH + G - 1) / G; ++it) {
      const int item = it * G + g;
      const bool valid = item < NH;
      const int pr = valid ? item : NH - 1;
      const int r1 = pr, r2 = (N - pr) % N;
      const bool self = r1 == r2;
      cx<T> a[1][8];
#pragma unroll
      for (int m = 0; m < 8; ++m) a[0][m] = pr < p.KF ? AT[(size_t)pr * ZS + t + m * NT] : cx<T>{T(0), T(0)};
      fft_run<T, N, -1, 1, false, N>(a, tw, buf, parity, t, sync);
      auto update = [&](int row, int col, cx<T> A, bool own) {
        if (!own) return;
        const int e = row * NH + col;
        const tab4<T> tb = p.tab[e];
        cx<T> F{tb.filt * A.x, tb.filt * A.y};
        if (p.fhat) F = F + p.fhat[e];
        const cx<T> w = wS[e];
        cx<T> h = F;
        if (rd_h) {
          const cx<T> ho = hS[e];
          h = cx<T>{fma_rn(ho.x, beta, F.x), fma_rn(ho.y, beta, F.y)};
        }
        if (wr_h) hS[e] = h;
        const T inv = rcp_cn(fma_rn(tb.lin, -mu, T(1)));
        const cx<T> lw{tb.lin * w.x, tb.lin * w.y};
        const cx<T> x{fma_rn(lw.x, mu, fma_rn(h.x, gdt, w.x)), fma_rn(lw.y, mu, fma_rn(h.y, gdt, w.y))};
        const cx<T> wn{inv * x.x, inv * x.y};
        wS[e] = wn;
        if (last_sub) {
          p.w_out[sb + e] = wn;
          if (want_dwdt) {
            const cx<T> o = p.w_in[sb + e];
            p.dwdt[sb + e] = cx<T>{p.inv_tdt * (wn.x - o.x), p.inv_tdt * (wn.y - o.y)};
          }
        }
      };
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int ky = t + m * NT;
        const bool lo = m < 4;
        update(lo ? r1 : r2, lo ? ky : N - ky, lo ? a[0][m] : conj(a[0][m]), valid && (lo || !self || (m == 4 && t == 0)));
      }
      if (t == 0) {
        update(r2, 0, conj(a[0][0]), valid && !self);
        update(r1, N / 2, a[0][4], valid && !self);
      }
    }
    __syncthreads();
  }
}

}  // namespace tcfd
