// Kernels of hot path B: the FNO3d / SFNO spectral convolution  y = irfftn( W (.) rfftn(x) )  on the
// four retained corner blocks (reference: fno/fno3d.py:86-116, fno/sfno.py:364-391, :433-457,
// fno/base.py:229-237).  fp32, x = (b, C, X, Y, T) with T innermost (the real-to-complex axis).
//
// Only 2mx * 2my * mt modes survive, so the transforms are PRUNED and the full spectrum is never
// materialised:
//   planes_fwd   per (b, c, x) plane [Y][T]: y-axis FFT of the real rows, two time samples per
//                complex transform and two transforms per thread (packed f32x2 lanes), kept ky only,
//                then the t-axis analysis as a small dense table product  ->  Z1 (b c, X, 2my, mt)
//   xaxis<FWD>   x-axis FFT of Z1 columns (two columns per packed transform), kept kx only
//                                                                   ->  Xh (b c, 2mx, 2my, mt)
//   mix_fwd      per mode  Yh[b,o] = sum_i Xh[b,i] W[i,o] (+ delta bias)   (complex, CUDA cores:
//                ~1 GFLOP against 3.4 GB of activations)
//   xaxis<INV>   zero-padded x-axis inverse                          ->  Z2 (b c, X, 2my, mt)
//   planes_inv   per plane: t-axis synthesis on the kept ky (table product), Hermitian part, y-axis
//                inverse FFT with two real outputs per complex transform  ->  y (b, C, X, Y, T')
// The t axis is table driven (any T, front zero padding, output resampling, normalisation and the
// C2R doubling all live in the host-built tables), which also makes the BACKWARD pass the same five
// kernels with conjugate-transposed tables (see sconv_api.cu).
#pragma once
#include "fft_core.cuh"
#include "tma.cuh"

namespace tcfd {

struct CtaSyncS {
  TCFD_D void operator()() const { __syncthreads(); }
};

struct SconvDims {
  int X, Y;            // spatial grid (powers of two)
  int mx, my, mt;      // retained modes; NKX = 2 mx, NKY = 2 my
  int Tin, Tout;       // time samples read / written by the plane kernels
  int nplanes_c;       // number of (b, c) slabs
};

TCFD_HD int kept_index(int k, int n, int m) {  // index of frequency k in the kept set, or -1
  if (k < m) return k;
  if (k >= n - m) return k - (n - 2 * m);
  return -1;
}
TCFD_HD int kept_freq(int ki, int n, int m) { return ki < m ? ki : n - 2 * m + ki; }

// ------------------------------------------------------------------------------------------
// planes_fwd.  CTA = GP groups of NT = Y/8 threads; group g works on plane blockIdx.x*GP + g.
// A(kt, t): analysis table [mt][Tin] complex.
// shared memory per group: tile [Y*Tin] floats | exchange Y cx<f2> | E (2my+1) cx<f2> | Xy [2my][Tq*4] cx<float>
template <int Y>
struct PlanesSmem {
  static constexpr int NT = Y / 8;
  static constexpr int GP = (128 / NT) > 0 ? (128 / NT) : 1;
  TCFD_HD static int tq(int T) { return (T + 3) / 4; }
  TCFD_HD static size_t group_bytes(int T, int my) {
    size_t b = (size_t)Y * T * 4;                         // plane tile
    b = (b + 15) / 16 * 16 + (size_t)Y * 16;              // exchange
    b += (size_t)(2 * my + 1) * 16;                       // E / Dh entries
    b += (size_t)(2 * my) * tq(T) * 4 * 8;                // Xy / D
    return (b + 15) / 16 * 16;
  }
};

template <int Y>
__global__ void __launch_bounds__(PlanesSmem<Y>::GP * (Y / 8))
sconv_planes_fwd_kernel(const float* __restrict__ x, cx<float>* __restrict__ Z1, const cx<float>* __restrict__ A,
                        const cx<float>* __restrict__ twtab, SconvDims d, int nplanes) {
  typedef PlanesSmem<Y> S;
  constexpr int NT = S::NT, GP = S::GP;
  const int T = d.Tin, my = d.my, mt = d.mt, NKY = 2 * my, TQ = S::tq(T);
  TCFD_DYN_SMEM(smem_raw);
  const int g = threadIdx.x / NT, t = threadIdx.x % NT;
  unsigned char* base = smem_raw + (size_t)g * S::group_bytes(T, my);
  float* tile = reinterpret_cast<float*>(base);
  cx<f2>* buf = reinterpret_cast<cx<f2>*>(base + ((size_t)Y * T * 4 + 15) / 16 * 16);
  cx<f2>* Es = buf + Y;
  cx<float>* Xy = reinterpret_cast<cx<float>*>(Es + (2 * my + 1));  // [NKY][TQ*4]
  FftTwiddles<float, Y> tw;
  tw.load(twtab, t);
  CtaSyncS sync;
  int parity = 0;
  const int plane = blockIdx.x * GP + g;
  const bool valid = plane < nplanes;
  const int pl = valid ? plane : nplanes - 1;

  // stage the plane (contiguous Y*T floats)
  const float* src = x + (size_t)pl * Y * T;
  for (int i = t; i < Y * T; i += NT) tile[i] = src[i];
  __syncthreads();
  for (int q = 0; q < TQ; ++q) {
    cx<f2> z[1][8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const float* r = tile + (t + m * NT) * T + 4 * q;
      const float v0 = r[0], v1 = (4 * q + 1 < T) ? r[1] : 0.f, v2 = (4 * q + 2 < T) ? r[2] : 0.f,
                  v3 = (4 * q + 3 < T) ? r[3] : 0.f;
      z[0][m] = cx<f2>{f2(v0, v2), f2(v1, v3)};  // lane lo: x[t0] + i x[t0+1]; lane hi: x[t0+2] + i x[t0+3]
    }
    fft_run<f2, Y, -1, 1, false, Y>(z, tw, buf, parity, t, sync);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int ky = t + m * NT;
      if (ky <= my) Es[ky] = z[0][m];
      else if (ky >= Y - my) Es[my + 1 + ky - (Y - my)] = z[0][m];
    }
    __syncthreads();
    // separate the two real transforms of each lane at the kept ky
    for (int kyi = t; kyi < NKY; kyi += NT) {
      const int ky = kept_freq(kyi, Y, my), kn = (Y - ky) % Y;
      const cx<f2> e = Es[ky <= my ? ky : my + 1 + ky - (Y - my)];
      const cx<f2> n = Es[kn <= my ? kn : my + 1 + kn - (Y - my)];
      // X_a = (E(k) + conj E(-k)) / 2 ;  X_b = (E(k) - conj E(-k)) / (2i)
      const f2 ar = 0.5f * (e.x + n.x), ai = 0.5f * (e.y - n.y);
      const f2 br = 0.5f * (e.y + n.y), bi = 0.5f * (n.x - e.x);
      cx<float>* o = Xy + (size_t)kyi * TQ * 4 + 4 * q;
      o[0] = cx<float>{ar.lo, ai.lo};
      o[1] = cx<float>{br.lo, bi.lo};
      o[2] = cx<float>{ar.hi, ai.hi};
      o[3] = cx<float>{br.hi, bi.hi};
    }
    __syncthreads();
  }
  // t-axis analysis on the kept ky:  Z1[kyi][kt] = sum_t A[kt][t] Xy[kyi][t]
  if (valid) {
    cx<float>* dst = Z1 + (size_t)plane * NKY * mt;
    for (int j = t; j < NKY * mt; j += NT) {
      const int kyi = j / mt, kt = j % mt;
      const cx<float>* xr = Xy + (size_t)kyi * TQ * 4;
      const cx<float>* ar = A + (size_t)kt * T;
      float sr = 0.f, si = 0.f;
      for (int tt = 0; tt < T; ++tt) {
        const cx<float> a = ar[tt], v = xr[tt];
        sr = fmaf(a.x, v.x, sr); sr = fmaf(-a.y, v.y, sr);
        si = fmaf(a.x, v.y, si); si = fmaf(a.y, v.x, si);
      }
      dst[j] = cx<float>{sr, si};
    }
  }
}

// ------------------------------------------------------------------------------------------
// planes_inv.  Sy(t, kt): synthesis table [Tout][mt] complex;  y[t] = Re( sum_kt Sy[t][kt] c[kt] ).
template <int Y>
__global__ void __launch_bounds__(PlanesSmem<Y>::GP * (Y / 8))
sconv_planes_inv_kernel(const cx<float>* __restrict__ Z2, float* __restrict__ y, const cx<float>* __restrict__ Sy,
                        const cx<float>* __restrict__ twtab, SconvDims d, int nplanes) {
  typedef PlanesSmem<Y> S;
  constexpr int NT = S::NT, GP = S::GP;
  const int T = d.Tout, my = d.my, mt = d.mt, NKY = 2 * my, TQ = S::tq(T);
  TCFD_DYN_SMEM(smem_raw);
  const int g = threadIdx.x / NT, t = threadIdx.x % NT;
  unsigned char* base = smem_raw + (size_t)g * S::group_bytes(T, my);
  float* tile = reinterpret_cast<float*>(base);
  cx<f2>* buf = reinterpret_cast<cx<f2>*>(base + ((size_t)Y * T * 4 + 15) / 16 * 16);
  cx<f2>* Dh = buf + Y;                                             // (2my+1) entries of the current quad
  cx<float>* D = reinterpret_cast<cx<float>*>(Dh + (2 * my + 1));   // [NKY][TQ*4]
  FftTwiddles<float, Y> tw;
  tw.load(twtab, t);
  CtaSyncS sync;
  int parity = 0;
  const int plane = blockIdx.x * GP + g;
  const bool valid = plane < nplanes;
  const int pl = valid ? plane : nplanes - 1;

  // t-axis synthesis on the kept ky:  D[kyi][t] = sum_kt Sy[t][kt] Z2[kyi][kt]   (complex)
  const cx<float>* src = Z2 + (size_t)pl * NKY * mt;
  for (int j = t; j < NKY * TQ * 4; j += NT) {
    const int kyi = j / (TQ * 4), tt = j % (TQ * 4);
    float sr = 0.f, si = 0.f;
    if (tt < T) {
      const cx<float>* zr = src + (size_t)kyi * mt;
      const cx<float>* sy = Sy + (size_t)tt * mt;
      for (int kt = 0; kt < mt; ++kt) {
        const cx<float> a = sy[kt], v = zr[kt];
        sr = fmaf(a.x, v.x, sr); sr = fmaf(-a.y, v.y, sr);
        si = fmaf(a.x, v.y, si); si = fmaf(a.y, v.x, si);
      }
    }
    D[j] = cx<float>{sr, si};
  }
  __syncthreads();
  for (int q = 0; q < TQ; ++q) {
    // Hermitian part along ky of the four time samples of this quad, packed two per lane:
    // lane lo: Dh[t0] + i Dh[t0+1], lane hi: Dh[t0+2] + i Dh[t0+3]   (support: ky in [0,my] u [Y-my,Y-1])
    for (int e = t; e < 2 * my + 1; e += NT) {
      const int ky = e <= my ? e : Y - my + (e - my - 1);
      const int kn = (Y - ky) % Y;
      const int i1 = kept_index(ky, Y, my), i2 = kept_index(kn, Y, my);
      cx<float> h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const cx<float> a = i1 >= 0 ? D[(size_t)i1 * TQ * 4 + 4 * q + j] : cx<float>{0.f, 0.f};
        const cx<float> b = i2 >= 0 ? D[(size_t)i2 * TQ * 4 + 4 * q + j] : cx<float>{0.f, 0.f};
        h[j] = cx<float>{0.5f * (a.x + b.x), 0.5f * (a.y - b.y)};
      }
      // (h0 + i h1, h2 + i h3)
      Dh[e] = cx<f2>{f2(h[0].x - h[1].y, h[2].x - h[3].y), f2(h[0].y + h[1].x, h[2].y + h[3].x)};
    }
    __syncthreads();
    cx<f2> z[1][8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int ky = t + m * NT;
      const cx<f2> zero{f2(0.f), f2(0.f)};
      z[0][m] = ky <= my ? Dh[ky] : (ky >= Y - my ? Dh[my + 1 + ky - (Y - my)] : zero);
    }
    fft_run<f2, Y, +1, 1, false, Y>(z, tw, buf, parity, t, sync);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      float* r = tile + (t + m * NT) * T + 4 * q;
      r[0] = z[0][m].x.lo;
      if (4 * q + 1 < T) r[1] = z[0][m].y.lo;
      if (4 * q + 2 < T) r[2] = z[0][m].x.hi;
      if (4 * q + 3 < T) r[3] = z[0][m].y.hi;
    }
    __syncthreads();
  }
  if (valid) {
    float* dst = y + (size_t)plane * Y * T;
    for (int i = t; i < Y * T; i += NT) dst[i] = tile[i];
  }
}

// ------------------------------------------------------------------------------------------
// Pipelined plane kernels (second generation).  Same arithmetic as the two kernels above, but
//   * persistent: a group walks over planes, the grid is sized to the SM count;
//   * the plane (forward) / its truncated spectrum block (inverse) of the NEXT plane is staged by a bulk
//     copy (TMA engine, mbarrier completion) while the current plane is transformed; the inverse kernel
//     writes its output plane with a bulk store from a double-buffered tile;
//   * groups synchronise among themselves only (a group is one warp for Y <= 256: __syncwarp; a named
//     barrier for Y = 512), so the groups of a CTA de-synchronise and hide each other's latencies;
//   * the t-axis tables live in shared memory.
template <int NT>
struct GroupSync {
  int id;
  TCFD_D void operator()() const {
#ifndef TCFD_EMU
    if constexpr (NT <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(NT) : "memory");
#else
    __syncthreads();
#endif
  }
};

// Small dense complex product of one thread group (the t-axis table products of the plane kernels):
//   out(i, c) = sum_k A[c * as + k] * V[i * vs + k]      i < NI rows, c < NC columns, k < NK   (shared memory operands)
// Thread (pl, kl) -- pl = t % NP, kl = t / NP, NP = ceil(NC / 2), KL = NT / NP, computed once per kernel by the caller --
// owns the column pair (2 pl, 2 pl + 1) and the rows kl, kl + KL, kl + 2 KL, ... in register tiles of RK rows: a
// table entry it loads serves RK rows, an input entry two columns (2 + RK loads per 2 RK multiply-adds instead of 2
// per multiply-add), the accumulators are packed (re, im) pairs (two FFMA2 per complex multiply-add), and no index
// of the loop nest needs a division.  store(i, c, value) writes one result.
template <int RK, class Store>
TCFD_D void group_cproduct(const cx<float>* A, int as, const cx<float>* V, int vs, int NI, int NC, int NK, int pl, int kl, int KL,
                           bool active, Store store) {
  if (!active) return;
  const int c0 = 2 * pl, c1 = c0 + 1;
  const bool has1 = c1 < NC;
  const cx<float>* a0p = A + (size_t)c0 * as;
  const cx<float>* a1p = A + (size_t)(has1 ? c1 : c0) * as;
  for (int i0 = kl; i0 < NI; i0 += KL * RK) {
    f2 acc0[RK], acc1[RK];
    const cx<float>* vp[RK];
#pragma unroll
    for (int r = 0; r < RK; ++r) {
      acc0[r] = f2(0.f);
      acc1[r] = f2(0.f);
      const int i = i0 + r * KL;
      vp[r] = V + (size_t)(i < NI ? i : i0) * vs;  // rows past the end repeat row i0 (results discarded)
    }
    for (int k = 0; k < NK; ++k) {
      const cx<float> a0 = a0p[k], a1 = a1p[k];
#pragma unroll
      for (int r = 0; r < RK; ++r) {
        const cx<float> v = vp[r][k];
        const f2 v1(v.x, v.y), v2(-v.y, v.x);
        acc0[r] = fma_rn(v2, a0.y, fma_rn(v1, a0.x, acc0[r]));
        acc1[r] = fma_rn(v2, a1.y, fma_rn(v1, a1.x, acc1[r]));
      }
    }
#pragma unroll
    for (int r = 0; r < RK; ++r) {
      const int i = i0 + r * KL;
      if (i < NI) {
        store(i, c0, cx<float>{acc0[r].lo, acc0[r].hi});
        if (has1) store(i, c1, cx<float>{acc1[r].lo, acc1[r].hi});
      }
    }
  }
}

template <int Y>
struct Planes2Smem {
  static constexpr int NT = Y / 8;
  // two groups per CTA (at least one warp): shared memory, not the CTA shape, bounds the residency
  static constexpr int GP = (64 / NT) > 0 ? (64 / NT) : 1;
  TCFD_HD static int tq(int T) { return (T + 3) / 4; }
  // padded row strides (complex entries): an odd stride spreads a column walk over all banks
  TCFD_HD static int xs(int T) { return tq(T) * 4 + 1; }
  TCFD_HD static int ts(int n) { return n | 1; }
  TCFD_HD static size_t a16(size_t b) { return (b + 15) / 16 * 16; }
  TCFD_HD static size_t tile_bytes(int T) { return a16((size_t)Y * T * 4); }
  TCFD_HD static size_t zin_bytes(int my, int mt) { return a16((size_t)2 * my * mt * 8); }
  // group: tile | exchange | E/Dh | Xy/D | zin (inverse only) | barrier
  // (forward: the E entries live in the exchange buffer, which is idle between two transforms -- C4: 18.1 KB per group,
  // six CTAs of two groups per SM instead of five)
  TCFD_HD static size_t group_bytes(int T, int my, int mt, bool inv) {
    size_t b = tile_bytes(T) + (size_t)Y * 16 + (size_t)(2 * my) * xs(T) * 8;
    if (inv) b += (size_t)(2 * my + 1) * 16 + zin_bytes(my, mt);
    else b += 16;  // my = Y/2 keeps Y + 1 entries
    return a16(b + 16);
  }
  TCFD_HD static size_t table_bytes(int T, int mt) {  // [mt][ts(T)] (analysis) or [T][ts(mt)] (synthesis)
    const int a = T * ts(mt), b = mt * ts(T);
    return a16((size_t)(a > b ? a : b) * 8);
  }
};

template <int Y>
__global__ void __launch_bounds__(Planes2Smem<Y>::GP * (Y / 8))
sconv_planes_fwd2_kernel(const float* __restrict__ x, cx<float>* __restrict__ Z1, const cx<float>* __restrict__ A,
                         const cx<float>* __restrict__ twtab, SconvDims d, int nplanes) {
  typedef Planes2Smem<Y> S;
  constexpr int NT = S::NT, GP = S::GP;
  const int T = d.Tin, my = d.my, mt = d.mt, NKY = 2 * my, TQ = S::tq(T), XS = S::xs(T), AS = S::ts(T);
  TCFD_DYN_SMEM(smem_raw);
  const int g = threadIdx.x / NT, t = threadIdx.x % NT;
  cx<float>* As = reinterpret_cast<cx<float>*>(smem_raw);  // [mt][T], CTA-shared
  unsigned char* base = smem_raw + S::table_bytes(T, mt) + (size_t)g * S::group_bytes(T, my, mt, false);
  const size_t TB = S::tile_bytes(T);
  const float* tile = reinterpret_cast<const float*>(base);
  cx<f2>* buf = reinterpret_cast<cx<f2>*>(base + TB);
  cx<f2>* Es = buf;  // kept ky of the transform just finished: every thread has left the last exchange (its closing barrier)
  cx<float>* Xy = reinterpret_cast<cx<float>*>(buf + Y + 1);  // [NKY][XS]
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(Xy) + S::a16((size_t)NKY * XS * 8));
  FftTwiddles<float, Y> tw;
  tw.load(twtab, t);
  GroupSync<NT> sync{1 + g};
  int parity = 0;
  const bool teven = (T & 1) == 0;
  // thread mapping of the register-tiled t-axis product (one division per kernel, none per plane)
  const int NP = (mt + 1) / 2, KL = NT / (NP > 0 ? NP : 1), pl = t % NP, kl = t / NP;
  const bool tiled = NP <= NT;
  const int rk1 = tiled ? (NKY + KL - 1) / KL : 0;
  for (int i = threadIdx.x; i < mt * T; i += blockDim.x) As[(i / T) * AS + i % T] = A[i];
  const int stride = (int)gridDim.x * GP;
  const int iters = (nplanes + stride - 1) / stride;
  auto plane_of = [&](int it) { return (it * (int)gridDim.x + (int)blockIdx.x) * GP + g; };
  auto issue = [&](int it) {  // one thread of the group
    int pl = plane_of(it);
    if (pl >= nplanes) pl = nplanes - 1;
    stage_expect(bar, (unsigned)((size_t)Y * T * 4));
    bulk_load(base, x + (size_t)pl * Y * T, (unsigned)((size_t)Y * T * 4), bar);
  };
  if (t == 0) stage_barrier_init(bar);
  __syncthreads();  // tables and barriers visible
  if (t == 0 && iters > 0) issue(0);

  for (int it = 0; it < iters; ++it) {
    const int plane = plane_of(it);
    const bool valid = plane < nplanes;
    tile_load_wait(bar, (unsigned)(it & 1));
    for (int q = 0; q < TQ; ++q) {
      cx<f2> z[1][8];
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const float* r = tile + (t + m * NT) * T + 4 * q;
        float v0, v1, v2 = 0.f, v3 = 0.f;
        if (teven) {  // even T: rows and quads are 8-byte aligned
          const cx<float> p0 = *reinterpret_cast<const cx<float>*>(r);
          v0 = p0.x; v1 = p0.y;
          if (4 * q + 2 < T) {
            const cx<float> p1 = *reinterpret_cast<const cx<float>*>(r + 2);
            v2 = p1.x; v3 = p1.y;
          }
        } else {
          v0 = r[0];
          v1 = (4 * q + 1 < T) ? r[1] : 0.f;
          v2 = (4 * q + 2 < T) ? r[2] : 0.f;
          v3 = (4 * q + 3 < T) ? r[3] : 0.f;
        }
        // lane lo: x[t0] + i x[t0+2]; lane hi: x[t0+1] + i x[t0+3] -- the two 8-byte loads ARE the packed operands (no repacking moves)
        z[0][m] = cx<f2>{f2(v0, v1), f2(v2, v3)};
      }
      if (q == TQ - 1) {
        // the plane is in registers: the next one is staged under the rest of this plane's work
        sync();
        if (t == 0 && it + 1 < iters) issue(it + 1);
      }
      fft_run<f2, Y, -1, 1, false, Y>(z, tw, buf, parity, t, sync);
      if (my < NT) {
        // (group-uniform) only the first and the last register row hold kept frequencies: two predicated stores instead of
        // sixteen divergent tests, and the other outputs of the closing pass are dead code
        if (t <= my) Es[t] = z[0][0];
        if (t >= NT - my) Es[2 * my + 1 + t - NT] = z[0][7];
      } else {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int ky = t + m * NT;
          if (ky <= my) Es[ky] = z[0][m];
          else if (ky >= Y - my) Es[my + 1 + ky - (Y - my)] = z[0][m];
        }
      }
      sync();
      // separation of the two real transforms of each lane, one mirror pair (ky = p, Y - p) per thread: both outputs
      // come from the same two entries and are complex conjugates of each other (a single pass over p = 0..my)
      for (int p = t; p <= my; p += NT) {
        const cx<f2> e = Es[p];
        const int kn = (Y - p) % Y;  // (kn == my only when my = Y/2: that frequency then has a single entry, Es[my])
        const cx<f2> n = Es[kn <= my ? kn : my + 1 + kn - (Y - my)];
        // X_a = (E(k) + conj E(-k)) / 2 ;  X_b = (E(k) - conj E(-k)) / (2i)
        const f2 ar = 0.5f * (e.x + n.x), ai = 0.5f * (e.y - n.y);
        const f2 br = 0.5f * (e.y + n.y), bi = 0.5f * (n.x - e.x);
        if (p < my) {  // ky = p (ky = my is not a kept output, only the partner of Y - my)
          cx<float>* o = Xy + (size_t)p * XS + 4 * q;
          o[0] = cx<float>{ar.lo, ai.lo};  // t0     (real part of lane lo)
          o[1] = cx<float>{ar.hi, ai.hi};  // t0 + 1 (real part of lane hi)
          o[2] = cx<float>{br.lo, bi.lo};  // t0 + 2 (imaginary part of lane lo)
          o[3] = cx<float>{br.hi, bi.hi};  // t0 + 3
        }
        if (p > 0) {  // ky = Y - p, kept index 2 my - p: the conjugates
          cx<float>* o = Xy + (size_t)(2 * my - p) * XS + 4 * q;
          o[0] = cx<float>{ar.lo, -ai.lo};
          o[1] = cx<float>{ar.hi, -ai.hi};
          o[2] = cx<float>{br.lo, -bi.lo};
          o[3] = cx<float>{br.hi, -bi.hi};
        }
      }
      sync();
    }
    if (valid) {
      cx<float>* dst = Z1 + (size_t)plane * NKY * mt;
      if (tiled) {
        // t-axis analysis: Z1[kyi][kt] = sum_tt A[kt][tt] Xy[kyi][tt], register-tiled (group_cproduct)
        // rows per thread so that ONE pass covers the NKY rows when that takes at most 7 (C4: 40 rows / 8 = 5)
        auto put = [&](int kyi, int kt, cx<float> v) { dst[(size_t)kyi * mt + kt] = v; };
        if (rk1 == 5) group_cproduct<5>(As, AS, Xy, XS, NKY, mt, T, pl, kl, KL, kl < KL, put);
        else if (rk1 == 6 || rk1 == 7) group_cproduct<7>(As, AS, Xy, XS, NKY, mt, T, pl, kl, KL, kl < KL, put);
        else group_cproduct<4>(As, AS, Xy, XS, NKY, mt, T, pl, kl, KL, kl < KL, put);
      } else {
        for (int j = t; j < NKY * mt; j += NT) {
          const int kyi = j / mt, kt = j % mt;
          const cx<float>* xr = Xy + (size_t)kyi * XS;
          const cx<float>* ar = As + (size_t)kt * AS;
          float sr = 0.f, si = 0.f;
          for (int tt = 0; tt < T; ++tt) {
            const cx<float> a = ar[tt], v = xr[tt];
            sr = fmaf(a.x, v.x, sr); sr = fmaf(-a.y, v.y, sr);
            si = fmaf(a.x, v.y, si); si = fmaf(a.y, v.x, si);
          }
          dst[j] = cx<float>{sr, si};
        }
      }
    }
    sync();  // Xy / Es are re-used by the next plane
  }
}

template <int Y>
__global__ void __launch_bounds__(Planes2Smem<Y>::GP * (Y / 8))
sconv_planes_inv2_kernel(const cx<float>* __restrict__ Z2, float* __restrict__ y, const cx<float>* __restrict__ Sy,
                         const cx<float>* __restrict__ twtab, SconvDims d, int nplanes) {
  typedef Planes2Smem<Y> S;
  constexpr int NT = S::NT, GP = S::GP;
  const int T = d.Tout, my = d.my, mt = d.mt, NKY = 2 * my, TQ = S::tq(T), XS = S::xs(T), SS = S::ts(mt);
  TCFD_DYN_SMEM(smem_raw);
  const int g = threadIdx.x / NT, t = threadIdx.x % NT;
  cx<float>* Ss = reinterpret_cast<cx<float>*>(smem_raw);  // [T][mt], CTA-shared
  unsigned char* base = smem_raw + S::table_bytes(T, mt) + (size_t)g * S::group_bytes(T, my, mt, true);
  const size_t TB = S::tile_bytes(T), ZB = S::zin_bytes(my, mt);
  float* tile = reinterpret_cast<float*>(base);
  cx<f2>* buf = reinterpret_cast<cx<f2>*>(base + TB);
  cx<f2>* Dh = buf + Y;
  cx<float>* D = reinterpret_cast<cx<float>*>(Dh + (2 * my + 1));  // [NKY][XS]
  unsigned char* zin = reinterpret_cast<unsigned char*>(D) + S::a16((size_t)NKY * XS * 8);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(zin + ZB);
  FftTwiddles<float, Y> tw;
  tw.load(twtab, t);
  GroupSync<NT> sync{1 + g};
  int parity = 0;
  const bool teven = (T & 1) == 0;
  // thread mapping of the register-tiled t-axis product: column pairs are pairs of output time samples here
  const int NP = (T + 1) / 2, KL = NT / (NP > 0 ? NP : 1), pl = t % NP, kl = t / NP;
  const bool tiled = NP <= NT;
  for (int i = threadIdx.x; i < mt * T; i += blockDim.x) Ss[(i / mt) * SS + i % mt] = Sy[i];
  const int stride = (int)gridDim.x * GP;
  const int iters = (nplanes + stride - 1) / stride;
  auto plane_of = [&](int it) { return (it * (int)gridDim.x + (int)blockIdx.x) * GP + g; };
  auto issue = [&](int it) {  // one thread of the group
    int pl = plane_of(it);
    if (pl >= nplanes) pl = nplanes - 1;
    stage_expect(bar, (unsigned)((size_t)NKY * mt * 8));
    bulk_load(zin, Z2 + (size_t)pl * NKY * mt, (unsigned)((size_t)NKY * mt * 8), bar);
  };
  if (t == 0) stage_barrier_init(bar);
  __syncthreads();
  if (t == 0 && iters > 0) issue(0);

  for (int it = 0; it < iters; ++it) {
    const int plane = plane_of(it);
    const bool valid = plane < nplanes;
    tile_load_wait(bar, (unsigned)(it & 1));
    // t-axis synthesis on the kept ky:  D[kyi][t] = sum_kt Sy[t][kt] Z2[kyi][kt]   (complex)
    const cx<float>* src = reinterpret_cast<const cx<float>*>(zin);
    if (tiled) {
      // D[kyi][tt] = sum_kt Sy[tt][kt] Z2[kyi][kt], register-tiled (group_cproduct); the padding entries tt in [T, 4 TQ) stay 0
      group_cproduct<4>(Ss, SS, src, mt, NKY, T, mt, pl, kl, KL, kl < KL,
                        [&](int kyi, int tt, cx<float> v) { D[(size_t)kyi * XS + tt] = v; });
      if (T < TQ * 4)
        for (int kyi = t; kyi < NKY; kyi += NT)
          for (int tt = T; tt < TQ * 4; ++tt) D[(size_t)kyi * XS + tt] = cx<float>{0.f, 0.f};
    } else {
      for (int j = t; j < NKY * TQ * 4; j += NT) {
        const int kyi = j / (TQ * 4), tt = j % (TQ * 4);
        float sr = 0.f, si = 0.f;
        if (tt < T) {
          const cx<float>* zr = src + (size_t)kyi * mt;
          const cx<float>* sy = Ss + (size_t)tt * SS;
          for (int kt = 0; kt < mt; ++kt) {
            const cx<float> a = sy[kt], v = zr[kt];
            sr = fmaf(a.x, v.x, sr); sr = fmaf(-a.y, v.y, sr);
            si = fmaf(a.x, v.y, si); si = fmaf(a.y, v.x, si);
          }
        }
        D[(size_t)kyi * XS + tt] = cx<float>{sr, si};
      }
    }
    sync();  // D complete, zin consumed
    if (t == 0) {
      if (it + 1 < iters) issue(it + 1);  // the next block is staged under this plane's transforms
      bulk_store_wait_read<0>();          // the previous plane's store has read the tile (thread 0 joins the next
                                          // group barrier only after this, and the tile is first written after it)
    }
    for (int q = 0; q < TQ; ++q) {
      for (int e = t; e < 2 * my + 1; e += NT) {
        const int ky = e <= my ? e : Y - my + (e - my - 1);
        const int kn = (Y - ky) % Y;
        const int i1 = kept_index(ky, Y, my), i2 = kept_index(kn, Y, my);
        cx<float> h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const cx<float> a = i1 >= 0 ? D[(size_t)i1 * XS + 4 * q + j] : cx<float>{0.f, 0.f};
          const cx<float> b = i2 >= 0 ? D[(size_t)i2 * XS + 4 * q + j] : cx<float>{0.f, 0.f};
          h[j] = cx<float>{0.5f * (a.x + b.x), 0.5f * (a.y - b.y)};
        }
        Dh[e] = cx<f2>{f2(h[0].x - h[1].y, h[2].x - h[3].y), f2(h[0].y + h[1].x, h[2].y + h[3].x)};
      }
      sync();
      cx<f2> z[1][8];
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int ky = t + m * NT;
        const cx<f2> zero{f2(0.f), f2(0.f)};
        z[0][m] = ky <= my ? Dh[ky] : (ky >= Y - my ? Dh[my + 1 + ky - (Y - my)] : zero);
      }
      fft_run<f2, Y, +1, 1, false, Y>(z, tw, buf, parity, t, sync);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        float* r = tile + (t + m * NT) * T + 4 * q;
        if (teven) {
          *reinterpret_cast<cx<float>*>(r) = cx<float>{z[0][m].x.lo, z[0][m].y.lo};
          if (4 * q + 2 < T) *reinterpret_cast<cx<float>*>(r + 2) = cx<float>{z[0][m].x.hi, z[0][m].y.hi};
        } else {
          r[0] = z[0][m].x.lo;
          if (4 * q + 1 < T) r[1] = z[0][m].y.lo;
          if (4 * q + 2 < T) r[2] = z[0][m].x.hi;
          if (4 * q + 3 < T) r[3] = z[0][m].y.hi;
        }
      }
      sync();
    }
    fence_async_smem();  // this thread's tile writes -> visible to the bulk store
    sync();
    if (valid && t == 0) bulk_store(y + (size_t)plane * Y * T, tile, (unsigned)((size_t)Y * T * 4));
  }
  if (t == 0) bulk_store_wait_read<0>();  // shared memory must outlive the reads of the last store
}

// ------------------------------------------------------------------------------------------
// Third-generation inverse plane kernel: the y-axis inverse as a PRUNED transform without shared-memory exchanges.
// Only the 2 my + 1 frequencies |k| <= my of the Y inputs are non-zero.  With Y = 8 NT and n = t + NT h:
//     y[t + NT h] = sum_b e^{+2 pi i b h / 8} U_t[b],      U_t[b] = sum_{|k| <= my, k = b (mod 8)} X[k] e^{+2 pi i k t / Y}
// so thread t accumulates its eight U_t[b] from broadcast reads of the 2 my + 1 inputs (its twiddles e^{2 pi i j t / Y},
// j = 1..my, live in registers; negative k take the conjugates) and one radix-8 butterfly in registers yields the
// same eight rows t + m NT the full transform leaves in the thread.  Against the Stockham transform (two exchanges =
// 128 shared-memory wavefronts per packed transform) this reads 2 my + 1 broadcast entries (one wavefront each) for a
// similar number of packed multiply-adds; measured with ncu the second-generation kernel kept the shared-memory
// data pipe 78 % busy (profiles/r2h_*).  The truncated spectrum block is staged row by row into rows padded to an
// even stride that spreads the rows a warp reads at once over the banks (mt even; odd mt keeps one contiguous copy).
template <int Y>
struct Planes3Smem {
  static constexpr int NT = Y / 8;
  static constexpr int GP = (64 / NT) > 0 ? (64 / NT) : 1;
  TCFD_HD static int tq(int T) { return (T + 3) / 4; }
  TCFD_HD static int xs(int T) { return tq(T) * 4 + 1; }
  TCFD_HD static int ts(int n) { return n | 1; }
  TCFD_HD static int zs(int mt) { return (mt & 1) ? mt : mt + 2; }  // padded row of the staged spectrum block (entries)
  TCFD_HD static size_t a16(size_t b) { return (b + 15) / 16 * 16; }
  TCFD_HD static size_t tile_bytes(int T) { return a16((size_t)Y * T * 4); }
  TCFD_HD static size_t dh_bytes(int my) { return (size_t)(2 * my + 1) * 16; }
  TCFD_HD static size_t zin_bytes(int my, int mt) { return a16((size_t)2 * my * zs(mt) * 8); }
  // group: tile | Dh | D | zin | barrier  (C4: 18.3 KB -> six CTAs of two groups per SM)
  TCFD_HD static size_t group_bytes(int T, int my, int mt) {
    return a16(tile_bytes(T) + dh_bytes(my) + a16((size_t)(2 * my + 1) * xs(T) * 8) + zin_bytes(my, mt) + 16);  // D: + one zero row
  }
  TCFD_HD static size_t table_bytes(int T, int mt) { return a16((size_t)T * ts(mt) * 8); }
};

// Pruned inverse transform of one thread: in(k) for k in [-nneg, npos) are the non-zero inputs (group-uniform
// addresses: broadcast reads), w[j-1] = e^{+2 pi i j t / N}; z[h] = output t + h N/8.
// EXACT: npos - 1 <= MT == nneg is known to the caller (the benchmark geometries: my == MYT), no bound checks
template <int MT, bool EXACT = false, class In>
TCFD_D void pruned_inverse(In in, int npos, int nneg, const cx<float> (&w)[MT], cx<f2> (&z)[8]) {
#pragma unroll
  for (int b = 1; b < 8; ++b) z[b] = cx<f2>{f2(0.f), f2(0.f)};
  z[0] = in(0);
#pragma unroll
  for (int j = 1; j <= MT; ++j) {
    if (!EXACT && j >= npos && j > nneg) break;  // group-uniform
    const int bp = j & 7, bn = (8 - (j & 7)) & 7;  // compile-time after unrolling
    const cx<float> wj = w[j - 1];
    if (EXACT || j < npos) {  // in(+j) * w
      const cx<f2> xp = in(j);
      z[bp].x = fma_rn(xp.y, -wj.y, fma_rn(xp.x, wj.x, z[bp].x));
      z[bp].y = fma_rn(xp.y, wj.x, fma_rn(xp.x, wj.y, z[bp].y));
    }
    if (EXACT || j <= nneg) {  // in(-j) * conj(w)
      const cx<f2> xn = in(-j);
      z[bn].x = fma_rn(xn.y, wj.y, fma_rn(xn.x, wj.x, z[bn].x));
      z[bn].y = fma_rn(xn.x, -wj.y, fma_rn(xn.y, wj.x, z[bn].y));
    }
  }
  radix8<+1>(z);
}
// the y-axis instance; Dh: entries k = 0..my, then k = -my..-1 (the layout of the Hermitian step)
template <int MYT, bool EXACT = false>
TCFD_D void pruned_inverse_y(const cx<f2>* Dh, int my, const cx<float> (&w)[MYT], cx<f2> (&z)[8]) {
  if constexpr (EXACT)  // my == MYT: entries +1..+MYT and -1..-MYT all exist
    pruned_inverse<MYT, true>([&](int kk) { return Dh[kk >= 0 ? kk : 2 * MYT + 1 + kk]; }, MYT + 1, MYT, w, z);
  else
    pruned_inverse<MYT>([&](int kk) { return Dh[kk >= 0 ? kk : 2 * my + 1 + kk]; }, my + 1, my, w, z);
}

template <int Y, int MYT, bool EXACT>
__global__ void __launch_bounds__(Planes3Smem<Y>::GP * (Y / 8))
sconv_planes_inv3_kernel(const cx<float>* __restrict__ Z2, float* __restrict__ y, const cx<float>* __restrict__ Sy,
                         const cx<float>* __restrict__ twtab, SconvDims d, int nplanes) {
  typedef Planes3Smem<Y> S;
  constexpr int NT = S::NT, GP = S::GP;
  const int T = d.Tout, my = d.my, mt = d.mt, NKY = 2 * my, TQ = S::tq(T), XS = S::xs(T), SS = S::ts(mt), ZS = S::zs(mt);
  TCFD_DYN_SMEM(smem_raw);
  const int g = threadIdx.x / NT, t = threadIdx.x % NT;
  cx<float>* Ss = reinterpret_cast<cx<float>*>(smem_raw);  // [T][mt], CTA-shared
  unsigned char* base = smem_raw + S::table_bytes(T, mt) + (size_t)g * S::group_bytes(T, my, mt);
  float* tile = reinterpret_cast<float*>(base);
  cx<f2>* Dh0 = reinterpret_cast<cx<f2>*>(base + S::tile_bytes(T));
  cx<float>* D = reinterpret_cast<cx<float>*>(Dh0 + (2 * my + 1));  // [NKY][XS]
  unsigned char* zin = reinterpret_cast<unsigned char*>(D) + S::a16((size_t)(NKY + 1) * XS * 8);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(zin + S::zin_bytes(my, mt));
  GroupSync<NT> sync{1 + g};
  // row NKY of D stays zero: the partner of the two entries (ky = my, ky = Y - my) whose mirror is not a kept frequency
  for (int i = t; i < XS; i += NT) D[(size_t)NKY * XS + i] = cx<float>{0.f, 0.f};
  // e^{+2 pi i j t / Y}, j = 1..MYT (the table holds the forward sign)
  cx<float> w[MYT];
#pragma unroll
  for (int j = 1; j <= MYT; ++j) w[j - 1] = conj(twtab[(j * t) & (Y - 1)]);
  const bool teven = (T & 1) == 0;
  const bool rowcopy = (mt & 1) == 0;
  const int NP = (T + 1) / 2, KL = NT / (NP > 0 ? NP : 1), pl = t % NP, kl = t / NP;
  const bool tiled = NP <= NT;
  const int rk1 = tiled ? (NKY + KL - 1) / KL : 0;  // rows per thread for a single pass of the t-axis product
  for (int i = threadIdx.x; i < mt * T; i += blockDim.x) Ss[(i / mt) * SS + i % mt] = Sy[i];
  const int stride = (int)gridDim.x * GP;
  const int iters = (nplanes + stride - 1) / stride;
  auto plane_of = [&](int it) { return (it * (int)gridDim.x + (int)blockIdx.x) * GP + g; };
  // staging of a spectrum block: even mt -- every thread copies 16-byte pieces into the padded rows (LDGSTS, completion
  // by wait_group + the group barrier); odd mt -- one contiguous bulk copy by thread 0 (mbarrier completion)
  const int upr = mt >> 1;  // 16-byte units per row (even mt)
  const int rows_per_it = (upr > 0 && NT % upr == 0) ? NT / upr : 0, row0 = upr > 0 ? t / upr : 0, wi0 = upr > 0 ? t % upr : 0;
  auto stage_block = [&](int it) {
    int p = plane_of(it);
    if (p >= nplanes) p = nplanes - 1;
    const cx<float>* src = Z2 + (size_t)p * NKY * mt;
    if (rowcopy) {
      if (rows_per_it > 0) {  // NT is a multiple of the units per row: (row, unit) advance without a division
        for (int row = row0; row < NKY; row += rows_per_it)
          ldgsts16(zin + (size_t)row * ZS * 8 + wi0 * 16, reinterpret_cast<const unsigned char*>(src) + ((size_t)row * upr + wi0) * 16);
      } else {
        for (int u = t; u < NKY * upr; u += NT) {
          const int row = u / upr, wi = u - row * upr;
          ldgsts16(zin + (size_t)row * ZS * 8 + wi * 16, reinterpret_cast<const unsigned char*>(src) + (size_t)u * 16);
        }
      }
      ldgsts_commit();
    } else if (t == 0) {
      stage_expect(bar, (unsigned)((size_t)NKY * mt * 8));
      bulk_load(zin, src, (unsigned)((size_t)NKY * mt * 8), bar);
    }
  };
  if (t == 0) stage_barrier_init(bar);
  __syncthreads();
  if (iters > 0) stage_block(0);

  for (int it = 0; it < iters; ++it) {
    const int plane = plane_of(it);
    const bool valid = plane < nplanes;
    if (rowcopy) {
      ldgsts_wait_all();
      sync();
    } else {
      tile_load_wait(bar, (unsigned)(it & 1));
    }
    const cx<float>* src = reinterpret_cast<const cx<float>*>(zin);
    if (tiled) {
      auto put = [&](int kyi, int tt, cx<float> v) { D[(size_t)kyi * XS + tt] = v; };
      if (rk1 == 5) group_cproduct<5>(Ss, SS, src, ZS, NKY, T, mt, pl, kl, KL, kl < KL, put);
      else if (rk1 == 6 || rk1 == 7) group_cproduct<7>(Ss, SS, src, ZS, NKY, T, mt, pl, kl, KL, kl < KL, put);
      else group_cproduct<4>(Ss, SS, src, ZS, NKY, T, mt, pl, kl, KL, kl < KL, put);
      if (T < TQ * 4)
        for (int kyi = t; kyi < NKY; kyi += NT)
          for (int tt = T; tt < TQ * 4; ++tt) D[(size_t)kyi * XS + tt] = cx<float>{0.f, 0.f};
    } else {
      for (int j = t; j < NKY * TQ * 4; j += NT) {
        const int kyi = j / (TQ * 4), tt = j % (TQ * 4);
        float sr = 0.f, si = 0.f;
        if (tt < T) {
          const cx<float>* zr = src + (size_t)kyi * ZS;
          const cx<float>* sy = Ss + (size_t)tt * SS;
          for (int kt = 0; kt < mt; ++kt) {
            const cx<float> a = sy[kt], v = zr[kt];
            sr = fmaf(a.x, v.x, sr); sr = fmaf(-a.y, v.y, sr);
            si = fmaf(a.x, v.y, si); si = fmaf(a.y, v.x, si);
          }
        }
        D[(size_t)kyi * XS + tt] = cx<float>{sr, si};
      }
    }
    if (t == 0) bulk_store_wait_read<0>();  // the previous plane's store has read the tile (first written after the next barrier)
    sync();  // D complete, zin consumed
    if (it + 1 < iters) stage_block(it + 1);  // the next block is staged under this plane's transforms
    for (int q = 0; q < TQ; ++q) {
      cx<f2>* Dh = Dh0;
      if (q > 0) sync();  // the previous quad's reads of Dh
      // Hermitian part along ky, one mirror pair (ky = p, Y - p) per thread: the two packed entries come from the same
      // eight loads (the mirror's time samples are the complex conjugates); a single pass over p = 0..my
      for (int p = t; p <= my; p += NT) {
        int i1 = kept_index(p, Y, my), i2 = kept_index((Y - p) % Y, Y, my);
        i1 = i1 >= 0 ? i1 : NKY;  // the zero row (ky = my is not a kept input)
        i2 = i2 >= 0 ? i2 : NKY;
        cx<float> h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const cx<float> a = D[(size_t)i1 * XS + 4 * q + j];
          const cx<float> b = D[(size_t)i2 * XS + 4 * q + j];
          h[j] = cx<float>{0.5f * (a.x + b.x), 0.5f * (a.y - b.y)};
        }
        // lane lo: h0 + i h2, lane hi: h1 + i h3 -- the transform then leaves (y[t0], y[t0+1]) and (y[t0+2], y[t0+3]) as its
        // packed real and imaginary parts, stored without repacking
        Dh[p] = cx<f2>{f2(h[0].x - h[2].y, h[1].x - h[3].y), f2(h[0].y + h[2].x, h[1].y + h[3].x)};
        if (p > 0)  // entry of ky = Y - p: conj(h0) + i conj(h2), conj(h1) + i conj(h3)
          Dh[2 * my + 1 - p] = cx<f2>{f2(h[0].x + h[2].y, h[1].x + h[3].y), f2(h[2].x - h[0].y, h[3].x - h[1].y)};
      }
      sync();
      cx<f2> z[8];
      pruned_inverse_y<MYT, EXACT>(Dh, my, w, z);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        float* r = tile + (t + m * NT) * T + 4 * q;
        if (teven) {
          *reinterpret_cast<f2*>(r) = z[m].x;
          if (4 * q + 2 < T) *reinterpret_cast<f2*>(r + 2) = z[m].y;
        } else {
          r[0] = z[m].x.lo;
          if (4 * q + 1 < T) r[1] = z[m].x.hi;
          if (4 * q + 2 < T) r[2] = z[m].y.lo;
          if (4 * q + 3 < T) r[3] = z[m].y.hi;
        }
      }
    }
    fence_async_smem();  // this thread's tile writes -> visible to the bulk store
    sync();              // also: every thread has finished reading Dh / D of this plane
    if (valid && t == 0) bulk_store(y + (size_t)plane * Y * T, tile, (unsigned)((size_t)Y * T * 4));
  }
  if (t == 0) bulk_store_wait_read<0>();  // shared memory must outlive the reads of the last store
}

// ------------------------------------------------------------------------------------------
// x-axis transforms on column pairs.  In: FWD  Z [bc][X][ncol]  ->  Xh [bc][2mx][ncol] (kept kx)
//                                      INV  Yh [bc][2mx][ncol] ->  Z  [bc][X][ncol]  (zero padded)
// CTA = GP groups of NT = X/8 threads; group g transforms column pair (blockIdx.x * GP + g) of slab
// blockIdx.y; the [X][GP] tile of 16-byte entries makes the strided side 128-byte rows.
template <int X>
struct XaxisSmem {
  static constexpr int NT = X / 8;
  static constexpr int GP = (256 / NT) > 8 ? 8 : ((256 / NT) > 0 ? (256 / NT) : 1);
  static constexpr int RS = GP + 1;  // padded tile row (entries)
  static constexpr size_t BYTES = (size_t)X * RS * 16 + (size_t)GP * X * 16;
};

template <int X, bool FWD>
__global__ void __launch_bounds__(XaxisSmem<X>::GP * (X / 8))
sconv_xaxis_kernel(const cx<float>* __restrict__ in, cx<float>* __restrict__ out,
                   const cx<float>* __restrict__ twtab, SconvDims d, int ncol) {
  typedef XaxisSmem<X> S;
  constexpr int NT = S::NT, GP = S::GP, RS = S::RS;
  const int mx = d.mx, NKX = 2 * mx;
  TCFD_DYN_SMEM(smem_raw);
  cx<f2>* tile = reinterpret_cast<cx<f2>*>(smem_raw);  // [X][RS]
  cx<f2>* bufs = tile + (size_t)X * RS;
  const int g = threadIdx.x / NT, t = threadIdx.x % NT;
  cx<f2>* buf = bufs + (size_t)g * X;
  FftTwiddles<float, X> tw;
  tw.load(twtab, t);
  CtaSyncS sync;
  int parity = 0;
  const int npairs = (ncol + 1) / 2;
  const int pair0 = blockIdx.x * GP;
  const size_t slab = blockIdx.y;
  const int pair = pair0 + g;
  const int c0 = 2 * pair, c1 = 2 * pair + 1;
  const bool v0 = pair < npairs, v1 = v0 && c1 < ncol;
  cx<f2> z[1][8];
  const cx<f2> zero{f2(0.f), f2(0.f)};

  if (FWD) {
    // cooperative load of the [X][GP pairs] tile (rows of up to GP*16 contiguous bytes)
    const cx<float>* src = in + slab * (size_t)X * ncol;
    for (int i = threadIdx.x; i < X * GP; i += GP * NT) {
      const int row = i / GP, gg = i % GP, ca = 2 * (pair0 + gg), cb = ca + 1;
      const cx<float> a = ca < ncol ? src[(size_t)row * ncol + ca] : cx<float>{0.f, 0.f};
      const cx<float> b = cb < ncol ? src[(size_t)row * ncol + cb] : cx<float>{0.f, 0.f};
      tile[row * RS + gg] = cx<f2>{f2(a.x, b.x), f2(a.y, b.y)};
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 8; ++m) z[0][m] = tile[(t + m * NT) * RS + g];
    fft_run<f2, X, -1, 1, false, X>(z, tw, buf, parity, t, sync);
    cx<float>* dst = out + slab * (size_t)NKX * ncol;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int kxi = kept_index(t + m * NT, X, mx);
      if (kxi >= 0) {
        if (v0) dst[(size_t)kxi * ncol + c0] = cx<float>{z[0][m].x.lo, z[0][m].y.lo};
        if (v1) dst[(size_t)kxi * ncol + c1] = cx<float>{z[0][m].x.hi, z[0][m].y.hi};
      }
    }
  } else {
    const cx<float>* src = in + slab * (size_t)NKX * ncol;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int kxi = kept_index(t + m * NT, X, mx);
      z[0][m] = zero;
      if (kxi >= 0) {
        const cx<float> a = v0 ? src[(size_t)kxi * ncol + c0] : cx<float>{0.f, 0.f};
        const cx<float> b = v1 ? src[(size_t)kxi * ncol + c1] : cx<float>{0.f, 0.f};
        z[0][m] = cx<f2>{f2(a.x, b.x), f2(a.y, b.y)};
      }
    }
    fft_run<f2, X, +1, 1, false, X>(z, tw, buf, parity, t, sync);
#pragma unroll
    for (int m = 0; m < 8; ++m) tile[(t + m * NT) * RS + g] = z[0][m];
    __syncthreads();
    cx<float>* dst = out + slab * (size_t)X * ncol;
    for (int i = threadIdx.x; i < X * GP; i += GP * NT) {
      const int row = i / GP, gg = i % GP, ca = 2 * (pair0 + gg), cb = ca + 1;
      const cx<f2> v = tile[row * RS + gg];
      if (ca < ncol) dst[(size_t)row * ncol + ca] = cx<float>{v.x.lo, v.y.lo};
      if (cb < ncol) dst[(size_t)row * ncol + cb] = cx<float>{v.x.hi, v.y.hi};
    }
  }
}

// ------------------------------------------------------------------------------------------
// Second-generation x-axis kernels: persistent CTAs (twiddles and tables are loaded once), the next [X][GP] tile is in
// flight (LDGSTS into the other tile buffer) while the current one is transformed.
//   fwd2: the tile just consumed doubles as the exchange buffer of its transforms (no separate buffers);
//   inv2: the 2 mx kept inputs of a column pair are broadcast-read by every thread of its group and the transform is the
//         pruned one (pruned_inverse, no exchanges); the output tile is written back in 128-byte rows.
template <int X>
struct Xaxis2Smem {
  static constexpr int NT = X / 8;
  static constexpr int GP = (256 / NT) > 8 ? 8 : ((256 / NT) > 0 ? (256 / NT) : 1);
  static constexpr int RS = GP + 1;  // padded tile row (16-byte entries)
  static constexpr size_t TILE = (size_t)X * RS * 16;
  static constexpr size_t FWD_BYTES = 2 * TILE;
  TCFD_HD static size_t kin_bytes(int mx) { return (size_t)2 * mx * GP * 16; }
  TCFD_HD static size_t inv_bytes(int mx) { return TILE + 2 * kin_bytes(mx); }
};

template <int X>
__global__ void __launch_bounds__(Xaxis2Smem<X>::GP * (X / 8), 3)
sconv_xaxis_fwd2_kernel(const cx<float>* __restrict__ in, cx<float>* __restrict__ out, const cx<float>* __restrict__ twtab,
                        SconvDims d, int ncol, int ntx, int ntiles) {
  typedef Xaxis2Smem<X> S;
  constexpr int NT = S::NT, GP = S::GP, RS = S::RS;
  const int mx = d.mx, NKX = 2 * mx;
  TCFD_DYN_SMEM(smem_raw);
  const int g = threadIdx.x / NT, t = threadIdx.x % NT;
  FftTwiddles<float, X> tw;
  tw.load(twtab, t);
  GroupSync<NT> sync{1 + g};
  int parity = 0;
  const int npairs = ncol / 2;  // ncol is even (2 my mt)
  // raw tile entry = the two adjacent columns of a pair as they lie in memory: (a.re, a.im, b.re, b.im)
  auto issue = [&](int tile_id, unsigned char* dstb) {
    const int pair0 = (tile_id % ntx) * GP;
    const cx<float>* src = in + (size_t)(tile_id / ntx) * X * ncol;
    for (int i = threadIdx.x; i < X * GP; i += GP * NT) {
      const int row = i / GP, gg = i % GP, pr = pair0 + gg;
      unsigned char* dst = dstb + ((size_t)row * RS + gg) * 16;
      if (pr < npairs) ldgsts16(dst, src + (size_t)row * ncol + 2 * pr);
      else *reinterpret_cast<cx<f2>*>(dst) = cx<f2>{f2(0.f), f2(0.f)};
    }
    ldgsts_commit();
  };
  int cur = 0;
  if ((int)blockIdx.x < ntiles) issue(blockIdx.x, smem_raw);
  for (int tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
    unsigned char* tb = smem_raw + (size_t)cur * S::TILE;
    ldgsts_wait_all();
    __syncthreads();  // the tile is complete; the other buffer (exchange buffer of the previous tile) is free
    const cx<f2>* tile = reinterpret_cast<const cx<f2>*>(tb);
    cx<f2> z[1][8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const cx<f2> r = tile[(t + m * NT) * RS + g];  // (a.re, a.im), (b.re, b.im)
      z[0][m] = cx<f2>{f2(r.x.lo, r.y.lo), f2(r.x.hi, r.y.hi)};
    }
    __syncthreads();  // every group holds its column: the tile becomes the exchange buffer
    const int nxt = tile_id + gridDim.x;
    if (nxt < ntiles) issue(nxt, smem_raw + (size_t)(cur ^ 1) * S::TILE);
    fft_run<f2, X, -1, 1, false, X>(z, tw, reinterpret_cast<cx<f2>*>(tb) + (size_t)g * X, parity, t, sync);
    const int pair = (tile_id % ntx) * GP + g;
    if (pair < npairs) {
      cx<float>* dst = out + (size_t)(tile_id / ntx) * NKX * ncol + 2 * pair;
      auto put = [&](int kxi, const cx<f2>& v) {  // both columns of the pair in one 16-byte store
        *reinterpret_cast<cx<f2>*>(dst + (size_t)kxi * ncol) = cx<f2>{f2(v.x.lo, v.y.lo), f2(v.x.hi, v.y.hi)};
      };
      if (mx <= NT) {  // (uniform) kept rows live in the first and the last register row only
        if (t < mx) put(t, z[0][0]);
        if (t >= NT - mx) put(2 * mx + t - NT, z[0][7]);
      } else {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int kxi = kept_index(t + m * NT, X, mx);
          if (kxi >= 0) put(kxi, z[0][m]);
        }
      }
    }
    cur ^= 1;
  }
}

template <int X, int MXT>
__global__ void __launch_bounds__(Xaxis2Smem<X>::GP * (X / 8))
sconv_xaxis_inv2_kernel(const cx<float>* __restrict__ in, cx<float>* __restrict__ out, const cx<float>* __restrict__ twtab,
                        SconvDims d, int ncol, int ntx, int ntiles) {
  typedef Xaxis2Smem<X> S;
  constexpr int NT = S::NT, GP = S::GP, RS = S::RS;
  const int mx = d.mx, NKX = 2 * mx;
  TCFD_DYN_SMEM(smem_raw);
  cx<f2>* tile = reinterpret_cast<cx<f2>*>(smem_raw);  // [X][RS], raw pairs
  unsigned char* kin0 = smem_raw + S::TILE;
  const int g = threadIdx.x / NT, t = threadIdx.x % NT;
  cx<float> w[MXT];
#pragma unroll
  for (int j = 1; j <= MXT; ++j) w[j - 1] = conj(twtab[(j * t) & (X - 1)]);
  const int npairs = ncol / 2;
  auto issue = [&](int tile_id, unsigned char* dstb) {  // the kept rows of the tile's GP column pairs: [NKX][GP] raw pairs
    const int pair0 = (tile_id % ntx) * GP;
    const cx<float>* src = in + (size_t)(tile_id / ntx) * NKX * ncol;
    for (int i = threadIdx.x; i < NKX * GP; i += GP * NT) {
      const int row = i / GP, gg = i % GP, pr = pair0 + gg;
      unsigned char* dst = dstb + (size_t)i * 16;
      if (pr < npairs) ldgsts16(dst, src + (size_t)row * ncol + 2 * pr);
      else *reinterpret_cast<cx<f2>*>(dst) = cx<f2>{f2(0.f), f2(0.f)};
    }
    ldgsts_commit();
  };
  int cur = 0;
  if ((int)blockIdx.x < ntiles) issue(blockIdx.x, kin0);
  for (int tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
    const cx<f2>* kin = reinterpret_cast<const cx<f2>*>(kin0 + (size_t)cur * S::kin_bytes(mx));
    ldgsts_wait_all();
    __syncthreads();  // inputs complete; the previous tile has been written out, the other input buffer is free
    const int nxt = tile_id + gridDim.x;
    if (nxt < ntiles) issue(nxt, kin0 + (size_t)(cur ^ 1) * S::kin_bytes(mx));
    cx<f2> z[8];
    pruned_inverse<MXT>(
        [&](int kk) {
          const cx<f2> r = kin[(kk >= 0 ? kk : NKX + kk) * GP + g];
          return cx<f2>{f2(r.x.lo, r.y.lo), f2(r.x.hi, r.y.hi)};
        },
        mx, mx, w, z);
#pragma unroll
    for (int m = 0; m < 8; ++m)
      tile[(t + m * NT) * RS + g] = cx<f2>{f2(z[m].x.lo, z[m].y.lo), f2(z[m].x.hi, z[m].y.hi)};  // back to raw pairs
    __syncthreads();
    const int pair0 = (tile_id % ntx) * GP;
    cx<float>* dst = out + (size_t)(tile_id / ntx) * X * ncol;
    for (int i = threadIdx.x; i < X * GP; i += GP * NT) {
      const int row = i / GP, gg = i % GP, pr = pair0 + gg;
      if (pr < npairs) *reinterpret_cast<cx<f2>*>(dst + (size_t)row * ncol + 2 * pr) = tile[row * RS + gg];
    }
    cur ^= 1;
  }
}

// ------------------------------------------------------------------------------------------
// mode mixing.  Modes are indexed k = (kxi * NKY + kyi) * mt + kt; corner = (kxi >= mx) + 2 (kyi >= my);
// weights of a corner: [Ci][Co][mx][my][mt] complex (the reference's parameter layout).
struct MixArgs {
  const cx<float>* w[4];
  const cx<float>* bias[4];  // [mx][my][mt] or null
  cx<float>* gw[4];
  cx<float>* gbias[4];
  int B, Ci, Co;
  float delta;
};

TCFD_HD void mode_split(int k, const SconvDims& d, int& corner, int& widx) {
  const int kt = k % d.mt, r = k / d.mt, kyi = r % (2 * d.my), kxi = r / (2 * d.my);
  const int cx_ = kxi >= d.mx, cy = kyi >= d.my;
  corner = cx_ + 2 * cy;
  widx = ((kxi - cx_ * d.mx) * d.my + (kyi - cy * d.my)) * d.mt + kt;
}
TCFD_D cx<float> cmul(cx<float> a, cx<float> b) {
  return cx<float>{fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x)};
}
TCFD_D cx<float> cmul_conj(cx<float> a, cx<float> b) {  // a * conj(b)
  return cx<float>{fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -(a.x * b.y))};
}

constexpr int MIX_BT = 8;  // batch tile held in registers

constexpr int MIX_OT = 4;  // output (forward) / input (backward) channels per thread

// Yh[b][o][k] = sum_i Xh[b][i][k] W[i][o][k] (+ delta bias[k]); thread = (k, tile of MIX_OT channels o),
// loop over b tiles: an Xh entry is read Co / MIX_OT times instead of Co times
__global__ void __launch_bounds__(128)
sconv_mix_fwd_kernel(const cx<float>* __restrict__ Xh, cx<float>* __restrict__ Yh, MixArgs a, SconvDims d) {
  const int K = 4 * d.mx * d.my * d.mt, msz = d.mx * d.my * d.mt;
  const int k = blockIdx.x * blockDim.x + threadIdx.x, o0 = blockIdx.y * MIX_OT;
  if (k >= K) return;
  int corner, widx;
  mode_split(k, d, corner, widx);
  const cx<float>* w = a.w[corner] + widx;
  cx<float> bias{0.f, 0.f};
  if (a.bias[corner]) {
    const cx<float> bb = a.bias[corner][widx];
    bias = cx<float>{a.delta * bb.x, a.delta * bb.y};
  }
  // batch tiles are spread over blockIdx.z (few modes x many samples would otherwise leave most SMs idle)
  for (int b0 = (int)blockIdx.z * MIX_BT; b0 < a.B; b0 += (int)gridDim.z * MIX_BT) {
    cx<float> acc[MIX_BT][MIX_OT];
#pragma unroll
    for (int j = 0; j < MIX_BT; ++j)
#pragma unroll
      for (int q = 0; q < MIX_OT; ++q) acc[j][q] = cx<float>{0.f, 0.f};
    for (int i = 0; i < a.Ci; ++i) {
      cx<float> wi[MIX_OT], xv[MIX_BT];
#pragma unroll
      for (int q = 0; q < MIX_OT; ++q)
        wi[q] = (o0 + q < a.Co) ? w[((size_t)i * a.Co + o0 + q) * msz] : cx<float>{0.f, 0.f};
#pragma unroll
      for (int j = 0; j < MIX_BT; ++j)
        xv[j] = (b0 + j < a.B) ? Xh[((size_t)(b0 + j) * a.Ci + i) * K + k] : cx<float>{0.f, 0.f};
#pragma unroll
      for (int j = 0; j < MIX_BT; ++j)
#pragma unroll
        for (int q = 0; q < MIX_OT; ++q) acc[j][q] = acc[j][q] + cmul(xv[j], wi[q]);
    }
#pragma unroll
    for (int j = 0; j < MIX_BT; ++j)
#pragma unroll
      for (int q = 0; q < MIX_OT; ++q)
        if (b0 + j < a.B && o0 + q < a.Co) Yh[((size_t)(b0 + j) * a.Co + o0 + q) * K + k] = acc[j][q] + bias;
  }
}

// gXh[b][i][k] = sum_o gYh[b][o][k] conj(W[i][o][k]); thread = (k, tile of MIX_OT channels i)
__global__ void __launch_bounds__(128)
sconv_mix_bwd_x_kernel(const cx<float>* __restrict__ gYh, cx<float>* __restrict__ gXh, MixArgs a, SconvDims d) {
  const int K = 4 * d.mx * d.my * d.mt, msz = d.mx * d.my * d.mt;
  const int k = blockIdx.x * blockDim.x + threadIdx.x, i0 = blockIdx.y * MIX_OT;
  if (k >= K) return;
  int corner, widx;
  mode_split(k, d, corner, widx);
  const cx<float>* w = a.w[corner] + widx;
  for (int b0 = (int)blockIdx.z * MIX_BT; b0 < a.B; b0 += (int)gridDim.z * MIX_BT) {
    cx<float> acc[MIX_BT][MIX_OT];
#pragma unroll
    for (int j = 0; j < MIX_BT; ++j)
#pragma unroll
      for (int q = 0; q < MIX_OT; ++q) acc[j][q] = cx<float>{0.f, 0.f};
    for (int o = 0; o < a.Co; ++o) {
      cx<float> wo[MIX_OT], gv[MIX_BT];
#pragma unroll
      for (int q = 0; q < MIX_OT; ++q)
        wo[q] = (i0 + q < a.Ci) ? w[((size_t)(i0 + q) * a.Co + o) * msz] : cx<float>{0.f, 0.f};
#pragma unroll
      for (int j = 0; j < MIX_BT; ++j)
        gv[j] = (b0 + j < a.B) ? gYh[((size_t)(b0 + j) * a.Co + o) * K + k] : cx<float>{0.f, 0.f};
#pragma unroll
      for (int j = 0; j < MIX_BT; ++j)
#pragma unroll
        for (int q = 0; q < MIX_OT; ++q) acc[j][q] = acc[j][q] + cmul_conj(gv[j], wo[q]);
    }
#pragma unroll
    for (int j = 0; j < MIX_BT; ++j)
#pragma unroll
      for (int q = 0; q < MIX_OT; ++q)
        if (b0 + j < a.B && i0 + q < a.Ci) gXh[((size_t)(b0 + j) * a.Ci + i0 + q) * K + k] = acc[j][q];
  }
}

// gW[i][o][k] = sum_b conj(Xh[b][i][k]) gYh[b][o][k]; thread = (k, o), blockIdx.z = i.
// gbias[k] = delta * sum_{b,o} gYh[b][o][k]  (computed by the i == 0, o == 0 threads)
__global__ void __launch_bounds__(128)
sconv_mix_bwd_w_kernel(const cx<float>* __restrict__ Xh, const cx<float>* __restrict__ gYh, MixArgs a, SconvDims d) {
  const int K = 4 * d.mx * d.my * d.mt, msz = d.mx * d.my * d.mt;
  const int k = blockIdx.x * blockDim.x + threadIdx.x, o = blockIdx.y, i = blockIdx.z;
  if (k >= K) return;
  int corner, widx;
  mode_split(k, d, corner, widx);
  cx<float> acc{0.f, 0.f};
  for (int b = 0; b < a.B; ++b) {
    const cx<float> xv = Xh[((size_t)b * a.Ci + i) * K + k];
    const cx<float> gv = gYh[((size_t)b * a.Co + o) * K + k];
    acc = acc + cmul_conj(gv, xv);  // g * conj(x)
  }
  a.gw[corner][((size_t)i * a.Co + o) * msz + widx] = acc;
  if (i == 0 && o == 0 && a.gbias[corner]) {
    cx<float> s{0.f, 0.f};
    for (int b = 0; b < a.B; ++b)
      for (int oo = 0; oo < a.Co; ++oo) s = s + gYh[((size_t)b * a.Co + oo) * K + k];
    a.gbias[corner][widx] = cx<float>{a.delta * s.x, a.delta * s.y};
  }
}

// ------------------------------------------------------------------------------------------
// Second-generation mode mixing.  The first-generation kernels above re-read every Xh entry once per output-channel
// tile and every weight once per batch tile through L2 (C4: 0.5 GB of L2 traffic for 0.17 GB of data, 12 warps per
// SM at 160 registers) and, in the weight gradient, every entry Ci (or Co) times (2.6 GB).  Here
//   mix2<BWD>   CTA = 32 consecutive modes x one warp per tile of MIX_OT output channels; the batch tile of the INPUT
//               spectrum (MIX_BT x P channels x 32 modes) is staged in shared memory once and shared by all warps;
//               BWD = false:  Yh[b][o] = sum_i Xh[b][i] W[i][o] (+ delta bias);   BWD = true:  gXh[b][i] = sum_o gYh[b][o] conj(W[i][o])
//   mix_bwd_w2  thread = (mode, tile of 4 i, tile of 4 o): 8 loads per 16 multiply-adds instead of 2 per 1.
template <bool BWD>
__global__ void __launch_bounds__(256)
sconv_mix2_kernel(const cx<float>* __restrict__ in, cx<float>* __restrict__ out, MixArgs a, SconvDims d) {
  const int K = 4 * d.mx * d.my * d.mt, msz = d.mx * d.my * d.mt;
  const int P = BWD ? a.Co : a.Ci, Q = BWD ? a.Ci : a.Co;  // input / output channels of this product
  TCFD_DYN_SMEM(smem_raw);
  cx<float>* xs = reinterpret_cast<cx<float>*>(smem_raw);  // [MIX_BT][P][32]
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5, q0 = wq * MIX_OT;
  const int k = blockIdx.x * 32 + lane;
  const bool kv = k < K;
  const int kc = kv ? k : K - 1;
  int corner, widx;
  mode_split(kc, d, corner, widx);
  const cx<float>* w = a.w[corner] + widx;
  cx<float> bias{0.f, 0.f};
  if (!BWD && a.bias[corner]) {
    const cx<float> bb = a.bias[corner][widx];
    bias = cx<float>{a.delta * bb.x, a.delta * bb.y};
  }
  for (int b0 = (int)blockIdx.y * MIX_BT; b0 < a.B; b0 += (int)gridDim.y * MIX_BT) {
    // stage the input tile: rows of 32 consecutive modes (256 contiguous bytes) as 16-byte asynchronous copies, all
    // in flight at once (K is even and the tile starts at a multiple of 32 modes: pairs never straddle the end)
    for (int e = threadIdx.x; e < MIX_BT * P * 16; e += blockDim.x) {
      const int l2 = e & 15, r = e >> 4, p = r % P, j = r / P;
      const int kk = blockIdx.x * 32 + 2 * l2;
      cx<float>* dst = xs + (size_t)r * 32 + 2 * l2;
      if (b0 + j < a.B && kk < K) ldgsts16(dst, in + ((size_t)(b0 + j) * P + p) * K + kk);
      else dst[0] = dst[1] = cx<float>{0.f, 0.f};
    }
    ldgsts_commit();
    ldgsts_wait_all();
    __syncthreads();
    if (q0 < Q) {
      cx<float> acc[MIX_BT][MIX_OT];
#pragma unroll
      for (int j = 0; j < MIX_BT; ++j)
#pragma unroll
        for (int q = 0; q < MIX_OT; ++q) acc[j][q] = cx<float>{0.f, 0.f};
#pragma unroll 4
      for (int p = 0; p < P; ++p) {
        cx<float> wv[MIX_OT], xv[MIX_BT];
#pragma unroll
        for (int q = 0; q < MIX_OT; ++q) {
          const int qq = q0 + q < Q ? q0 + q : Q - 1;
          wv[q] = BWD ? w[((size_t)qq * a.Co + p) * msz] : w[((size_t)p * a.Co + qq) * msz];
        }
#pragma unroll
        for (int j = 0; j < MIX_BT; ++j) xv[j] = xs[(j * P + p) * 32 + lane];
#pragma unroll
        for (int j = 0; j < MIX_BT; ++j)
#pragma unroll
          for (int q = 0; q < MIX_OT; ++q) acc[j][q] = acc[j][q] + (BWD ? cmul_conj(xv[j], wv[q]) : cmul(xv[j], wv[q]));
      }
      if (kv) {
#pragma unroll
        for (int j = 0; j < MIX_BT; ++j)
#pragma unroll
          for (int q = 0; q < MIX_OT; ++q)
            if (b0 + j < a.B && q0 + q < Q) out[((size_t)(b0 + j) * Q + q0 + q) * K + k] = acc[j][q] + bias;
      }
    }
    __syncthreads();  // the tile is re-used by the next batch tile
  }
}

constexpr int MIX_WT = 4;  // channel tile (both i and o) of the weight-gradient kernel

// gW[i][o][k] = sum_b conj(Xh[b][i][k]) gYh[b][o][k]; thread = (k, tile of MIX_WT o, tile of MIX_WT i)
// gbias[k] = delta * sum_{b,o} gYh[b][o][k]  (computed by the threads of the first tiles)
__global__ void __launch_bounds__(128)
sconv_mix_bwd_w2_kernel(const cx<float>* __restrict__ Xh, const cx<float>* __restrict__ gYh, MixArgs a, SconvDims d) {
  const int K = 4 * d.mx * d.my * d.mt, msz = d.mx * d.my * d.mt;
  const int k = blockIdx.x * blockDim.x + threadIdx.x, o0 = blockIdx.y * MIX_WT, i0 = blockIdx.z * MIX_WT;
  if (k >= K) return;
  int corner, widx;
  mode_split(k, d, corner, widx);
  cx<float> acc[MIX_WT][MIX_WT];
#pragma unroll
  for (int i = 0; i < MIX_WT; ++i)
#pragma unroll
    for (int o = 0; o < MIX_WT; ++o) acc[i][o] = cx<float>{0.f, 0.f};
#pragma unroll 2
  for (int b = 0; b < a.B; ++b) {
    cx<float> xv[MIX_WT], gv[MIX_WT];
#pragma unroll
    for (int i = 0; i < MIX_WT; ++i) xv[i] = Xh[((size_t)b * a.Ci + (i0 + i < a.Ci ? i0 + i : a.Ci - 1)) * K + k];
#pragma unroll
    for (int o = 0; o < MIX_WT; ++o) gv[o] = gYh[((size_t)b * a.Co + (o0 + o < a.Co ? o0 + o : a.Co - 1)) * K + k];
#pragma unroll
    for (int i = 0; i < MIX_WT; ++i)
#pragma unroll
      for (int o = 0; o < MIX_WT; ++o) acc[i][o] = acc[i][o] + cmul_conj(gv[o], xv[i]);  // g * conj(x)
  }
#pragma unroll
  for (int i = 0; i < MIX_WT; ++i)
#pragma unroll
    for (int o = 0; o < MIX_WT; ++o)
      if (i0 + i < a.Ci && o0 + o < a.Co) a.gw[corner][((size_t)(i0 + i) * a.Co + o0 + o) * msz + widx] = acc[i][o];
  if (i0 == 0 && o0 == 0 && a.gbias[corner]) {
    cx<float> s{0.f, 0.f};
    for (int b = 0; b < a.B; ++b)
      for (int oo = 0; oo < a.Co; ++oo) s = s + gYh[((size_t)b * a.Co + oo) * K + k];
    a.gbias[corner][widx] = cx<float>{a.delta * s.x, a.delta * s.y};
  }
}

}  // namespace tcfd
