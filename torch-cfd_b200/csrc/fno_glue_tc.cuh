// FNO3d layer glue on the 5th-generation tensor cores (tcgen05, sm_100a only; not part of the host-emulation build).
//
//   y = act( W2 gelu(W1 c + b1) + b2 + Ww x + bw )        (reference: fno/fno3d.py:223-230, MLP :119-130)
//
// Per point this is three C x C products (C = width <= 32) around two GELUs: on the CUDA cores the products alone
// are 3 C^2 FMAs per point -- with C = 20 as many FP32 lane-cycles as the whole HBM budget of the point (8 C bytes)
// -- and the kernel ran at 22 % of the HBM roofline (round 1).  Here a tile of 128 consecutive points is ONE UMMA
// M-block:
//   * thread t of the 128-thread CTA owns point t of the tile = TMEM lane t.  It loads the point's C channel values
//     of c and x (a warp reads 128 contiguous bytes per channel plane), splits each into a TF32-exact high part and
//     the fp32 remainder, and stores both into TMEM as the A operand of the products (tcgen05.st.32x32b: a thread
//     writes its own lane -- the operand layout needs no shuffle and no shared memory);
//   * the weights (B operands, K-major, split the same way on the host) sit in shared memory in the canonical
//     no-swizzle core-matrix layout for the life of the persistent CTA;
//   * one elected thread issues tcgen05.mma kind::tf32 with A from TMEM: D = A_hi W_hi + A_lo W_hi + A_hi W_lo
//     ("3xTF32": the dropped A_lo W_lo term is 2^-22 relative) accumulated in fp32 in TMEM; completion arrives on
//     an mbarrier (tcgen05.commit);
//   * every thread reads its accumulator row back (tcgen05.ld), adds the bias, applies the erf-form GELU (the only
//     arithmetic left on the CUDA cores), and either feeds the result back as the A operand of the second product
//     or writes the output channel planes.
// TMEM budget per CTA: 32 accumulator columns + 2 x (2 KP) operand columns (KP = C padded to 8) = 128 for C <= 24,
// 256 for C <= 32.  Resident CTAs per SM: 3 for C <= 24 (80 registers, no spills: measured 1.37 ms per C5 layer
// against 1.66 ms with 4 CTAs at 64 registers and 1.51 ms with 2), 2 above.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace tcfd {
namespace gluetc {

#ifndef TCFD_GLUE_MINB
#define TCFD_GLUE_MINB 3
#endif
constexpr int NP = 32;  // accumulator columns = UMMA N (outputs padded to 32)

template <int KP>
struct TcWeights {
  // shared-memory images of the six B operands (hi / lo parts of W1, W2, Ww), canonical K-major no-swizzle layout:
  // [k chunk of 4][n group of 8][8 rows][4 elements]  ->  element (n, k) at 128 * ((k/4) * 4 + n/8) + 16 * (n%8) + 4 * (k%4)
  float img[6][KP / 4][NP / 8][8][4];
  float b1[NP], b2[NP];  // b2 holds b2 + bw
};

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(unsigned* slot, unsigned ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_st8(unsigned taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, float (&v)[8]) {
  unsigned r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32, M = 128, N = 32, K = 8
__device__ __forceinline__ void umma_tf32_ts(unsigned d_tmem, unsigned a_tmem, unsigned long long b_desc, unsigned idesc,
                                             unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init1(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Wait for the phase.  The retry loop is four instructions (try_wait with a suspend-time hint, branch, count, compare):
// the first version re-read the SM clock on every retry and spent ~ 18 % of the kernel's issue slots spinning (ncu
// source view).  Bounded by the retry count (each retry suspends the warp for up to the hint) so a lost completion
// traps instead of hanging the GPU.
__device__ __forceinline__ bool mbar_try_wait(unsigned addr, unsigned parity) {
  unsigned done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(addr), "r"(parity), "r"(2000u)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait_parity(unsigned long long* bar, unsigned parity) {
  const unsigned addr = smem_addr(bar);
  if (mbar_try_wait(addr, parity)) return;
  for (unsigned spin = 0; !mbar_try_wait(addr, parity); ++spin)
    if (spin > (1u << 22)) asm volatile("trap;");
}

// shared-memory matrix descriptor, no swizzle: start address, leading-dimension (K direction) and stride (N direction)
// byte offsets between 8 x 16-byte core matrices, descriptor version 1 (sm_100)
__device__ __forceinline__ unsigned long long make_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
  unsigned long long d = 0;
  d |= (unsigned long long)((saddr & 0x3FFFFu) >> 4);
  d |= (unsigned long long)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (unsigned long long)((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}

// GELU, erf form, with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7 absolute: below the fp32 rounding of
// 1 + erf): one reciprocal, one exponential, six fused multiply-adds instead of the ~30 instructions of erff --
// the GELUs are the only arithmetic this kernel leaves on the CUDA cores, 2 C of them per point.
//   gelu(v) = v/2 (1 + erf(v / sqrt 2)) = v/2 + |v|/2 (1 - poly(t) exp(-v^2/2)),  t = 1 / (1 + p |v| / sqrt 2)
// The function returns 2 gelu(v) = v + |v| erf(|v| / sqrt 2).
__device__ __forceinline__ float gelu2_erf_tc(float v) {
  const float av = fabsf(v);
  float t, ex;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(av, 0.3275911f * 0.70710678118654752440f, 1.0f)));  // argument in [1, inf)
  float pl = fmaf(t, 1.061405429f, -1.453152027f);
  pl = fmaf(pl, t, 1.421413741f);
  pl = fmaf(pl, t, -0.284496736f);
  pl = fmaf(pl, t, 0.254829592f);
  pl *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(av * av * (-0.5f * 1.4426950408889634f)));
  const float e = fmaf(-pl, ex, 1.0f);  // erf(|v| / sqrt 2)
  return fmaf(av, e, v);                // = 2 gelu(v): the caller folds the factor 1/2 (into W2, or one multiply)
}

// n consecutive 32-bit columns of this thread's lane, in pieces of 8 / 4 / 2 / 1
template <int N>
__device__ __forceinline__ void tmem_st(unsigned taddr, const float* v) {
  if constexpr (N >= 8) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
    tmem_st<N - 8>(taddr + 8, v + 8);
  } else if constexpr (N >= 4) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
                 : "memory");
    tmem_st<N - 4>(taddr + 4, v + 4);
  } else if constexpr (N >= 2) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1]))
                 : "memory");
    tmem_st<N - 2>(taddr + 2, v + 2);
  } else if constexpr (N == 1) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v[0])) : "memory");
  }
}
template <int N>
__device__ __forceinline__ void tmem_ld_issue(unsigned taddr, unsigned* r) {
  if constexpr (N >= 8) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    tmem_ld_issue<N - 8>(taddr + 8, r + 8);
  } else if constexpr (N >= 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
    tmem_ld_issue<N - 4>(taddr + 4, r + 4);
  } else if constexpr (N >= 2) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
    tmem_ld_issue<N - 2>(taddr + 2, r + 2);
  } else if constexpr (N == 1) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr) : "memory");
  }
}
template <int N>
__device__ __forceinline__ void tmem_ld(unsigned taddr, float* v) {
  unsigned r[N];
  tmem_ld_issue<N>(taddr, r);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = __uint_as_float(r[i]);
}

// C (even, compile time): channels; KP = C padded to a multiple of 8 (K steps of the tf32 UMMA).  CTA = 256 threads =
// TWO threads per point of the 128-point tile: warps w and w + 4 share TMEM lanes 32 (w % 4) .. + 31 and each takes
// CH = C / 2 channels (operand columns in, accumulator columns out), so a tile's GELUs and its loads / stores are
// spread over twice the warps for the same TMEM footprint.  The next tile's channel values are requested before this
// tile's products are waited for.  TAIL: the number of points per sample is not a multiple of 128 (predicated
// accesses).  SWAP: exchange the two descriptor strides (bring-up knob: gives garbage).
template <int C, bool TAIL, bool SWAP>
__global__ void __launch_bounds__(256, (C <= 24 ? TCFD_GLUE_MINB : 2))
fno_layer_glue_tc_kernel(const float* __restrict__ c, const float* __restrict__ x, float* __restrict__ y,
                         const __grid_constant__ TcWeights<(C + 7) / 8 * 8> W, int act, size_t npts, size_t tiles_per_sample,
                         size_t ntiles) {
  constexpr int KP = (C + 7) / 8 * 8;
  constexpr int KC = KP / 4;                       // 16-byte chunks along K
  constexpr int CH = C / 2;                        // channels per thread
  constexpr unsigned TCOLS = (NP + 4 * KP) <= 128 ? 128 : 256;
  constexpr unsigned COL_D = 0, COL_A0 = NP, COL_AX = NP + 2 * KP;  // A0: c (then gelu(h)) hi | lo; AX: x hi | lo
  constexpr unsigned IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(NP >> 3) << 17) | ((128u >> 4) << 24);
  constexpr unsigned MAT_BYTES = KC * (NP / 8) * 128;
  constexpr unsigned LBO = SWAP ? 128u : (NP / 8) * 128u, SBO = SWAP ? (NP / 8) * 128u : 128u;
  static_assert(C % 2 == 0 && C >= 2 && C <= 32, "even channel counts up to 32");
  __shared__ __align__(128) float bimg[6 * KC * (NP / 8) * 32];
  __shared__ float b1s[NP], b2s[NP];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ unsigned tmem_slot;
  // (the shuffle tells the compiler the warp index is warp-uniform: TMEM addresses are then formed on the uniform
  // datapath instead of per lane + R2UR)
  const int t = threadIdx.x, warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int half = warp >> 2, pt = (warp & 3) * 32 + (t & 31);  // channel half, point of the tile
  const int ch0 = half * CH;
  {
    const float* src = &W.img[0][0][0][0][0];
    for (int i = t; i < 6 * KC * (NP / 8) * 32; i += 256) bimg[i] = src[i];
    if (t < NP) { b1s[t] = W.b1[t]; b2s[t] = W.b2[t]; }
  }
  if (warp == 0) tmem_alloc(&tmem_slot, TCOLS);
  if (t == 0) mbar_init1(&bar);
  // B images were written with generic stores and are read by the tensor core through the async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before();
  __syncthreads();
  fence_after();
  const unsigned tbase = __shfl_sync(0xffffffffu, tmem_slot, 0);
  const unsigned trow = tbase + ((unsigned)((warp & 3) * 32) << 16);  // this warp's 32 lanes
  const unsigned bbase = smem_addr(bimg);
  unsigned parity = 0;
  // KP > C: the K padding columns of the four operand blocks are written once for the life of the CTA -- zero, except the
  // first padding column of the hi block of operand A0, which is ONE: the matching rows of the W1 / W2 images hold the
  // biases (b1; b2 + bw), so the tensor core adds them (no per-channel bias registers or adds on the CUDA cores)
  constexpr bool BIAS_MMA = KP > C;
  if constexpr (KP > C) {
    if (half == 0) {
      float z[KP - C] = {};
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        z[0] = blk == 0 ? 1.0f : 0.0f;
        tmem_st<KP - C>(trow + COL_A0 + blk * KP + C, z);
      }
    }
  }
  float b1r[BIAS_MMA ? 1 : CH], b2r[BIAS_MMA ? 1 : CH];
  if constexpr (!BIAS_MMA) {
#pragma unroll
    for (int j = 0; j < CH; ++j) { b1r[j] = b1s[ch0 + j]; b2r[j] = b2s[ch0 + j]; }
  }

  auto split_store = [&](const float (&v)[CH], unsigned col) {  // hi -> col + ch0 .., lo -> col + KP + ch0 ..
    float hi[CH], lo[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      hi[j] = __uint_as_float(__float_as_uint(v[j]) & 0xFFFFE000u);  // exactly representable in TF32
      lo[j] = v[j] - hi[j];                                           // exact in fp32
    }
    tmem_st<CH>(trow + col + ch0, hi);
    tmem_st<CH>(trow + col + KP + ch0, lo);
  };
  // one product: D (+)= A_hi B_hi + A_lo B_hi + A_hi B_lo, A at TMEM columns acol (hi) / acol + KP (lo), B images m (hi), m + 1 (lo)
  auto issue_product = [&](unsigned acol, int m, bool first) {
    unsigned acc = first ? 0u : 1u;
#pragma unroll
    for (int part = 0; part < 3; ++part) {
      const unsigned a = tbase + acol + (part == 1 ? KP : 0);
      const unsigned b = bbase + (unsigned)(m + (part == 2 ? 1 : 0)) * MAT_BYTES;
#pragma unroll
      for (int s = 0; s < KP / 8; ++s) {
        umma_tf32_ts(tbase + COL_D, a + 8 * s, make_desc(b + (unsigned)(2 * s) * (NP / 8) * 128u, LBO, SBO), IDESC, acc);
        acc = 1u;
      }
    }
  };
  // element offset of (this thread's first channel, its point) of tile (sample bq, tile qt of the sample); valid: the
  // point exists.  The walk over tiles is incremental (tile += gridDim.x): a 64-bit division per tile cost ~ 10 % of
  // the kernel's instructions (ncu source view, round 2).
  const size_t step = gridDim.x;
  size_t bq = blockIdx.x / tiles_per_sample, qt = blockIdx.x % tiles_per_sample;  // of the tile being PREFETCHED
  auto locate = [&](size_t tile, size_t& off, bool& valid) {
    const size_t q = qt * 128 + pt;
    valid = tile < ntiles && (!TAIL || q < npts);
    // without a ragged tail every thread of an existing tile has a point: a prefetch past the last tile simply re-reads
    // the current one (no predicated loads, no zero fill)
    if (TAIL || tile < ntiles) off = (bq * C + ch0) * npts + q;
    qt += step;
    while (qt >= tiles_per_sample) { qt -= tiles_per_sample; ++bq; }
  };
  // channel planes are npts elements apart: 32-bit element offsets i * npts from the tile's 64-bit base (the host
  // guarantees 16 * npts < 2^31), one IMAD.WIDE per access
  const unsigned stride = (unsigned)npts;
  auto load_tile = [&](size_t off, bool valid, float (&cv)[CH], float (&xv)[CH]) {
    const float* cp = c + off;
    const float* xp = x + off;
    if constexpr (TAIL) {
#pragma unroll
      for (int i = 0; i < CH; ++i) cv[i] = valid ? __ldcs(cp + (unsigned)i * stride) : 0.f;
#pragma unroll
      for (int i = 0; i < CH; ++i) xv[i] = valid ? __ldcs(xp + (unsigned)i * stride) : 0.f;
    } else {  // running pointers: one 64-bit add per access instead of a multiply and an add
#pragma unroll
      for (int i = 0; i < CH; ++i) { cv[i] = __ldcs(cp); cp += stride; }
#pragma unroll
      for (int i = 0; i < CH; ++i) { xv[i] = __ldcs(xp); xp += stride; }
    }
  };

  float cv[CH], xv[CH];
  size_t off = ((size_t)0 * C + ch0) * npts + pt;  // (a CTA beyond the last tile never enters the loop; any valid address)
  bool valid;
  locate(blockIdx.x, off, valid);
  load_tile(off, valid, cv, xv);
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t cur_off = off;
    const bool cur_valid = valid;
    split_store(cv, COL_A0);
    split_store(xv, COL_AX);
    tmem_wait_st();
    fence_before();
    __syncthreads();
    if (t == 0) {
      fence_after();
      issue_product(COL_A0, 0, true);  // h = W1 c
      umma_commit(&bar);
    }
    locate(tile + gridDim.x, off, valid);  // the next tile's values land under this tile's products and GELUs
    load_tile(off, valid, cv, xv);
    mbar_wait_parity(&bar, parity);
    parity ^= 1u;
    fence_after();
    float g[CH];
    tmem_ld<CH>(trow + COL_D + ch0, g);
#pragma unroll
    for (int j = 0; j < CH; ++j) g[j] = gelu2_erf_tc(BIAS_MMA ? g[j] : g[j] + b1r[j]);  // 2 gelu(h): W2's image carries the 1/2
    split_store(g, COL_A0);  // the first product is complete: its operand columns are free
    tmem_wait_st();
    fence_before();
    __syncthreads();
    if (t == 0) {
      fence_after();
      issue_product(COL_A0, 2, true);   // W2 gelu(h)
      issue_product(COL_AX, 4, false);  // + Ww x
      umma_commit(&bar);
    }
    mbar_wait_parity(&bar, parity);
    parity ^= 1u;
    fence_after();
    float d[CH];
    tmem_ld<CH>(trow + COL_D + ch0, d);
    if (!TAIL || cur_valid) {
      float* yp = y + cur_off;
      if (act) {  // (kernel-uniform: one branch per tile, not one select per channel)
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          __stcs(yp, 0.5f * gelu2_erf_tc(BIAS_MMA ? d[j] : d[j] + b2r[j]));
          yp += stride;
        }
      } else {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          __stcs(yp, BIAS_MMA ? d[j] : d[j] + b2r[j]);
          yp += stride;
        }
      }
    }
    // (the next tile's operand stores and products are ordered after this tile's tcgen05.ld by the barrier that
    // follows its split_store: every thread's loads above completed -- tcgen05.wait::ld -- before it gets there)
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, TCOLS);
}

template <int KP>
TcWeights<KP> pack_tc(const float* w1, const float* b1, const float* w2, const float* b2, const float* ww, const float* bw, int C) {
  TcWeights<KP> W;
  const float* mats[3] = {w1, w2, ww};
  for (int m = 0; m < 3; ++m)
    for (int n = 0; n < NP; ++n)
      for (int k = 0; k < KP; ++k) {
        // torch Conv3d layout [out][in] = B[n][k]; W2 (m = 1) is halved: the kernel feeds it 2 gelu(h) (exact scaling)
        const float v = (n < C && k < C) ? (m == 1 ? 0.5f : 1.0f) * mats[m][n * C + k] : 0.f;
        uint32_t bits;
        memcpy(&bits, &v, 4);
        bits &= 0xFFFFE000u;
        float hi;
        memcpy(&hi, &bits, 4);
        W.img[2 * m][k / 4][n / 8][n % 8][k % 4] = hi;
        W.img[2 * m + 1][k / 4][n / 8][n % 8][k % 4] = v - hi;
      }
  for (int n = 0; n < NP; ++n) {
    W.b1[n] = (b1 && n < C) ? b1[n] : 0.f;
    W.b2[n] = n < C ? (b2 ? b2[n] : 0.f) + (bw ? bw[n] : 0.f) : 0.f;
  }
  if (KP > C) {  // the kernel's operand A0 carries a constant 1 in K column C: that column of W1 / W2 is the bias
    for (int m = 0; m < 2; ++m)
      for (int n = 0; n < NP; ++n) {
        const float v = m == 0 ? W.b1[n] : W.b2[n];
        uint32_t bits;
        memcpy(&bits, &v, 4);
        bits &= 0xFFFFE000u;
        float hi;
        memcpy(&hi, &bits, 4);
        W.img[2 * m][C / 4][n / 8][n % 8][C % 4] = hi;
        W.img[2 * m + 1][C / 4][n / 8][n % 8][C % 4] = v - hi;
      }
  }
  return W;
}

}  // namespace gluetc
}  // namespace tcfd
