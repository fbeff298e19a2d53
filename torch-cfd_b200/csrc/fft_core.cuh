// Register-resident Stockham FFT building block (sm_100a; also compiled by g++ under TCFD_EMU).
//
// One N-point complex FFT is owned by a "group" of NT = N/8 threads.  Thread t holds the eight
// elements  t + m*NT  (m = 0..7) in registers, before AND after the transform (self-sorting).
// The transform is a decimation-in-time Stockham sequence of radix-8 passes, closed by one
// radix-2/4 pass when log2(N) is not a multiple of 3.  Between passes the group exchanges data
// through a shared-memory buffer of N elements: each thread writes its butterfly outputs to the
// Stockham-permuted positions and reads back the strided set  t + m*NT.  XOR swizzles keep both
// sides of every exchange bank-conflict free (checked by tests/test_bank_conflicts.py).
// Twiddles come from a per-plan table w[j] = exp(-2 pi i j / N) (computed in double on the host)
// and are held in registers for the lifetime of the kernel, so the inner loop touches shared
// memory only for the exchanges.
#pragma once
#include "tcfd_common.cuh"
#include "packed.cuh"

namespace tcfd {

// ---------------------------------------------------------------- plan (compile-time)
TCFD_HD constexpr int fft_num_passes(int n) {
  int p = 0;
  while (n > 1) { n /= (n >= 8 ? 8 : n); ++p; }
  return p;
}
TCFD_HD constexpr int fft_pass_radix(int n, int p) {
  int r = 1;
  for (int i = 0; i <= p; ++i) { r = (n >= 8 ? 8 : n); n /= r; }
  return r;
}
TCFD_HD constexpr int fft_pass_ns(int n, int p) {  // product of the radices of the passes before p
  int ns = 1;
  for (int i = 0; i < p; ++i) { int r = (n >= 8 ? 8 : n); ns *= r; n /= r; }
  return ns;
}
TCFD_HD constexpr int fft_pass_ntw(int n, int p) {  // twiddle registers (complex) used by pass p
  if (p == 0) return 0;
  int r = fft_pass_radix(n, p);
  return r == 8 ? 7 : (8 / r) * (r - 1);
}
TCFD_HD constexpr int fft_tw_offset(int n, int p) {
  int o = 0;
  for (int i = 0; i < p; ++i) o += fft_pass_ntw(n, i);
  return o;
}
TCFD_HD constexpr int fft_num_tw(int n) { return fft_tw_offset(n, fft_num_passes(n)); }

// ---------------------------------------------------------------- swizzle of the exchange buffer
// Logical index i in [0, N) -> physical slot.  Only low bits are changed, using higher bits, so it
// is a bijection of [0, N) for every N >= 16.  NS is the stride of the pass that WRITES.
template <class T, int NS>
TCFD_HD int fft_swz(int i) {
  if (sizeof(T) == 4) {
    if (NS == 1) return i ^ ((i >> 4) & 7);
    if (NS == 8) return i ^ (((i >> 6) & 1) << 3);
    return i;
  } else {
    if (NS == 1) return i ^ ((i >> 3) & 7);
    return i;
  }
}

// ---------------------------------------------------------------- butterflies
// DIR = -1: forward (exp(-i...)), DIR = +1: inverse (exp(+i...)), un-normalised.
template <int DIR, class T>
TCFD_D cx<T> mul_i_dir(cx<T> a) {  // a * (DIR * i)
  return DIR < 0 ? cx<T>{a.y, -a.x} : cx<T>{-a.y, a.x};
}
template <int DIR, class T, class S>
TCFD_D cx<T> tw_mul(cx<T> a, cx<S> w) {  // a * w (forward) or a * conj(w) (inverse); w has forward sign
  // the library is compiled with -fmad=false: these are the only fused multiply-adds of the FFT,
  // so every kernel instantiation rounds identically
  if (DIR < 0) return cx<T>{fma_rn(a.x, w.x, -(a.y * w.y)), fma_rn(a.x, w.y, a.y * w.x)};
  return cx<T>{fma_rn(a.x, w.x, a.y * w.y), fma_rn(a.y, w.x, -(a.x * w.y))};
}

template <int DIR, class T>
TCFD_D void radix2(cx<T>& a, cx<T>& b) {
  cx<T> t = a - b;
  a = a + b;
  b = t;
}

template <int DIR, class T>
TCFD_D void radix4(cx<T>& x0, cx<T>& x1, cx<T>& x2, cx<T>& x3) {
  cx<T> t0 = x0 + x2, t1 = x0 - x2, t2 = x1 + x3, t3 = mul_i_dir<DIR>(x1 - x3);
  x0 = t0 + t2;
  x2 = t0 - t2;
  x1 = t1 + t3;
  x3 = t1 - t3;
}

template <int DIR, class T>
TCFD_D void radix8(cx<T> (&v)[8]) {
  const typename lane_traits<T>::scalar h = (typename lane_traits<T>::scalar)0.70710678118654752440;
  cx<T> b0 = v[0] + v[4], c0 = v[0] - v[4];
  cx<T> b1 = v[1] + v[5], c1 = v[1] - v[5];
  cx<T> b2 = v[2] + v[6], c2 = v[2] - v[6];
  cx<T> b3 = v[3] + v[7], c3 = v[3] - v[7];
  // c_j *= W8^j (direction DIR)
  if (DIR < 0) {
    c1 = cx<T>{(c1.x + c1.y) * h, (c1.y - c1.x) * h};
    c3 = cx<T>{(c3.y - c3.x) * h, -(c3.x + c3.y) * h};
  } else {
    c1 = cx<T>{(c1.x - c1.y) * h, (c1.x + c1.y) * h};
    c3 = cx<T>{-(c3.x + c3.y) * h, (c3.x - c3.y) * h};
  }
  c2 = mul_i_dir<DIR>(c2);
  radix4<DIR>(b0, b1, b2, b3);  // even outputs 0,2,4,6
  radix4<DIR>(c0, c1, c2, c3);  // odd outputs 1,3,5,7
  v[0] = b0; v[2] = b1; v[4] = b2; v[6] = b3;
  v[1] = c0; v[3] = c1; v[5] = c2; v[7] = c3;
}

// ---------------------------------------------------------------- twiddle registers
template <class T, int N>
struct FftTwiddles {  // T: the SCALAR type (float twiddles serve both float and f2 data)
  static constexpr int NTW = fft_num_tw(N) > 0 ? fft_num_tw(N) : 1;
  cx<T> w[NTW];

  template <int P>
  TCFD_D void load_pass(const cx<T>* __restrict__ table, int t) {
    constexpr int NT = N / 8;
    constexpr int r = fft_pass_radix(N, P), ns = fft_pass_ns(N, P), off = fft_tw_offset(N, P);
    if constexpr (r == 8) {
      const int base = (t % ns) * (N / (ns * 8));
#pragma unroll
      for (int m = 1; m < 8; ++m) w[off + m - 1] = table[base * m];
    } else {
#pragma unroll
      for (int i = 0; i < 8 / r; ++i) {
        const int b = t + i * NT;
#pragma unroll
        for (int j = 1; j < r; ++j) w[off + i * (r - 1) + (j - 1)] = table[b * j];
      }
    }
    if constexpr (P + 1 < fft_num_passes(N)) load_pass<P + 1>(table, t);
  }
  TCFD_D void load(const cx<T>* __restrict__ table, int t) {
    if constexpr (fft_num_passes(N) > 1) load_pass<1>(table, t);
  }
};

// ---------------------------------------------------------------- the transform
// V transforms are carried side by side by the same thread (shared twiddles and index math).
// buf: V*N elements of shared memory private to the group (transform j uses [j*N, (j+1)*N)).
// sync(): barrier over (at least) the group's threads.
// Every exchange is  write - sync - read;  a second sync protects the buffer before the next
// write unless the caller alternates between two buffers (PP = true: buf holds two halves HALF
// elements apart and `parity`, toggled at EVERY exchange of the kernel, selects one; all transforms
// of a kernel must use the same HALF.  A write to half P at exchange e only needs every thread to
// have finished reading half P at exchange e-2, which it did before the sync of exchange e-1).
template <class T, int N, int DIR, int V, bool PP, int HALF, int P, class Sync>
TCFD_D void fft_pass(cx<T> (&v)[V][8], const FftTwiddles<typename lane_traits<T>::scalar, N>& tw, cx<T>* buf,
                     int& parity, int t, Sync& sync) {
  constexpr int NT = N / 8;
  constexpr int NP = fft_num_passes(N);
  constexpr int r = fft_pass_radix(N, P), ns = fft_pass_ns(N, P), off = fft_tw_offset(N, P);
  if constexpr (r == 8) {
    if constexpr (P > 0) {
#pragma unroll
      for (int j = 0; j < V; ++j)
#pragma unroll
        for (int m = 1; m < 8; ++m) v[j][m] = tw_mul<DIR>(v[j][m], tw.w[off + m - 1]);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) radix8<DIR>(v[j]);
    if constexpr (P + 1 < NP) {
      cx<T>* b = buf + (PP ? parity * HALF : 0);
      if (PP) parity ^= 1;
      const int base = (t % ns) + 8 * ns * (t / ns);
#pragma unroll
      for (int j = 0; j < V; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) b[j * N + fft_swz<T, ns>(base + ns * k)] = v[j][k];
      sync();
#pragma unroll
      for (int j = 0; j < V; ++j)
#pragma unroll
        for (int m = 0; m < 8; ++m) v[j][m] = b[j * N + fft_swz<T, ns>(t + m * NT)];
      if (!PP) sync();
    }
  } else {
    // closing radix-2/4 pass: 8/r butterflies per thread, in place in the register file
    static_assert(P + 1 == NP, "small radix must be the last pass");
    constexpr int nb = 8 / r;
#pragma unroll
    for (int j = 0; j < V; ++j)
#pragma unroll
      for (int i = 0; i < nb; ++i) {
#pragma unroll
        for (int q = 1; q < r; ++q)
          v[j][i + q * nb] = tw_mul<DIR>(v[j][i + q * nb], tw.w[off + i * (r - 1) + (q - 1)]);
        if constexpr (r == 2) radix2<DIR>(v[j][i], v[j][i + nb]);
        if constexpr (r == 4) radix4<DIR>(v[j][i], v[j][i + nb], v[j][i + 2 * nb], v[j][i + 3 * nb]);
      }
  }
  if constexpr (P + 1 < NP) fft_pass<T, N, DIR, V, PP, HALF, P + 1>(v, tw, buf, parity, t, sync);
}

template <class T, int N, int DIR, int V, bool PP, int HALF, class Sync>
TCFD_D void fft_run(cx<T> (&v)[V][8], const FftTwiddles<typename lane_traits<T>::scalar, N>& tw, cx<T>* buf,
                    int& parity, int t, Sync& sync) {
  static_assert(HALF >= V * N, "ping-pong half too small");
  fft_pass<T, N, DIR, V, PP, HALF, 0>(v, tw, buf, parity, t, sync);
}

}  // namespace tcfd
