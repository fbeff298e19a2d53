// Two fp32 lanes processed in lock-step by Blackwell's packed FP32 instructions (FADD2 / FMUL2 /
// FFMA2, PTX add/mul/fma.rn.f32x2).  On sm_100a the FP32 pipe retires 128 lanes/clk/SM either way
// (scripts/microbench/fp32_pipe.cu), but a packed instruction takes ONE issue slot for two lanes,
// which leaves the other slot to the shared-memory, integer and global-memory instructions of the
// FFT.  Scalar second operands are broadcast for free (SASS `R.F32` operand), negations fold into
// operand modifiers.  Under TCFD_EMU the same type is two plain floats.
#pragma once
#include "tcfd_common.cuh"

namespace tcfd {

struct alignas(8) f2 {
  float lo, hi;
  f2() = default;
  TCFD_HD f2(float a, float b) : lo(a), hi(b) {}
  TCFD_HD explicit f2(float s) : lo(s), hi(s) {}
  TCFD_HD explicit f2(double s) : lo((float)s), hi((float)s) {}
};

#ifndef TCFD_EMU
typedef unsigned long long u64_t;
TCFD_D u64_t f2_pk(f2 a) {
  u64_t r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a.lo), "f"(a.hi));
  return r;
}
TCFD_D f2 f2_upk(u64_t v) {
  f2 r;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(r.lo), "=f"(r.hi) : "l"(v));
  return r;
}
TCFD_D f2 operator+(f2 a, f2 b) {
  u64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pk(a)), "l"(f2_pk(b)));
  return f2_upk(r);
}
TCFD_D f2 operator-(f2 a, f2 b) {
  u64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pk(a)), "l"(f2_pk(b)));
  return f2_upk(r);
}
TCFD_D f2 operator*(f2 a, f2 b) {
  u64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pk(a)), "l"(f2_pk(b)));
  return f2_upk(r);
}
TCFD_D f2 fma_rn(f2 a, f2 b, f2 c) {
  u64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_pk(a)), "l"(f2_pk(b)), "l"(f2_pk(c)));
  return f2_upk(r);
}
#else
TCFD_HD f2 operator+(f2 a, f2 b) { return f2(a.lo + b.lo, a.hi + b.hi); }
TCFD_HD f2 operator-(f2 a, f2 b) { return f2(a.lo - b.lo, a.hi - b.hi); }
TCFD_HD f2 operator*(f2 a, f2 b) { return f2(a.lo * b.lo, a.hi * b.hi); }
TCFD_HD f2 fma_rn(f2 a, f2 b, f2 c) { return f2(fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)); }
#endif
TCFD_HD f2 operator-(f2 a) { return f2(-a.lo, -a.hi); }
// scalar operands are broadcast to both lanes
TCFD_D f2 operator*(f2 a, float s) { return a * f2(s); }
TCFD_D f2 operator*(float s, f2 a) { return a * f2(s); }
TCFD_D f2 fma_rn(f2 a, float s, f2 c) { return fma_rn(a, f2(s), c); }

TCFD_D f2 operator/(f2 a, f2 b) { return f2(a.lo / b.lo, a.hi / b.hi); }  // IEEE div.rn per lane

// The fp64 counterpart: two doubles handled by scalar instructions (sm_100a has no packed fp64
// arithmetic); same interface, so the kernels are written once over a 2-lane type.
struct alignas(16) d2 {
  double lo, hi;
  d2() = default;
  TCFD_HD d2(double a, double b) : lo(a), hi(b) {}
  TCFD_HD explicit d2(double s) : lo(s), hi(s) {}
};
TCFD_HD d2 operator+(d2 a, d2 b) { return d2(a.lo + b.lo, a.hi + b.hi); }
TCFD_HD d2 operator-(d2 a, d2 b) { return d2(a.lo - b.lo, a.hi - b.hi); }
TCFD_HD d2 operator*(d2 a, d2 b) { return d2(a.lo * b.lo, a.hi * b.hi); }
TCFD_HD d2 operator/(d2 a, d2 b) { return d2(a.lo / b.lo, a.hi / b.hi); }
TCFD_HD d2 operator-(d2 a) { return d2(-a.lo, -a.hi); }
TCFD_HD d2 fma_rn(d2 a, d2 b, d2 c) { return d2(fma(a.lo, b.lo, c.lo), fma(a.hi, b.hi, c.hi)); }
TCFD_HD d2 operator*(d2 a, double s) { return a * d2(s); }
TCFD_HD d2 operator*(double s, d2 a) { return a * d2(s); }
TCFD_HD d2 fma_rn(d2 a, double s, d2 c) { return fma_rn(a, d2(s), c); }

// scalar type behind a lane type (tables, twiddles)
template <class T> struct lane_traits { typedef T scalar; static constexpr int width = 1; };
template <> struct lane_traits<f2> { typedef float scalar; static constexpr int width = 2; };
template <> struct lane_traits<d2> { typedef double scalar; static constexpr int width = 2; };
// the 2-lane type of a scalar type
template <class T> struct pack2;
template <> struct pack2<float> { typedef f2 type; };
template <> struct pack2<double> { typedef d2 type; };

}  // namespace tcfd
