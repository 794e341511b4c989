// Small in-register DFTs (forward, e^{-i theta}), natural-order in and out.
// Used by the warp-level Stockham passes of the fused frontend kernel.
#pragma once
#include <cuda_runtime.h>

namespace sfe {

typedef float2 cpx;

// Packed fp32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2): one instruction for both halves of a
// complex value.  ptxas folds half swaps, per-half negation and scalar broadcast into the operand
// modifiers, so a complex add is 1 issue slot and a complex multiply 2 (FMUL2 + FFMA2) instead of 2 and 4.
#define SFE_F2OP(name, op)                                                                               \
  __device__ __forceinline__ float2 name(float2 a, float2 b) {                                           \
    float2 r;                                                                                            \
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; " op                        \
        " rc, ra, rb; mov.b64 {%0, %1}, rc; }"                                                           \
        : "=f"(r.x), "=f"(r.y)                                                                           \
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));                                                       \
    return r;                                                                                            \
  }
SFE_F2OP(f2add, "add.rn.f32x2")
SFE_F2OP(f2sub, "sub.rn.f32x2")
SFE_F2OP(f2mul, "mul.rn.f32x2")
#undef SFE_F2OP
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}

__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return f2add(a, b); }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return f2sub(a, b); }
// (a.x*b.x - a.y*b.y, a.y*b.x + a.x*b.y): same roundings as mul + fma on each component
__device__ __forceinline__ cpx cmul(cpx a, cpx b) {
  return f2fma(make_float2(a.y, a.x), make_float2(-b.y, b.y), f2mul(a, make_float2(b.x, b.x)));
}
// a * (-i)
__device__ __forceinline__ cpx mul_mi(cpx a) { return make_float2(a.y, -a.x); }
// a * (+i)
__device__ __forceinline__ cpx mul_pi(cpx a) { return make_float2(-a.y, a.x); }
// a * h*(1-i)  and  a * h*(-1-i)   (h = sqrt(1/2)): the W8 / W16 odd eighth-turn twiddles
__device__ __forceinline__ cpx mul_w8_1(cpx a) {
  const float h = 0.70710678118654752440f;
  return f2mul(f2add(a, make_float2(a.y, -a.x)), make_float2(h, h));
}
__device__ __forceinline__ cpx mul_w8_3(cpx a) {
  const float h = 0.70710678118654752440f;
  return f2mul(f2sub(make_float2(a.y, -a.x), a), make_float2(h, h));
}

__device__ __forceinline__ void fft2(cpx& a, cpx& b) {
  cpx t = csub(a, b);
  a = cadd(a, b);
  b = t;
}

__device__ __forceinline__ void fft3(cpx& a0, cpx& a1, cpx& a2) {
  const float c = -0.5f, s = 0.86602540378443864676f;
  cpx t1 = cadd(a1, a2), t2 = csub(a1, a2);
  cpx m = f2fma(make_float2(c, c), t1, a0);
  cpx n = f2mul(make_float2(s, s), t2);  // X1 = m - i n, X2 = m + i n
  a0 = cadd(a0, t1);
  a1 = cadd(m, mul_mi(n));
  a2 = csub(m, mul_mi(n));
}

__device__ __forceinline__ void fft4(cpx& a0, cpx& a1, cpx& a2, cpx& a3) {
  cpx s0 = cadd(a0, a2), d0 = csub(a0, a2);
  cpx s1 = cadd(a1, a3), d1 = csub(a1, a3);
  cpx md1 = mul_mi(d1);
  a0 = cadd(s0, s1);
  a2 = csub(s0, s1);
  a1 = cadd(d0, md1);
  a3 = csub(d0, md1);
}

__device__ __forceinline__ void fft5(cpx& a0, cpx& a1, cpx& a2, cpx& a3, cpx& a4) {
  const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
  const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
  cpx t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
  cpx m1 = f2fma(make_float2(c2, c2), t2, f2fma(make_float2(c1, c1), t1, a0));
  cpx m2 = f2fma(make_float2(c1, c1), t2, f2fma(make_float2(c2, c2), t1, a0));
  cpx n1 = f2fma(make_float2(s2, s2), t4, f2mul(make_float2(s1, s1), t3));
  cpx n2 = f2fma(make_float2(-s1, -s1), t4, f2mul(make_float2(s2, s2), t3));
  a0 = cadd(a0, cadd(t1, t2));
  // X1 = m1 - i n1, X4 = m1 + i n1, X2 = m2 - i n2, X3 = m2 + i n2
  a1 = cadd(m1, mul_mi(n1));
  a4 = csub(m1, mul_mi(n1));
  a2 = cadd(m2, mul_mi(n2));
  a3 = csub(m2, mul_mi(n2));
}

__device__ __forceinline__ void fft8(cpx (&u)[8]) {
  // n = 2*n1 + n2, k = k1 + 4*k2
  fft4(u[0], u[2], u[4], u[6]);  // A[0][k1] -> u[0],u[2],u[4],u[6]
  fft4(u[1], u[3], u[5], u[7]);  // A[1][k1] -> u[1],u[3],u[5],u[7]
  // twiddle A[1][k1] *= W8^k1
  cpx b1 = mul_w8_1(u[3]);  // * (1-i)/sqrt2
  cpx b2 = mul_mi(u[5]);
  cpx b3 = mul_w8_3(u[7]);  // * (-1-i)/sqrt2
  cpx a0 = u[0], a1 = u[2], a2 = u[4], a3 = u[6], b0 = u[1];
  u[0] = cadd(a0, b0); u[4] = csub(a0, b0);
  u[1] = cadd(a1, b1); u[5] = csub(a1, b1);
  u[2] = cadd(a2, b2); u[6] = csub(a2, b2);
  u[3] = cadd(a3, b3); u[7] = csub(a3, b3);
}

__device__ __forceinline__ void fft16(cpx (&u)[16]) {
  // n = 4*n1 + n2, k = k1 + 4*k2
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) fft4(u[n2], u[4 + n2], u[8 + n2], u[12 + n2]);
  // now u[4*k1 + n2] = A[n2][k1]; multiply by W16^(n2*k1)
  u[5] = cmul(u[5], make_float2(c1, -s1));                              // m=1
  u[6] = mul_w8_1(u[6]);                                                // m=2
  u[7] = cmul(u[7], make_float2(s1, -c1));                              // m=3
  u[9] = mul_w8_1(u[9]);                                                // m=2
  u[10] = mul_mi(u[10]);                                                // m=4
  u[11] = mul_w8_3(u[11]);                                              // m=6
  u[13] = cmul(u[13], make_float2(s1, -c1));                            // m=3
  u[14] = mul_w8_3(u[14]);                                              // m=6
  u[15] = cmul(u[15], make_float2(-c1, s1));                            // m=9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) fft4(u[4 * k1 + 0], u[4 * k1 + 1], u[4 * k1 + 2], u[4 * k1 + 3]);
  // u[4*k1 + k2] = X[k1 + 4*k2] -> transpose to natural order
  cpx t;
#define SFE_SWAP(a, b) t = u[a]; u[a] = u[b]; u[b] = t;
  SFE_SWAP(1, 4) SFE_SWAP(2, 8) SFE_SWAP(3, 12) SFE_SWAP(6, 9) SFE_SWAP(7, 13) SFE_SWAP(11, 14)
#undef SFE_SWAP
}

template <int R>
__device__ __forceinline__ void bfly(cpx (&u)[R]);
template <>
__device__ __forceinline__ void bfly<2>(cpx (&u)[2]) { fft2(u[0], u[1]); }
template <>
__device__ __forceinline__ void bfly<3>(cpx (&u)[3]) { fft3(u[0], u[1], u[2]); }
template <>
__device__ __forceinline__ void bfly<4>(cpx (&u)[4]) { fft4(u[0], u[1], u[2], u[3]); }
template <>
__device__ __forceinline__ void bfly<5>(cpx (&u)[5]) { fft5(u[0], u[1], u[2], u[3], u[4]); }
template <>
__device__ __forceinline__ void bfly<8>(cpx (&u)[8]) { fft8(u); }
template <>
__device__ __forceinline__ void bfly<16>(cpx (&u)[16]) { fft16(u); }

}  // namespace sfe
