// Small in-register DFTs (forward, e^{-i theta}), natural-order in and out.
// Used by the warp-level Stockham passes of the fused frontend kernel.
#pragma once
#include <cuda_runtime.h>

namespace sfe {

typedef float2 cpx;

__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cpx cmul(cpx a, cpx b) {
  return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.y, b.x, a.x * b.y));
}
// a * (-i)
__device__ __forceinline__ cpx mul_mi(cpx a) { return make_float2(a.y, -a.x); }
// a * (+i)
__device__ __forceinline__ cpx mul_pi(cpx a) { return make_float2(-a.y, a.x); }

__device__ __forceinline__ void fft2(cpx& a, cpx& b) {
  cpx t = csub(a, b);
  a = cadd(a, b);
  b = t;
}

__device__ __forceinline__ void fft3(cpx& a0, cpx& a1, cpx& a2) {
  const float c = -0.5f, s = 0.86602540378443864676f;
  cpx t1 = cadd(a1, a2), t2 = csub(a1, a2);
  cpx m = make_float2(fmaf(c, t1.x, a0.x), fmaf(c, t1.y, a0.y));
  cpx n = make_float2(s * t2.x, s * t2.y);  // X1 = m - i n, X2 = m + i n
  a0 = cadd(a0, t1);
  a1 = make_float2(m.x + n.y, m.y - n.x);
  a2 = make_float2(m.x - n.y, m.y + n.x);
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
  cpx m1 = make_float2(fmaf(c2, t2.x, fmaf(c1, t1.x, a0.x)), fmaf(c2, t2.y, fmaf(c1, t1.y, a0.y)));
  cpx m2 = make_float2(fmaf(c1, t2.x, fmaf(c2, t1.x, a0.x)), fmaf(c1, t2.y, fmaf(c2, t1.y, a0.y)));
  cpx n1 = make_float2(fmaf(s2, t4.x, s1 * t3.x), fmaf(s2, t4.y, s1 * t3.y));
  cpx n2 = make_float2(fmaf(-s1, t4.x, s2 * t3.x), fmaf(-s1, t4.y, s2 * t3.y));
  a0 = cadd(a0, cadd(t1, t2));
  // X1 = m1 - i n1, X4 = m1 + i n1, X2 = m2 - i n2, X3 = m2 + i n2
  a1 = make_float2(m1.x + n1.y, m1.y - n1.x);
  a4 = make_float2(m1.x - n1.y, m1.y + n1.x);
  a2 = make_float2(m2.x + n2.y, m2.y - n2.x);
  a3 = make_float2(m2.x - n2.y, m2.y + n2.x);
}

__device__ __forceinline__ void fft8(cpx (&u)[8]) {
  // n = 2*n1 + n2, k = k1 + 4*k2
  const float h = 0.70710678118654752440f;
  fft4(u[0], u[2], u[4], u[6]);  // A[0][k1] -> u[0],u[2],u[4],u[6]
  fft4(u[1], u[3], u[5], u[7]);  // A[1][k1] -> u[1],u[3],u[5],u[7]
  // twiddle A[1][k1] *= W8^k1
  cpx b1 = make_float2(h * (u[3].x + u[3].y), h * (u[3].y - u[3].x));   // * (1-i)/sqrt2
  cpx b2 = mul_mi(u[5]);
  cpx b3 = make_float2(h * (u[7].y - u[7].x), -h * (u[7].x + u[7].y));  // * (-1-i)/sqrt2
  cpx a0 = u[0], a1 = u[2], a2 = u[4], a3 = u[6], b0 = u[1];
  u[0] = cadd(a0, b0); u[4] = csub(a0, b0);
  u[1] = cadd(a1, b1); u[5] = csub(a1, b1);
  u[2] = cadd(a2, b2); u[6] = csub(a2, b2);
  u[3] = cadd(a3, b3); u[7] = csub(a3, b3);
}

__device__ __forceinline__ void fft16(cpx (&u)[16]) {
  // n = 4*n1 + n2, k = k1 + 4*k2
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
  const float h = 0.70710678118654752440f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) fft4(u[n2], u[4 + n2], u[8 + n2], u[12 + n2]);
  // now u[4*k1 + n2] = A[n2][k1]; multiply by W16^(n2*k1)
  u[5] = cmul(u[5], make_float2(c1, -s1));                              // m=1
  u[6] = make_float2(h * (u[6].x + u[6].y), h * (u[6].y - u[6].x));     // m=2
  u[7] = cmul(u[7], make_float2(s1, -c1));                              // m=3
  u[9] = make_float2(h * (u[9].x + u[9].y), h * (u[9].y - u[9].x));     // m=2
  u[10] = mul_mi(u[10]);                                                // m=4
  u[11] = make_float2(h * (u[11].y - u[11].x), -h * (u[11].x + u[11].y));  // m=6
  u[13] = cmul(u[13], make_float2(s1, -c1));                            // m=3
  u[14] = make_float2(h * (u[14].y - u[14].x), -h * (u[14].x + u[14].y));  // m=6
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
