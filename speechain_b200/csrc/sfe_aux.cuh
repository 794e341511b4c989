// Light kernels around the fused frontend: statistics, normalise (+warp +mask), delta, standalone
// mel projection.  All HBM-bound streaming passes with coalesced, vectorised accesses.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sfe_fused.cuh"

namespace sfe {

// ---------------------------------------------------------------------------------------------
// per-utterance sums over valid frames of an arbitrary (B, F, D) tensor      feat_norm.py:115-133
// grid (chunks, B); each CTA reduces FR frames of one utterance, then fp64 atomics
// ---------------------------------------------------------------------------------------------
constexpr int STATS_FR = 64;

__global__ void __launch_bounds__(256) stats_accumulate_kernel(const float* __restrict__ feat,
                                                               const long long* __restrict__ flen, long long F, int D,
                                                               double* __restrict__ stats) {
  const int b = blockIdx.y;
  const long long n = min(flen[b], F);
  const long long t0 = static_cast<long long>(blockIdx.x) * STATS_FR;
  if (t0 >= n) return;
  const long long t1 = min(n, t0 + STATS_FR);
  const float* src = feat + (static_cast<long long>(b) * F) * D;
  // thread -> (row group, dim): consecutive threads walk consecutive dims (coalesced)
  extern __shared__ double sred[];  // [rows][D][2]
  const int rows = max(1, 256 / D);
  const int r = threadIdx.x / D, d = threadIdx.x % D;
  double s1 = 0.0, s2 = 0.0;
  if (r < rows) {
    for (long long t = t0 + r; t < t1; t += rows) {
      const double x = static_cast<double>(__ldg(src + t * D + d));
      s1 += x;
      s2 = fma(x, x, s2);
    }
    sred[(r * D + d) * 2] = s1;
    sred[(r * D + d) * 2 + 1] = s2;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) {
    double s = 0.0;
    for (int q = 0; q < rows; ++q) s += sred[q * D * 2 + i];
    atomicAdd(stats + static_cast<long long>(b) * D * 2 + i, s);
  }
}

// wide-D variant (D > 256): one thread per dim, loop
__global__ void __launch_bounds__(256) stats_accumulate_wide_kernel(const float* __restrict__ feat,
                                                                    const long long* __restrict__ flen, long long F,
                                                                    int D, double* __restrict__ stats) {
  const int b = blockIdx.y;
  const long long n = min(flen[b], F);
  const long long t0 = static_cast<long long>(blockIdx.x) * STATS_FR;
  if (t0 >= n) return;
  const long long t1 = min(n, t0 + STATS_FR);
  const float* src = feat + (static_cast<long long>(b) * F) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double s1 = 0.0, s2 = 0.0;
    for (long long t = t0; t < t1; ++t) {
      const double x = static_cast<double>(__ldg(src + t * D + d));
      s1 += x;
      s2 = fma(x, x, s2);
    }
    atomicAdd(stats + (static_cast<long long>(b) * D + d) * 2, s1);
    atomicAdd(stats + (static_cast<long long>(b) * D + d) * 2 + 1, s2);
  }
}

// sums -> (mean, unbiased std clamped).  Two encodings of the (B, D, 2) sums buffer: fp64 (written by
// stats_accumulate_kernel) or the fused kernel's fixed point (sum x * 2^32, sum x^2 * 2^24).
__device__ __forceinline__ void load_moments(const void* stats, int fixed, long long idx, double& s1, double& s2) {
  if (fixed) {
    const long long* q = reinterpret_cast<const long long*>(stats) + idx * 2;
    s1 = static_cast<double>(q[0]) * (1.0 / 4294967296.0);
    s2 = static_cast<double>(q[1]) * (1.0 / 16777216.0);
  } else {
    const double* q = reinterpret_cast<const double*>(stats) + idx * 2;
    s1 = q[0];
    s2 = q[1];
  }
}

__device__ __forceinline__ void moments_to_mean_std(double s1, double s2, long long n, float clamp, float& mean,
                                                    float& sd) {
  const double dn = static_cast<double>(n);
  const double m = s1 / dn;
  // unbiased variance; n == 1 gives NaN like torch.std — stated explicitly: the
  // fixed-point sums leave a rounding residual in the numerator, and residual / 0 would be +-inf, not NaN
  double var = (n == 1) ? __longlong_as_double(0x7ff8000000000000LL) : (s2 - s1 * m) / (dn - 1.0);
  if (var < 0.0) var = 0.0;
  float s = static_cast<float>(sqrt(var));
  if (s < clamp) s = clamp;  // NaN stays NaN, like torch.clamp
  mean = static_cast<float>(m);
  sd = s;
}

// mean/std (B, D) and packed [sum mean | sum std | B]                      feat_norm.py:115-133,239-266
// A CTA owns Dc columns and cuts the utterances into S (<= blockDim / Dc) interleaved slices (thread = (slice,
// column)); the slices' partial sums are combined in slice order, so the result does not depend on timing.
__global__ void stats_finalize_kernel(const void* __restrict__ stats, int fixed, const long long* __restrict__ flen,
                                      int B, int D, int Dc, int S, float clamp, float* __restrict__ mean,
                                      float* __restrict__ sd, float* __restrict__ packed,
                                      long long* __restrict__ zero_sums) {
  pdl_launch_dependents();
  pdl_wait();  // (a no-op unless launched with programmatic stream serialisation)
  extern __shared__ float fin_red[];  // [S][Dc][2]
  const int sl = threadIdx.x / Dc, dl = threadIdx.x - sl * Dc;
  const int d = blockIdx.x * Dc + dl;
  float sm = 0.f, ss = 0.f;
  if (sl < S && d < D) {
    for (int b = sl; b < B; b += S) {
      float m, s;
      double s1, s2;
      if (zero_sums) {  // the other half of the double-buffered sums, for the next call (sfe_chain_forward)
        zero_sums[(static_cast<long long>(b) * D + d) * 2] = 0;
        zero_sums[(static_cast<long long>(b) * D + d) * 2 + 1] = 0;
      }
      load_moments(stats, fixed, static_cast<long long>(b) * D + d, s1, s2);
      moments_to_mean_std(s1, s2, flen[b], clamp, m, s);
      if (mean) mean[static_cast<long long>(b) * D + d] = m;
      if (sd) sd[static_cast<long long>(b) * D + d] = s;
      sm += m;
      ss += s;
    }
  }
  if (sl < S) {
    fin_red[(sl * Dc + dl) * 2] = sm;
    fin_red[(sl * Dc + dl) * 2 + 1] = ss;
  }
  __syncthreads();
  if (packed && sl == 0 && d < D) {
    for (int j = 1; j < S; ++j) {
      sm += fin_red[(j * Dc + dl) * 2];
      ss += fin_red[(j * Dc + dl) * 2 + 1];
    }
    packed[d] = sm;
    packed[D + d] = ss;
    if (d == 0) packed[2 * D] = static_cast<float>(B);
  }
}

// batch statistics + running average                                       feat_norm.py:250,265,483-508
__global__ void running_update_kernel(const float* __restrict__ packed, int D, float w, int first, int update,
                                      float* __restrict__ run_mean, float* __restrict__ run_std,
                                      float* __restrict__ bmean, float* __restrict__ bstd,
                                      long long* __restrict__ counter, long long count) {
  pdl_launch_dependents();
  pdl_wait();  // (a no-op unless launched with programmatic stream serialisation)
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d == 0 && counter) *counter = count;  // the `{prefix}_batch` buffer (feat_norm.py:469-480)
  if (d >= D) return;
  const float n = packed[2 * D];
  const float m = packed[d] / n, s = packed[D + d] / n;
  if (bmean) bmean[d] = m;
  if (bstd) bstd[d] = s;
  if (update) {
    if (run_mean) run_mean[d] = first ? m : w * m + (1.f - w) * run_mean[d];
    if (run_std) run_std[d] = first ? s : w * s + (1.f - w) * run_std[d];
  }
}

// ---------------------------------------------------------------------------------------------
// normalise (+ time warp) (+ SpecAugment masks): one read, one write
// ---------------------------------------------------------------------------------------------
struct NormParams {
  const float* src;
  float* dst;
  unsigned short* dst16;  // if set, the result is written here as bfloat16 (round to nearest even) instead of dst
  const long long* flen;
  int B;
  long long F;
  int D;
  const void* stats;
  int stats_fixed;
  float clamp;
  int use_mean, use_std;
  const float* mean;
  const float* sd;
  int per_utt;
  float* mean_out;
  float* std_out;
  int src_has_pad;
  float pad_value;
  int warp_on, warp_center, warp_pos, warp_minlen;
  MaskDesc mk;
  float fill;
  const float* fill_dev;
  int frames_per_cta;
  long long* zero_sums;  // optional (B, D, 2) sums buffer to clear for the next step (double-buffered scratch)
};

// fp32 -> bf16 bits, round to nearest even (NaN stays NaN): what torch's .to(torch.bfloat16) does
__device__ __forceinline__ unsigned int bf16_bits(float f) {
  const unsigned int u = __float_as_uint(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (u >> 16) | 0x40u;
  return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}
__device__ __forceinline__ uint2 bf16x4(float4 x) {
  return make_uint2(bf16_bits(x.x) | (bf16_bits(x.y) << 16), bf16_bits(x.z) | (bf16_bits(x.w) << 16));
}

__device__ __forceinline__ void cubic_w(float t, float (&w)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

#ifndef SFE_NORM_MINB
#define SFE_NORM_MINB 4
#endif
// ---------------------------------------------------------------------------------------------
// SpecAugment time warp IN PLACE, rows [0, minlen) only (+ time masks): the last launch of a training step whose
// statistics were known up front.  The frontend kernel has already written every row normalised and frequency-
// masked; rows at or past minlen are final (the warp and the time masks live below the batch's shortest length,
// specaug.py:96-127,178-199), so this kernel touches minlen of F rows instead of reading and re-writing the tensor.
// A CTA owns `cols` Mel bins of one utterance for ALL warped rows: it stages them in shared memory, then gathers the
// four taps from there — no other CTA reads what it overwrites.  (A frequency-masked bin holds the constant fill 0
// in every row, and any interpolation of zeros is zero: the frequency masks need not be applied again.)
// ---------------------------------------------------------------------------------------------
struct WarpInPlaceParams {
  float* x;
  int B;
  long long F;
  int D, cols;
  int mode, center, pos, minlen;  // mode 1 bicubic, 2 bilinear (align_corners=True)
  MaskDesc mk;                    // time masks only
  float fill;
};

__global__ void __launch_bounds__(256) warp_inplace_kernel(const __grid_constant__ WarpInPlaceParams p) {
  pdl_launch_dependents();
  extern __shared__ __align__(16) float wsm[];  // [minlen][cols]
  const int b = blockIdx.y, c0 = blockIdx.x * p.cols;
  const int cols4 = min(p.cols, p.D - c0) / 4;
  const int n_tmask = p.mk.n_tmask;
  __shared__ int2 s_tiv[64];
  if (p.mk.inl_on)
    for (int q = threadIdx.x; q < n_tmask; q += blockDim.x)
      s_tiv[q] = make_int2(p.mk.inl[b * n_tmask + q], max(p.mk.inl[p.B * n_tmask + b * n_tmask + q], 0));
  pdl_wait();
  if (!p.mk.inl_on)
    for (int q = threadIdx.x; q < n_tmask; q += blockDim.x)
      s_tiv[q] = make_int2(p.mk.tpos[b * n_tmask + q], max(p.mk.tlen[b * n_tmask + q], 0));
  float* base = p.x + static_cast<long long>(b) * p.F * p.D + c0;
  const int items = p.minlen * cols4;
  float4* sm4 = reinterpret_cast<float4*>(wsm);
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int t = i / cols4, c = i - t * cols4;
    sm4[i] = *(reinterpret_cast<const float4*>(base + static_cast<long long>(t) * p.D) + c);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int t = i / cols4, c = i - t * cols4;
    int lo, in_len, out_len, t2;
    if (t < p.pos) { lo = 0; in_len = p.center; out_len = p.pos; t2 = t; }
    else { lo = p.center; in_len = p.minlen - p.center; out_len = p.minlen - p.pos; t2 = t - p.pos; }
    const float scale = out_len > 1 ? static_cast<float>(in_len - 1) / static_cast<float>(out_len - 1) : 0.f;
    const float real = scale * static_cast<float>(t2);
    const int i0 = static_cast<int>(floorf(real));
    float cw[4];
    if (p.mode == 2) {
      const float f = real - static_cast<float>(i0);
      cw[0] = 0.f; cw[1] = 1.f - f; cw[2] = f; cw[3] = 0.f;
    } else {
      cubic_w(real - static_cast<float>(i0), cw);
    }
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = lo + min(max(i0 - 1 + q, 0), in_len - 1);
      const float4 v = sm4[idx * cols4 + c];
      x.x = fmaf(cw[q], v.x, x.x); x.y = fmaf(cw[q], v.y, x.y); x.z = fmaf(cw[q], v.z, x.z); x.w = fmaf(cw[q], v.w, x.w);
    }
    bool tmask = false;
    for (int q = 0; q < n_tmask; ++q) {
      const int2 iv = s_tiv[q];
      tmask |= static_cast<unsigned>(t - iv.x) < static_cast<unsigned>(iv.y);
    }
    if (tmask) x = make_float4(p.fill, p.fill, p.fill, p.fill);
    *(reinterpret_cast<float4*>(base + static_cast<long long>(t) * p.D) + c) = x;
  }
}

// WARP: the instance with the time-warp gather (more registers: 3 CTAs per SM); the plain instance keeps 4 per SM
template <bool WARP>
__global__ void __launch_bounds__(256, WARP ? SFE_NORM_MINB - 1 : SFE_NORM_MINB) normalize_kernel(const __grid_constant__ NormParams p) {
  pdl_launch_dependents();
  extern __shared__ float nsm[];  // mean[D] | 1/std[D] | fmask[D]
  const int b = blockIdx.y, D = p.D;
  float* s_mean = nsm;
  float* s_rstd = nsm + D;
  float* s_fm = nsm + 2 * D;
  const int n_fmask = p.mk.n_fmask, n_tmask = p.mk.n_tmask;
  // this utterance's time-mask intervals [lo, hi) -> shared memory (n_tmask <= 64)
  __shared__ int s_tlo[64], s_thi[64];
  __shared__ int2 s_tiv[64];  // {first frame, length}: one unsigned compare per interval in the streaming loop
  // the same intervals as one bit per row of this CTA's slice of the time axis: the streaming loop tests a bit instead
  // of walking the intervals (slices longer than 32 * NORM_TB_WORDS rows walk them)
  constexpr int NORM_TB_WORDS = 128;
  __shared__ unsigned s_tbits[NORM_TB_WORDS];
  const int t0i = static_cast<int>(blockIdx.x) * static_cast<int>(p.frames_per_cta);
  const bool tb_on = p.frames_per_cta <= 32 * NORM_TB_WORDS;
  auto fill_bits = [&](const int* tpos, const int* tlen) {
    if (!tb_on) return;
    const int nw = (static_cast<int>(p.frames_per_cta) + 31) >> 5;
    for (int w = threadIdx.x; w < nw; w += blockDim.x) {
      unsigned bits = 0u;
      const int base = t0i + 32 * w;
      for (int q = 0; q < n_tmask; ++q) {
        const int ps = tpos[b * n_tmask + q], ln = tlen[b * n_tmask + q];
        const int lo = max(ps - base, 0), hi = min(ps + ln - base, 32);
        if (hi > lo) bits |= (hi - lo >= 32 ? 0xffffffffu : ((1u << (hi - lo)) - 1u)) << lo;
      }
      s_tbits[w] = bits;
    }
  };
  if (p.mk.inl_on) {
    // descriptors that travel in the parameter block need nothing from the previous kernel: before the wait
    const int* m_fpos = p.mk.inl;
    const int* m_flen = p.mk.inl + p.B * n_fmask;
    const int* m_tpos = p.mk.inl + 2 * p.B * n_fmask;
    const int* m_tlen = p.mk.inl + 2 * p.B * n_fmask + p.B * n_tmask;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      bool fm = false;
      for (int q = 0; q < n_fmask; ++q) {
        const int ps = m_fpos[b * n_fmask + q], ln = m_flen[b * n_fmask + q];
        fm |= (d >= ps) && (d < ps + ln);
      }
      s_fm[d] = fm ? 1.f : 0.f;
    }
    for (int q = threadIdx.x; q < n_tmask; q += blockDim.x) {
      const int ps = m_tpos[b * n_tmask + q];
      s_tlo[q] = ps;
      s_thi[q] = ps + m_tlen[b * n_tmask + q];
      s_tiv[q] = make_int2(ps, max(m_tlen[b * n_tmask + q], 0));
    }
    fill_bits(m_tpos, m_tlen);
  }
  pdl_wait();
  const long long n_valid = min(p.flen[b], p.F);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float m = 0.f, s = 1.f;
    if (p.stats) {
      double s1, s2;
      load_moments(p.stats, p.stats_fixed, static_cast<long long>(b) * D + d, s1, s2);
      moments_to_mean_std(s1, s2, p.flen[b], p.clamp, m, s);
      if (blockIdx.x == 0) {
        if (p.mean_out) p.mean_out[static_cast<long long>(b) * D + d] = m;
        if (p.std_out) p.std_out[static_cast<long long>(b) * D + d] = s;
        if (p.zero_sums) {
          p.zero_sums[(static_cast<long long>(b) * D + d) * 2] = 0;
          p.zero_sums[(static_cast<long long>(b) * D + d) * 2 + 1] = 0;
        }
      }
      if (!p.use_mean) m = 0.f;
      if (!p.use_std) s = 1.f;
    } else {
      if (p.mean) m = p.per_utt ? p.mean[static_cast<long long>(b) * D + d] : p.mean[d];
      if (p.sd) s = p.per_utt ? p.sd[static_cast<long long>(b) * D + d] : p.sd[d];
    }
    s_mean[d] = m;
    s_rstd[d] = 1.0f / s;  // (x-m)*(1/s): within 1 ulp of the reference's division
    if (!p.mk.inl_on) {
      bool fm = false;
      for (int q = 0; q < n_fmask; ++q) {
        const int ps = p.mk.fpos[b * n_fmask + q], ln = p.mk.flen[b * n_fmask + q];
        fm |= (d >= ps) && (d < ps + ln);
      }
      s_fm[d] = fm ? 1.f : 0.f;
    }
  }
  if (!p.mk.inl_on) {
    for (int q = threadIdx.x; q < n_tmask; q += blockDim.x) {
      const int ps = p.mk.tpos[b * n_tmask + q];
      s_tlo[q] = ps;
      s_thi[q] = ps + p.mk.tlen[b * n_tmask + q];
      s_tiv[q] = make_int2(ps, max(p.mk.tlen[b * n_tmask + q], 0));
    }
    fill_bits(p.mk.tpos, p.mk.tlen);
  }
  __syncthreads();
  const float fillv = p.fill_dev ? *p.fill_dev : p.fill;
  const long long t0 = static_cast<long long>(blockIdx.x) * p.frames_per_cta;
  const long long t1 = min(p.F, t0 + p.frames_per_cta);
  const float* src = p.src + static_cast<long long>(b) * p.F * D;
  float* dst = p.dst + static_cast<long long>(b) * p.F * D;
  unsigned short* dst16 = p.dst16 ? p.dst16 + static_cast<long long>(b) * p.F * D : nullptr;
  const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.src) & 15) == 0) &&
                   (p.dst16 ? ((reinterpret_cast<uintptr_t>(p.dst16) & 7) == 0) : ((reinterpret_cast<uintptr_t>(p.dst) & 15) == 0));
  const int W = vec ? D / 4 : D;  // work items per frame
  // thread -> (row within a group of rows, column chunk): no per-element index arithmetic
  const int rows = max(1, static_cast<int>(blockDim.x) / W);
  const int r0 = (W <= static_cast<int>(blockDim.x)) ? threadIdx.x / W : 0;
  const bool active = (W <= static_cast<int>(blockDim.x)) ? (r0 < rows) : true;
  if (!active) return;
  for (int c = (W <= static_cast<int>(blockDim.x)) ? threadIdx.x % W : threadIdx.x; c < W; c += blockDim.x) {
    float mv[4], rv[4], fv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = vec ? c * 4 + e : c;
      mv[e] = s_mean[d];
      rv[e] = s_rstd[d];
      fv[e] = s_fm[d];
    }
    if (vec) {
      // common case: four independent rows in flight per thread (the pass is latency bound otherwise);
      // 32-bit row indices, one compare per time-mask interval, 64-bit arithmetic only in the address.
      // A time warp (on by default in the recipes' SpecAugment blocks) turns the load of a row below warp_minlen
      // into the four-tap gather; every other row streams as usual.
      const int it1 = static_cast<int>(t1), nv = static_cast<int>(n_valid);
      const float4* sp = reinterpret_cast<const float4*>(src) + c;
      float4* dp = reinterpret_cast<float4*>(dst) + c;
      auto finish = [&](int t, float4 x) {
        bool tmask = false;
        if (tb_on) {
          const int r = t - t0i;
          tmask = ((s_tbits[r >> 5] >> (r & 31)) & 1u) != 0u;
        } else {
          for (int q = 0; q < n_tmask; ++q) {
            const int2 iv = s_tiv[q];
            tmask |= static_cast<unsigned>(t - iv.x) < static_cast<unsigned>(iv.y);
          }
        }
        float* xv = reinterpret_cast<float*>(&x);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = (xv[e] - mv[e]) * rv[e];
          if (tmask || fv[e] != 0.f) v = fillv;
          xv[e] = v;
        }
        if (dst16) __stcs(reinterpret_cast<uint2*>(dst16) + static_cast<long long>(t) * W + c, bf16x4(x));
        else __stcs(dp + static_cast<long long>(t) * W, x);
      };
      const float4 padv = make_float4(p.pad_value, p.pad_value, p.pad_value, p.pad_value);
      const bool synth = !p.src_has_pad;
      int t = static_cast<int>(t0) + r0;
      if (!WARP || !p.warp_on) {
        for (; t + 3 * rows < it1; t += 4 * rows) {
          float4 x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int tt = t + u * rows;
            x[u] = (synth && tt >= nv) ? padv : __ldcs(sp + static_cast<long long>(tt) * W);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) finish(t + u * rows, x[u]);
        }
        for (; t < it1; t += rows) finish(t, (synth && t >= nv) ? padv : __ldcs(sp + static_cast<long long>(t) * W));
        continue;
      }
      if constexpr (WARP) {
        const int wmin = (WARP && p.warp_on) ? p.warp_minlen : 0;
        auto load_row = [&](int tt) -> float4 {
          if (synth && tt >= nv) return padv;
          if (tt < wmin) {
            int lo, in_len, out_len, t2;
            if (tt < p.warp_pos) { lo = 0; in_len = p.warp_center; out_len = p.warp_pos; t2 = tt; }
            else { lo = p.warp_center; in_len = p.warp_minlen - p.warp_center; out_len = p.warp_minlen - p.warp_pos; t2 = tt - p.warp_pos; }
            const float scale = out_len > 1 ? static_cast<float>(in_len - 1) / static_cast<float>(out_len - 1) : 0.f;
            const float real = scale * static_cast<float>(t2);
            const int i0 = static_cast<int>(floorf(real));
            float cw[4];
            if (p.warp_on == 2) {  // bilinear (align_corners=True): taps i0, i0 + 1 with weights 1 - f, f
              const float f = real - static_cast<float>(i0);
              cw[0] = 0.f; cw[1] = 1.f - f; cw[2] = f; cw[3] = 0.f;
            } else {
              cubic_w(real - static_cast<float>(i0), cw);
            }
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int idx = lo + min(max(i0 - 1 + q, 0), in_len - 1);
              const float4 v = __ldg(sp + static_cast<long long>(idx) * W);
              x.x = fmaf(cw[q], v.x, x.x); x.y = fmaf(cw[q], v.y, x.y); x.z = fmaf(cw[q], v.z, x.z); x.w = fmaf(cw[q], v.w, x.w);
            }
            return x;
          }
          return __ldcs(sp + static_cast<long long>(tt) * W);
        };
        for (; t + 3 * rows < it1; t += 4 * rows) {
          float4 x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) x[u] = load_row(t + u * rows);
#pragma unroll
          for (int u = 0; u < 4; ++u) finish(t + u * rows, x[u]);
        }
        for (; t < it1; t += rows) finish(t, load_row(t));
      }
      continue;
    }
    for (long long t = t0 + r0; t < t1; t += rows) {
      bool tmask = false;
      for (int q = 0; q < n_tmask; ++q) tmask |= (t >= s_tlo[q]) && (t < s_thi[q]);
      // source taps (time warp) ------------------------------------------------------------
      int idx[4];
      float cw[4];
      bool warped = false;
      if (WARP && p.warp_on && t < p.warp_minlen) {
        warped = true;
        int lo, in_len, out_len, tt;
        if (t < p.warp_pos) { lo = 0; in_len = p.warp_center; out_len = p.warp_pos; tt = static_cast<int>(t); }
        else { lo = p.warp_center; in_len = p.warp_minlen - p.warp_center; out_len = p.warp_minlen - p.warp_pos; tt = static_cast<int>(t) - p.warp_pos; }
        const float scale = out_len > 1 ? static_cast<float>(in_len - 1) / static_cast<float>(out_len - 1) : 0.f;
        const float real = scale * static_cast<float>(tt);
        const int i0 = static_cast<int>(floorf(real));
        if (p.warp_on == 2) {  // bilinear (align_corners=True): taps i0, i0 + 1 with weights 1 - f, f
          const float f = real - static_cast<float>(i0);
          cw[0] = 0.f; cw[1] = 1.f - f; cw[2] = f; cw[3] = 0.f;
        } else {
          cubic_w(real - static_cast<float>(i0), cw);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) idx[q] = lo + min(max(i0 - 1 + q, 0), in_len - 1);
      }
      const bool is_pad = (t >= n_valid) && !p.src_has_pad;
      {  // (scalar columns: feature widths that are not a multiple of four, or unaligned tensors)
        float v;
        if (is_pad) v = p.pad_value;
        else if (warped) {
          v = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) v = fmaf(cw[q], __ldg(src + static_cast<long long>(idx[q]) * D + c), v);
        } else v = src[t * D + c];
        v = (v - mv[0]) * rv[0];
        if (tmask || fv[0] != 0.f) v = fillv;
        if (dst16) dst16[t * D + c] = static_cast<unsigned short>(bf16_bits(v));
        else dst[t * D + c] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// two-sided linear spectrogram (torch.stft(onesided=False), speech2linear.py:197-246): the one-sided POWER spectrum
// of the frontend kernel mirrored to n_fft bins (|X[n_fft - k]| = |X[k]| for a real signal), the frame energy over
// all n_fft bins, then the reference's point-wise tail (sqrt for mag_spec, clamp + log) — rows at or past feat_len
// hold zero power already, which is what the reference feeds to that tail.  One warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) twosided_kernel(const float* __restrict__ power, const long long* __restrict__ flen,
                                                       long long rows, long long F, int K, int n_fft, int mag, int logging,
                                                       float clampv, float lmul, float* __restrict__ out,
                                                       float* __restrict__ energy) {
  const int lane = threadIdx.x & 31;
  const long long warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long r = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    const float* src = power + r * K;
    float* dst = out + r * n_fft;
    float sum = 0.f;
    for (int k = lane; k < n_fft; k += 32) {
      float x = src[k < K ? k : n_fft - k];
      sum += x;
      if (mag) x = sqrtf(x);
      if (logging) x = fast_log(x, clampv, lmul);
      dst[k] = x;
    }
    if (energy) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) energy[r] = ((r % F) < flen[r / F]) ? sqrtf(fmaxf(sum, 1e-10f)) : 0.f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// delta / delta-delta                                                      delta_feat.py:57-81
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) delta_kernel(const float* __restrict__ x, int B, long long F, int D, int order,
                                                    int N, float denom, float* __restrict__ out) {
  const long long total = static_cast<long long>(B) * F * D;
  const int OD = D * (order + 1);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    const long long bt = i / D;
    const long long t = bt % F, b = bt / F;
    const float* row = x + b * F * D + d;
    float* o = out + bt * OD;
    o[d] = row[t * D];
    // first order at time u (zero outside [0,F))
    auto delta1 = [&](long long u) -> float {
      if (u < 0 || u >= F) return 0.f;
      float acc = 0.f;
      for (int j = -N; j <= N; ++j) {
        const long long v = u + j;
        if (v >= 0 && v < F) acc = fmaf(static_cast<float>(j) / denom, __ldg(row + v * D), acc);
      }
      return acc;
    };
    o[D + d] = delta1(t);
    if (order == 2) {
      float acc = 0.f;
      for (int j = -N; j <= N; ++j) acc = fmaf(static_cast<float>(j) / denom, delta1(t + j), acc);
      o[2 * D + d] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// mean of all elements (two-stage, fp64 accumulate)                         specaug.py:203
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sum_all_kernel(const float* __restrict__ x, long long n, double* __restrict__ acc) {
  double s = 0.0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    s += static_cast<double>(__ldg(x + i));
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(acc, sh[0]);
}
__global__ void mean_finish_kernel(double* acc, long long n, float* out) {
  out[0] = static_cast<float>(acc[0] / static_cast<double>(n));
  acc[0] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// standalone banded mel projection of an existing linear spectrogram      linear2mel.py:163-189
// each warp converts FPW rows per iteration through the same banded tables as the fused kernel
// ---------------------------------------------------------------------------------------------
struct L2MParams {
  const float* lin;
  float* mel;
  long long n_rows;
  int K, D;
  int mel_log;
  float mel_clamp, mel_div, mel_lmul;
  const float* tables;
  int tab_floats, o_mtab, o_mw, mel_rounds, mel_tail;
  unsigned int mel_split_mask;
  int2 mel_ph[MAX_MEL_ROUNDS];
  int sm_warp, sm_warp_bytes;
};

__global__ void __launch_bounds__(NTHREADS) linear2mel_kernel(const __grid_constant__ L2MParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* tab = reinterpret_cast<float*>(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* P2 = reinterpret_cast<float*>(smem + p.sm_warp + warp * p.sm_warp_bytes);
  for (int i = tid; i < p.tab_floats; i += NTHREADS) tab[i] = __ldg(p.tables + i);
  const int p2_floats = (p.K + MEL_PAD_BINS) * FPW;  // the pad bins stay zero
  for (int i = lane; i < p2_floats; i += 32) P2[i] = 0.f;
  __syncthreads();
  const int2* mtab = reinterpret_cast<const int2*>(tab + p.o_mtab) + lane;
  const float4* mw4 = reinterpret_cast<const float4*>(tab + p.o_mw);
  const float2* P2v = reinterpret_cast<const float2*>(P2);
  const long long pairs = (p.n_rows + FPW - 1) / FPW;
  for (long long pr = static_cast<long long>(blockIdx.x) * NWARPS + warp; pr < pairs;
       pr += static_cast<long long>(gridDim.x) * NWARPS) {
    const long long r0 = pr * FPW;
#pragma unroll
    for (int s = 0; s < FPW; ++s) {
      const long long r = r0 + s;
      for (int k = lane; k < p.K; k += 32) P2[k * FPW + s] = (r < p.n_rows) ? __ldg(p.lin + r * p.K + k) : 0.f;
    }
    __syncwarp();
    for (int ph = 0; ph < p.mel_rounds; ++ph) {
      const int2 so = mtab[ph * 32];
      const int oinfo = so.y;
      const float4* pp = reinterpret_cast<const float4*>(P2v + so.x);  // units start on even bins
      const int n4 = p.mel_ph[ph].x;
      const float4* wp = mw4 + p.mel_ph[ph].y + lane;
      float a0 = 0.f, a1 = 0.f;
      for (int q = 0; q < n4; ++q) {
        const float4 wq = wp[q * 32];
        const float4 pa = pp[2 * q], pb = pp[2 * q + 1];
        a0 = fmaf(wq.x, pa.x, a0); a1 = fmaf(wq.x, pa.y, a1);
        a0 = fmaf(wq.y, pa.z, a0); a1 = fmaf(wq.y, pa.w, a1);
        a0 = fmaf(wq.z, pb.x, a0); a1 = fmaf(wq.z, pb.y, a1);
        a0 = fmaf(wq.w, pb.z, a0); a1 = fmaf(wq.w, pb.w, a1);
      }
      if ((p.mel_split_mask >> ph) & 1u) {
        const float o0 = __shfl_xor_sync(0xffffffffu, a0, 1), o1 = __shfl_xor_sync(0xffffffffu, a1, 1);
        if (oinfo & 0x10000) { a0 += o0; a1 += o1; }
      }
      if (oinfo >= 0) {
        const int d = oinfo & 0xffff;
        if (p.mel_log) {
          a0 = fast_log(a0, p.mel_clamp, p.mel_lmul);
          a1 = fast_log(a1, p.mel_clamp, p.mel_lmul);
        }
        if (r0 < p.n_rows) p.mel[r0 * p.D + d] = a0;
        if (r0 + 1 < p.n_rows) p.mel[(r0 + 1) * p.D + d] = a1;
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// mel -> linear recovery (TTS inference): rows of exp'd mel features times pinv(W)^T.
// One CTA owns M2L_ROWS rows: their features sit in shared memory transposed ([d][row]) so that the inner
// loop reads four rows per LDS.128 broadcast; thread <-> output bin k, pinv_t (D, K) read coalesced from L2.
// ---------------------------------------------------------------------------------------------
constexpr int M2L_ROWS = 32;
constexpr int M2L_THREADS = 288;

__global__ void __launch_bounds__(M2L_THREADS) mel2linear_kernel(const float* __restrict__ mel, const long long* __restrict__ feat_len,
                                                                long long n_rows, long long F, int D, int K,
                                                                const float* __restrict__ pinv_t, int log_mode, float log_base,
                                                                long long max_len, float clampv, float* __restrict__ lin) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* E = reinterpret_cast<float*>(smem);  // [D][M2L_ROWS]
  const long long r0 = static_cast<long long>(blockIdx.x) * M2L_ROWS;
  for (int i = threadIdx.x; i < D * M2L_ROWS; i += M2L_THREADS) {
    const int r = i % M2L_ROWS, d = i / M2L_ROWS;
    float v = 0.f;
    if (r0 + r < n_rows) {
      v = __ldg(mel + (r0 + r) * D + d);
      if (log_mode == 1) v = expf(v);
      else if (log_mode == 2) v = powf(log_base, v);
    }
    E[i] = v;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += M2L_THREADS) {
    float acc[M2L_ROWS];
#pragma unroll
    for (int r = 0; r < M2L_ROWS; ++r) acc[r] = 0.f;
    for (int d = 0; d < D; ++d) {
      const float w = __ldg(pinv_t + static_cast<long long>(d) * K + k);
      const float4* e4 = reinterpret_cast<const float4*>(E + d * M2L_ROWS);
#pragma unroll
      for (int q = 0; q < M2L_ROWS / 4; ++q) {
        const float4 e = e4[q];
        acc[4 * q + 0] = fmaf(e.x, w, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(e.y, w, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(e.z, w, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(e.w, w, acc[4 * q + 3]);
      }
    }
#pragma unroll
    for (int r = 0; r < M2L_ROWS; ++r) {
      const long long row = r0 + r;
      if (row < n_rows) {
        const long long b = row / F, f = row - b * F;
        const long long len = feat_len[b];
        float v = acc[r];
        if (f >= len && len != max_len) v = 0.f;  // only utterances shorter than the longest are silenced
        lin[row * K + k] = fmaxf(v, clampv);
      }
    }
  }
}

// 16-bit PCM rows -> fp32 rows (x / 32768: what soundfile's float32 read does, utilbox/data_loading_util.py:71-78).
// grid (chunks, B); a thread converts eight samples: one 128-bit load, two 128-bit stores.
__global__ void __launch_bounds__(256) pcm16_to_float_kernel(const long long* __restrict__ len, const short* __restrict__ src,
                                                             long long src_pitch, float* __restrict__ dst,
                                                             long long dst_pitch, int vec_ok) {
  const int b = blockIdx.y;
  const long long n = len[b];
  const short* s = src + static_cast<long long>(b) * src_pitch;
  float* d = dst + static_cast<long long>(b) * dst_pitch;
  const float k = 1.f / 32768.f;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x * 8) {
    if (vec_ok && i + 8 <= n) {
      const int4 v = *reinterpret_cast<const int4*>(s + i);
      const int w[4] = {v.x, v.y, v.z, v.w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[2 * j] = static_cast<float>(static_cast<short>(w[j] & 0xffff)) * k;
        o[2 * j + 1] = static_cast<float>(static_cast<short>(w[j] >> 16)) * k;
      }
      *reinterpret_cast<float4*>(d + i) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(d + i + 4) = make_float4(o[4], o[5], o[6], o[7]);
    } else {
      for (long long j = i; j < n && j < i + 8; ++j) d[j] = static_cast<float>(s[j]) * k;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// waveform pre-normalisation before the STFT                               speech2linear.py:165-187
// y[n] = x[n] - a x[n-1] (n < len, else 0), statistics over the WHOLE padded row like the reference
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float prenorm_sample(const float* __restrict__ x, long long n, long long len, int has_pre, float a) {
  if (n >= len) return 0.f;
  float v = x[n];
  if (has_pre && n > 0) v = v - a * x[n - 1];
  return v;
}

// grid (B): mode 1 -> scratch[b] = {sum, sum of squares} (fp64), mode 2 -> {min, max}
__global__ void __launch_bounds__(1024) prenorm_stats_kernel(const float* __restrict__ wav, long long pitch,
                                                             const long long* __restrict__ len, long long T, int has_pre,
                                                             float a, int mode, double* __restrict__ scratch) {
  const int b = blockIdx.x;
  const float* x = wav + static_cast<long long>(b) * pitch;
  const long long n_valid = min(len[b], T);
  double u = (mode == 1) ? 0.0 : 3.0e38, v = (mode == 1) ? 0.0 : -3.0e38;
  for (long long n = threadIdx.x; n < T; n += blockDim.x) {
    const double y = static_cast<double>(prenorm_sample(x, n, n_valid, has_pre, a));
    if (mode == 1) { u += y; v = fma(y, y, v); }
    else { u = fmin(u, y); v = fmax(v, y); }
  }
  __shared__ double su[32], sv[32];
  for (int o = 16; o > 0; o >>= 1) {
    const double ou = __shfl_xor_sync(0xffffffffu, u, o), ov = __shfl_xor_sync(0xffffffffu, v, o);
    if (mode == 1) { u += ou; v += ov; } else { u = fmin(u, ou); v = fmax(v, ov); }
  }
  if ((threadIdx.x & 31) == 0) { su[threadIdx.x >> 5] = u; sv[threadIdx.x >> 5] = v; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    u = (threadIdx.x < nw) ? su[threadIdx.x] : ((mode == 1) ? 0.0 : 3.0e38);
    v = (threadIdx.x < nw) ? sv[threadIdx.x] : ((mode == 1) ? 0.0 : -3.0e38);
    for (int o = 16; o > 0; o >>= 1) {
      const double ou = __shfl_xor_sync(0xffffffffu, u, o), ov = __shfl_xor_sync(0xffffffffu, v, o);
      if (mode == 1) { u += ou; v += ov; } else { u = fmin(u, ou); v = fmax(v, ov); }
    }
    if (threadIdx.x == 0) { scratch[2 * b] = u; scratch[2 * b + 1] = v; }
  }
}

// grid (chunks, B): mode 1 (y - mean) / std (unbiased), mode 2 (y - min) / (max - min) * 2 - 1, whole row
__global__ void __launch_bounds__(256) prenorm_apply_kernel(const float* __restrict__ wav, long long pitch,
                                                            const long long* __restrict__ len, long long T, int has_pre,
                                                            float a, int mode, const double* __restrict__ scratch,
                                                            float* __restrict__ out, long long out_pitch) {
  const int b = blockIdx.y;
  const float* x = wav + static_cast<long long>(b) * pitch;
  float* o = out + static_cast<long long>(b) * out_pitch;
  const long long n_valid = min(len[b], T);
  float c0, c1;
  if (mode == 1) {
    const double s1 = scratch[2 * b], s2 = scratch[2 * b + 1], dn = static_cast<double>(T);
    const double mean = s1 / dn;
    c0 = static_cast<float>(mean);
    c1 = static_cast<float>(sqrt(fmax((s2 - s1 * mean) / (dn - 1.0), 0.0)));
  } else {
    c0 = static_cast<float>(scratch[2 * b]);
    c1 = static_cast<float>(scratch[2 * b + 1]) - c0;
  }
  for (long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; n < T;
       n += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float y = prenorm_sample(x, n, n_valid, has_pre, a);
    o[n] = (mode == 1) ? (y - c0) / c1 : (y - c0) / c1 * 2.f - 1.f;
  }
}

}  // namespace sfe
