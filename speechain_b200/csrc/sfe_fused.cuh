// Fused waveform -> (log-)mel / linear spectrogram kernel for sm_100a.
//
// One persistent CTA per resident slot.  A CTA repeatedly claims a *tile* of TILE_F consecutive
// frames of one utterance (dynamic atomic scheduler), stages the (TILE_F-1)*hop + win samples
// the tile touches into shared memory with ONE 1-D bulk async copy (cp.async.bulk -> UBLKCP,
// completion on an mbarrier) — every sample is read from HBM once although it belongs to
// win/hop frames — and then each warp turns two frames into features:
//
//   window (+ pre-emphasis on the fly)  ->  real FFT as an n_fft/2-point complex Stockham FFT
//   (radix 16/8/5/4/2 butterflies in registers, passes exchanged through a warp-private,
//   padded shared-memory buffer)  ->  real-FFT unpack + |X|^2  ->  banded (sparse) mel filter
//   ->  clamp/log  ->  [global (x-mean)/std]  ->  [SpecAugment masks]  ->  coalesced stores,
//   plus fp64 per-utterance sum / sum-of-squares accumulation for the normaliser.
//
// Reference arithmetic replaced: speechain/module/frontend/speech2linear.py:165-245,
// linear2mel.py:177-187 (see include/sfe.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sfe_fft.cuh"

namespace sfe {

constexpr int NWARPS = 8;
constexpr int NTHREADS = NWARPS * 32;
constexpr int FPW = 2;                    // frames per warp per tile
constexpr int TILE_F = NWARPS * FPW;      // frames per tile
constexpr int MAX_MEL_ROUNDS = 8;         // n_mels <= 256
constexpr int MAX_MASKS = 16;

struct FwdParams {
  // input
  const float* wav;
  long long pitch;
  long long total_floats;  // extent of the wav allocation in floats (for the bulk-copy bound)
  const long long* len;
  int B;
  long long T_max;
  int F;
  int tiles_per_utt;
  int num_tiles;
  // outputs
  float* mel;
  float* lin;
  float* energy;
  long long* feat_len;
  double* stats;
  const float* norm_mean;
  const float* norm_std;
  // masks
  const int* fpos;
  const int* flen;
  int n_fmask;
  const int* tpos;
  const int* tlen;
  int n_tmask;
  float fill;
  const float* fill_dev;
  int skip_pad;
  // config
  int n_fft, win, hop, K, D, off, center;
  int has_pre;
  float pre;
  int mag;
  int lin_log;
  float lin_clamp, lin_div;
  int mel_log;
  float mel_clamp, mel_div;
  float pad_mel, pad_lin;
  // tables (device blob, float units) and shared-memory carve-up (byte offsets)
  const float* tables;
  int tab_floats;
  int o_win, o_tw2, o_tw3, o_wpost, o_mstart, o_mw;  // float offsets into the blob
  int mel_rounds;
  int mel_taps[MAX_MEL_ROUNDS];
  int mel_tail;                                      // max taps of any bin (P2 tail padding)
  int sm_stage, sm_stage_floats, sm_warp, sm_warp_bytes, sm_z_bytes, sm_red, sm_bar;
  // scheduler
  unsigned int* sched;  // [0] next tile, [1] CTAs done
};

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA engine, no tensor map needed)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// FFT plan (compile time)
// ---------------------------------------------------------------------------------------------
template <int M_, int G_, int R1_, int R2_, int R3_>
struct FftPlan {
  static constexpr int M = M_;     // complex FFT length = n_fft / 2
  static constexpr int G = G_;     // lanes cooperating on one frame (16 or 32)
  static constexpr int R1 = R1_, R2 = R2_, R3 = R3_;
  static constexpr int C = 32 / G_;        // frames in flight per warp
  static constexpr int S = FPW / C;        // sequential sub-iterations per tile
  static constexpr int PADQ = R1_;         // one padding slot every PADQ complex values
  static constexpr int MP = M_ + M_ / R1_ + 1;
  static_assert(R1_ * R2_ * R3_ == M_, "radices must multiply to M");
  static_assert(C * S == FPW, "bad lane grouping");
};

template <int PADQ>
__device__ __forceinline__ int padidx(int i) { return i + i / PADQ; }

// One Stockham pass of radix R with P = product of the previous radices.  FIRST passes read the
// windowed samples; the others read the warp-private Z buffer.  In place: all inputs of all
// rounds are held in registers before anything is written.
template <class PL, int R, int P, bool FIRST>
__device__ __forceinline__ void stockham_pass(cpx* __restrict__ Z, const cpx* __restrict__ tw, int l,
                                              const float* __restrict__ fr, const float* __restrict__ wtab,
                                              int win, bool vec_ok, bool apply_pre, float pre) {
  constexpr int M = PL::M, G = PL::G;
  constexpr int T = M / R;
  constexpr int NR = (T + G - 1) / G;
  cpx u[NR][R];
#pragma unroll
  for (int rd = 0; rd < NR; ++rd) {
    const int j = l + rd * G;
    if (T % G == 0 || j < T) {
      if (FIRST) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int s = 2 * (j + r * T);
          float y0 = 0.f, y1 = 0.f;
          if (s < win) {
            float x0, x1;
            if (vec_ok) {
              const float2 v = *reinterpret_cast<const float2*>(fr + s);
              x0 = v.x;
              x1 = v.y;
            } else {
              x0 = fr[s];
              x1 = fr[s + 1];
            }
            if (apply_pre) {
              const float xm = fr[s - 1];
              y0 = fmaf(-pre, xm, x0);
              y1 = fmaf(-pre, x0, x1);
            } else {
              y0 = x0;
              y1 = x1;
            }
            const float2 w = *reinterpret_cast<const float2*>(wtab + s);
            y0 *= w.x;
            y1 = (s + 1 < win) ? y1 * w.y : 0.f;
          }
          u[rd][r] = make_float2(y0, y1);
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) u[rd][r] = Z[padidx<PL::PADQ>(j + r * T)];
#pragma unroll
        for (int r = 1; r < R; ++r) u[rd][r] = cmul(u[rd][r], tw[(r - 1) * T + j]);
      }
    }
  }
  if (!FIRST) __syncwarp();
#pragma unroll
  for (int rd = 0; rd < NR; ++rd) {
    const int j = l + rd * G;
    if (T % G == 0 || j < T) {
      bfly<R>(u[rd]);
      const int k = j % P;
      const int jo = (j - k) * R + k;
#pragma unroll
      for (int r = 0; r < R; ++r) Z[padidx<PL::PADQ>(jo + r * P)] = u[rd][r];
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <class PL>
__global__ void __launch_bounds__(NTHREADS) fused_frontend_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int M = PL::M, G = PL::G;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane / G, l = lane % G;

  float* tab = reinterpret_cast<float*>(smem);
  float* stage = reinterpret_cast<float*>(smem + p.sm_stage);
  unsigned char* wbase = smem + p.sm_warp + warp * p.sm_warp_bytes;
  cpx* Zw = reinterpret_cast<cpx*>(wbase) + g * PL::MP;
  float* P2 = reinterpret_cast<float*>(wbase + p.sm_z_bytes);  // [k][FPW] interleaved
  double* red = reinterpret_cast<double*>(smem + p.sm_red);    // [NWARPS][D][2]
  unsigned long long* barp = reinterpret_cast<unsigned long long*>(smem + p.sm_bar);
  __shared__ int s_tile;

  // ---- one-time setup: tables -> smem, zero the scratch tails, init the mbarrier ----
  for (int i = tid; i < p.tab_floats; i += NTHREADS) tab[i] = __ldg(p.tables + i);
  for (int i = tid; i < p.sm_stage_floats; i += NTHREADS) stage[i] = 0.f;
  {
    const int p2_floats = (p.K + p.mel_tail + 2) * FPW;
    for (int i = lane; i < p2_floats; i += 32) P2[i] = 0.f;
  }
  const uint32_t bar = smem_u32(barp);
  if (tid == 0) mbar_init(bar, 1);
  uint32_t bar_parity = 0;
  __syncthreads();

  const float* wtab = tab + p.o_win;
  const cpx* tw2 = reinterpret_cast<const cpx*>(tab + p.o_tw2);
  const cpx* tw3 = reinterpret_cast<const cpx*>(tab + p.o_tw3);
  const cpx* wpost = reinterpret_cast<const cpx*>(tab + p.o_wpost);
  const int* mstart = reinterpret_cast<const int*>(tab + p.o_mstart);
  const float* mw = tab + p.o_mw;
  const int hop = p.hop, win = p.win, D = p.D, K = p.K;
  const bool base_aligned = ((reinterpret_cast<uintptr_t>(p.wav) & 15) == 0);

  for (;;) {
    if (tid == 0) s_tile = static_cast<int>(atomicAdd(p.sched, 1u));
    __syncthreads();  // also: every warp is done with the previous tile's staging buffer
    const int tile = s_tile;
    if (tile >= p.num_tiles) break;
    const int b = tile / p.tiles_per_utt;
    const int t0 = (tile - b * p.tiles_per_utt) * TILE_F;
    const long long len = p.len[b];
    const long long flen_ll = (p.center ? (len + 2 * (p.n_fft / 2)) - p.n_fft : len - p.n_fft) / hop + 1;
    const int flen = static_cast<int>(flen_ll < 0 ? 0 : (flen_ll > p.F ? p.F : flen_ll));
    if (t0 == 0 && tid == 0 && p.feat_len) p.feat_len[b] = flen_ll;
    const int nfr = min(TILE_F, p.F - t0);  // frames of this tile that exist in the output

    // ---------------- pad tile: nothing to transform ----------------
    if (t0 >= flen) {
      if (p.mel && !p.skip_pad) {
        float* dst = p.mel + (static_cast<long long>(b) * p.F + t0) * D;
        for (int i = tid; i < nfr * D; i += NTHREADS) {
          const int d = i % D;
          float v = p.pad_mel;
          if (p.norm_mean) v -= p.norm_mean[d];
          if (p.norm_std) v /= p.norm_std[d];
          dst[i] = v;  // masks on pad rows are applied below only for freq masks
        }
        if (p.n_fmask > 0) {
          const float fillv = p.fill_dev ? *p.fill_dev : p.fill;
          for (int i = tid; i < nfr * D; i += NTHREADS) {
            const int d = i % D;
            bool m = false;
            for (int q = 0; q < p.n_fmask; ++q) {
              const int ps = p.fpos[b * p.n_fmask + q], ln = p.flen[b * p.n_fmask + q];
              m |= (d >= ps) && (d < ps + ln);
            }
            if (m) dst[i] = fillv;
          }
        }
      }
      if (p.lin) {
        float* dst = p.lin + (static_cast<long long>(b) * p.F + t0) * K;
        for (int i = tid; i < nfr * K; i += NTHREADS) dst[i] = p.pad_lin;
      }
      if (p.energy) {
        for (int i = tid; i < nfr; i += NTHREADS) p.energy[static_cast<long long>(b) * p.F + t0 + i] = 0.f;
      }
      continue;
    }

    // ---------------- stage the samples of this tile ----------------
    const long long o_start = static_cast<long long>(t0) * hop + p.off;   // first sample of frame t0
    const long long o_end = o_start + static_cast<long long>(TILE_F - 1) * hop + win;
    const int pre1 = p.has_pre ? 1 : 0;
    int shift;                 // staging index of sample o_start
    bool staged_raw;           // true: staging holds x (apply pre-emphasis on the fly)
    {
      const long long gi = static_cast<long long>(b) * p.pitch + o_start - pre1;
      const long long ga = gi & ~3LL;
      const long long ge = (static_cast<long long>(b) * p.pitch + o_end + 3) & ~3LL;
      const bool interior = base_aligned && (o_start - pre1 >= 0) && (o_end <= len) && (ge <= p.total_floats);
      if (interior) {
        shift = static_cast<int>(gi - ga) + pre1;
        staged_raw = true;
        if (tid == 0) {
          const uint32_t bytes = static_cast<uint32_t>((ge - ga) * 4);
          fence_proxy_async();
          mbar_expect_tx(bar, bytes);
          bulk_g2s(smem_u32(stage), p.wav + ga, bytes, bar);
        }
        mbar_wait(bar, bar_parity);
        bar_parity ^= 1u;
      } else {
        // edge tile: explicit gather with batch-level reflection, zero beyond the length and
        // pre-emphasis applied here (staging then holds the final samples)
        shift = 0;
        staged_raw = false;
        const float* row = p.wav + static_cast<long long>(b) * p.pitch;
        const int n = static_cast<int>(o_end - o_start);
        for (int i = tid; i < n; i += NTHREADS) {
          long long o = o_start + i;
          if (p.center) {
            if (o < 0) o = -o;
            if (o >= p.T_max) o = 2 * (p.T_max - 1) - o;
          }
          float v = 0.f;
          if (o >= 0 && o < len) {
            v = __ldg(row + o);
            if (p.has_pre && o > 0) v = fmaf(-p.pre, __ldg(row + o - 1), v);
          }
          stage[i] = v;
        }
        __syncthreads();
      }
    }

    // ---------------- per-warp: FFT of FPW frames ----------------
    float esum[PL::S];
#pragma unroll
    for (int sub = 0; sub < PL::S; ++sub) {
      const int slot = sub * PL::C + g;
      const int tl = warp * FPW + slot;   // frame index inside the tile
      const int t = t0 + tl;
      const int fo = tl * hop + shift;
      const float* fr = stage + fo;
      const bool vec_ok = ((fo & 1) == 0);
      stockham_pass<PL, PL::R1, 1, true>(Zw, nullptr, l, fr, wtab, win, vec_ok, staged_raw && p.has_pre, p.pre);
      stockham_pass<PL, PL::R2, PL::R1, false>(Zw, tw2, l, nullptr, nullptr, 0, false, false, 0.f);
      if (PL::R3 > 1)
        stockham_pass<PL, (PL::R3 > 1 ? PL::R3 : 2), PL::R1 * PL::R2, false>(Zw, tw3, l, nullptr, nullptr, 0, false,
                                                                             false, 0.f);
      // real-FFT unpack: X[k] = E + W^k O, X[M-k] = conj(E - W^k O); the 1/2 lives in the window
      const bool live = (t < flen);
      float es = 0.f;
      for (int k = l; k <= M / 2; k += G) {
        const cpx a = Zw[padidx<PL::PADQ>(k)];
        const cpx bb = Zw[padidx<PL::PADQ>(k == 0 ? 0 : M - k)];
        const cpx w = wpost[k];
        const cpx E = make_float2(a.x + bb.x, a.y - bb.y);
        const cpx O = make_float2(a.y + bb.y, bb.x - a.x);
        const cpx Tt = cmul(w, O);
        const float xr = E.x + Tt.x, xi = E.y + Tt.y, yr = E.x - Tt.x, yi = E.y - Tt.y;
        float pk = fmaf(xr, xr, xi * xi), pm = fmaf(yr, yr, yi * yi);
        if (!live) { pk = 0.f; pm = 0.f; }
        es += pk;
        if (2 * k != M) es += pm;
        if (p.mag) { pk = sqrtf(pk); pm = sqrtf(pm); }
        P2[k * FPW + slot] = pk;
        P2[(M - k) * FPW + slot] = pm;
      }
      esum[sub] = es;
      __syncwarp();
    }

    // ---------------- energy (sum of the power spectrum of each frame) ----------------
    if (p.energy) {
#pragma unroll
      for (int sub = 0; sub < PL::S; ++sub) {
        float es = esum[sub];
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) es += __shfl_xor_sync(0xffffffffu, es, o);
        const int slot = sub * PL::C + g;
        const int t = t0 + warp * FPW + slot;
        if (l == 0 && t < p.F) p.energy[static_cast<long long>(b) * p.F + t] = (t < flen) ? sqrtf(fmaxf(es, 1e-10f)) : 0.f;
      }
    }

    // ---------------- linear spectrogram output ----------------
    if (p.lin) {
#pragma unroll
      for (int slot = 0; slot < FPW; ++slot) {
        const int t = t0 + warp * FPW + slot;
        if (t < p.F) {
          float* dst = p.lin + (static_cast<long long>(b) * p.F + t) * K;
          for (int k = lane; k < K; k += 32) {
            float v = P2[k * FPW + slot];
            if (p.lin_log) v = logf(fmaxf(v, p.lin_clamp)) / p.lin_div;
            dst[k] = v;
          }
        }
      }
    }

    // ---------------- banded mel + log + epilogue ----------------
    if (p.mel) {
      const int tA = t0 + warp * FPW;
      const float fillv = (p.n_fmask > 0 || p.n_tmask > 0) ? (p.fill_dev ? *p.fill_dev : p.fill) : 0.f;
      bool tm[FPW];
#pragma unroll
      for (int s = 0; s < FPW; ++s) {
        tm[s] = false;
        for (int q = 0; q < p.n_tmask; ++q) {
          const int ps = p.tpos[b * p.n_tmask + q], ln = p.tlen[b * p.n_tmask + q];
          tm[s] |= (tA + s >= ps) && (tA + s < ps + ln);
        }
      }
      const float2* P2v = reinterpret_cast<const float2*>(P2);
      int wofs = 0;
      for (int rd = 0; rd < p.mel_rounds; ++rd) {
        const int d = rd * 32 + lane;
        const int st = mstart[d];
        const int nt = p.mel_taps[rd];
        float acc0 = 0.f, acc1 = 0.f;
        for (int q = 0; q < nt; ++q) {
          const float wgt = mw[wofs + q * 32 + lane];
          const float2 pv = P2v[st + q];
          acc0 = fmaf(wgt, pv.x, acc0);
          acc1 = fmaf(wgt, pv.y, acc1);
        }
        wofs += nt * 32;
        if (d < D) {
          float v[FPW] = {acc0, acc1};
          bool fm = false;
          for (int q = 0; q < p.n_fmask; ++q) {
            const int ps = p.fpos[b * p.n_fmask + q], ln = p.flen[b * p.n_fmask + q];
            fm |= (d >= ps) && (d < ps + ln);
          }
          double s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int s = 0; s < FPW; ++s) {
            const int t = tA + s;
            float x = v[s];
            if (p.mel_log) x = logf(fmaxf(x, p.mel_clamp)) / p.mel_div;
            if (p.stats && t < flen) {
              const double xd = static_cast<double>(x);
              s1 += xd;
              s2 = fma(xd, xd, s2);
            }
            if (p.norm_mean) x -= p.norm_mean[d];
            if (p.norm_std) x /= p.norm_std[d];
            if (fm || tm[s]) x = fillv;
            if (t < p.F && !(p.skip_pad && t >= flen)) p.mel[(static_cast<long long>(b) * p.F + t) * D + d] = x;
          }
          if (p.stats) {
            red[(warp * D + d) * 2 + 0] = s1;
            red[(warp * D + d) * 2 + 1] = s2;
          }
        }
      }
    }

    // ---------------- per-utterance sums: CTA reduce, then one fp64 atomic per (dim, moment) ----------------
    if (p.stats) {
      __syncthreads();
      for (int i = tid; i < 2 * D; i += NTHREADS) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) s += red[w * D * 2 + i];
        atomicAdd(p.stats + static_cast<long long>(b) * D * 2 + i, s);
      }
    }
  }

  // self-resetting scheduler: the last CTA to leave re-arms the counters for the next launch
  if (tid == 0) {
    __threadfence();
    const unsigned done = atomicAdd(p.sched + 1, 1u);
    if (done == gridDim.x - 1) {
      p.sched[0] = 0u;
      p.sched[1] = 0u;
      __threadfence();
    }
  }
}

}  // namespace sfe
