// Fused waveform -> (log-)mel / linear spectrogram kernel for sm_100a.
//
// One CTA per SM: one producer warp and up to 23 consumer warps; there is no CTA-wide barrier after the
// one-time setup.  The work is cut into *units* of FPW = 2 consecutive frames of one utterance, laid out
// utterance-major on a cost axis (units that hold only frames past an utterance's length are either skipped
// or cost 1/8 of a real one); every CTA owns a contiguous equal-cost range of it.
//
//   producer   lane i serves consumer warp i: when the warp hands its staging buffer back ("empty" mbarrier)
//              the lane takes the next unit of the CTA's range from a shared-memory counter, writes a 32-byte
//              descriptor into the warp's slot and starts ONE 1-D bulk async copy (cp.async.bulk -> UBLKCP)
//              of the hop + win samples the unit touches onto the warp's "full" mbarrier.  The buffer comes
//              back as soon as the first FFT pass of unit i has consumed it, so the copy of unit i+1 flies
//              under the rest of unit i.  Neighbouring units overlap by win - hop samples: those re-reads
//              hit L2, HBM is read once.
//   consumer   window (+ pre-emphasis on the fly) -> real FFT as an n_fft/2-point complex Stockham FFT (radix
//              16/8/5/4/3/2 butterflies in registers, passes exchanged through a warp-private padded shared
//              buffer) -> real-FFT unpack + |X|^2 -> banded (sparse) mel filter -> clamp/log -> [global
//              (x-mean)/std] -> [SpecAugment masks] -> coalesced stores, plus exact fixed-point per-utterance
//              sum / sum-of-squares accumulation for the normaliser in warp-private shared memory (flushed
//              when the utterance changes).
//
// Reference arithmetic replaced: speechain/module/frontend/speech2linear.py:165-245,
// linear2mel.py:177-187 (see include/sfe.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "sfe_fft.cuh"

// Timing experiments only (tools/variant.py): SFE_ABLATE=1 no mel stage, 2 no unpack either, 3 first FFT pass only,
// 4 staging only.  Never defined in the shipped build.
#ifndef SFE_ABLATE
#define SFE_ABLATE 0
#endif
#ifndef SFE_MEL_UNROLL  // tuning experiments (tools/variant.py)
#define SFE_MEL_UNROLL 1
#endif

namespace sfe {

constexpr int NWARPS = 8;                 // warps per CTA of the small standalone kernels
constexpr int NTHREADS = NWARPS * 32;
constexpr int FPW = 2;                    // frames per unit (per warp step)
constexpr int MAX_MEL_ROUNDS = 16;        // phases of 32 filter units (n_mels <= 256)
constexpr int kMelUnroll = SFE_MEL_UNROLL;
constexpr int MEL_PAD_BINS = 4;           // zeroed bins behind the power spectrum (filter units read whole 4-tap blocks)
constexpr int MASK_INLINE_MAX = 2304;     // SFE_MASK_INLINE_MAX (the recipes' 4 + 5 masks for up to 128 utterances)
constexpr int UNIT_TABLE_MAX_B = 256;     // batches up to this size get a per-utterance cost table
constexpr int REAL_UNIT_COST = 8;         // a unit with real frames costs this many pad-only units
constexpr int HOST_SPLIT_MAX_GRID = 160;  // tile kernel: grids up to this size can take their work split from the host

// SpecAugment interval descriptors: device arrays, or (small batches) carried in the parameter block
struct MaskDesc {
  const int* fpos;
  const int* flen;
  const int* tpos;
  const int* tlen;
  int n_fmask, n_tmask;
  int inl_on;
  int inl[MASK_INLINE_MAX];
};

struct FwdParams {
  // input
  const float* wav;
  long long pitch;
  long long total_floats;  // extent of the wav allocation in floats (for the bulk-copy bound)
  const long long* len;
  int B;
  long long T_max;
  int F;
  int units_per_utt;       // ceil(F / FPW)
  // outputs
  float* mel;
  float* lin;
  float* energy;
  long long* feat_len;
  unsigned long long* stats;   // fixed point: [..][0] = sum x * 2^32, [..][1] = sum x^2 * 2^24
  const float* norm_mean;
  const float* norm_std;
  // masks
  MaskDesc mk;
  float fill;
  const float* fill_dev;
  int skip_pad;
  // config
  int n_fft, win, hop, K, D, off, center;
  int has_pre;
  float pre;
  int mag;
  int lin_log;
  float lin_clamp, lin_div, lin_lmul;   // lin_lmul = ln2 / ln(base): log_b(x) = log2(x) * lmul
  int mel_log;
  float mel_clamp, mel_div, mel_lmul;
  float pad_mel, pad_lin;
  // tables (device blob, float units) and shared-memory carve-up (byte offsets)
  const float* tables;
  int tab_floats;
  int o_win, o_tw2, o_tw3, o_wpost, o_mtab, o_mw;  // float offsets into the blob
  int mel_rounds;                                    // phases of 32 filter units
  int2 mel_ph[MAX_MEL_ROUNDS];                       // per phase: {blocks of 4 taps, float4 offset of its weights}
  unsigned int mel_split_mask;                       // phases that hold bins split over two lanes
  int mel_tail;                                      // zeroed bins behind the spectrum (MEL_PAD_BINS)
  // per-CTA: [tables | table mbarrier | warp 0 | warp 1 | ...]; per warp: [FFT scratch (P2 may alias it) | P2 |
  // staging | sums | mbarrier]
  int sm_tbar, sm_warp, sm_warp_bytes, sm_p2_off, sm_stage_off, sm_acc_off, sm_bar_off;
  // ---- tile kernel (sfe_tile.cuh): shared-memory carve-up (byte offsets) and the Mel emit schedule ----
  // [tables (window, twiddles) | emit schedule | control block | stage x2 | power-spectrum tiles x2 | warp scratch]
  int t2_ctrl, t2_stage, t2_stage_floats, t2_p2, t2_pp, t2_rowf, t2_warp, t2_warp_bytes;
  int t2_o_sched_g, t2_sched_floats;  // float offset of the schedule in the device blob; its length
  // Work split made on the host (the caller knew the lengths there: sfe_frontend_forward's wav_len_host): the CTA
  // starts without the length load -> scan -> search chain.  hs_cta[c] = {first pair, number of pairs, utterance
  // of the first pair}; hs_len / hs_pc = samples of every utterance / pair index at which it starts.
  int hs_on;
  int hs_cta[HOST_SPLIT_MAX_GRID][3];
  int hs_len[UNIT_TABLE_MAX_B];
  int hs_pc[UNIT_TABLE_MAX_B + 1];
  int ms_ng, ms_two, ms_b_ld, ms_d_split, ms_b_off, ms_c_split, ms_tail_a, ms_tail_b;  // sfe_melsched.h
};

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA engine, no tensor map needed)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
// Programmatic dependent launch (no-ops when the launch does not carry the attribute): let the next kernel of
// the stream start its prologue on SMs this grid has left / wait until the previous grid has completed and its
// memory is visible.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// FFT plan (compile time)
// ---------------------------------------------------------------------------------------------
// WIN_ > 0 fixes the window length at compile time (the common (n_fft, win_length) pairs): the rows of the
// first pass that fall into the window's zero padding, and every test against `win`, vanish from the code.
template <int M_, int G_, int R1_, int R2_, int R3_, int PADQ_ = R1_, int WIN_ = 0>
struct FftPlan {
  static constexpr int M = M_;     // complex FFT length = n_fft / 2
  static constexpr int WIN = WIN_;
  static constexpr int G = G_;     // lanes cooperating on one frame (16 or 32)
  static constexpr int R1 = R1_, R2 = R2_, R3 = R3_;
  static constexpr int C = 32 / G_;        // frames in flight per warp
  static constexpr int S = FPW / C;        // sequential sub-iterations per tile
  static constexpr int PADQ = PADQ_;       // one padding slot every PADQ complex values
  static constexpr int MP = M_ + M_ / PADQ_ + 1;
  // the power spectrum may overwrite the FFT scratch only if no later FFT of the tile needs it
  static constexpr bool ALIAS_P2 = (S == 1);
  // one CTA per SM; consumer warps per CTA at most, plus one producer warp: 24 / 16 / 8 warps in all (register
  // files are per SM sub-partition, so a 25th warp would cost every thread 8 registers), 80 / 128 / 255 registers
#ifdef SFE_MAXW  // tuning experiments (tools/variant.py)
  static constexpr int MAXW = SFE_MAXW;
#else
  static constexpr int MAXW = (M_ <= 256) ? 23 : ((M_ <= 512) ? 15 : 7);
#endif
  static constexpr int RL = (R3_ > 1) ? R3_ : R2_;               // radix of the last pass
  // the last pass leaves Z[l + G*r] in register r of lane l: the real-FFT unpack then pairs k with M-k
  // by one shuffle instead of a shared-memory round trip
  static constexpr bool LASTREG = (M_ / RL == G_);
  static_assert(R1_ * R2_ * R3_ == M_, "radices must multiply to M");
  static_assert(C * S == FPW, "bad lane grouping");
};

template <int PADQ>
__device__ __forceinline__ int padidx(int i) { return i + i / PADQ; }

// One Stockham pass of radix R with P = product of the previous radices.  FIRST passes read the
// windowed samples; the others read the warp-private Z buffer.  In place: all inputs of all
// rounds are held in registers before anything is written.
// SYNC1: the samples of a FIRST pass live in the scratch the pass writes (edge units of the tile kernel).
template <class PL, int R, int P, bool FIRST, bool VEC, bool PRE, bool SYNC1 = false>
__device__ __forceinline__ void stockham_pass(cpx* __restrict__ Z, const cpx* __restrict__ tw, int l,
                                              const float* __restrict__ fr, const float* __restrict__ wtab,
                                              int win, float pre) {
  constexpr int M = PL::M, G = PL::G, Q = PL::PADQ;
  constexpr int T = M / R;
  constexpr int NR = (T + G - 1) / G;
  constexpr bool FULL = (T % G == 0);
  if (PL::WIN > 0) win = PL::WIN;  // compile-time window length: the row tests below fold away
  cpx u[NR][R];
#pragma unroll
  for (int rd = 0; rd < NR; ++rd) {
    const int j = l + rd * G;
    if (FULL || j < T) {
      if (FIRST) {
        const float* f0 = fr + 2 * j;
        const float* w0 = wtab + 2 * j;
        // 1) raw sample pairs (rows entirely inside the window's zero padding are not touched)
        cpx raw[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          raw[r] = make_float2(0.f, 0.f);
          if (2 * r * T < win) {  // warp-uniform
            if (VEC) raw[r] = *reinterpret_cast<const float2*>(f0 + 2 * r * T);
            else raw[r] = make_float2(f0[2 * r * T], f0[2 * r * T + 1]);
          }
        }
        // 2) pre-emphasis needs x[2n-1]: the .y of the pair one lane to the left (same row), or of the
        //    last lane of the previous row for lane 0 — one select + one shuffle instead of a second,
        //    bank-conflicting load.  Kept outside any branch so the shuffles are unconditional.
        if (PRE) {
          if (T == G && NR == 1) {
            float carry = fr[-1];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const float give = (l == G - 1) ? carry : raw[r].y;
              const float xm = __shfl_sync(0xffffffffu, give, (l + G - 1) & (G - 1), G);
              carry = raw[r].y;
              raw[r] = f2fma(make_float2(-pre, -pre), make_float2(xm, raw[r].x), raw[r]);
            }
          } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
              if (2 * r * T < win) {
                const float xm = f0[2 * r * T - 1];
                raw[r] = f2fma(make_float2(-pre, -pre), make_float2(xm, raw[r].x), raw[r]);
              }
            }
          }
        }
        // 3) window
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (2 * r * T < win) {
            const float2 w = *reinterpret_cast<const float2*>(w0 + 2 * r * T);
            const float2 yw = f2mul(raw[r], w);
            float y0 = yw.x, y1 = yw.y;
            if (2 * r * T + 2 * T > win) {
              // (warp-uniform) this row straddles the end of the window: samples past it may be anything
              // — memory beyond an utterance's length is never assumed to be initialised — so select, don't scale
              const int sidx = 2 * (j + r * T);
              if (sidx >= win) y0 = 0.f;
              if (sidx + 1 >= win) y1 = 0.f;
            }
            u[rd][r] = make_float2(y0, y1);
          } else {
            u[rd][r] = make_float2(0.f, 0.f);
          }
        }
      } else {
        if (T % Q == 0) {
          const cpx* zr = Z + padidx<Q>(j);
#pragma unroll
          for (int r = 0; r < R; ++r) u[rd][r] = zr[r * (T + T / Q)];
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) u[rd][r] = Z[padidx<Q>(j + r * T)];
        }
        const cpx* twj = tw + j;
#pragma unroll
        for (int r = 1; r < R; ++r) u[rd][r] = cmul(u[rd][r], twj[(r - 1) * T]);
      }
    }
  }
  if (!FIRST || SYNC1) __syncwarp();
#pragma unroll
  for (int rd = 0; rd < NR; ++rd) {
    const int j = l + rd * G;
    if (FULL || j < T) {
      bfly<R>(u[rd]);
      if (P == 1 && R == Q) {
        cpx* zw = Z + j * (R + 1);
#pragma unroll
        for (int r = 0; r < R; ++r) zw[r] = u[rd][r];
      } else {
        const int k = (P >= G * NR) ? j : j % P;  // when every butterfly index is < P the twiddle index is j itself
        const int jo = (j - k) * R + k;
        if (P % Q == 0) {
          cpx* zw = Z + padidx<Q>(jo);
#pragma unroll
          for (int r = 0; r < R; ++r) zw[r * (P + P / Q)] = u[rd][r];
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) Z[padidx<Q>(jo + r * P)] = u[rd][r];
        }
      }
    }
  }
  __syncwarp();
}

// Last pass with T == G: one butterfly per lane, result kept in registers (u[r] = Z[l + G*r]).
template <class PL, int R, int P>
__device__ __forceinline__ void stockham_last_reg(const cpx* __restrict__ Z, const cpx* __restrict__ tw, int l,
                                                  cpx (&u)[R]) {
  constexpr int M = PL::M, Q = PL::PADQ;
  constexpr int T = M / R;
  static_assert(T == PL::G, "register-resident last pass needs one butterfly per lane");
  if (T % Q == 0) {
    const cpx* zr = Z + padidx<Q>(l);
#pragma unroll
    for (int r = 0; r < R; ++r) u[r] = zr[r * (T + T / Q)];
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) u[r] = Z[padidx<Q>(l + r * T)];
  }
  const cpx* twj = tw + l;
#pragma unroll
  for (int r = 1; r < R; ++r) u[r] = cmul(u[r], twj[(r - 1) * T]);
  bfly<R>(u);
}

// all passes that go through the shared-memory scratch
template <class PL, bool VEC, bool PRE>
__device__ __forceinline__ void fft_frame(cpx* Zw, const cpx* tw2, const cpx* tw3, int l, const float* fr,
                                          const float* wtab, int win, float pre) {
  stockham_pass<PL, PL::R1, 1, true, VEC, PRE>(Zw, nullptr, l, fr, wtab, win, pre);
  if (PL::R3 > 1) {
    stockham_pass<PL, PL::R2, PL::R1, false, false, false>(Zw, tw2, l, nullptr, nullptr, 0, 0.f);
    if (!PL::LASTREG)
      stockham_pass<PL, (PL::R3 > 1 ? PL::R3 : 2), PL::R1 * PL::R2, false, false, false>(Zw, tw3, l, nullptr, nullptr,
                                                                                         0, 0.f);
  } else if (!PL::LASTREG) {
    stockham_pass<PL, PL::R2, PL::R1, false, false, false>(Zw, tw2, l, nullptr, nullptr, 0, 0.f);
  }
}

// the passes after the first one (the staging buffer is free again once the first pass has returned)
template <class PL>
__device__ __forceinline__ void fft_rest(cpx* Zw, const cpx* tw2, const cpx* tw3, int l) {
  if (PL::R3 > 1) {
    stockham_pass<PL, PL::R2, PL::R1, false, false, false>(Zw, tw2, l, nullptr, nullptr, 0, 0.f);
    if (!PL::LASTREG)
      stockham_pass<PL, (PL::R3 > 1 ? PL::R3 : 2), PL::R1 * PL::R2, false, false, false>(Zw, tw3, l, nullptr, nullptr,
                                                                                         0, 0.f);
  } else if (!PL::LASTREG) {
    stockham_pass<PL, PL::R2, PL::R1, false, false, false>(Zw, tw2, l, nullptr, nullptr, 0, 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// log_b(max(x, clamp)) through the SFU: MUFU.LG2 is accurate to 2^-22 absolute on [0.5, 2] and 2 ulp
// elsewhere — three orders of magnitude inside the 1e-4 / 1e-3 tolerance of the features — and costs 3
// instructions where logf + an IEEE division cost ~45.
// The clamp floor is a normal number (sfe_plan_create rejects smaller ones), so the flush-to-zero form is
// exact here and drops the denormal pre-scaling that __log2f carries (3 instructions instead of 7).
__device__ __forceinline__ float fast_log(float x, float clampv, float lmul) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(x, clampv)));
  return r * lmul;
}

template <class PL, bool PRE>
__global__ void __launch_bounds__((PL::MAXW + 1) * 32, 1) fused_frontend_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int M = PL::M, G = PL::G, Q = PL::PADQ;
  // magnitude spectra (mag_spec, no shipped recipe): sqrt in one sweep after the unpack where the spectrum
  // aliases the FFT scratch, inside the unpack loop elsewhere — whichever measured faster for the power path
  constexpr bool MAG_SWEEP = PL::ALIAS_P2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane / G, l = lane % G;
  const int NW = (blockDim.x >> 5) - 1;     // consumer warps; the last warp of the CTA is the producer
  const bool producer = (warp == NW);

  float* tab = reinterpret_cast<float*>(smem);
  unsigned char* wbase = smem + p.sm_warp + (producer ? 0 : warp) * p.sm_warp_bytes;
  cpx* Zw = reinterpret_cast<cpx*>(wbase) + g * PL::MP;
  float* P2 = reinterpret_cast<float*>(wbase + p.sm_p2_off);          // [k][FPW] interleaved
  float* stage = reinterpret_cast<float*>(wbase + p.sm_stage_off);    // this warp's samples
  long long* wacc = reinterpret_cast<long long*>(wbase + p.sm_acc_off);  // [D][2] fixed-point sums of the current utterance
  const uint32_t bar = smem_u32(wbase + p.sm_bar_off);                // this warp's "full" barrier ("empty" at +8)
  const uint32_t tbar = smem_u32(smem + p.sm_tbar);                   // tables
  __shared__ int s_cum[UNIT_TABLE_MAX_B + 1];  // cost coordinate at which every utterance starts
  __shared__ int s_next;                       // next chunk of this CTA's range

  const int hop = p.hop, win = p.win, D = p.D, K = p.K, F = p.F, B = p.B;
  const int upu = p.units_per_utt;
  // Units that hold only frames past an utterance's length have nothing to compute.  When they have nothing
  // to write either (the normalise pass synthesises those rows) they are not enumerated at all.
  const bool compact = p.skip_pad && !p.lin && !p.energy;
  const bool tabled = (B <= UNIT_TABLE_MAX_B);
  auto frames_of = [&](long long len) -> long long {
    return (p.center ? (len + 2 * (p.n_fft / 2)) - p.n_fft : len - p.n_fft) / hop + 1;
  };
  auto clampF = [&](long long fl) -> int { return static_cast<int>(fl < 0 ? 0 : (fl > F ? F : fl)); };

  // ---- one-time setup: barriers, tables -> smem with one bulk copy, cost table, zeroed sums ----
  // (the staging buffers need no initialisation: every read of them is either covered by a copy or
  //  replaced by a select, see the first Stockham pass)
  if (tid == 0) {
    s_next = 0;
    mbar_init(tbar, 1);
    fence_mbar_init();
    const uint32_t bytes = static_cast<uint32_t>(p.tab_floats) * 4u;
    mbar_expect_tx(tbar, bytes);
    bulk_g2s(smem_u32(tab), p.tables, bytes, tbar);
  }
  pdl_launch_dependents();
  pdl_wait();  // (the tables above are constants of the plan; everything below may depend on the previous kernel)
  if (lane == 0 && !producer) {
    mbar_init(bar, 1);       // "full": the unit's descriptor is written and its samples have landed
    mbar_init(bar + 8, 1);   // "empty": the warp has consumed its staging buffer
    fence_mbar_init();
  }
  if (!producer)
    for (int i = lane; i < 2 * D; i += 32) wacc[i] = 0;
  if (blockIdx.x == 0 && p.feat_len)
    for (int b = tid; b < B; b += blockDim.x) p.feat_len[b] = frames_of(__ldg(p.len + b));
  if (tabled && warp == 0) {
    // exclusive scan of the utterances' costs (warp shuffles, 32 utterances per round)
    int carry = 0;
    for (int b0 = 0; b0 < B; b0 += 32) {
      const int b = b0 + lane;
      int c = 0;
      if (b < B) {
        const int nr = (clampF(frames_of(__ldg(p.len + b))) + FPW - 1) / FPW;
        c = REAL_UNIT_COST * nr + (compact ? 0 : upu - nr);
      }
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (b < B) s_cum[b] = carry + incl - c;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_cum[B] = carry;
  }
  __syncthreads();               // barrier inits and the cost table are visible
  mbar_wait(tbar, 0u);           // tables have landed

  const float* wtab = tab + p.o_win;
  const cpx* tw2 = reinterpret_cast<const cpx*>(tab + p.o_tw2);
  const cpx* tw3 = reinterpret_cast<const cpx*>(tab + p.o_tw3);
  const cpx* wpost = reinterpret_cast<const cpx*>(tab + p.o_wpost);
  const int2* mtab = reinterpret_cast<const int2*>(tab + p.o_mtab) + lane;  // {first bin, output info} per (phase, lane)
  const float4* mw4 = reinterpret_cast<const float4*>(tab + p.o_mw);
  const bool base_aligned = ((reinterpret_cast<uintptr_t>(p.wav) & 15) == 0);
  const int n_fmask = p.mk.n_fmask, n_tmask = p.mk.n_tmask;
  const int* m_fpos = p.mk.inl_on ? p.mk.inl : p.mk.fpos;
  const int* m_flen = p.mk.inl_on ? p.mk.inl + B * n_fmask : p.mk.flen;
  const int* m_tpos = p.mk.inl_on ? p.mk.inl + 2 * B * n_fmask : p.mk.tpos;
  const int* m_tlen = p.mk.inl_on ? p.mk.inl + 2 * B * n_fmask + B * n_tmask : p.mk.tlen;
  constexpr int pre1 = PRE ? 1 : 0;
  constexpr int EDGE_SHIFT = 4;  // gathered units leave room for the (unused) predecessor sample

  // ---- work distribution over the cost axis [0, ctot) ----
  // One CTA per SM.  The axis is cut into one contiguous, equal-cost range per CTA (SMs are equally fast).
  // Inside a CTA the LAST WARP IS A PRODUCER: lane i serves consumer warp i.  Whenever warp i reports its
  // staging buffer free, lane i takes the next real unit (REAL_UNIT_COST) of the CTA's range from a
  // shared-memory counter, works out which utterance / frames it is, writes a 32-byte descriptor into the
  // warp's slot and starts the bulk copy of the unit's samples onto the warp's "full" barrier.  The consumer
  // warps carry none of that arithmetic (≈ 300 of 1300 instructions per unit) or state: they wait, read five
  // words, and transform.  A consumer that the hardware scheduler favours simply asks more often, so the tail
  // of an SM is one unit long.
  const int ctot = tabled ? s_cum[B] : B * upu;
  const int real_step = tabled ? REAL_UNIT_COST : 1;  // without the table every unit counts the same
  const int cta_lo = static_cast<int>(static_cast<long long>(ctot) * blockIdx.x / gridDim.x);
  const int cta_hi = static_cast<int>(static_cast<long long>(ctot) * (blockIdx.x + 1) / gridDim.x);

  enum { U_DONE = 0, U_PAD = 1, U_BULK = 2, U_EDGE = 3 };
  struct Slot {               // what a consumer needs to know about its unit
    int kind, b, tu, flen;    // frames FPW*tu, FPW*tu+1 of utterance b, which has flen frames
    int shift, pad_;          // staging index of the unit's first sample (bulk units); pad units in a row (pad kind)
    long long len;            // samples of the utterance (gathered units)
  };
  static_assert(sizeof(Slot) == 32, "slot layout");

  if (producer) {
    // ======================= producer warp: lane i <-> consumer warp i =======================
    if (lane >= NW) return;
    unsigned char* cbase = smem + p.sm_warp + lane * p.sm_warp_bytes;
    Slot* slot = reinterpret_cast<Slot*>(cbase + p.sm_bar_off + 16);
    const uint32_t c_full = smem_u32(cbase + p.sm_bar_off), c_empty = c_full + 8;
    const uint32_t c_stage = smem_u32(cbase + p.sm_stage_off);
    // position on the axis: unit tu of utterance b starts at cost coordinate c; the lane's chunk ends at c_hi
    int qb = -1, qtu = 0, qflen = 0, qnunit = 0, qc = 0, qc_hi = 0;
    long long qlen = 0;
    auto load_utt = [&]() {
      qlen = __ldg(p.len + qb);
      qflen = clampF(frames_of(qlen));
      qnunit = (compact && tabled) ? (qflen + FPW - 1) / FPW : upu;
    };
    // step over utterances that own no unit (zero frames in compact mode) / past the end of one
    auto settle = [&]() {
      while (qtu >= qnunit && qb + 1 < B) {
        ++qb;
        qtu = 0;
        load_utt();
        if (tabled) qc = s_cum[qb];
      }
    };
    bool first = true, live = true;
    uint32_t pe = 0u;  // parity of the consumer's "empty" barrier
    while (live) {
      bool ready = first;
      if (!first) {
        uint32_t ok;
        asm volatile("{ .reg .pred q; mbarrier.test_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                     : "=r"(ok) : "r"(c_empty), "r"(pe) : "memory");
        ready = (ok != 0u);
      }
      if (!ready) {
        __nanosleep(40);
        continue;
      }
      if (!first) pe ^= 1u;
      // ---- the unit after the previous one of this consumer ----
      bool have = false;  // (qb, qtu) is a unit inside the lane's current chunk
      if (!first && qtu * FPW >= qflen) {
        // after a pad unit the chunk may hold more of them (a real unit always uses its chunk up)
        qc += 1;
        ++qtu;
        settle();
        have = (qtu < qnunit && qc < qc_hi);
      }
      first = false;
      while (!have) {
        // next chunk of the CTA's range
        unsigned int k;
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(k) : "r"(smem_u32(&s_next)) : "memory");
        const long long lo64 = static_cast<long long>(cta_lo) + static_cast<long long>(k) * real_step;
        if (lo64 >= cta_hi) { live = false; break; }
        const int c0 = static_cast<int>(lo64);
        qc_hi = min(c0 + real_step, cta_hi);
        // first unit whose start coordinate is >= c0 (chunks move forward: usually still the same utterance)
        if (tabled) {
          int lo = qb;
          if (lo < 0 || s_cum[lo] > c0) lo = 0;
          if (lo + 1 < B && s_cum[lo + 1] <= c0) {
            int hi = B - 1;  // last utterance that starts at or before c0
            ++lo;
            while (lo < hi) {
              const int mid = (lo + hi + 1) >> 1;
              if (s_cum[mid] <= c0) lo = mid; else hi = mid - 1;
            }
          }
          if (lo != qb) {
            qb = lo;
            load_utt();
          }
          const int nreal = (qflen + FPW - 1) / FPW;
          const int off = c0 - s_cum[qb];
          const int real_cost = REAL_UNIT_COST * nreal;
          if (off <= real_cost) {
            qtu = (off + REAL_UNIT_COST - 1) / REAL_UNIT_COST;
            qc = c0 - off + qtu * REAL_UNIT_COST;
          } else {
            qtu = nreal + (off - real_cost);
            qc = c0;
          }
        } else {
          int nb = c0 / upu;
          if (nb >= B) nb = B - 1;
          qtu = c0 - nb * upu;
          qc = c0;
          if (nb != qb) {
            qb = nb;
            load_utt();
          }
        }
        settle();
        have = (qtu < qnunit && qc < qc_hi);
      }
      // ---- describe it, start its copy, publish ----
      Slot sl;
      sl.kind = U_DONE; sl.b = qb; sl.tu = qtu; sl.flen = qflen; sl.shift = 0; sl.pad_ = 0; sl.len = qlen;
      bool bulk = false;
      long long ga = 0;
      uint32_t bytes = 0;
      if (live) {
        sl.kind = U_PAD;
        if (qtu * FPW < qflen) {
          const long long o_start = static_cast<long long>(qtu) * FPW * hop + p.off;
          const long long o_end = o_start + static_cast<long long>(FPW - 1) * hop + win;
          const long long row0 = static_cast<long long>(qb) * p.pitch;
          const long long gi = row0 + o_start - pre1;
          ga = gi & ~3LL;
          const long long ge = (row0 + o_end + 3) & ~3LL;
          bulk = base_aligned && (o_start - pre1 >= 0) && (o_end <= qlen) && (ge <= p.total_floats);
          sl.shift = static_cast<int>(gi - ga) + pre1;
          bytes = static_cast<uint32_t>((ge - ga) * 4);
          sl.kind = bulk ? U_BULK : U_EDGE;
        } else {
          // a run of pad units of this utterance inside the chunk goes out as ONE descriptor (their rows are
          // contiguous): a pad unit is too cheap to pay a producer round trip each
          const int n = min(qc_hi - qc, qnunit - qtu);
          sl.pad_ = n;
          qtu += n - 1;
          qc += n - 1;
        }
      }
      *slot = sl;
      if (bulk) {
        // the consumer released the buffer through its "empty" barrier; order that before the async-proxy write
        fence_proxy_async();
        mbar_expect_tx(c_full, bytes);   // (arrive with release: the slot is visible to whoever sees the phase flip)
        bulk_g2s(c_stage, p.wav + ga, bytes, c_full);
      } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(c_full) : "memory");
      }
    }
    return;
  }

  // ============================== consumer warps ==============================
  const Slot* slot = reinterpret_cast<const Slot*>(wbase + p.sm_bar_off + 16);
  auto flush_sums = [&](int bb) {
    __syncwarp();
    for (int i = lane; i < 2 * D; i += 32) {
      const long long v = wacc[i];
      if (v != 0) {
        wacc[i] = 0;
        atomicAdd(p.stats + static_cast<long long>(bb) * D * 2 + i, static_cast<unsigned long long>(v));
      }
    }
    __syncwarp();
  };
  auto release_stage = [&]() {  // every lane is past its last read of the staging buffer and of the slot
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar + 8) : "memory");
  };

  uint32_t parity = 0u;
  int prev_b = -1;

  while (true) {
    mbar_wait(bar, parity);   // descriptor written, samples landed
    parity ^= 1u;
    const int kind = slot->kind;
    const int b = slot->b, flen = slot->flen;
    const int tA = slot->tu * FPW;
    if (p.stats && prev_b >= 0 && (kind == U_DONE || b != prev_b)) flush_sums(prev_b);
    if (kind == U_DONE) break;
    prev_b = b;
    const bool real = (kind != U_PAD);

    if (!real) {
      // ---------------- pad unit: nothing to transform ----------------
      const int nfr = min(FPW * slot->pad_, F - tA);  // rows of the whole run of pad units
      if (p.mel && !p.skip_pad) {
        float* dst = p.mel + (static_cast<long long>(b) * F + tA) * D;
        const float fillv = p.fill_dev ? *p.fill_dev : p.fill;
        for (int i = lane; i < nfr * D; i += 32) {
          const int d = i % D;
          float v = p.pad_mel;
          if (p.norm_mean) v -= p.norm_mean[d];
          if (p.norm_std) v /= p.norm_std[d];
          bool m = false;
          for (int q = 0; q < n_fmask; ++q) {
            const int ps = m_fpos[b * n_fmask + q], ln = m_flen[b * n_fmask + q];
            m |= (d >= ps) && (d < ps + ln);
          }
          dst[i] = m ? fillv : v;
        }
      }
      if (p.lin) {
        float* dst = p.lin + (static_cast<long long>(b) * F + tA) * K;
        for (int i = lane; i < nfr * K; i += 32) dst[i] = p.pad_lin;
      }
      if (p.energy)
        for (int i = lane; i < nfr; i += 32) p.energy[static_cast<long long>(b) * F + tA + i] = 0.f;
      release_stage();
    } else {
      // ---------------- samples of this unit ----------------
      const bool staged_raw = (kind == U_BULK);
      const int shift = staged_raw ? slot->shift : EDGE_SHIFT;
      if (!staged_raw) {
        // edge unit: explicit gather with batch-level reflection, zero beyond the length and
        // pre-emphasis applied here (staging then holds the final samples)
        const long long len = slot->len;
        const long long o_start = static_cast<long long>(tA) * hop + p.off;
        const float* row = p.wav + static_cast<long long>(b) * p.pitch;
        const int n = (FPW - 1) * hop + win;
        for (int i0 = lane; i0 < n; i0 += 32 * 4) {
          float v[4], pv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 32;
            long long o = o_start + i;
            if (p.center) {
              if (o < 0) o = -o;
              if (o >= p.T_max) o = 2 * (p.T_max - 1) - o;
            }
            const bool ok = (i < n) && (o >= 0) && (o < len);
            v[u] = ok ? __ldg(row + o) : 0.f;
            pv[u] = (ok && p.has_pre && o > 0) ? __ldg(row + o - 1) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 32;
            if (i < n) stage[EDGE_SHIFT + i] = fmaf(-p.pre, pv[u], v[u]);
          }
        }
        if (lane == 0) stage[EDGE_SHIFT - 1] = 0.f;  // read (times a zero coefficient) as the predecessor of sample 0
        __syncwarp();
      }

      // the staging buffer is handed back to the producer as soon as the first FFT pass has consumed it:
      // the next unit's descriptor and copy are prepared under the rest of this unit
      auto prefetch_next = [&]() { release_stage(); };

      // ---------------- FFT + |X|^2 of FPW frames ----------------
      float esum[PL::S];
#pragma unroll
      for (int sub = 0; sub < ((SFE_ABLATE >= 4) ? 0 : PL::S); ++sub) {
        const int slot = sub * PL::C + g;   // frame index inside the unit
        const int t = tA + slot;
        const int fo = slot * hop + shift;
        const float* fr = stage + fo;
        // gathered units were pre-emphasised by the gather: same code path with a zero coefficient
        const float pre = (PRE && staged_raw) ? p.pre : 0.f;
        if ((fo & 1) == 0) stockham_pass<PL, PL::R1, 1, true, true, PRE>(Zw, nullptr, l, fr, wtab, win, pre);
        else stockham_pass<PL, PL::R1, 1, true, false, PRE>(Zw, nullptr, l, fr, wtab, win, pre);
        // (the pass ends with a __syncwarp: every lane has consumed its samples)
        if (sub == PL::S - 1) prefetch_next();
        if (SFE_ABLATE >= 3) { esum[sub] = 0.f; continue; }
        fft_rest<PL>(Zw, tw2, tw3, l);

        // real-FFT unpack: X[k] = E + W^k O, X[M-k] = conj(E - W^k O); the 1/2 lives in the window.
        // Lane l owns k = l + G*i (i <= NI-1) and its mirror M-k.
        constexpr int RL = PL::RL;
        constexpr int NI = PL::LASTREG ? (RL / 2 + 1) : ((M / 2) / G + 1);
        cpx za[NI], zb[NI];
        if constexpr (PL::LASTREG) {
          cpx u[RL];
          stockham_last_reg<PL, RL, M / RL>(Zw, (PL::R3 > 1) ? tw3 : tw2, l, u);
          if (SFE_ABLATE == 2) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < RL; ++i) a += u[i].x + u[i].y;
            __syncwarp();
            P2[l * FPW + slot] = a;
            esum[sub] = 0.f;
            __syncwarp();
            continue;
          }
          const int src = (G - l) & (G - 1);
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            za[i] = u[i];
            const float bx = __shfl_sync(0xffffffffu, u[RL - 1 - i].x, src, G);
            const float by = __shfl_sync(0xffffffffu, u[RL - 1 - i].y, src, G);
            zb[i] = (l == 0) ? u[(RL - i) % RL] : make_float2(bx, by);
          }
          // (the shuffles are convergence points: every lane has finished reading the scratch)
        } else {
          const cpx* zl = Zw + padidx<Q>(l);
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            const int k = l + i * G;
            if (i < NI - 1 || k <= M / 2) {
              za[i] = (G % Q == 0) ? zl[i * (G + G / Q)] : Zw[padidx<Q>(k)];
              zb[i] = Zw[padidx<Q>(k == 0 ? 0 : M - k)];
            }
          }
          if (PL::ALIAS_P2) __syncwarp();  // every lane has its Z values: the buffer may now hold P2
        }
        float es = 0.f;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int k = l + i * G;
          if (i < NI - 1 || k <= M / 2) {
            const cpx a = za[i], bb = zb[i];
            const cpx w = wpost[k];
            const cpx E = f2add(a, make_float2(bb.x, -bb.y));
            const cpx O = f2add(make_float2(a.y, -a.x), make_float2(bb.y, bb.x));
            const cpx Tt = cmul(O, w);
            const cpx X = f2add(E, Tt), Y = f2sub(E, Tt);
            float pk = fmaf(X.x, X.x, X.y * X.y), pm = fmaf(Y.x, Y.x, Y.y * Y.y);
            es += pk;
            if (2 * k != M) es += pm;
            if (!MAG_SWEEP && p.mag) { pk = sqrtf(pk); pm = sqrtf(pm); }
            P2[k * FPW + slot] = pk;
            P2[(M - k) * FPW + slot] = pm;
          }
        }
        if (t >= flen) {
          // a frame beyond the utterance (last unit only): its spectrum is defined as zero
          es = 0.f;
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            const int k = l + i * G;
            if (i < NI - 1 || k <= M / 2) {
              P2[k * FPW + slot] = 0.f;
              P2[(M - k) * FPW + slot] = 0.f;
            }
          }
        }
        esum[sub] = es;
        __syncwarp();
      }
      if (SFE_ABLATE >= 4) prefetch_next();
      if (lane < MEL_PAD_BINS) reinterpret_cast<float2*>(P2)[K + lane] = make_float2(0.f, 0.f);  // (see the mel stage)
      __syncwarp();
      if (MAG_SWEEP && p.mag) {  // magnitude spectrum: one extra sweep, off the power-spectrum path every recipe takes
        for (int i = lane; i < K * FPW; i += 32) P2[i] = sqrtf(P2[i]);
        __syncwarp();
      }

      // ---------------- energy (sum of the power spectrum of each frame) ----------------
      if (p.energy) {
#pragma unroll
        for (int sub = 0; sub < PL::S; ++sub) {
          float es = esum[sub];
#pragma unroll
          for (int o = G / 2; o > 0; o >>= 1) es += __shfl_xor_sync(0xffffffffu, es, o);
          const int t = tA + sub * PL::C + g;
          if (l == 0 && t < F) p.energy[static_cast<long long>(b) * F + t] = (t < flen) ? sqrtf(fmaxf(es, 1e-10f)) : 0.f;
        }
      }

      // ---------------- linear spectrogram output ----------------
      if (p.lin) {
#pragma unroll
        for (int slot = 0; slot < FPW; ++slot) {
          const int t = tA + slot;
          if (t < F) {
            float* dst = p.lin + (static_cast<long long>(b) * F + t) * K;
            for (int k = lane; k < K; k += 32) {
              float v = P2[k * FPW + slot];
              if (p.lin_log) v = fast_log(v, p.lin_clamp, p.lin_lmul);
              dst[k] = v;
            }
          }
        }
      }

      // ---------------- banded mel + log + epilogue ----------------
      // Filter *units* (a mel bin, or half of a wide one) are spread over phases of 32 lanes with similar
      // tap counts; a lane walks its unit two taps at a time over both frames of the warp.
      if (p.mel && SFE_ABLATE < 1) {
        const bool fancy = (p.norm_mean != nullptr) || (p.norm_std != nullptr) || (n_fmask > 0) || (n_tmask > 0);
        const float2* P2v = reinterpret_cast<const float2*>(P2);
        float* mel_row = p.mel + (static_cast<long long>(b) * F + tA) * D;
        const bool w0 = (tA < F) && !(p.skip_pad && tA >= flen);          // row tA is written
        const bool w1 = (tA + 1 < F) && !(p.skip_pad && tA + 1 >= flen);  // row tA+1 is written
        const float v0 = (tA < flen) ? 1.f : 0.f, v1 = (tA + 1 < flen) ? 1.f : 0.f;  // rows that count for the sums
        float fillv = 0.f;
        bool tm0 = false, tm1 = false;
        if (fancy) {
          fillv = p.fill_dev ? *p.fill_dev : p.fill;
          for (int q = 0; q < n_tmask; ++q) {
            const int ps = m_tpos[b * n_tmask + q], ln = m_tlen[b * n_tmask + q];
            tm0 |= (tA >= ps) && (tA < ps + ln);
            tm1 |= (tA + 1 >= ps) && (tA + 1 < ps + ln);
          }
        }
        // The option flags are kernel constants: pick one of eight straight-line variants per unit instead
        // of testing them for every bin.
        auto mel_stage = [&](auto kLogT, auto kStatsT, auto kFancyT) {
          constexpr bool kLog = decltype(kLogT)::value, kStats = decltype(kStatsT)::value,
                         kFancy = decltype(kFancyT)::value;
          const float clampv = p.mel_clamp, lmul = p.mel_lmul;
          const float4* wl = mw4 + lane;
          float* mel_row1 = mel_row + D;
          auto phase = [&](const int ph, const int n4, const int wofs) {
            const int2 so = mtab[ph * 32];   // .x first bin of this lane's unit (even), .y output info
            const float4* pp = reinterpret_cast<const float4*>(P2v + so.x);  // {bin k: frame 0, 1; bin k+1: frame 0, 1}
            const float4* wp = wl + wofs;
            float2 accA = make_float2(0.f, 0.f), accB = make_float2(0.f, 0.f);  // (frame 0, frame 1) pairs
#pragma unroll kMelUnroll
            for (int q = 0; q < n4; ++q) {
              const float4 wq = wp[q * 32];
              const float4 pa = pp[2 * q], pb = pp[2 * q + 1];
              accA = f2fma(make_float2(wq.x, wq.x), make_float2(pa.x, pa.y), accA);
              accB = f2fma(make_float2(wq.y, wq.y), make_float2(pa.z, pa.w), accB);
              accA = f2fma(make_float2(wq.z, wq.z), make_float2(pb.x, pb.y), accA);
              accB = f2fma(make_float2(wq.w, wq.w), make_float2(pb.z, pb.w), accB);
            }
            float2 acc = f2add(accA, accB);
            // bins wider than a unit: add the partner lane's half (cheaper done always than tested per phase)
            const float o0 = __shfl_xor_sync(0xffffffffu, acc.x, 1), o1 = __shfl_xor_sync(0xffffffffu, acc.y, 1);
            if (so.y & 0x10000) acc = f2add(acc, make_float2(o0, o1));
            if (so.y >= 0) {  // this lane publishes bin d
              const int d = so.y & 0xffff;
              float x0 = acc.x, x1 = acc.y;
              if (kLog) {
                x0 = fast_log(x0, clampv, lmul);
                x1 = fast_log(x1, clampv, lmul);
              }
              if (kStats) {
                // pair sums in fp32, then exact order-independent fixed-point accumulation; bin d belongs to
                // exactly one lane of one phase, so the warp-private sums need no atomics
                const float s1 = x0 * v0 + x1 * v1;
                const float s2 = fmaf(x0 * v0, x0, x1 * v1 * x1);
                longlong2* ap = reinterpret_cast<longlong2*>(wacc) + d;
                longlong2 a = *ap;
                a.x += __float2ll_rn(s1 * 4294967296.f);
                a.y += __float2ll_rn(s2 * 16777216.f);
                *ap = a;
              }
              if (kFancy) {
                if (p.norm_mean) { const float nm = p.norm_mean[d]; x0 -= nm; x1 -= nm; }
                if (p.norm_std) { const float ns = p.norm_std[d]; x0 /= ns; x1 /= ns; }
                bool fm = false;
                for (int q = 0; q < n_fmask; ++q) {
                  const int ps = m_fpos[b * n_fmask + q], ln = m_flen[b * n_fmask + q];
                  fm |= (d >= ps) && (d < ps + ln);
                }
                if (fm || tm0) x0 = fillv;
                if (fm || tm1) x1 = fillv;
              }
              if (w0) mel_row[d] = x0;
              if (w1) mel_row1[d] = x1;
            }
          };
          const int nph = p.mel_rounds;
          if (nph <= 4) {
            // the usual filterbanks need at most four phases: straight-line code, every table offset an immediate
#pragma unroll
            for (int ph = 0; ph < 4; ++ph)
              if (ph < nph) phase(ph, p.mel_ph[ph].x, p.mel_ph[ph].y);
          } else {
            for (int ph = 0; ph < nph; ++ph) phase(ph, p.mel_ph[ph].x, p.mel_ph[ph].y);
          }
        };
        using T1 = std::true_type;
        using T0 = std::false_type;
        switch ((p.mel_log ? 1 : 0) | (p.stats ? 2 : 0) | (fancy ? 4 : 0)) {
          case 0: mel_stage(T0{}, T0{}, T0{}); break;
          case 1: mel_stage(T1{}, T0{}, T0{}); break;
          case 2: mel_stage(T0{}, T1{}, T0{}); break;
          case 3: mel_stage(T1{}, T1{}, T0{}); break;
          case 4: mel_stage(T0{}, T0{}, T1{}); break;
          case 5: mel_stage(T1{}, T0{}, T1{}); break;
          case 6: mel_stage(T0{}, T1{}, T1{}); break;
          default: mel_stage(T1{}, T1{}, T1{}); break;
        }
      }
      __syncwarp();  // P2 (and the scratch it may alias) is free for the next unit
    }

  }
}

}  // namespace sfe
