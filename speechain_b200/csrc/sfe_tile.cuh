// Tile variant of the fused waveform -> (log-)mel kernel for sm_100a: the kernel of the shipped recipes.
//
// Same arithmetic as sfe_fused.cuh up to the power spectrum (window, pre-emphasis, Stockham FFT, real-FFT
// unpack).  What changes is WHO does the Mel stage and HOW the work flows: one CTA per SM, warp-specialised,
// every hand-over an mbarrier, no CTA-wide barrier after the set-up:
//
//   tiles        A CTA walks its contiguous, equal-cost range of 2-frame units ("pairs") in TILES of TS pairs.
//   producer     (1 warp) describes tile j — one lane per pair: utterance, frames, kind — and starts the bulk
//                async copies (cp.async.bulk -> UBLKCP) of its samples into stage buffer j & 1 as soon as the
//                FFT warps have read tile j - 2 out of it: 2*hop samples per pair plus the overlap tail of the
//                last one, so HBM is read once with no overlap re-reads inside a tile.  Pairs that touch the
//                batch-level reflection, the end of their utterance, or another utterance than the tile's
//                first are gathered by their FFT warp instead, exactly like in sfe_fused.cuh.
//   FFT warps    (16) claim pairs from a shared counter; pair s of tile j: window (+ pre-emphasis) -> FFT ->
//                unpack -> |X|^2 into ROWS 2s, 2s+1 of power-spectrum buffer j % NPB.
//   Mel warps    (NPB, one per power-spectrum buffer) take a complete tile and run the Mel stage for all 2*TS
//                frames at once with ONE LANE PER FRAME: the lane walks the bins of its row in order with one
//                packed FMA per bin into an (even filter, odd filter) accumulator pair and stores a filter when
//                its last bin is in — straight-line code generated for the recipe's filterbank
//                (sfe_melprog.h): no flag test, no branch, no address arithmetic; the weights are one
//                broadcast 128-bit load per two bins; no lane idles on a narrow filter.  The gather-style Mel
//                stage of sfe_fused.cuh spent 290 of its 960 instructions per pair here, ~90 % of them address
//                arithmetic and predicates; this one ~50.  Then the read-out: two rows per step, 16 lanes per
//                row, logarithm, per-utterance sums, optional (x-mean)/std and SpecAugment masks, 128-bit
//                coalesced stores.
//
// Reference arithmetic replaced: speechain/module/frontend/speech2linear.py:165-245, linear2mel.py:177-187.
#pragma once
#include "sfe_fused.cuh"
#include "sfe_melprog.h"

namespace sfe {

template <int BANK> struct MelBank { typedef void type; };
template <> struct MelBank<1> { typedef melprog::Bank1 type; };
template <> struct MelBank<2> { typedef melprog::Bank2 type; };
template <> struct MelBank<3> { typedef melprog::Bank3 type; };

template <class PL>
struct TileCfg {
  // The tile kernel exists for the (n_fft, win_length) pairs of the shipped recipes, each with the straight-line
  // Mel program of the recipe's filterbank (sfe_melprog.h); sfe_plan_create checks that the plan's filterbank has
  // that band structure, every other configuration runs the gather kernel (sfe_fused.cuh).
  static constexpr int BANK = (PL::M == 256 && PL::WIN == 400) ? 1 : (PL::M == 200 && PL::WIN == 400) ? 2
                            : (PL::M == 512 && PL::WIN == 1024) ? 3 : 0;
  static constexpr bool SUPPORTED = (BANK != 0);
  static constexpr int TS = (PL::M <= 256) ? 16 : 8;   // pairs per tile: 2*TS rows, one lane per row in the Mel stage
  static constexpr int NPB = (PL::M <= 256) ? 3 : 2;   // power-spectrum buffers = Mel warps
#ifdef SFE_T2_NF  // tuning experiments (tools/variant.py)
  static constexpr int NF = SFE_T2_NF;
#else
  static constexpr int NF = 16;                        // FFT warps
#endif
  static constexpr int NW = NF + 1 + NPB;              // + producer + Mel warps: <= 5 warps per SM sub-partition, 96 registers
};
constexpr int T2_ROWQ = 8;  // tiles whose row descriptors are kept (the producer runs ahead of the Mel warps)

// Waiting warps must not eat the issue slots of the warps they wait for: sleep between polls.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "WAIT_%=:\n"
      "nanosleep.u32 64;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

struct T2Slot {  // a pair as its FFT warp sees it
  int kind, b, tu, flen;  // frames 2*tu, 2*tu+1 of utterance b, which has flen frames
  int soff, pad_;         // bulk pairs: index of the pair's first sample in the stage buffer
  long long len;          // samples of the utterance (gathered pairs)
};
struct T2Row {  // a pair as the Mel stage sees it
  int b, tA, flen, kind;
};
static_assert(sizeof(T2Slot) == 32 && sizeof(T2Row) == 16, "descriptor layout");
// control block: barriers and counters (128 bytes), slots[2][TS], rows[T2_ROWQ][TS]
constexpr int T2_CTRL_BYTES(int TS) { return 128 + 2 * TS * 32 + T2_ROWQ * TS * 16; }

template <class PL, bool PRE, bool ENERGY>
__global__ void __launch_bounds__(TileCfg<PL>::NW * 32, 1) tile_frontend_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int M = PL::M, G = PL::G, Q = PL::PADQ, TS = TileCfg<PL>::TS, NPB = TileCfg<PL>::NPB, NF = TileCfg<PL>::NF;
  constexpr unsigned FULLM = 0xffffffffu;
  enum { U_DONE = 0, U_PAD = 1, U_BULK = 2, U_EDGE = 3 };
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane / G, l = lane % G;

  float* tab = reinterpret_cast<float*>(smem);
  unsigned char* ctrl = smem + p.t2_ctrl;
  const uint32_t tbar = smem_u32(ctrl);          // tables landed
  const uint32_t full0 = smem_u32(ctrl + 8);     // [2]   stage buffer holds its tile (descriptors + samples)   producer -> FFT
  const uint32_t sfree0 = smem_u32(ctrl + 24);   // [2]   every pair of the tile has read its samples            FFT -> producer
  const uint32_t pfull0 = smem_u32(ctrl + 40);   // [NPB] every pair of the tile has stored its rows             FFT -> Mel
  const uint32_t pfree0 = smem_u32(ctrl + 64);   // [NPB] the Mel stage has consumed the buffer                  Mel -> FFT
  int* s_misc = reinterpret_cast<int*>(ctrl + 88);  // [0] next pair of the CTA, [1] first pair, [2] number of pairs
  // A barrier's phase bit cannot tell tile j from the tile two rounds later on the same buffer, and claims are only
  // a counter — FFT warps may race ahead of a slow one: they also check WHICH tile a buffer holds / may take.
  volatile int* stile = reinterpret_cast<volatile int*>(ctrl + 104);  // [2]   tile whose samples stage buffer sb holds
  volatile int* pgen = stile + 2;                                      // [NPB] tile that may write buffer pb next
  T2Slot* slots = reinterpret_cast<T2Slot*>(ctrl + 128);                // [2][TS]
  T2Row* rows = reinterpret_cast<T2Row*>(ctrl + 128 + 2 * TS * 32);     // [T2_ROWQ][TS]
  float* stage0 = reinterpret_cast<float*>(smem + p.t2_stage);
  float* p2_0 = reinterpret_cast<float*>(smem + p.t2_p2);
  __shared__ int s_cum[UNIT_TABLE_MAX_B + 1];  // cost coordinate at which every utterance starts
  __shared__ int s_pc[UNIT_TABLE_MAX_B + 1];   // pair index at which every utterance starts

  const int hop = p.hop, win = p.win, K = p.K, F = p.F, B = p.B;
  const int upu = p.units_per_utt;
  const int PP = p.t2_pp;
  const bool compact = p.skip_pad && !p.lin && !p.energy;  // pairs past an utterance's last frame are not enumerated
  const bool tabled = (B <= UNIT_TABLE_MAX_B);
  auto frames_of = [&](long long len) -> long long {
    return (p.center ? (len + 2 * (p.n_fft / 2)) - p.n_fft : len - p.n_fft) / hop + 1;
  };
  auto clampF = [&](long long fl) -> int { return static_cast<int>(fl < 0 ? 0 : (fl > F ? F : fl)); };
  constexpr int pre1 = PRE ? 1 : 0;
  constexpr int EDGE_SHIFT = 4;
  const int p2_floats = 2 * TS * PP;

  // ---- one-time set-up ----
  if (tid == 0) {
    s_misc[0] = 0;
    stile[0] = stile[1] = -1;
    mbar_init(tbar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(sfree0 + 8 * i, TS);
    }
    for (int i = 0; i < NPB; ++i) {
      pgen[i] = i;
      mbar_init(pfull0 + 8 * i, TS);
      mbar_init(pfree0 + 8 * i, 1);
    }
    fence_mbar_init();
    const uint32_t b1 = static_cast<uint32_t>(p.o_mtab) * 4u, b2 = static_cast<uint32_t>(p.t2_sched_floats) * 4u;
    mbar_expect_tx(tbar, b1 + b2);
    bulk_g2s(smem_u32(tab), p.tables, b1, tbar);
    if (b2) bulk_g2s(smem_u32(tab + p.o_mtab), p.tables + p.t2_o_sched_g, b2, tbar);
  }
  // the bins behind the spectrum carry zero weights in the Mel program: keep them finite (no emit ever lands there)
  for (int i = tid; i < NPB * 2 * TS * (p.t2_rowf - K); i += blockDim.x) {
    const int r = i / (p.t2_rowf - K), c = i - r * (p.t2_rowf - K);
    p2_0[r * PP + K + c] = 0.f;
  }
  if (blockIdx.x == 0 && p.feat_len)
    for (int b = tid; b < B; b += blockDim.x) p.feat_len[b] = frames_of(__ldg(p.len + b));
  if (warp == 0) {
    if (tabled) {
      // exclusive scans of the utterances' costs (a pair with real frames costs REAL_UNIT_COST, one that only
      // holds rows past the utterance's length 1) and pair counts
      int carry_c = 0, carry_p = 0;
      for (int b0 = 0; b0 < B; b0 += 32) {
        const int b = b0 + lane;
        int c = 0, n = 0;
        if (b < B) {
          const int nr = (clampF(frames_of(__ldg(p.len + b))) + FPW - 1) / FPW;
          n = compact ? nr : upu;
          c = REAL_UNIT_COST * nr + (n - nr);
        }
        int ic = c, ip = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int vc = __shfl_up_sync(FULLM, ic, o), vp = __shfl_up_sync(FULLM, ip, o);
          if (lane >= o) { ic += vc; ip += vp; }
        }
        if (b < B) { s_cum[b] = carry_c + ic - c; s_pc[b] = carry_p + ip - n; }
        carry_c += __shfl_sync(FULLM, ic, 31);
        carry_p += __shfl_sync(FULLM, ip, 31);
      }
      if (lane == 0) { s_cum[B] = carry_c; s_pc[B] = carry_p; }
      __syncwarp();
    }
    if (lane == 0) {
      // this CTA's pairs: [P_lo, P_hi) on the utterance-major pair axis, an equal-cost share
      auto pair_at_cost = [&](int c) -> int {
        int lo = 0, hi = B - 1;  // smallest b whose range ends beyond c
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (s_cum[mid + 1] > c) hi = mid; else lo = mid + 1;
        }
        const int off = c - s_cum[lo];
        const int nr = (clampF(frames_of(__ldg(p.len + lo))) + FPW - 1) / FPW;
        const int tu = (off <= REAL_UNIT_COST * nr) ? (off + REAL_UNIT_COST - 1) / REAL_UNIT_COST
                                                    : nr + (off - REAL_UNIT_COST * nr);
        return s_pc[lo] + tu;
      };
      int plo, phi;
      if (tabled) {
        const long long ctot = s_cum[B];
        plo = pair_at_cost(static_cast<int>(ctot * blockIdx.x / gridDim.x));
        phi = pair_at_cost(static_cast<int>(ctot * (blockIdx.x + 1) / gridDim.x));
      } else {
        const long long tot = static_cast<long long>(B) * upu;
        plo = static_cast<int>(tot * blockIdx.x / gridDim.x);
        phi = static_cast<int>(tot * (blockIdx.x + 1) / gridDim.x);
      }
      s_misc[1] = plo;
      s_misc[2] = phi - plo;
    }
  }
  __syncthreads();  // barrier inits, the tables of the work split and this CTA's range are visible

  const int P_lo = s_misc[1], npairs = s_misc[2];
  const int ntiles = (npairs + TS - 1) / TS;

  if (warp == NF) {
    // =========================================== producer warp ===========================================
    const int stage_floats = p.t2_stage_floats;
    const bool bulk_ok = ((reinterpret_cast<uintptr_t>(p.wav) & 15) == 0) && ((hop & 1) == 0);
    for (int j = 0; j < ntiles; ++j) {
      const int sb = j & 1;
      if (j >= 2) mbar_wait_backoff(sfree0 + 8 * sb, static_cast<uint32_t>(((j >> 1) - 1) & 1));
      // describe tile j, one lane per pair
      const int nsl = min(TS, npairs - j * TS);
      T2Slot sl;
      sl.kind = U_DONE; sl.b = 0; sl.tu = 0; sl.flen = 0; sl.soff = 0; sl.pad_ = 0; sl.len = 0;
      bool elig = false;
      long long ga = 0;
      int need = 0, shiftfull = 0;
      if (lane < nsl) {
        const int P = P_lo + j * TS + lane;
        if (tabled) {
          int lo = 0, hi = B - 1;  // smallest b whose pairs end beyond P
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_pc[mid + 1] > P) hi = mid; else lo = mid + 1;
          }
          sl.b = lo;
          sl.tu = P - s_pc[lo];
        } else {
          sl.b = P / upu;
          sl.tu = P - sl.b * upu;
        }
        sl.len = __ldg(p.len + sl.b);
        sl.flen = clampF(frames_of(sl.len));
        sl.kind = U_PAD;
        if (sl.tu * FPW < sl.flen) {
          sl.kind = U_EDGE;
          const long long o_start = static_cast<long long>(sl.tu) * FPW * hop + p.off;
          const long long o_end = o_start + static_cast<long long>(FPW - 1) * hop + win;
          const long long gi = static_cast<long long>(sl.b) * p.pitch + o_start - pre1;
          ga = gi & ~3LL;
          shiftfull = static_cast<int>(gi - ga) + pre1;
          need = (shiftfull + (FPW - 1) * hop + win + 3) & ~3;  // floats from ga that cover the pair
          elig = bulk_ok && (o_start - pre1 >= 0) && (o_end <= sl.len) && (ga + need <= p.total_floats);
        }
      }
      // bulk pairs: ONE contiguous run of pairs of ONE utterance per tile (their samples are contiguous in HBM
      // and in the stage buffer: pair s at s * 2 hop); everything else is gathered by its FFT warp
      unsigned m = __ballot_sync(FULLM, elig);
      uint32_t bytes = 0;
      if (m) {
        const int f = __ffs(m) - 1;
        const int bstar = __shfl_sync(FULLM, sl.b, f);
        elig = elig && (sl.b == bstar);
        m = __ballot_sync(FULLM, elig);
        const unsigned run = m & ~(m + (1u << f));
        elig = (run >> lane) & 1u;
        if (elig) {
          const int last = 31 - __clz(run);
          sl.kind = U_BULK;
          sl.soff = lane * FPW * hop + shiftfull;
          bytes = static_cast<uint32_t>((lane == last ? need : FPW * hop) * 4);
        }
      }
      uint32_t total = bytes;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(FULLM, total, o);
      if (lane < nsl) {
        slots[sb * TS + lane] = sl;
        T2Row rd;
        rd.b = sl.b; rd.tA = sl.tu * FPW; rd.flen = sl.flen; rd.kind = sl.kind;
        rows[(j & (T2_ROWQ - 1)) * TS + lane] = rd;
      }
      __syncwarp();
      if (lane == 0) {
        stile[sb] = j;
        mbar_expect_tx(full0 + 8 * sb, total);  // arrive (release): the descriptors are visible to whoever sees the phase flip
      }
      __syncwarp();
      if (elig) {
        fence_proxy_async();  // the buffer's previous readers are done (sfree): order that before the async-proxy writes
        bulk_g2s(smem_u32(stage0 + sb * stage_floats + lane * FPW * hop), p.wav + ga, bytes, full0 + 8 * sb);
      }
    }
    return;
  }

  mbar_wait(tbar, 0u);  // tables have landed

  if (warp > NF) {
    // ============================== Mel warps: warp NF + 1 + pb owns buffer pb ==============================
    const int pb = warp - NF - 1;
    float* p2buf = p2_0 + pb * p2_floats;
    const float4* wprog = reinterpret_cast<const float4*>(tab + p.o_mtab);  // Mel weights in the program's layout
    const int n_fmask = p.mk.n_fmask, n_tmask = p.mk.n_tmask;
    const int* m_fpos = p.mk.inl_on ? p.mk.inl : p.mk.fpos;
    const int* m_flen = p.mk.inl_on ? p.mk.inl + B * n_fmask : p.mk.flen;
    const int* m_tpos = p.mk.inl_on ? p.mk.inl + 2 * B * n_fmask : p.mk.tpos;
    const int* m_tlen = p.mk.inl_on ? p.mk.inl + 2 * B * n_fmask + B * n_tmask : p.mk.tlen;
    for (int j = pb; j < ntiles; j += NPB) {
      mbar_wait_backoff(pfull0 + 8 * pb, static_cast<uint32_t>((j / NPB) & 1));
      const int nsl = min(TS, npairs - j * TS);
      const T2Row* trow = rows + (j & (T2_ROWQ - 1)) * TS;
      typedef typename MelBank<TileCfg<PL>::BANK>::type BK;
      static_assert(BK::D == 80, "the read-out below is laid out for 80 Mel bins");
      const int nrows = 2 * nsl;
      bool va = false, ex = false, tm = false;  // the row is a real frame / is written / is time-masked
      int myb = 0, myt = 0;
      if (lane < nrows) {
        const T2Row rd = trow[lane >> 1];
        myb = rd.b;
        myt = rd.tA + (lane & 1);
        va = (rd.kind != U_PAD) && (myt < rd.flen);
        ex = (myt < F) && !(p.skip_pad && !va);
        for (int qq = 0; qq < n_tmask; ++qq) {
          const int ps = m_tpos[myb * n_tmask + qq], ln = m_tlen[myb * n_tmask + qq];
          tm |= (myt >= ps) && (myt < ps + ln);
        }
      }
      const unsigned va_bits = __ballot_sync(FULLM, va), ex_bits = __ballot_sync(FULLM, ex);
      const unsigned tm_bits = __ballot_sync(FULLM, tm);

      // ---- linear spectrogram output (before the Mel stage overwrites the rows) ----
      if (p.lin) {
        for (int r = 0; r < nrows; ++r) {
          const int rb = __shfl_sync(FULLM, myb, r), rt = __shfl_sync(FULLM, myt, r);
          if (rt < F) {
            float* dst = p.lin + (static_cast<long long>(rb) * F + rt) * K;
            const bool v = (va_bits >> r) & 1u;
            const float* src = p2buf + r * PP;
            for (int k = lane; k < K; k += 32) {
              float x = v ? src[k] : 0.f;
              if (p.lin_log) x = fast_log(x, p.lin_clamp, p.lin_lmul);
              dst[k] = x;
            }
          }
        }
        __syncwarp();
      }

      // ---- banded Mel: the recipe filterbank's straight-line program (sfe_melprog.h) ----
      // One packed FMA per bin into (even filter, odd filter); a finished filter's raw sum is stored over a bin
      // the lane has already consumed (constant address, no flag test, no branch); two independent streams
      // (bins below / above B_LD) for instruction-level parallelism, joined at the end.  The logarithm is
      // taken by the read-out.  (A data-driven version of this loop — flag tests and branches around the
      // emits — cost 5 900 instructions and ~30 us of one warp per tile, the whole CTA waiting for it.)
      if (lane < nrows) {
        float* rowp = p2buf + lane * PP;
        const float4* pA = reinterpret_cast<const float4*>(rowp);
        const float4* pB = reinterpret_cast<const float4*>(rowp + BK::B_LD);
        float4 xA[2], xB[2], wA0[2], wA1[2], wB0[2], wB1[2];
        float2 aA = make_float2(0.f, 0.f), aB = make_float2(0.f, 0.f);
#define SFE_LD(g, buf)                                                                            \
  xA[buf] = pA[g]; xB[buf] = pB[g];                                                               \
  wA0[buf] = wprog[4 * (g)]; wA1[buf] = wprog[4 * (g) + 1]; wB0[buf] = wprog[4 * (g) + 2]; wB1[buf] = wprog[4 * (g) + 3];
#define SFE_W(u_, v_, i) ((i) == 0 ? make_float2(u_.x, u_.y) : (i) == 1 ? make_float2(u_.z, u_.w) : (i) == 2 ? make_float2(v_.x, v_.y) : make_float2(v_.z, v_.w))
#define SFE_X(v_, i) ((i) == 0 ? v_.x : (i) == 1 ? v_.y : (i) == 2 ? v_.z : v_.w)
#define SFE_FA(buf, i) aA = f2fma(SFE_W(wA0[buf], wA1[buf], i), make_float2(SFE_X(xA[buf], i), SFE_X(xA[buf], i)), aA);
#define SFE_FB(buf, i) aB = f2fma(SFE_W(wB0[buf], wB1[buf], i), make_float2(SFE_X(xB[buf], i), SFE_X(xB[buf], i)), aB);
#define SFE_EA(slot, pos) if (slot) { rowp[pos] = aA.y; aA.y = 0.f; } else { rowp[pos] = aA.x; aA.x = 0.f; }
#define SFE_EB(slot, pos) if (slot) { rowp[pos] = aB.y; aB.y = 0.f; } else { rowp[pos] = aB.x; aB.x = 0.f; }
#define SFE_JOIN(slot, pos) rowp[pos] += (slot) ? aA.y : aA.x; rowp[(pos) + 1] += (slot) ? aA.x : aA.y;
#define SFE_MOVE(from, to) rowp[to] = rowp[from];
        if constexpr (TileCfg<PL>::BANK == 1) { SFE_MELPROG_1(SFE_LD, SFE_FA, SFE_FB, SFE_EA, SFE_EB, SFE_JOIN, SFE_MOVE) }
        else if constexpr (TileCfg<PL>::BANK == 2) { SFE_MELPROG_2(SFE_LD, SFE_FA, SFE_FB, SFE_EA, SFE_EB, SFE_JOIN, SFE_MOVE) }
        else { SFE_MELPROG_3(SFE_LD, SFE_FA, SFE_FB, SFE_EA, SFE_EB, SFE_JOIN, SFE_MOVE) }
#undef SFE_LD
#undef SFE_W
#undef SFE_X
#undef SFE_FA
#undef SFE_FB
#undef SFE_EA
#undef SFE_EB
#undef SFE_JOIN
#undef SFE_MOVE
      }
      __syncwarp();

      // ---- read-out: two rows per step, 16 lanes per row; lane q owns Mel bins 4q..4q+3 and 64+q ----
      if (p.mel) {
        const int h = lane >> 4, q = lane & 15;
        const int so4 = (q < BK::C_SPLIT) ? 4 * q : BK::B_OFF + 4 * q;
        const int so1 = (64 + q < 4 * BK::C_SPLIT) ? 64 + q : BK::B_OFF + 64 + q;
        const bool fancy = (p.norm_mean != nullptr) || (p.norm_std != nullptr) || (n_fmask > 0) || (n_tmask > 0);
        float nm[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, rs[5] = {1.f, 1.f, 1.f, 1.f, 1.f};
        float fillv = 0.f;
        if (fancy) {
          fillv = p.fill_dev ? *p.fill_dev : p.fill;
#pragma unroll
          for (int e = 0; e < 5; ++e) {
            const int d = (e < 4) ? 4 * q + e : 64 + q;
            if (p.norm_mean) nm[e] = p.norm_mean[d];
            if (p.norm_std) rs[e] = 1.f / p.norm_std[d];
          }
        }
        const float padv = p.pad_mel, clampv = p.mel_clamp, lmul = p.mel_lmul;
        const bool dolog = p.mel_log != 0;
        const bool stats_on = (p.stats != nullptr);
        // per-utterance sums of this tile: fp32 in a fixed order (rows ascending inside each half-warp, then
        // the halves), converted ONCE to the exact, order-independent fixed point of the global accumulators —
        // run-to-run reproducible for a given batch and grid
        float s1[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, s2[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        auto flush = [&](int bb) {
#pragma unroll
          for (int e = 0; e < 5; ++e) {
            s1[e] += __shfl_xor_sync(FULLM, s1[e], 16);
            s2[e] += __shfl_xor_sync(FULLM, s2[e], 16);
          }
          if (h == 0) {
            unsigned long long* base = p.stats + static_cast<long long>(bb) * BK::D * 2;
#pragma unroll
            for (int e = 0; e < 5; ++e) {
              const int d = (e < 4) ? 4 * q + e : 64 + q;
              atomicAdd(base + 2 * d, static_cast<unsigned long long>(__float2ll_rn(s1[e] * 4294967296.f)));
              atomicAdd(base + 2 * d + 1, static_cast<unsigned long long>(__float2ll_rn(s2[e] * 16777216.f)));
            }
          }
#pragma unroll
          for (int e = 0; e < 5; ++e) s1[e] = s2[e] = 0.f;
        };
        int curb = -1;
        unsigned fm = 0u;
        for (int it = 0; it < nsl; ++it) {
          const T2Row rd = trow[it];
          if (rd.b != curb) {
            if (curb >= 0 && stats_on) flush(curb);
            curb = rd.b;
            fm = 0u;
            for (int qq = 0; qq < n_fmask; ++qq) {
              const int ps = m_fpos[curb * n_fmask + qq], ln = m_flen[curb * n_fmask + qq];
#pragma unroll
              for (int e = 0; e < 5; ++e) {
                const int d = (e < 4) ? 4 * q + e : 64 + q;
                if ((d >= ps) && (d < ps + ln)) fm |= 1u << e;
              }
            }
          }
          const int r = 2 * it + h;
          const bool v = (va_bits >> r) & 1u, w = (ex_bits >> r) & 1u;
          if (!w) continue;
          float x[5] = {padv, padv, padv, padv, padv};
          if (v) {
            const float4 x4 = *reinterpret_cast<const float4*>(p2buf + r * PP + so4);
            x[0] = x4.x; x[1] = x4.y; x[2] = x4.z; x[3] = x4.w;
            x[4] = p2buf[r * PP + so1];
            if (dolog) {
#pragma unroll
              for (int e = 0; e < 5; ++e) x[e] = fast_log(x[e], clampv, lmul);
            }
            if (stats_on) {
#pragma unroll
              for (int e = 0; e < 5; ++e) {
                s1[e] += x[e];
                s2[e] = fmaf(x[e], x[e], s2[e]);
              }
            }
          }
          if (fancy) {
            const bool tmr = (tm_bits >> r) & 1u;
#pragma unroll
            for (int e = 0; e < 5; ++e) {
              x[e] = (x[e] - nm[e]) * rs[e];
              if (tmr || (fm & (1u << e))) x[e] = fillv;
            }
          }
          float* dst = p.mel + (static_cast<long long>(rd.b) * F + rd.tA + h) * BK::D;
          *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(x[0], x[1], x[2], x[3]);
          dst[64 + q] = x[4];
        }
        if (curb >= 0 && stats_on) flush(curb);
      }

      // hand the buffer back: tile j + NPB may write its rows
      __syncwarp();
      if (lane == 0) {
        pgen[pb] = j + NPB;
        __threadfence_block();
        mbar_arrive(pfree0 + 8 * pb);
      }
    }
    return;
  }

  // =============================== FFT warps: one pair per iteration ===============================
  unsigned char* wbase = smem + p.t2_warp + warp * p.t2_warp_bytes;
  cpx* Zw = reinterpret_cast<cpx*>(wbase) + g * PL::MP;
  float* escratch = reinterpret_cast<float*>(wbase);  // gathered pairs stage their samples in the FFT scratch
  const float* wtab = tab + p.o_win;
  const cpx* tw2 = reinterpret_cast<const cpx*>(tab + p.o_tw2);
  const cpx* tw3 = reinterpret_cast<const cpx*>(tab + p.o_tw3);
  const cpx* wpost = reinterpret_cast<const cpx*>(tab + p.o_wpost);
  const int stage_floats = p.t2_stage_floats;
  const int qend = ntiles * TS;  // (the last tile's missing pairs still count for its barriers)

  while (true) {
    int q = 0;
    if (lane == 0) q = atomicAdd(&s_misc[0], 1);
    q = __shfl_sync(FULLM, q, 0);
    if (q >= qend) break;
    const int j = q / TS, s = q - j * TS, sb = j & 1, pb = j % NPB;
    // rows of buffer pb may be overwritten once the Mel stage of tile j - NPB has consumed them
    auto p2_wait = [&]() {
      while (true) {
        mbar_wait_backoff(pfree0 + 8 * pb, static_cast<uint32_t>(((j / NPB) & 1) ^ 1));
        if (pgen[pb] == j) break;
        __nanosleep(200);  // the phase that completed belongs to an earlier tile of this buffer
      }
    };
    auto publish = [&]() {
      __syncwarp();
      if (lane == 0) mbar_arrive(pfull0 + 8 * pb);
    };
    if (q >= npairs) {  // padding of the CTA's last tile
      p2_wait();
      publish();
      continue;
    }
    while (true) {
      mbar_wait_backoff(full0 + 8 * sb, static_cast<uint32_t>((j >> 1) & 1));
      if (stile[sb] == j) break;
      __nanosleep(200);
    }
    const T2Slot sl = slots[sb * TS + s];
    const int kind = sl.kind, b = sl.b, flen = sl.flen, tA = sl.tu * FPW;
    float* p2buf = p2_0 + pb * p2_floats;
    // every lane is past its last read of the stage buffer and of the slot
    auto stage_release = [&]() {
      __syncwarp();
      if (lane == 0) mbar_arrive(sfree0 + 8 * sb);
    };

    if (kind == U_PAD) {
      stage_release();
      p2_wait();
      publish();
      continue;
    }
    // ---------------- samples of this pair ----------------
    const bool bulk = (kind == U_BULK);
    const float* sbase = bulk ? (stage0 + sb * stage_floats) : escratch;
    auto gather = [&](int t0, int n) {
      // explicit gather with batch-level reflection, zero beyond the length and pre-emphasis applied here
      const long long len = sl.len;
      const long long o_start = static_cast<long long>(t0) * hop + p.off;
      const float* row = p.wav + static_cast<long long>(b) * p.pitch;
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
          if (i < n) escratch[EDGE_SHIFT + i] = fmaf(-p.pre, pv[u], v[u]);
        }
      }
      if (lane == 0) escratch[EDGE_SHIFT - 1] = 0.f;  // read (times a zero coefficient) as the predecessor of sample 0
      __syncwarp();
    };
    if (!bulk && PL::S == 1) gather(tA, (FPW - 1) * hop + win);

    float esum[PL::S];
#pragma unroll
    for (int sub = 0; sub < PL::S; ++sub) {
      const int fslot = sub * PL::C + g;  // frame index inside the pair
      const int t = tA + fslot;
      if (!bulk && PL::S > 1) {
        if (sub > 0) __syncwarp();  // the previous frame's transform is out of the scratch
        gather(t, win);
      }
      const int fo = bulk ? (fslot * hop + sl.soff) : ((PL::S == 1 ? fslot * hop : 0) + EDGE_SHIFT);
      const float* fr = sbase + fo;
      const float pre = (PRE && bulk) ? p.pre : 0.f;  // gathered pairs were pre-emphasised by the gather
      if ((fo & 1) == 0) stockham_pass<PL, PL::R1, 1, true, true, PRE, true>(Zw, nullptr, l, fr, wtab, win, pre);
      else stockham_pass<PL, PL::R1, 1, true, false, PRE, true>(Zw, nullptr, l, fr, wtab, win, pre);
      if (sub == PL::S - 1) stage_release();
      fft_rest<PL>(Zw, tw2, tw3, l);

      // real-FFT unpack: X[k] = E + W^k O, X[M-k] = conj(E - W^k O); the 1/2 lives in the window.
      constexpr int RL = PL::RL;
      constexpr int NI = PL::LASTREG ? (RL / 2 + 1) : ((M / 2) / G + 1);
      cpx za[NI], zb[NI];
      if constexpr (PL::LASTREG) {
        cpx u[RL];
        stockham_last_reg<PL, RL, M / RL>(Zw, (PL::R3 > 1) ? tw3 : tw2, l, u);
        const int srcl = (G - l) & (G - 1);
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          za[i] = u[i];
          const float bx = __shfl_sync(FULLM, u[RL - 1 - i].x, srcl, G);
          const float by = __shfl_sync(FULLM, u[RL - 1 - i].y, srcl, G);
          zb[i] = (l == 0) ? u[(RL - i) % RL] : make_float2(bx, by);
        }
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
      }
      if (sub == 0) p2_wait();
      float* row = p2buf + (2 * s + fslot) * PP;
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
          const float pk = fmaf(X.x, X.x, X.y * X.y), pm = fmaf(Y.x, Y.x, Y.y * Y.y);
          if (ENERGY) {
            es += pk;
            if (2 * k != M) es += pm;
          }
          row[k] = pk;
          row[M - k] = pm;
        }
      }
      if (p.mag) {  // magnitude spectrum (TTS recipes): one sweep over the frame's row, off the power path
        __syncwarp();
        for (int i = l; i < K; i += G) row[i] = sqrtf(row[i]);
      }
      esum[sub] = es;
      __syncwarp();  // (S > 1: the scratch is free for the next frame)
    }

    // ---------------- energy (sum of the power spectrum of each frame) ----------------
    if (ENERGY && p.energy) {
#pragma unroll
      for (int sub = 0; sub < PL::S; ++sub) {
        float es = esum[sub];
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) es += __shfl_xor_sync(FULLM, es, o);
        const int t = tA + sub * PL::C + g;
        if (l == 0 && t < F) p.energy[static_cast<long long>(b) * F + t] = (t < flen) ? sqrtf(fmaxf(es, 1e-10f)) : 0.f;
      }
    }
    publish();
  }
}

}  // namespace sfe
