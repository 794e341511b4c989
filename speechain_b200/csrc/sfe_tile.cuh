// Tile variant of the fused waveform -> (log-)mel kernel for sm_100a: the kernel of the shipped recipes.
//
// Same arithmetic as sfe_fused.cuh up to the power spectrum (window, pre-emphasis, Stockham FFT, real-FFT
// unpack).  What changes is WHO does the Mel stage: one CTA per SM, warp-specialised, every hand-over an
// mbarrier, no CTA-wide barrier after the set-up:
//
//   FFT warps    (NF) claim the CTA's 2-frame units ("pairs") one by one from a shared counter.  A warp stages the
//                samples of its NEXT pair itself — ONE 1-D bulk async copy (cp.async.bulk -> UBLKCP) into its
//                private staging buffer, issued right after the first FFT pass of the current pair has
//                consumed the buffer, so that it lands under the rest of the pair; pairs that touch the
//                batch-level reflection or the end of their utterance are gathered by the warp instead.
//                Window (+ pre-emphasis) -> FFT -> unpack -> |X|^2 goes into ROWS 2s, 2s+1 of a CTA-wide TILE of
//                TS pairs (pair q of the CTA: tile q / TS, slot s = q % TS, buffer tile % NPB).
//                (A shared ring of stage buffers filled by a producer warp was measured first: the round trip
//                buffer drained -> producer woken -> copy issued -> samples landed -> FFT warps woken took
//                longer than the FFT of the buffered pairs, 33 us for C1 with the Mel stage switched off.)
//   Mel warps    (NPB, one per tile buffer) take a complete tile and run the Mel stage for all 2*TS frames at
//                once with ONE LANE PER FRAME: the lane walks the bins of its row in order with one packed FMA
//                per bin into an (even filter, odd filter) accumulator pair and stores a filter when its last
//                bin is in — straight-line code generated for the recipe's filterbank (sfe_melprog.h): no flag
//                test, no branch, no address arithmetic; the weights are one broadcast 128-bit load per two
//                bins; no lane idles on a narrow filter.  The gather-style Mel stage of sfe_fused.cuh spent 290
//                of its 960 instructions per pair here, ~90 % of them address arithmetic and predicates; this
//                one ~50.  Then the read-out: two rows per step, 16 lanes per row, logarithm, per-utterance
//                sums, optional (x-mean)/std and SpecAugment masks, 128-bit coalesced stores.
//
// Reference arithmetic replaced: speechain/module/frontend/speech2linear.py:165-245, linear2mel.py:177-187.
#pragma once
#include "sfe_fused.cuh"
#include "sfe_melprog.h"

// Timing experiments only (tools/variant.py), never defined in the shipped build: 1 Mel warps hand their tiles back
// untouched, 2 additionally the FFT warps skip their transform, 3 FFT warps skip the transform but the Mel stage runs.
#ifndef SFE_T2_ABLATE
#define SFE_T2_ABLATE 0
#endif

namespace sfe {

// Development only (tools/trace_tile.py): -DSFE_T2_TRACE records (tag, warp, index, globaltimer) events of CTA 0.
#ifdef SFE_T2_TRACE
__device__ unsigned long long g_t2_trace_t[1 << 15];
__device__ unsigned int g_t2_trace_tag[1 << 15];
__device__ unsigned int g_t2_trace_n;
__device__ __forceinline__ void t2_trace(int tag, int idx) {
  if (blockIdx.x != 0 || (threadIdx.x & 31) != 0) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  const unsigned int i = atomicAdd(&g_t2_trace_n, 1u);
  if (i < (1u << 15)) {
    g_t2_trace_t[i] = t;
    g_t2_trace_tag[i] = (static_cast<unsigned>(tag) << 24) | ((threadIdx.x >> 5) << 16) | (static_cast<unsigned>(idx) & 0xffffu);
  }
}
#define SFE_TR(tag, idx) t2_trace(tag, idx)
// ... and when every warp of every CTA entered / left the kernel (slot = 2 * (32 * cta + warp))
__device__ unsigned long long g_t2_span[2 * 32 * 256];
__device__ __forceinline__ void t2_span(int which) {
  if ((threadIdx.x & 31) != 0 || blockIdx.x >= 256) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  g_t2_span[2 * (32 * blockIdx.x + (threadIdx.x >> 5)) + which] = t;
}
#define SFE_SPAN(which) t2_span(which)
#else
#define SFE_TR(tag, idx)
#define SFE_SPAN(which)
#endif
// -DSFE_T2_PROF: per-warp cycle totals of the phases of CTA 0 (registers only while running)
#ifdef SFE_T2_PROF
__device__ long long g_t2_prof[32 * 16];
#define SFE_PF_DECL long long pf_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long pf_t = clock64();
#define SFE_PF(k) { const long long t_ = clock64(); pf_acc[k] += t_ - pf_t; pf_t = t_; }
#define SFE_PF_DUMP(w) if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { for (int k_ = 0; k_ < 8; ++k_) g_t2_prof[(w) * 16 + k_] = pf_acc[k_]; \
  unsigned long long te_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(te_)); g_t2_prof[(w) * 16 + 8] = pf_g0; g_t2_prof[(w) * 16 + 9] = pf_g1; g_t2_prof[(w) * 16 + 10] = pf_g2; g_t2_prof[(w) * 16 + 11] = te_; }
#define SFE_PF_G(v) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); v = t_; }
#else
#define SFE_PF_DECL
#define SFE_PF(k)
#define SFE_PF_DUMP(w)
#define SFE_PF_G(v)
#endif

template <int BANK> struct MelBank { typedef void type; };
template <> struct MelBank<1> { typedef melprog::Bank1 type; };
template <> struct MelBank<2> { typedef melprog::Bank2 type; };

template <class PL>
struct TileCfg {
  // The tile kernel exists for the (n_fft, win_length) pairs of the shipped recipes, each with the straight-line
  // Mel program of the recipe's filterbank (sfe_melprog.h); sfe_plan_create checks that the plan's filterbank has
  // that band structure, every other configuration runs the gather kernel (sfe_fused.cuh).
  // (n_fft 1024 — fewer, fatter warps — measured slower than the gather kernel on this design: 285 vs 232 us for C3)
  static constexpr int BANK = (PL::M == 256 && PL::WIN == 400) ? 1 : (PL::M == 200 && PL::WIN == 400) ? 2 : 0;
  static constexpr bool SUPPORTED = (BANK != 0);
  static constexpr int TS = (PL::M <= 256) ? 16 : 8;   // pairs per tile: 2*TS rows, one lane per row in the Mel stage
#ifdef SFE_T2_NPB  // tuning experiments (tools/variant.py)
  static constexpr int NPB = SFE_T2_NPB;
#else
  static constexpr int NPB = (PL::M <= 256) ? 3 : 2;   // power-spectrum buffers = Mel warps
#endif
#ifdef SFE_T2_MS  // tuning experiments (tools/variant.py)
  static constexpr int MS = SFE_T2_MS;
#else
  // Mel warps per tile buffer.  2: the Mel program's two streams (bins below / above B_LD) run on two warps and each
  // reads out half of the rows — what matters is the LATENCY of a tile's Mel stage, not its instruction count: it is
  // the tail of every launch (the last tile's Mel stage runs on an otherwise idle SM) and the depth of the
  // FFT -> Mel pipeline (FFT warps wait for a buffer only the Mel stage can free).
  static constexpr int MS = 2;
#endif
#ifdef SFE_T2_NF  // tuning experiments (tools/variant.py)
  static constexpr int NF = SFE_T2_NF;
#else
  // FFT warps (measured, C1 / C2 us with two Mel warps per buffer: 13: 34.8 / 53.5, 14: 34.3 / 53.6, 15: 34.1 / 54.9,
  // 16: 33.9 / 53.5; 16 + 6 warps run at 80 registers)
#ifdef SFE_T2_NF11  // tuning experiments (tools/variant.py)
  static constexpr int NF = (PL::M == 256) ? 16 : SFE_T2_NF11;
#else
  // (n_fft 400 — radix 8 x 5 x 5, three passes — is short of registers, not of warps: measured C1-shaped 67.3 / 62.2 /
  //  61.3 / 56.5 us with 16 / 14 / 12 / 10 FFT warps)
  static constexpr int NF = (PL::M == 256) ? 16 : 10;
#endif
#endif
  // Two frames in flight per warp (G = 16) store their rows at the same time: rows 16 banks apart (4 rows with the
  // pitch below) make that store conflict-free, so pair s owns rows 8 (s / 4) + s % 4 and that + 4
  static constexpr bool PERM = (PL::C == 2);
  static_assert(NPB >= 2 && NPB <= 4, "control block layout (T2_CTRL_BYTES) and named barrier ids");
  static constexpr int NW = NF + MS * NPB;             // + Mel warps: <= 20 warps = 5 per SM sub-partition at <= 96 registers
  static_assert(MS == 1 || (MS == 2 && (TS / 2) % 4 == 0), "a Mel warp reads out whole groups of four pairs");
};
// Row pitch of a power-spectrum tile: the spectrum plus the zero-weight bins its Mel program rounds up to, an ODD
// multiple of 16 bytes (one lane per row reads 128 bits at a time: conflict-free).
__host__ __device__ constexpr int t2_row_pitch(int rowf) { return (((rowf + 3) / 4) & 1) ? (rowf + 3) / 4 * 4 : (rowf + 3) / 4 * 4 + 4; }

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

// acquire / release on a shared-memory word (a membar.cta — what __threadfence_block() is — is a sequentially
// consistent fence: it drains every outstanding access of the warp)
__device__ __forceinline__ int ld_acquire_shared(const volatile int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(const_cast<const int*>(p))) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_shared(volatile int* p, int v) {
  asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(smem_u32(const_cast<int*>(p))), "r"(v) : "memory");
}

// named barriers between the two Mel warps of a tile buffer (whole warps arrive: .aligned)
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <bool PERM> __device__ __forceinline__ int t2_row(int s, int f) {  // tile row of frame f of pair s
  return PERM ? (8 * (s >> 2) + (s & 3) + 4 * f) : (2 * s + f);
}
struct T2Row {  // a pair as the Mel stage sees it: frames tA, tA+1 of utterance b, which has flen frames
  int b, tA, flen, kind;
};
static_assert(sizeof(T2Row) == 16, "descriptor layout");
// control block: barriers and counters (96 bytes), rows[NPB <= 4][TS]
constexpr int T2_MAX_NPB = 4;
constexpr int T2_CTRL_BYTES(int TS) { return 96 + T2_MAX_NPB * TS * 16; }

// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase flips (or ~1 us passes) — it
// neither eats the issue slots of the warps it waits for nor adds a polling interval to the hand-over
#ifndef SFE_T2_FULLBAR
#define SFE_T2_FULLBAR 1
#endif
#ifndef SFE_T2_WAIT  // tuning experiments: 0 plain try_wait loop, 1 sleep between polls, 2 hardware suspend hint
#define SFE_T2_WAIT 2
#endif
__device__ __forceinline__ void mbar_wait_hint(uint32_t bar, uint32_t parity) {
  if (SFE_T2_WAIT == 0) { mbar_wait(bar, parity); return; }
  if (SFE_T2_WAIT == 1) { mbar_wait_backoff(bar, parity); return; }
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 1000;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

template <class PL, bool PRE, bool ENERGY>
__global__ void __launch_bounds__(TileCfg<PL>::NW * 32, 1) tile_frontend_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int M = PL::M, G = PL::G, Q = PL::PADQ, TS = TileCfg<PL>::TS, NPB = TileCfg<PL>::NPB, NF = TileCfg<PL>::NF;
  constexpr unsigned FULLM = 0xffffffffu;
  constexpr bool PERM = TileCfg<PL>::PERM;
  enum { U_DONE = 0, U_PAD = 1, U_BULK = 2, U_EDGE = 3 };
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane / G, l = lane % G;

  long long pf_g0 = 0, pf_g1 = 0, pf_g2 = 0;  // (profile builds: kernel entry, set-up done, tables landed)
  SFE_PF_G(pf_g0)
  SFE_SPAN(0);
  SFE_TR(1, 0);
  float* tab = reinterpret_cast<float*>(smem);
  unsigned char* ctrl = smem + p.t2_ctrl;
  const uint32_t tbar = smem_u32(ctrl);          // tables landed
  const uint32_t pfull0 = smem_u32(ctrl + 8);    // [NPB] every pair of the tile has stored its rows    FFT -> Mel
  int* s_misc = reinterpret_cast<int*>(ctrl + 56);  // [0] next pair of the CTA, [1] first pair, [2] number of pairs
  // [NPB] the tile that may write buffer pb next (Mel -> FFT): a counter rather than a barrier, because claims are
  // only a counter too — an FFT warp may race two rounds ahead of a slow one, which a phase bit cannot tell apart
  volatile int* pgen = reinterpret_cast<volatile int*>(ctrl + 72);
  T2Row* rows = reinterpret_cast<T2Row*>(ctrl + 96);                  // [NPB][TS]
  float* p2_0 = reinterpret_cast<float*>(smem + p.t2_p2);
  __shared__ int s_cum[UNIT_TABLE_MAX_B + 1];  // cost coordinate at which every utterance starts
  __shared__ int s_pc[UNIT_TABLE_MAX_B + 1];   // pair index at which every utterance starts
  __shared__ int s_len[UNIT_TABLE_MAX_B];      // samples of every utterance (the host keeps longer inputs on the gather kernel)
  __shared__ int s_flen[UNIT_TABLE_MAX_B];     // frames of every utterance

  const int hop = p.hop, win = p.win, K = p.K, F = p.F, B = p.B;
  const int upu = p.units_per_utt;
  constexpr int PP = t2_row_pitch(MelBank<TileCfg<PL>::BANK>::type::ROWF);
  const bool tabled = (B <= UNIT_TABLE_MAX_B);
  // Pairs past an utterance's last frame are not enumerated: those rows hold a constant, which the FFT warps write
  // straight to global memory once they have run out of pairs (`padfill`, below) — or nobody does, when a normalise
  // pass follows that synthesises them (`skip_pad`).  Only batches beyond the tables still send them through the tiles.
  const bool compact = p.skip_pad || tabled;
  const bool padfill = !p.skip_pad && tabled;
  int* s_padc = s_cum;  // (after the set-up) row index at which every utterance's rows past its pairs start
  __shared__ int s_pad[2];  // [0] next chunk of those rows, [1] s_padc is ready
  auto frames_of = [&](long long len) -> long long {
    return (p.center ? (len + 2 * (p.n_fft / 2)) - p.n_fft : len - p.n_fft) / hop + 1;
  };
  auto clampF = [&](long long fl) -> int { return static_cast<int>(fl < 0 ? 0 : (fl > F ? F : fl)); };
  constexpr int pre1 = PRE ? 1 : 0;
  constexpr int EDGE_SHIFT = 4;
  const int p2_floats = 2 * TS * PP;
  // hand-over of a full tile through a hardware named barrier (TS arriving pair warps + the buffer's Mel warps) instead
  // of an mbarrier the Mel warps poll
  constexpr int FULLBAR_ID0 = 1 + 3 * T2_MAX_NPB;
  constexpr bool FULLBAR = (SFE_T2_FULLBAR != 0) && (FULLBAR_ID0 + NPB <= 16) && (TS + TileCfg<PL>::MS <= TileCfg<PL>::NW);

  // ---- one-time set-up ----
  if (tid == 0) {
    s_misc[0] = 0;
    s_pad[0] = 0;
    s_pad[1] = 0;
    mbar_init(tbar, 1);
    for (int i = 0; i < NPB; ++i) {
      pgen[i] = i;
      mbar_init(pfull0 + 8 * i, TS);
    }
    fence_mbar_init();
    const uint32_t b1 = static_cast<uint32_t>(p.o_mtab) * 4u, b2 = static_cast<uint32_t>(p.t2_sched_floats) * 4u;
    mbar_expect_tx(tbar, b1 + b2);
    bulk_g2s(smem_u32(tab), p.tables, b1, tbar);
    if (b2) bulk_g2s(smem_u32(tab + p.o_mtab), p.tables + p.t2_o_sched_g, b2, tbar);
  }
  pdl_launch_dependents();
  // the bins behind the spectrum carry zero weights in the Mel program: keep them finite (no emit ever lands there)
  for (int i = tid; i < NPB * 2 * TS * (p.t2_rowf - K); i += blockDim.x) {
    const int r = i / (p.t2_rowf - K), c = i - r * (p.t2_rowf - K);
    p2_0[r * PP + K + c] = 0.f;
  }
  if (warp < NF && lane == 0) {  // every FFT warp's "samples landed" barrier
    mbar_init(smem_u32(smem + p.t2_warp + warp * p.t2_warp_bytes + p.t2_stage), 1);
    fence_mbar_init();
  }
  if (p.hs_on) {
    // The host knew the lengths and cut the batch: tables and this CTA's range come with the parameter block —
    // nothing here waits for global memory (the length load -> scan -> search chain below is ~2 us of every launch)
    for (int b = tid; b < B; b += blockDim.x) {
      const int len = p.hs_len[b];
      s_len[b] = len;
      const int fl = (p.center ? (len + 2 * (p.n_fft / 2)) - p.n_fft : len - p.n_fft) / hop + 1;
      s_flen[b] = fl < 0 ? 0 : (fl > F ? F : fl);
      s_pc[b] = p.hs_pc[b];
    }
    if (tid == 0) {
      s_pc[B] = p.hs_pc[B];
      s_misc[1] = p.hs_cta[blockIdx.x][0];
      s_misc[2] = p.hs_cta[blockIdx.x][1];
      s_misc[3] = p.hs_cta[blockIdx.x][2];
    }
  }
  pdl_wait();  // (so far only shared memory, the parameter block and the plan's constant tables were touched)
  if (blockIdx.x == 0 && p.feat_len)
    for (int b = tid; b < B; b += blockDim.x) p.feat_len[b] = frames_of(__ldg(p.len + b));
  if (warp == 0 && !p.hs_on) {
    if (tabled) {
      // exclusive scans of the utterances' costs (a pair with real frames costs REAL_UNIT_COST, one that only
      // holds rows past the utterance's length 1) and pair counts
      int carry_c = 0, carry_p = 0;
      for (int b0 = 0; b0 < B; b0 += 32) {
        const int b = b0 + lane;
        int c = 0, n = 0;
        if (b < B) {
          const long long len = __ldg(p.len + b);
          s_len[b] = static_cast<int>(len);
          const int fl = clampF(frames_of(len));
          s_flen[b] = fl;
          const int nr = (fl + FPW - 1) / FPW;
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
      int b0 = 0;  // utterance of the CTA's first pair: where every FFT warp starts looking
      if (tabled)
        while (b0 + 1 < B && s_pc[b0 + 1] <= plo) ++b0;
      s_misc[3] = b0;
    }
  }
  __syncthreads();  // barrier inits, the tables of the work split and this CTA's range are visible
  SFE_PF_G(pf_g1)
  SFE_TR(2, 0);

  const int P_lo = s_misc[1], npairs = s_misc[2];
  const int ntiles = (npairs + TS - 1) / TS;

  if (padfill && warp == NF) {
    // (first Mel warp, idle until the first tile is complete) prefix over the utterances of the rows that no pair
    // covers; the FFT warps need it at the very end of the launch
    int carry = 0;
    for (int b0 = 0; b0 < B; b0 += 32) {
      const int b = b0 + lane;
      const int n = (b < B) ? F - min(F, FPW * ((s_flen[b] + FPW - 1) / FPW)) : 0;
      int inc = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULLM, inc, o);
        if (lane >= o) inc += v;
      }
      if (b < B) s_padc[b] = carry + inc - n;
      carry += __shfl_sync(FULLM, inc, 31);
    }
    __syncwarp();
    if (lane == 0) {
      s_padc[B] = carry;
      st_release_shared(s_pad + 1, 1);
    }
  }
  if (warp >= NF) mbar_wait(tbar, 0u);  // (Mel warps) tables have landed

  if (warp >= NF) {
    // ================== Mel warps: warps NF + MS pb ... NF + MS pb + MS - 1 own buffer pb ==================
    constexpr int MS = TileCfg<PL>::MS, TSH = TS / MS;
    const int pb = (warp - NF) / MS, half = (warp - NF) % MS;  // half 0: stream A (+ the tail), half 1: stream B
    const int bid = 1 + 3 * pb;                                // named barriers bid, bid + 1, bid + 2 (MS == 2)
    float* p2buf = p2_0 + pb * p2_floats;
    const float4* wprog = reinterpret_cast<const float4*>(tab + p.o_mtab);  // Mel weights in the program's layout
    const int n_fmask = p.mk.n_fmask, n_tmask = p.mk.n_tmask;
    const int* m_fpos = p.mk.inl_on ? p.mk.inl : p.mk.fpos;
    const int* m_flen = p.mk.inl_on ? p.mk.inl + B * n_fmask : p.mk.flen;
    const int* m_tpos = p.mk.inl_on ? p.mk.inl + 2 * B * n_fmask : p.mk.tpos;
    const int* m_tlen = p.mk.inl_on ? p.mk.inl + 2 * B * n_fmask + B * n_tmask : p.mk.tlen;
    SFE_PF_DECL
    for (int j = pb; j < ntiles; j += NPB) {
      SFE_PF(4)
      if (FULLBAR) named_sync(FULLBAR_ID0 + pb, 32 * TS + 32 * MS);
      else mbar_wait_hint(pfull0 + 8 * pb, static_cast<uint32_t>((j / NPB) & 1));
      SFE_PF(0)
      SFE_TR(10, j);
      const int nsl = min(TS, npairs - j * TS);
      const T2Row* trow = rows + pb * TS;
      typedef typename MelBank<TileCfg<PL>::BANK>::type BK;
      static_assert(BK::D == 80, "the read-out below is laid out for 80 Mel bins");
      // lane = row: the pair and frame behind it
      const int srow = PERM ? (4 * (lane >> 3) + (lane & 3)) : (lane >> 1), grow = PERM ? ((lane >> 2) & 1) : (lane & 1);
      const bool rowact = (lane < 2 * TS) && (srow < nsl) && !(SFE_T2_ABLATE == 1 || SFE_T2_ABLATE == 2);
      bool va = false, ex = false, tm = false;  // the row is a real frame / is written / is time-masked
      int myb = 0, myt = 0;
      if (rowact) {
        const T2Row rd = trow[srow];
        myb = rd.b;
        myt = rd.tA + grow;
        va = (rd.kind != U_PAD) && (rd.kind != U_DONE) && (myt < rd.flen);
        ex = (rd.kind != U_DONE) && (myt < F) && !(p.skip_pad && !va);
        for (int qq = 0; qq < n_tmask; ++qq) {
          const int ps = m_tpos[myb * n_tmask + qq], ln = m_tlen[myb * n_tmask + qq];
          tm |= (myt >= ps) && (myt < ps + ln);
        }
      }
      const unsigned va_bits = __ballot_sync(FULLM, va), ex_bits = __ballot_sync(FULLM, ex);
      const unsigned tm_bits = __ballot_sync(FULLM, tm);

      SFE_PF(1)
      // ---- banded Mel: the recipe filterbank's straight-line program (sfe_melprog.h) ----
      // One packed FMA per bin into (even filter, odd filter); a finished filter's raw sum is stored over a bin
      // the lane has already consumed (constant address, no flag test, no branch); two independent streams
      // (bins below / above B_LD) for instruction-level parallelism, joined at the end.  The logarithm is
      // taken by the read-out.  (A data-driven version of this loop — flag tests and branches around the
      // emits — cost 5 900 instructions and ~30 us of one warp per tile, the whole CTA waiting for it.)
      float* rowp = p2buf + lane * PP;  // (dereferenced by active rows only)
      float2 aA = make_float2(0.f, 0.f), aB = make_float2(0.f, 0.f);
      {
        const float4* pA = reinterpret_cast<const float4*>(rowp);
        const float4* pB = reinterpret_cast<const float4*>(rowp + BK::B_LD);
        float4 xA[2], xB[2], wA0[2], wA1[2], wB0[2], wB1[2];
        float oa[4] = {0.f, 0.f, 0.f, 0.f}, ob[4] = {0.f, 0.f, 0.f, 0.f};  // finished filters waiting for a 128-bit store
        // HF: which streams this warp runs — 0 stream A, 1 stream B, 2 both (one Mel warp per buffer)
#define SFE_LD(g, buf)                                                                                              \
  if (HF != 1) { xA[buf] = pA[g]; wA0[buf] = wprog[4 * (g)]; wA1[buf] = wprog[4 * (g) + 1]; }                       \
  if (HF != 0) { xB[buf] = pB[g]; wB0[buf] = wprog[4 * (g) + 2]; wB1[buf] = wprog[4 * (g) + 3]; }
#define SFE_W(u_, v_, i) ((i) == 0 ? make_float2(u_.x, u_.y) : (i) == 1 ? make_float2(u_.z, u_.w) : (i) == 2 ? make_float2(v_.x, v_.y) : make_float2(v_.z, v_.w))
#define SFE_X(v_, i) ((i) == 0 ? v_.x : (i) == 1 ? v_.y : (i) == 2 ? v_.z : v_.w)
#define SFE_FA(buf, i) if (HF != 1) aA = f2fma(SFE_W(wA0[buf], wA1[buf], i), make_float2(SFE_X(xA[buf], i), SFE_X(xA[buf], i)), aA);
#define SFE_FB(buf, i) if (HF != 0) aB = f2fma(SFE_W(wB0[buf], wB1[buf], i), make_float2(SFE_X(xB[buf], i), SFE_X(xB[buf], i)), aB);
#define SFE_EA(slot, k) if (HF != 1) { if (slot) { oa[k] = aA.y; aA.y = 0.f; } else { oa[k] = aA.x; aA.x = 0.f; } }
#define SFE_EB(slot, k) if (HF != 0) { if (slot) { ob[k] = aB.y; aB.y = 0.f; } else { ob[k] = aB.x; aB.x = 0.f; } }
#define SFE_SA(pos) if (HF != 1) *reinterpret_cast<float4*>(rowp + (pos)) = make_float4(oa[0], oa[1], oa[2], oa[3]);
#define SFE_SB(pos) if (HF != 0) *reinterpret_cast<float4*>(rowp + (pos)) = make_float4(ob[0], ob[1], ob[2], ob[3]);
#define SFE_EA1(slot, pos) if (HF != 1) { if (slot) { rowp[pos] = aA.y; aA.y = 0.f; } else { rowp[pos] = aA.x; aA.x = 0.f; } }
#define SFE_EB1(slot, pos) if (HF != 0) { if (slot) { rowp[pos] = aB.y; aB.y = 0.f; } else { rowp[pos] = aB.x; aB.x = 0.f; } }
#define SFE_PA(k, pos) if (HF != 1) rowp[pos] = oa[k];
#define SFE_PROG                                                                                                       \
  if constexpr (TileCfg<PL>::BANK == 1) { SFE_MELPROG_1(SFE_LD, SFE_FA, SFE_FB, SFE_EA, SFE_EB, SFE_SA, SFE_SB, SFE_EA1, SFE_EB1, SFE_PA) } \
  else { SFE_MELPROG_2(SFE_LD, SFE_FA, SFE_FB, SFE_EA, SFE_EB, SFE_SA, SFE_SB, SFE_EA1, SFE_EB1, SFE_PA) }
        if (rowact) {
          if (MS == 1) { constexpr int HF = 2; SFE_PROG }
          else if (half == 0) { constexpr int HF = 0; SFE_PROG }
          else { constexpr int HF = 1; SFE_PROG }
        }
#undef SFE_PROG
#undef SFE_SA
#undef SFE_SB
#undef SFE_EA1
#undef SFE_EB1
#undef SFE_PA
#undef SFE_LD
#undef SFE_W
#undef SFE_X
#undef SFE_FA
#undef SFE_FB
#undef SFE_EA
#undef SFE_EB
      }
      // tail: filters fed by both streams are joined, stream A's last outputs move next to stream B's (the read-out's
      // layout) — after BOTH streams are through
      __syncwarp();
      if (MS == 2) named_sync(bid, 64);
      if (rowact && half == 0) {
#define SFE_JOIN(slot, pos) rowp[pos] += (slot) ? aA.y : aA.x; rowp[(pos) + 1] += (slot) ? aA.x : aA.y;
#define SFE_MOVE(from, to) rowp[to] = rowp[from];
        if constexpr (TileCfg<PL>::BANK == 1) { SFE_MELTAIL_1(SFE_JOIN, SFE_MOVE) }
        else { SFE_MELTAIL_2(SFE_JOIN, SFE_MOVE) }
#undef SFE_JOIN
#undef SFE_MOVE
      }
      __syncwarp();
      if (MS == 2) named_sync(bid + 1, 64);

      SFE_TR(11, j);
      SFE_PF(2)
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
#ifdef SFE_T2_NOATOM  // (timing experiment: sums without their global atomics)
              if (s1[e] == 123.456f) base[2 * d] = 0;
#else
              atomicAdd(base + 2 * d, static_cast<unsigned long long>(__float2ll_rn(s1[e] * 4294967296.f)));
              atomicAdd(base + 2 * d + 1, static_cast<unsigned long long>(__float2ll_rn(s2[e] * 16777216.f)));
#endif
            }
          }
#pragma unroll
          for (int e = 0; e < 5; ++e) s1[e] = s2[e] = 0.f;
        };
        // The usual tile: 2*TS consecutive real frames of ONE utterance, nothing but the logarithm (and the sums)
        // between the row and HBM — straight-line, every offset an immediate, the loads of all steps in flight.
        // (this warp's share of the tile: pairs it0 ... it0 + TSH - 1)
        const int it0 = half * TSH;
        const int b0 = __shfl_sync(FULLM, myb, 0), t0 = __shfl_sync(FULLM, myt, 0);
        const bool plain = !fancy && nsl == TS && dolog &&
                           __all_sync(FULLM, lane >= 2 * TS || (va && ex && myb == b0 && myt == t0 + 2 * srow + grow));
        if (plain) {
          float* dst0 = p.mel + (static_cast<long long>(b0) * F + t0 + 2 * it0) * BK::D + h * BK::D;
          const float* src0 = p2buf + (t2_row<PERM>(it0, 0) + t2_row<PERM>(0, 1) * h) * PP;
          if (stats_on) {
#pragma unroll
            for (int it = 0; it < TSH; ++it) {
              const float4 x4 = *reinterpret_cast<const float4*>(src0 + t2_row<PERM>(it, 0) * PP + so4);
              float x[5] = {x4.x, x4.y, x4.z, x4.w, src0[t2_row<PERM>(it, 0) * PP + so1]};
#pragma unroll
              for (int e = 0; e < 5; ++e) {
                x[e] = fast_log(x[e], clampv, lmul);
                s1[e] += x[e];
                s2[e] = fmaf(x[e], x[e], s2[e]);
              }
              *reinterpret_cast<float4*>(dst0 + 2 * it * BK::D + 4 * q) = make_float4(x[0], x[1], x[2], x[3]);
              dst0[2 * it * BK::D + 64 + q] = x[4];
            }
            flush(b0);
          } else {
#pragma unroll
            for (int it = 0; it < TSH; ++it) {
              const float4 x4 = *reinterpret_cast<const float4*>(src0 + t2_row<PERM>(it, 0) * PP + so4);
              float x[5] = {x4.x, x4.y, x4.z, x4.w, src0[t2_row<PERM>(it, 0) * PP + so1]};
#pragma unroll
              for (int e = 0; e < 5; ++e) x[e] = fast_log(x[e], clampv, lmul);
              *reinterpret_cast<float4*>(dst0 + 2 * it * BK::D + 4 * q) = make_float4(x[0], x[1], x[2], x[3]);
              dst0[2 * it * BK::D + 64 + q] = x[4];
            }
          }
        }
        int curb = -1;
        unsigned fm = 0u;
        for (int it = it0; it < (plain ? 0 : min(nsl, it0 + TSH)); ++it) {
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
          const int r = t2_row<PERM>(it, h);
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

      // hand the buffer back: tile j + NPB may write its rows (MS == 2: once the other warp is through as well —
      // it only arrives, the warp that publishes waits)
      __syncwarp();
      SFE_TR(12, j);
      SFE_PF(3)
      if (MS == 2) {
        if (half == 0) named_sync(bid + 2, 64);
        else named_arrive(bid + 2, 64);
      }
      // (release: every lane's reads of the tile — __syncwarp / barrier above — are done)
      if (half == 0 && lane == 0) st_release_shared(pgen + pb, j + NPB);
    }
    SFE_PF_DUMP(warp)
    SFE_SPAN(1);
    return;
  }

  // =============================== FFT warps: one pair per iteration ===============================
  unsigned char* wbase = smem + p.t2_warp + warp * p.t2_warp_bytes;
  cpx* Zw = reinterpret_cast<cpx*>(wbase) + g * PL::MP;
  float* escratch = reinterpret_cast<float*>(wbase);  // gathered pairs stage their samples in the FFT scratch
  float* stage = reinterpret_cast<float*>(wbase + p.t2_stage + 16);  // this warp's bulk-copied samples
  const uint32_t sbar = smem_u32(wbase + p.t2_stage);                // ... have landed
  const float* wtab = tab + p.o_win;
  const cpx* tw2 = reinterpret_cast<const cpx*>(tab + p.o_tw2);
  const cpx* tw3 = reinterpret_cast<const cpx*>(tab + p.o_tw3);
  const cpx* wpost = reinterpret_cast<const cpx*>(tab + p.o_wpost);
  const int qend = ntiles * TS;  // (the last tile's missing pairs still count for its barrier)
  const bool bulk_ok = ((reinterpret_cast<uintptr_t>(p.wav) & 15) == 0);

  // A pair as its warp sees it.  Everything here is warp-uniform: every lane computes the same values (no
  // broadcast).  Claims of a warp move forward, and mostly stay inside one utterance: its pair range, length,
  // frame count and row base are cached in registers, so a claim is ~30 instructions (the general search with
  // table loads and 64-bit arithmetic cost as much as an FFT pass per pair).
  struct Pair {
    int q, kind, b, tA, flen, shift;
    long long len;
  };
  int c_b = s_misc[3], c_p0 = 0, c_p1 = -1, c_len = 0, c_flen = 0;  // cached utterance: pairs [c_p0, c_p1)
  long long c_row = 0;                                             // ... and the float index of its first sample
  auto claim = [&]() -> Pair {
    Pair c;
    int q = 0;
    if (lane == 0) q = atomicAdd(&s_misc[0], 1);
    c.q = __shfl_sync(FULLM, q, 0);
    c.kind = U_DONE; c.b = 0; c.tA = 0; c.flen = 0; c.shift = 0; c.len = 0;
    if (c.q >= npairs) return c;  // nothing to stage (past the end, or padding of the CTA's last tile)
    const int P = P_lo + c.q;
    if (P >= c_p1) {  // another utterance
      if (tabled) {
        while (c_b + 1 < B && s_pc[c_b + 1] <= P) ++c_b;  // (also steps over utterances without pairs)
        c_p0 = s_pc[c_b];
        c_p1 = s_pc[c_b + 1];
        c_len = s_len[c_b];
        c_flen = s_flen[c_b];
      } else {
        c_b = P / upu;
        c_p0 = c_b * upu;
        c_p1 = c_p0 + upu;
        c_len = static_cast<int>(__ldg(p.len + c_b));
        c_flen = clampF(frames_of(c_len));
      }
      c_row = static_cast<long long>(c_b) * p.pitch;
    }
    c.b = c_b;
    c.len = c_len;
    c.flen = c_flen;
    c.tA = (P - c_p0) * FPW;
    c.kind = U_PAD;
    if (c.tA < c_flen) {
      c.kind = U_EDGE;
      const int o_start = c.tA * hop + p.off;               // (the host keeps inputs of 2^31 samples off this kernel)
      const int o_end = o_start + (FPW - 1) * hop + win;
      const long long gi = c_row + (o_start - pre1);
      const long long ga = gi & ~3LL;
      const int shiftfull = static_cast<int>(gi - ga) + pre1;
      const int need = (shiftfull + (FPW - 1) * hop + win + 3) & ~3;  // floats from ga that cover the pair
      if (bulk_ok && (o_start - pre1 >= 0) && (o_end <= c_len) && (ga + need <= p.total_floats)) {
        c.kind = U_BULK;
        c.shift = shiftfull;
        if (lane == 0) {
          // (the warp is past its last read of the buffer — __syncwarp before the call; order that before the
          //  async-proxy write)
          fence_proxy_async();
          mbar_expect_tx(sbar, static_cast<uint32_t>(need) * 4u);
          bulk_g2s(smem_u32(stage), p.wav + ga, static_cast<uint32_t>(need) * 4u, sbar);
        }
      }
    }
    return c;
  };

  uint32_t sparity = 0u;  // phase of sbar the next bulk pair completes
  Pair nx = claim();      // the first pair's samples fly while the tables land
  mbar_wait(tbar, 0u);
  SFE_PF_G(pf_g2)
  SFE_TR(3, 0);
  SFE_PF_DECL
  while (true) {
    const Pair cu = nx;
    const int q = cu.q;
    SFE_TR(20, q);
    if (q >= qend) break;
    const int j = q / TS, s = q - j * TS, pb = j % NPB;  // tile, slot and buffer of the Mel stage
    const int kind = cu.kind, b = cu.b, flen = cu.flen, tA = cu.tA;
    float* p2buf = p2_0 + pb * p2_floats;
    // rows of buffer pb may be overwritten once the Mel stage of tile j - NPB has consumed them
    auto p2_wait = [&]() {
      // (almost always already true: one shared-memory load; acquire: the Mel warps' reads of the rows are behind us)
      while (ld_acquire_shared(pgen + pb) != j) __nanosleep(64);
    };
    auto publish = [&]() {
      __syncwarp();
      if (lane == 0) {
        T2Row rd;
        rd.b = b; rd.tA = tA; rd.flen = flen; rd.kind = kind;
        rows[pb * TS + s] = rd;
        if (!FULLBAR) mbar_arrive(pfull0 + 8 * pb);  // (release: rows and descriptor are visible to the Mel warp)
      }
      if (FULLBAR) {
        __syncwarp();
        named_arrive(FULLBAR_ID0 + pb, 32 * TS + 32 * TileCfg<PL>::MS);
      }
    };
    if (kind == U_DONE || kind == U_PAD || SFE_T2_ABLATE >= 2) {  // (U_DONE: padding of the CTA's last tile)
      if (kind == U_BULK) { mbar_wait_hint(sbar, sparity); sparity ^= 1u; }
      __syncwarp();
      nx = claim();
      p2_wait();
      publish();
      SFE_PF(7)
      continue;
    }
    // ---------------- samples of this pair ----------------
    const bool bulk = (kind == U_BULK);
    const float* sbase = bulk ? stage : escratch;
    SFE_PF(6)
    if (bulk) {
      mbar_wait_hint(sbar, sparity);
      sparity ^= 1u;
    }
    SFE_PF(0)
    SFE_TR(21, q);
    auto gather = [&](int t0, int n) {
      // explicit gather with batch-level reflection, zero beyond the length and pre-emphasis applied here
      const long long len = cu.len;
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
      const int fo = bulk ? (fslot * hop + cu.shift) : ((PL::S == 1 ? fslot * hop : 0) + EDGE_SHIFT);
      const float* fr = sbase + fo;
      const float pre = (PRE && bulk) ? p.pre : 0.f;  // gathered pairs were pre-emphasised by the gather
      if ((fo & 1) == 0) stockham_pass<PL, PL::R1, 1, true, true, PRE, true>(Zw, nullptr, l, fr, wtab, win, pre);
      else stockham_pass<PL, PL::R1, 1, true, false, PRE, true>(Zw, nullptr, l, fr, wtab, win, pre);
      if (sub == PL::S - 1) {
        // the staging buffer is free (the pass ends with a __syncwarp): claim the next pair and start its copy,
        // which lands under the rest of this one
        SFE_PF(1)
        nx = claim();
        SFE_PF(2)
        SFE_TR(22, q);
      }
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
      if (sub == 0) { SFE_TR(23, q); SFE_PF(3) p2_wait(); SFE_PF(4) SFE_TR(24, q); }
      float* row = p2buf + t2_row<PERM>(s, fslot) * PP;
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
    SFE_PF(5)
    publish();
    SFE_TR(25, q);
  }
  // ---------------- rows past the utterances' pairs ----------------
  // A constant per utterance and Mel bin — log(clamp), normalised and masked like every other row (SURVEY App. A #10;
  // energy: 0) — written by the FFT warps that have run out of pairs, while the last tiles are still in the Mel
  // stage: the SM would idle otherwise.  (Through the tiles these rows cost more than the frames: 105 us against 47
  // for the C2 lengths without a normalise pass behind the kernel; the gather kernel took 90.)
  if (padfill) {
    typedef typename MelBank<TileCfg<PL>::BANK>::type BKp;
    constexpr int D4 = BKp::D / 4, CH = 16;
    while (ld_acquire_shared(s_pad + 1) == 0) __nanosleep(64);
    const long long R = s_padc[B];
    const int lo = static_cast<int>(R * blockIdx.x / gridDim.x), hi = static_cast<int>(R * (blockIdx.x + 1) / gridDim.x);
    const int n_fmask = p.mk.n_fmask, n_tmask = p.mk.n_tmask;
    const bool fancy = (p.norm_mean != nullptr) || (p.norm_std != nullptr) || (n_fmask > 0) || (n_tmask > 0);
    const int* m_fpos = p.mk.inl_on ? p.mk.inl : p.mk.fpos;
    const int* m_flen = p.mk.inl_on ? p.mk.inl + B * n_fmask : p.mk.flen;
    const int* m_tpos = p.mk.inl_on ? p.mk.inl + 2 * B * n_fmask : p.mk.tpos;
    const int* m_tlen = p.mk.inl_on ? p.mk.inl + 2 * B * n_fmask + B * n_tmask : p.mk.tlen;
    const float fillv = (fancy && p.fill_dev) ? *p.fill_dev : p.fill;
    int ub = 0, cached = -1;  // utterance cursor (chunks come in increasing order), utterance of `vrow`
    float4 vrow = make_float4(p.pad_mel, p.pad_mel, p.pad_mel, p.pad_mel);  // this lane's four Mel bins of a padded row
    while (true) {
      int c = 0;
      if (lane == 0) c = atomicAdd(&s_pad[0], 1);
      c = __shfl_sync(FULLM, c, 0);
      const int r0 = lo + c * CH;
      if (r0 >= hi) break;
      const int r1 = min(r0 + CH, hi);
      for (int r = r0; r < r1; ++r) {
        while (s_padc[ub + 1] <= r) ++ub;  // (steps over utterances without such rows)
        const int t = min(F, FPW * ((s_flen[ub] + FPW - 1) / FPW)) + (r - s_padc[ub]);
        if (fancy && ub != cached && lane < D4) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int d = 4 * lane + e;
            const float nm = p.norm_mean ? p.norm_mean[d] : 0.f;
            const float rs = p.norm_std ? 1.f / p.norm_std[d] : 1.f;
            v[e] = (p.pad_mel - nm) * rs;
            for (int qq = 0; qq < n_fmask; ++qq) {
              const int ps = m_fpos[ub * n_fmask + qq], ln = m_flen[ub * n_fmask + qq];
              if ((d >= ps) && (d < ps + ln)) v[e] = fillv;
            }
          }
          vrow = make_float4(v[0], v[1], v[2], v[3]);
        }
        cached = ub;
        bool tm = false;
        for (int qq = 0; qq < n_tmask; ++qq) {
          const int ps = m_tpos[ub * n_tmask + qq], ln = m_tlen[ub * n_tmask + qq];
          tm |= (t >= ps) && (t < ps + ln);
        }
        if (p.mel && lane < D4)
          *reinterpret_cast<float4*>(p.mel + (static_cast<long long>(ub) * F + t) * BKp::D + 4 * lane) =
              tm ? make_float4(fillv, fillv, fillv, fillv) : vrow;
        if (ENERGY && p.energy && lane == D4) p.energy[static_cast<long long>(ub) * F + t] = 0.f;
      }
    }
  }
  SFE_PF_DUMP(warp)
  SFE_SPAN(1);
}

}  // namespace sfe
