// libsfe.so — C ABI (include/sfe.h) over the sm_100a kernels.  No torch, no exceptions across
// the boundary, no per-call allocation: a plan owns its constant tables and nothing else.
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sfe.h"
#include "sfe_aux.cuh"
#include "sfe_fused.cuh"
#include "sfe_gl.cuh"
#include "sfe_melsched.h"
#include "sfe_registry.h"

using namespace sfe;

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define SFE_CUDA(expr)                                                                                 \
  do {                                                                                                 \
    cudaError_t e_ = (expr);                                                                           \
    if (e_ != cudaSuccess) return fail(SFE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// FFT plan registry: n_fft -> template instance
// ---------------------------------------------------------------------------------------------
// (the kernel instances live in sfe_plans.cu, one translation unit per FFT plan so they compile in parallel)
static const std::vector<FftEntry>& registry() {
  static const std::vector<FftEntry> r = [] {
    std::vector<FftEntry> v;
    for (int i = 0; i < kNumPlanTus; ++i) v.push_back(kPlanTus[i]());
    return v;
  }();
  return r;
}

struct sfe_plan {
  sfe_config cfg;
  const FftEntry* fft;
  int device;
  int n_sm;
  float* d_tables;
  double* d_acc;  // one fp64 scratch word for sfe_mean_all
  FwdParams proto;  // everything that does not change per call
  int smem_bytes;
  int ctas_per_sm;
  int warps_per_cta;  // warps never synchronise with each other: chosen for the most resident warps per SM
  // tile kernel (sfe_tile.cuh): usable for this plan / its dynamic shared memory
  int t2_ok = 0;
  int t2_smem = 0;
  // standalone mel kernel
  int l2m_smem;
  int l2m_sm_warp, l2m_warp_bytes;
  // Griffin-Lim (built on first use)
  std::vector<float> h_window;  // the caller's window (win_length floats)
  float* d_gl = nullptr;        // [ones | synthesis window / n_fft | window^2 | inverse pre-emphasis taps]
  int gl_o_synth = 0, gl_o_wsq = 0, gl_o_pw = 0, gl_o_awin = 0, gl_taps = -1;
};

extern "C" int sfe_abi_version(void) { return SFE_ABI_VERSION; }
extern "C" const char* sfe_last_error(void) { return g_err.c_str(); }

extern "C" int sfe_supported_nfft(int32_t* out, int cap) {
  const auto& r = registry();
  int n = 0;
  for (const auto& e : r)
    if (e.win == 0) {
      if (n < cap) out[n] = e.n_fft;
      ++n;
    }
  return n;
}

static int round_up(int x, int a) { return (x + a - 1) / a * a; }

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the FUNCTION, not to a plan: several plans share a
// kernel instance (same n_fft, different n_mels / hop; the no-mel plan recover() builds next to the training
// plan), so the attribute is only ever raised — a later, smaller plan must not shrink it under an earlier one.
static cudaError_t raise_dyn_smem(const void* fn, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, int> cur;  // (device, function) -> bytes granted so far
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  int& have = cur[std::make_pair(dev, fn)];
  if (bytes <= have) return cudaSuccess;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}

// Launch with programmatic stream serialisation: the kernel may start (its prologue, up to griddepcontrol.wait)
// while the previous kernel of the stream drains.  SFE_PDL=0 turns it off (A/B runs).
template <class Params>
static cudaError_t launch_pdl(void (*fn)(const Params), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const Params& P) {
  static const bool on = [] {
    const char* ev = getenv("SFE_PDL");
    return !(ev && ev[0] == '0');
  }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at;
  memset(&at, 0, sizeof(at));
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at;
  cfg.numAttrs = on ? 1 : 0;
  void* args[1] = {const_cast<Params*>(&P)};
  return cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(fn), args);
}

// (the same for kernels that take their arguments one by one)
template <class... KArgs, class... Args>
static cudaError_t launch_pdl_args(void (*fn)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  static const bool on = [] {
    const char* ev = getenv("SFE_PDL");
    return !(ev && ev[0] == '0');
  }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at;
  memset(&at, 0, sizeof(at));
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at;
  cfg.numAttrs = on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, fn, static_cast<KArgs>(args)...);
}

// SM count of the current device, looked up once per device (hot host path: no driver query per call)
static int cached_sm_count() {
  static std::mutex mu;
  static std::map<int, int> cnt;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cnt.find(dev);
  if (it != cnt.end()) return it->second;
  int n = 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
  cnt[dev] = n;
  return n;
}

extern "C" int sfe_plan_create(const sfe_config* cfg, sfe_plan** out) {
  if (!cfg || !out) return fail(SFE_ERR_BAD_ARG, "sfe_plan_create: NULL argument");
  *out = nullptr;
  if (cfg->n_fft <= 0 || (cfg->n_fft & 1)) return fail(SFE_ERR_BAD_CONFIG, "n_fft must be positive and even");
  if (cfg->win_length < 1 || cfg->win_length > cfg->n_fft)
    return fail(SFE_ERR_BAD_CONFIG, "win_length must be in [1, n_fft]");
  if (cfg->hop_length < 1) return fail(SFE_ERR_BAD_CONFIG, "hop_length must be >= 1");
  if (!cfg->window_host) return fail(SFE_ERR_BAD_CONFIG, "window_host is NULL");
  if (cfg->n_mels < 0 || cfg->n_mels > 32 * MAX_MEL_ROUNDS)
    return fail(SFE_ERR_BAD_CONFIG, "n_mels must be in [0, " + std::to_string(32 * MAX_MEL_ROUNDS) + "]");
  if (cfg->n_mels > 0 && !cfg->mel_weight_host) return fail(SFE_ERR_BAD_CONFIG, "mel_weight_host is NULL");
  // the kernels take logarithms with the flush-to-zero SFU form, exact only above the smallest normal float
  if ((cfg->n_mels > 0 && cfg->mel_logging && !(cfg->mel_clamp >= FLT_MIN)) ||
      (cfg->lin_logging && !(cfg->lin_clamp >= FLT_MIN)))
    return fail(SFE_ERR_BAD_CONFIG, "clamp must be >= 1.17549435e-38 when logging is on");
  const FftEntry* fe = nullptr;
  for (const auto& e : registry())
    if (e.n_fft == cfg->n_fft && (e.win == 0 || e.win == cfg->win_length) && (!fe || e.win != 0)) fe = &e;
  if (!fe) {
    std::string s = "unsupported n_fft " + std::to_string(cfg->n_fft) + "; compiled plans:";
    for (const auto& e : registry()) s += " " + std::to_string(e.n_fft);
    return fail(SFE_ERR_UNSUPPORTED_NFFT, s);
  }

  sfe_plan* pl = new sfe_plan();
  memset(&pl->proto, 0, sizeof(FwdParams));
  pl->cfg = *cfg;
  pl->h_window.assign(cfg->window_host, cfg->window_host + cfg->win_length);
  pl->cfg.window_host = nullptr;
  pl->cfg.mel_weight_host = nullptr;
  pl->fft = fe;
  const int M = fe->M, K = M + 1, D = cfg->n_mels, win = cfg->win_length;
  const double PI = 3.14159265358979323846;

  // ---- constant tables (fp64 -> fp32) ----
  std::vector<float> tab;
  auto align4 = [&]() { while (tab.size() % 4) tab.push_back(0.f); };
  FwdParams& P = pl->proto;
  // window, pre-scaled by 1/2 (the real-FFT unpack then needs no halving)
  P.o_win = static_cast<int>(tab.size());
  for (int i = 0; i < win; ++i) tab.push_back(0.5f * cfg->window_host[i]);
  for (int i = win; i < cfg->n_fft + 2; ++i) tab.push_back(0.f);  // zero padding up to the FFT length
  align4();
  // Stockham twiddles of pass 2 and 3: tw[(r-1)*T + j] = exp(-2 pi i r (j % P) / (P R))
  auto push_tw = [&](int R, int Pp) {
    const int T = M / R;
    for (int r = 1; r < R; ++r)
      for (int j = 0; j < T; ++j) {
        const double a = -2.0 * PI * static_cast<double>(r) * static_cast<double>(j % Pp) / (static_cast<double>(Pp) * R);
        tab.push_back(static_cast<float>(cos(a)));
        tab.push_back(static_cast<float>(sin(a)));
      }
  };
  P.o_tw2 = static_cast<int>(tab.size());
  push_tw(fe->R2, fe->R1);
  align4();
  P.o_tw3 = static_cast<int>(tab.size());
  if (fe->R3 > 1) push_tw(fe->R3, fe->R1 * fe->R2);
  align4();
  // real-FFT unpack twiddles W_N^k, k = 0..M/2
  P.o_wpost = static_cast<int>(tab.size());
  for (int k = 0; k <= M / 2; ++k) {
    const double a = -2.0 * PI * k / (2.0 * M);
    tab.push_back(static_cast<float>(cos(a)));
    tab.push_back(static_cast<float>(sin(a)));
  }
  align4();
  // banded mel filter -> filter *units* (a bin, or half of a wide bin) packed into phases of 32 lanes
  // with similar tap counts; weights lane-major, two taps per float2
  P.mel_tail = 0;
  P.mel_split_mask = 0;
  std::vector<int> mst(D > 0 ? D : 1, 0), mnt(D > 0 ? D : 1, 0);
  int maxn = 0;
  for (int d = 0; d < D; ++d) {
    const float* row = cfg->mel_weight_host + static_cast<size_t>(d) * K;
    int first = -1, last = -1;
    for (int k = 0; k < K; ++k)
      if (row[k] != 0.f) {
        if (first < 0) first = k;
        last = k;
      }
    if (first >= 0) {
      mst[d] = first;
      mnt[d] = last - first + 1;
    }
    if (mnt[d] > maxn) maxn = mnt[d];
  }
  // A bin wider than `ucap` taps is split over two neighbouring lanes.  Every phase costs a fixed prologue
  // and epilogue (~75 instructions) on top of ~9 per pair of taps of its widest unit, so the split threshold
  // is the one that minimises that estimate (80 mel / 257 bins: 3 phases of 12+8+4 taps instead of 4).
  struct Lane { int start, ntaps, d, out; int lead = 0; };  // out: -1 none, else d | combine flag
  auto pack = [&](int ucap, unsigned int* split_mask) {
    struct Item { int d, size; bool pair; };
    std::vector<Item> items;
    for (int d = 0; d < D; ++d) {
      const bool pair = mnt[d] > ucap;
      items.push_back(Item{d, pair ? (mnt[d] + 1) / 2 : mnt[d], pair});
    }
    std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.size > b.size; });
    std::vector<std::vector<Lane>> phs;
    *split_mask = 0;
    for (const Item& it : items) {
      if (phs.empty()) phs.emplace_back();
      if (it.pair && (phs.back().size() & 1)) phs.back().push_back(Lane{0, 0, -1, -1});
      if (phs.back().size() + (it.pair ? 2 : 1) > 32) phs.emplace_back();
      std::vector<Lane>& ph = phs.back();
      if (it.pair) {
        const int n1 = (mnt[it.d] + 1) / 2;
        ph.push_back(Lane{mst[it.d], n1, it.d, it.d | 0x10000});
        ph.push_back(Lane{mst[it.d] + n1, mnt[it.d] - n1, it.d, -1});
        if (phs.size() <= 32) *split_mask |= 1u << (phs.size() - 1);
      } else {
        ph.push_back(Lane{mst[it.d], mnt[it.d], it.d, it.d});
      }
    }
    return phs;
  };
  auto cost = [](const std::vector<std::vector<Lane>>& phs) {
    int c = 0;
    for (auto& ph : phs) {
      int nt = 0;
      for (auto& ln : ph) nt = ln.ntaps > nt ? ln.ntaps : nt;
      c += 60 + 11 * ((nt + 4) / 4);  // (+1 tap: a unit that starts on an odd bin is moved one bin down)
    }
    return c;
  };
  std::vector<std::vector<Lane>> phases = pack(maxn, &P.mel_split_mask);
  for (int ucap = maxn - 1; ucap >= (maxn + 1) / 2 && ucap >= 1; --ucap) {
    unsigned int sm = 0;
    std::vector<std::vector<Lane>> cand = pack(ucap, &sm);
    if (cost(cand) < cost(phases)) {
      phases.swap(cand);
      P.mel_split_mask = sm;
    }
  }
  if (static_cast<int>(phases.size()) > MAX_MEL_ROUNDS) {
    delete pl;
    return fail(SFE_ERR_BAD_CONFIG, "mel filterbank needs too many filter phases");
  }
  P.mel_rounds = static_cast<int>(phases.size());
  for (auto& ph : phases)
    while (ph.size() < 32) ph.push_back(Lane{0, 0, -1, -1});
  auto push_int = [&](int v) {
    float f;
    memcpy(&f, &v, 4);
    tab.push_back(f);
  };
  // A lane walks its unit four taps at a time with 128-bit loads: every unit starts on an even bin (one
  // leading zero weight if need be) and every lane of a phase walks the phase's tap count, a multiple of
  // four.  The power spectrum is followed by MEL_PAD_BINS zeroed bins; a unit that would run past them
  // starts earlier and leads with zero weights instead.
  std::vector<int> phase_nt(phases.size(), 0);
  for (size_t i = 0; i < phases.size(); ++i) {
    int nt = 0;
    for (auto& ln : phases[i]) {
      if (ln.d < 0) continue;
      const int a = ln.start & ~1;
      ln.lead = ln.start - a;
      ln.start = a;
      nt = std::max(nt, ln.lead + ln.ntaps);
    }
    nt = round_up(nt > 0 ? nt : 4, 4);
    phase_nt[i] = nt;
    for (auto& ln : phases[i]) {
      if (ln.d < 0) { ln.start = 0; ln.lead = 0; continue; }
      if (ln.start + nt > K + MEL_PAD_BINS) {
        int sh = round_up(ln.start + nt - (K + MEL_PAD_BINS), 2);
        if (sh > ln.start) sh = ln.start;
        ln.start -= sh;
        ln.lead += sh;
      }
    }
    if (nt > K + MEL_PAD_BINS) {
      delete pl;
      return fail(SFE_ERR_BAD_CONFIG, "mel filter wider than the spectrum");
    }
  }
  P.mel_tail = MEL_PAD_BINS;
  P.o_mtab = static_cast<int>(tab.size());  // int2 {first bin (even), output info} per (phase, lane)
  for (auto& ph : phases)
    for (auto& ln : ph) {
      push_int(ln.start);
      push_int(ln.out);
    }
  align4();
  P.o_mw = static_cast<int>(tab.size());
  for (size_t i = 0; i < phases.size(); ++i) {
    const int nt = phase_nt[i];
    P.mel_ph[i] = make_int2(nt / 4, (static_cast<int>(tab.size()) - P.o_mw) / 4);  // {blocks of 4 taps, float4 offset}
    for (int q4 = 0; q4 < nt / 4; ++q4)
      for (auto& ln : phases[i])
        for (int e = 0; e < 4; ++e) {
          const int q = q4 * 4 + e;
          tab.push_back((ln.d >= 0 && q >= ln.lead && q < ln.lead + ln.ntaps)
                            ? cfg->mel_weight_host[static_cast<size_t>(ln.d) * K + ln.start + q] : 0.f);
        }
  }
  align4();
  P.tab_floats = static_cast<int>(tab.size());

  // ---- tile kernel: weights in its Mel program's layout (behind the blob the gather kernel copies) + carve-up ----
  if (fe->t2[0] && D > 0) {
    const int TS = fe->t2_ts;
    const MelBankInfo bi = match_bank(cfg->mel_weight_host, D, K);
    if (bi.bank != 0 && bi.bank == fe->t2_bank) {  // the filterbank has the band structure compiled for this plan
      // a row holds the spectrum plus the zero-weight bins the program rounds up to; pitch = odd multiple of 16 bytes
      // (one lane per row reads 128 bits at a time: conflict-free)
      const int pp = t2_row_pitch(bi.ROWF);
      const std::vector<float> wt = program_weights(cfg->mel_weight_host, bi);
      P.t2_o_sched_g = static_cast<int>(tab.size());
      P.t2_sched_floats = static_cast<int>(wt.size());
      tab.insert(tab.end(), wt.begin(), wt.end());
      align4();
      P.t2_pp = pp;
      P.t2_rowf = bi.ROWF;
      const int hop = cfg->hop_length;
      // [tables | Mel weights | control block | NPB power-spectrum tiles | per FFT warp: FFT scratch, "samples
      //  landed" barrier, staging buffer]
      int o = round_up((P.o_mtab + P.t2_sched_floats) * 4, 16);
      P.t2_ctrl = o;
      o += round_up(T2_CTRL_BYTES(TS), 16);
      P.t2_p2 = o;
      o += fe->t2_npb * (2 * TS * pp) * 4;
      P.t2_warp = o;
      // gathered pairs stage their samples in the FFT scratch (a whole pair when both frames are in flight, one
      // frame at a time otherwise)
      const int gather_floats = (fe->C == FPW ? (FPW - 1) * hop + win : win) + 8;
      P.t2_stage = round_up(std::max(fe->C * fe->MP * 8, gather_floats * 4), 16);  // offset of the barrier in a warp's region
      P.t2_stage_floats = round_up((FPW - 1) * hop + win + 8, 4);
      P.t2_warp_bytes = P.t2_stage + 16 + P.t2_stage_floats * 4;
      pl->t2_smem = P.t2_warp + fe->t2_nf * P.t2_warp_bytes;
      pl->t2_ok = 1;
    }
  }

  // ---- shared-memory carve-up: [tables | table mbarrier | one private region per warp] ----
  // Warps never synchronise with each other, so the number of warps per CTA is free: take the one that
  // puts the most warps on an SM (big FFT plans need > 20 KB per warp).
  const bool alias = (fe->C == FPW);  // FftPlan::ALIAS_P2
  const int stage_floats = round_up((FPW - 1) * cfg->hop_length + cfg->n_fft + 16, 4);
  const int z_bytes = round_up(fe->C * fe->MP * 8, 16);
  const int p2_bytes = round_up((K + MEL_PAD_BINS) * FPW * 4, 16);
  int smem_limit = 227 * 1024, static_bytes = 2048;
  {
    int dev = 0;
    cudaDeviceProp pr;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&pr, dev) == cudaSuccess) {
      smem_limit = static_cast<int>(pr.sharedMemPerBlockOptin);
    }
    // the opt-in limit covers static + dynamic shared memory: leave room for the kernel's static part
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, reinterpret_cast<const void*>(fe->fn[cfg->has_preemphasis ? 1 : 0])) == cudaSuccess)
      static_bytes = static_cast<int>(fa.sharedSizeBytes);
    smem_limit -= static_bytes;
  }
  {
    int o = round_up(P.tab_floats * 4, 16);
    P.sm_tbar = o;
    o += 16;
    P.sm_warp = o;
    int wb;
    if (alias) {  // the power spectrum overwrites the FFT scratch
      P.sm_p2_off = 0;
      wb = z_bytes > p2_bytes ? z_bytes : p2_bytes;
    } else {      // the power spectrum sits behind the FFT scratch
      P.sm_p2_off = z_bytes;
      wb = z_bytes + p2_bytes;
    }
    P.sm_stage_off = wb;
    wb += stage_floats * 4;
    P.sm_acc_off = round_up(wb, 16);
    wb = P.sm_acc_off + (D > 0 ? D : 1) * 16;
    P.sm_bar_off = round_up(wb, 16);
    P.sm_warp_bytes = P.sm_bar_off + 16 + 32;  // two mbarriers, unit descriptor
  }
  // one CTA per SM with as many warps as shared memory and the plan's register budget allow
  pl->warps_per_cta = (smem_limit - P.sm_warp) / P.sm_warp_bytes;
  if (pl->warps_per_cta > fe->MAXW) pl->warps_per_cta = fe->MAXW;
  if (pl->warps_per_cta < 1) pl->warps_per_cta = 1;
  pl->smem_bytes = P.sm_warp + pl->warps_per_cta * P.sm_warp_bytes;
  // standalone mel kernel: tables + per-warp P2
  pl->l2m_sm_warp = round_up(P.tab_floats * 4, 16);
  pl->l2m_warp_bytes = p2_bytes;
  pl->l2m_smem = pl->l2m_sm_warp + NWARPS * p2_bytes;

  // ---- static part of the launch parameters ----
  P.n_fft = cfg->n_fft;
  P.win = win;
  P.hop = cfg->hop_length;
  P.K = K;
  P.D = D;
  P.center = cfg->center ? 1 : 0;
  P.off = (cfg->n_fft - win) / 2 - (cfg->center ? cfg->n_fft / 2 : 0);
  P.has_pre = cfg->has_preemphasis ? 1 : 0;
  P.pre = cfg->preemphasis;
  P.mag = cfg->mag_spec ? 1 : 0;
  P.lin_log = cfg->lin_logging ? 1 : 0;
  P.lin_clamp = cfg->lin_clamp;
  P.lin_div = cfg->lin_log_div;
  P.lin_lmul = static_cast<float>(0.69314718055994530942 / static_cast<double>(cfg->lin_log_div));
  P.mel_log = cfg->mel_logging ? 1 : 0;
  P.mel_clamp = cfg->mel_clamp;
  P.mel_div = cfg->mel_log_div;
  P.mel_lmul = static_cast<float>(0.69314718055994530942 / static_cast<double>(cfg->mel_log_div));
  P.pad_mel = cfg->mel_logging ? logf(cfg->mel_clamp) / cfg->mel_log_div : 0.f;
  P.pad_lin = cfg->lin_logging ? logf(cfg->lin_clamp) / cfg->lin_log_div : 0.f;

  // ---- device resources ----
  cudaError_t e = cudaGetDevice(&pl->device);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, pl->device);
  if (e == cudaSuccess && pl->smem_bytes > smem_limit) {
    delete pl;
    return fail(SFE_ERR_BAD_CONFIG, "configuration needs more shared memory per CTA than the device offers");
  }
  if (e == cudaSuccess) e = cudaMalloc(&pl->d_tables, tab.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(pl->d_tables, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&pl->d_acc, sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(pl->d_acc, 0, sizeof(double));
  for (int v = 0; v < 2; ++v)
    if (e == cudaSuccess)
      e = raise_dyn_smem(reinterpret_cast<const void*>(fe->fn[v]), pl->smem_bytes);
  if (e == cudaSuccess && pl->t2_ok) {
    cudaFuncAttributes fa;
    int t2_static = 4096;
    if (cudaFuncGetAttributes(&fa, reinterpret_cast<const void*>(fe->t2[0])) == cudaSuccess)
      t2_static = static_cast<int>(fa.sharedSizeBytes);
    if (pl->t2_smem + t2_static > static_cast<int>(prop.sharedMemPerBlockOptin)) pl->t2_ok = 0;
  }
  for (int v = 0; v < 3; ++v)
    if (e == cudaSuccess && pl->t2_ok)
      e = raise_dyn_smem(reinterpret_cast<const void*>(fe->t2[v]), pl->t2_smem);
  if (e == cudaSuccess)
    e = raise_dyn_smem(reinterpret_cast<const void*>(&linear2mel_kernel), pl->l2m_smem);
  int occ = 0;
  if (e == cudaSuccess)
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, reinterpret_cast<const void*>(fe->fn[cfg->has_preemphasis ? 1 : 0]),
                                                      (pl->warps_per_cta + 1) * 32, pl->smem_bytes);
  if (e != cudaSuccess) {
    std::string msg = std::string("sfe_plan_create: ") + cudaGetErrorString(e);
    cudaFree(pl->d_tables);
    cudaFree(pl->d_acc);
    delete pl;
    return fail(SFE_ERR_CUDA, msg);
  }
  pl->n_sm = prop.multiProcessorCount;
  pl->ctas_per_sm = occ < 1 ? 1 : occ;
  P.tables = pl->d_tables;
  *out = pl;
  return SFE_OK;
}

extern "C" void sfe_plan_destroy(sfe_plan* pl) {
  if (!pl) return;
  cudaFree(pl->d_tables);
  cudaFree(pl->d_acc);
  cudaFree(pl->d_gl);
  delete pl;
}

extern "C" int64_t sfe_plan_num_frames(const sfe_plan* pl, int64_t T_max) {
  if (!pl) return -1;
  const int64_t T = pl->cfg.center ? T_max + 2 * (pl->cfg.n_fft / 2) : T_max;
  if (T < pl->cfg.n_fft) return 0;
  return 1 + (T - pl->cfg.n_fft) / pl->cfg.hop_length;
}

extern "C" int sfe_plan_query(const sfe_plan* pl, int32_t* smem_bytes, int32_t* ctas_per_sm, int32_t* n_sm) {
  if (!pl) return fail(SFE_ERR_BAD_ARG, "sfe_plan_query: NULL plan");
  if (smem_bytes) *smem_bytes = pl->smem_bytes;
  if (ctas_per_sm) *ctas_per_sm = pl->ctas_per_sm;
  if (n_sm) *n_sm = pl->n_sm;
  return SFE_OK;
}

extern "C" int sfe_mel_schedule(const float* mel_weight_host, int32_t n_mels, int32_t n_bins, int32_t* meta,
                                float* table, int32_t table_cap) {
  if (!mel_weight_host || !meta || n_mels <= 0 || n_bins <= 0) return -fail(SFE_ERR_BAD_ARG, "sfe_mel_schedule: bad argument");
  const MelBankInfo bi = match_bank(mel_weight_host, n_mels, n_bins);
  const int vals[10] = {bi.bank, bi.K, bi.D, bi.NG, bi.B_LD, bi.D_SPLIT, bi.B_OFF, bi.C_SPLIT, bi.ROWF, bi.TAIL};
  for (int i = 0; i < 10; ++i) meta[i] = vals[i];
  if (!bi.bank) return 0;
  const std::vector<float> wt = program_weights(mel_weight_host, bi);
  const int n = static_cast<int>(wt.size());
  if (table && table_cap >= n) memcpy(table, wt.data(), sizeof(float) * n);
  return n;
}

// validate + translate the public mask descriptor (device arrays, or host arrays copied into the
// kernel parameter block)
static int fill_masks(const sfe_masks* m, int B, MaskDesc* out, float* fill, const float** fill_dev) {
  out->fpos = out->flen = out->tpos = out->tlen = nullptr;
  out->n_fmask = out->n_tmask = 0;
  out->inl_on = 0;
  if (!m) return SFE_OK;
  if (m->n_freq < 0 || m->n_time < 0 || m->n_freq > 64 || m->n_time > 64)
    return fail(SFE_ERR_BAD_ARG, "mask counts must be in [0, 64]");
  if ((m->n_freq > 0 && (!m->freq_pos || !m->freq_len)) || (m->n_time > 0 && (!m->time_pos || !m->time_len)))
    return fail(SFE_ERR_BAD_ARG, "mask descriptor pointers are NULL");
  out->n_fmask = m->n_freq;
  out->n_tmask = m->n_time;
  *fill = m->fill;
  *fill_dev = m->fill_dev;
  if (m->on_host) {
    const long long need = 2LL * (m->n_freq + m->n_time) * B;
    if (need > MASK_INLINE_MAX)
      return fail(SFE_ERR_BAD_ARG, "host mask descriptors exceed SFE_MASK_INLINE_MAX; upload them and pass on_host=0");
    int* q = out->inl;
    if (m->n_freq) {
      memcpy(q, m->freq_pos, sizeof(int) * B * m->n_freq);
      memcpy(q + B * m->n_freq, m->freq_len, sizeof(int) * B * m->n_freq);
    }
    q += 2 * B * m->n_freq;
    if (m->n_time) {
      memcpy(q, m->time_pos, sizeof(int) * B * m->n_time);
      memcpy(q + B * m->n_time, m->time_len, sizeof(int) * B * m->n_time);
    }
    out->inl_on = 1;
  } else {
    out->fpos = m->freq_pos;
    out->flen = m->freq_len;
    out->tpos = m->time_pos;
    out->tlen = m->time_len;
  }
  return SFE_OK;
}

static int frontend_forward_impl(const sfe_plan* pl, const float* wav, int64_t wav_pitch, const int64_t* wav_len,
                                 int32_t B, int64_t T_max, float* mel, float* lin, float* energy, int64_t* feat_len,
                                 int64_t* stats, const float* norm_mean, const float* norm_std, const sfe_masks* masks,
                                 int32_t skip_pad, void* stream, bool zero_stats, const int64_t* wav_len_host);

extern "C" int sfe_frontend_forward(const sfe_plan* pl, const float* wav, int64_t wav_pitch, const int64_t* wav_len,
                                    int32_t B, int64_t T_max, float* mel, float* lin, float* energy, int64_t* feat_len,
                                    int64_t* stats, const float* norm_mean, const float* norm_std,
                                    const sfe_masks* masks, int32_t skip_pad, void* stream,
                                    const int64_t* wav_len_host) {
  return frontend_forward_impl(pl, wav, wav_pitch, wav_len, B, T_max, mel, lin, energy, feat_len, stats, norm_mean,
                               norm_std, masks, skip_pad, stream, true, wav_len_host);
}

// The tile kernel's work split, made on the host when the caller has the lengths there: the same equal-cost cut of
// the utterance-major pair axis the kernel makes for itself (tile_frontend_kernel's set-up), written into the
// parameter block.  Returns false (kernel splits for itself) when the batch or the grid is beyond the tables.
static bool host_split(const sfe_plan* pl, const int64_t* len_host, int B, int64_t F, int upu, bool compact, int grid,
                       FwdParams* P) {
  if (!len_host || B > UNIT_TABLE_MAX_B || grid > HOST_SPLIT_MAX_GRID) return false;
  const int n_fft = pl->cfg.n_fft, hop = pl->cfg.hop_length;
  long long cum[UNIT_TABLE_MAX_B + 1];
  int nreal[UNIT_TABLE_MAX_B];
  long long c = 0, np = 0;
  for (int b = 0; b < B; ++b) {
    const long long len = len_host[b];
    if (len < 0 || len > 0x7ff00000LL) return false;
    long long fl = (pl->cfg.center ? (len + 2 * (n_fft / 2)) - n_fft : len - n_fft) / hop + 1;
    fl = fl < 0 ? 0 : (fl > F ? F : fl);
    const int nr = static_cast<int>((fl + FPW - 1) / FPW);
    const int n = compact ? nr : upu;
    nreal[b] = nr;
    cum[b] = c;
    P->hs_len[b] = static_cast<int>(len);
    P->hs_pc[b] = static_cast<int>(np);
    c += static_cast<long long>(REAL_UNIT_COST) * nr + (n - nr);
    np += n;
  }
  cum[B] = c;
  P->hs_pc[B] = static_cast<int>(np);
  auto pair_at_cost = [&](long long x) -> int {
    int lo = 0, hi = B - 1;  // smallest b whose range ends beyond x
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cum[mid + 1] > x) hi = mid; else lo = mid + 1;
    }
    const long long off = x - cum[lo];
    const long long nr = nreal[lo];
    const long long tu = (off <= REAL_UNIT_COST * nr) ? (off + REAL_UNIT_COST - 1) / REAL_UNIT_COST : nr + (off - REAL_UNIT_COST * nr);
    return P->hs_pc[lo] + static_cast<int>(tu);
  };
  int b0 = 0;
  for (int g = 0; g < grid; ++g) {
    const int plo = pair_at_cost(c * g / grid), phi = pair_at_cost(c * (g + 1) / grid);
    while (b0 + 1 < B && P->hs_pc[b0 + 1] <= plo) ++b0;
    P->hs_cta[g][0] = plo;
    P->hs_cta[g][1] = phi - plo;
    P->hs_cta[g][2] = b0;
  }
  P->hs_on = 1;
  return true;
}

static int frontend_forward_impl(const sfe_plan* pl, const float* wav, int64_t wav_pitch, const int64_t* wav_len,
                                 int32_t B, int64_t T_max, float* mel, float* lin, float* energy, int64_t* feat_len,
                                 int64_t* stats, const float* norm_mean, const float* norm_std, const sfe_masks* masks,
                                 int32_t skip_pad, void* stream, bool zero_stats, const int64_t* wav_len_host) {
  if (!pl) return fail(SFE_ERR_BAD_ARG, "sfe_frontend_forward: NULL plan");
  if (!wav || !wav_len) return fail(SFE_ERR_BAD_ARG, "sfe_frontend_forward: wav / wav_len is NULL");
  if (B <= 0 || T_max <= 0 || wav_pitch < T_max) return fail(SFE_ERR_BAD_ARG, "sfe_frontend_forward: bad B / T_max / pitch");
  if (reinterpret_cast<uintptr_t>(wav) & 3) return fail(SFE_ERR_BAD_ARG, "wav must be 4-byte aligned");
  if (mel && pl->cfg.n_mels == 0) return fail(SFE_ERR_BAD_ARG, "plan has no mel stage but mel output requested");
  if (!mel && (stats || norm_mean || norm_std || masks))
    return fail(SFE_ERR_BAD_ARG, "stats / normalisation / masks need the mel output");
  if (stats && !pl->cfg.mel_logging)
    return fail(SFE_ERR_BAD_ARG, "fused fixed-point sums need a log-mel plan; use sfe_stats_accumulate on the output");
  if (!mel && !lin && !energy) return fail(SFE_ERR_BAD_ARG, "no output requested");
  if (pl->cfg.center && T_max <= pl->cfg.n_fft / 2)
    return fail(SFE_ERR_BAD_ARG, "reflect padding needs T_max > n_fft/2 (torch.stft raises here too)");
  const int64_t F = sfe_plan_num_frames(pl, T_max);
  if (F <= 0) return fail(SFE_ERR_BAD_ARG, "input shorter than one frame");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  FwdParams P = pl->proto;
  P.wav = wav;
  P.pitch = wav_pitch;
  P.total_floats = static_cast<long long>(B - 1) * wav_pitch + T_max;
  P.len = reinterpret_cast<const long long*>(wav_len);
  P.B = B;
  P.T_max = T_max;
  P.F = static_cast<int>(F);
  P.units_per_utt = static_cast<int>((F + FPW - 1) / FPW);
  const long long nu = static_cast<long long>(P.units_per_utt) * B;
  if (nu * REAL_UNIT_COST > 0x7fffffffLL) return fail(SFE_ERR_BAD_ARG, "batch too large");
  P.mel = mel;
  P.lin = lin;
  P.energy = energy;
  P.feat_len = reinterpret_cast<long long*>(feat_len);
  P.stats = reinterpret_cast<unsigned long long*>(stats);
  P.norm_mean = norm_mean;
  P.norm_std = norm_std;
  P.skip_pad = skip_pad;
  if (int rc = fill_masks(masks, B, &P.mk, &P.fill, &P.fill_dev)) return rc;
  if (stats && zero_stats)
    SFE_CUDA(cudaMemsetAsync(stats, 0, static_cast<size_t>(B) * pl->cfg.n_mels * 2 * sizeof(int64_t), st));
  // one CTA per SM; every CTA takes an equal-cost contiguous range of the batch's units
  long long grid = static_cast<long long>(pl->n_sm) * pl->ctas_per_sm;
  if (grid > nu) grid = nu;
  // The tile kernel (sfe_tile.cuh) serves every call with a Mel output whose filterbank could be scheduled;
  // the gather kernel keeps the rest (linear-only plans, n_fft 2048, unaligned outputs, energy + pre-emphasis).
  static const int force = [] {  // SFE_KERNEL=gather|tile: A/B runs
    const char* ev = getenv("SFE_KERNEL");
    return !ev ? 0 : (ev[0] == 'g' ? 1 : (ev[0] == 't' ? 2 : 0));
  }();
  const bool tile = pl->t2_ok && mel && !lin && force != 1 && T_max < 0x7ff00000LL && !(reinterpret_cast<uintptr_t>(mel) & 15) &&
                    !(energy && pl->cfg.has_preemphasis) &&
                    (!norm_mean || !(reinterpret_cast<uintptr_t>(norm_mean) & 15)) &&
                    (!norm_std || !(reinterpret_cast<uintptr_t>(norm_std) & 15));
  if (force == 2 && !tile) return fail(SFE_ERR_BAD_ARG, "SFE_KERNEL=tile: this call cannot run on the tile kernel");
  if (tile) {
    long long g2 = pl->n_sm;
    if (g2 > nu) g2 = nu;
    const fused_fn fn = energy ? pl->fft->t2[2] : pl->fft->t2[pl->cfg.has_preemphasis ? 1 : 0];
    static const bool hs_off = [] {  // SFE_HOST_SPLIT=0: A/B runs
      const char* ev = getenv("SFE_HOST_SPLIT");
      return ev && ev[0] == '0';
    }();
    P.hs_on = 0;
    if (!hs_off)
      // (the kernel enumerates real pairs only whenever the batch fits its tables: same rule here)
      host_split(pl, wav_len_host, B, F, P.units_per_utt, true, static_cast<int>(g2), &P);
    SFE_CUDA(launch_pdl<FwdParams>(fn, dim3(static_cast<unsigned>(g2)), dim3(pl->fft->t2_nw * 32), pl->t2_smem, st, P));
  } else {
    SFE_CUDA(launch_pdl<FwdParams>(pl->fft->fn[pl->cfg.has_preemphasis ? 1 : 0], dim3(static_cast<unsigned>(grid)),
                                   dim3((pl->warps_per_cta + 1) * 32), pl->smem_bytes, st, P));
  }
  SFE_CUDA(cudaGetLastError());
  return SFE_OK;
}

extern "C" int sfe_linear2mel_forward(const sfe_plan* pl, const float* lin, int64_t n_rows, float* mel, void* stream) {
  if (!pl || !lin || !mel) return fail(SFE_ERR_BAD_ARG, "sfe_linear2mel_forward: NULL argument");
  if (pl->cfg.n_mels == 0) return fail(SFE_ERR_BAD_ARG, "plan has no mel stage");
  if (n_rows <= 0) return SFE_OK;
  L2MParams P;
  memset(&P, 0, sizeof(P));
  P.lin = lin;
  P.mel = mel;
  P.n_rows = n_rows;
  P.K = pl->proto.K;
  P.D = pl->proto.D;
  P.mel_log = pl->proto.mel_log;
  P.mel_clamp = pl->proto.mel_clamp;
  P.mel_div = pl->proto.mel_div;
  P.mel_lmul = pl->proto.mel_lmul;
  P.tables = pl->d_tables;
  P.tab_floats = pl->proto.tab_floats;
  P.o_mtab = pl->proto.o_mtab;
  P.o_mw = pl->proto.o_mw;
  P.mel_split_mask = pl->proto.mel_split_mask;
  P.mel_rounds = pl->proto.mel_rounds;
  P.mel_tail = pl->proto.mel_tail;
  for (int i = 0; i < MAX_MEL_ROUNDS; ++i) P.mel_ph[i] = pl->proto.mel_ph[i];
  P.sm_warp = pl->l2m_sm_warp;
  P.sm_warp_bytes = pl->l2m_warp_bytes;
  long long pairs = (n_rows + FPW - 1) / FPW;
  long long want = (pairs + NWARPS - 1) / NWARPS;
  int grid = static_cast<int>(want < pl->n_sm * 4LL ? want : pl->n_sm * 4LL);
  linear2mel_kernel<<<grid, NTHREADS, pl->l2m_smem, static_cast<cudaStream_t>(stream)>>>(P);
  SFE_CUDA(cudaGetLastError());
  return SFE_OK;
}

extern "C" int sfe_stats_accumulate(const float* feat, const int64_t* feat_len, int32_t B, int64_t F, int32_t D,
                                    double* stats, void* stream) {
  if (!feat || !feat_len || !stats) return fail(SFE_ERR_BAD_ARG, "sfe_stats_accumulate: NULL argument");
  if (B <= 0 || F <= 0 || D <= 0) return fail(SFE_ERR_BAD_ARG, "sfe_stats_accumulate: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SFE_CUDA(cudaMemsetAsync(stats, 0, static_cast<size_t>(B) * D * 2 * sizeof(double), st));
  dim3 grid(static_cast<unsigned>((F + STATS_FR - 1) / STATS_FR), static_cast<unsigned>(B));
  if (D <= 256) {
    const int rows = 256 / D > 0 ? 256 / D : 1;
    stats_accumulate_kernel<<<grid, 256, rows * D * 2 * sizeof(double), st>>>(
        feat, reinterpret_cast<const long long*>(feat_len), F, D, stats);
  } else {
    stats_accumulate_wide_kernel<<<grid, 256, 0, st>>>(feat, reinterpret_cast<const long long*>(feat_len), F, D, stats);
  }
  SFE_CUDA(cudaGetLastError());
  return SFE_OK;
}

static int stats_finalize_impl(const void* stats, int32_t stats_fixed, const int64_t* feat_len, int32_t B, int32_t D,
                               float clamp, float* mean, float* std, float* packed, void* stream, int64_t* zero_sums,
                               bool pdl) {
  if (!stats || !feat_len) return fail(SFE_ERR_BAD_ARG, "sfe_stats_finalize: NULL argument");
  if (B <= 0 || D <= 0) return fail(SFE_ERR_BAD_ARG, "sfe_stats_finalize: bad shape");
  // one CTA per 1024 columns; inside it the utterances are cut into interleaved slices (1024 / Dc of them)
  const int Dc = D < 1024 ? D : 1024;
  int S = 1024 / Dc;
  if (S > B) S = B;
  const int threads = ((S * Dc + 31) / 32) * 32;
  const dim3 grid((D + Dc - 1) / Dc), block(threads);
  const size_t smem = static_cast<size_t>(S) * Dc * 2 * sizeof(float);
  const long long* fl = reinterpret_cast<const long long*>(feat_len);
  long long* zs = reinterpret_cast<long long*>(zero_sums);
  if (pdl)
    SFE_CUDA(launch_pdl_args(stats_finalize_kernel, grid, block, smem, static_cast<cudaStream_t>(stream), stats,
                             static_cast<int>(stats_fixed), fl, static_cast<int>(B), static_cast<int>(D), Dc, S, clamp, mean,
                             std, packed, zs));
  else
    stats_finalize_kernel<<<grid, block, smem, static_cast<cudaStream_t>(stream)>>>(stats, stats_fixed, fl, B, D, Dc, S,
                                                                                    clamp, mean, std, packed, zs);
  SFE_CUDA(cudaGetLastError());
  return SFE_OK;
}

extern "C" int sfe_stats_finalize(const void* stats, int32_t stats_fixed, const int64_t* feat_len, int32_t B, int32_t D,
                                  float clamp, float* mean, float* std, float* packed, void* stream) {
  return stats_finalize_impl(stats, stats_fixed, feat_len, B, D, clamp, mean, std, packed, stream, nullptr, false);
}

extern "C" int sfe_running_update(const float* packed, int32_t D, float weight, int32_t first, int32_t update,
                                  float* run_mean, float* run_std, float* batch_mean_out, float* batch_std_out,
                                  void* stream) {
  if (!packed || D <= 0) return fail(SFE_ERR_BAD_ARG, "sfe_running_update: bad argument");
  running_update_kernel<<<(D + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      packed, D, weight, first, update, run_mean, run_std, batch_mean_out, batch_std_out, nullptr, 0);
  SFE_CUDA(cudaGetLastError());
  return SFE_OK;
}

static int normalize_impl(const float* src, float* dst, const int64_t* feat_len, int32_t B, int64_t F, int32_t D,
                          const void* stats, int32_t stats_fixed, float clamp, int32_t use_mean, int32_t use_std,
                          const float* mean, const float* std, int32_t per_utt, float* mean_out, float* std_out,
                          int32_t src_has_pad, float pad_value, const sfe_warp* warp, const sfe_masks* masks,
                          void* stream, int64_t* zero_sums, uint16_t* dst16 = nullptr);

extern "C" int sfe_normalize_bf16(const float* src, uint16_t* dst_bf16, const int64_t* feat_len, int32_t B, int64_t F,
                                  int32_t D, const void* stats, int32_t stats_fixed, float clamp, int32_t use_mean,
                                  int32_t use_std, const float* mean, const float* std, int32_t per_utt, int32_t src_has_pad,
                                  float pad_value, const sfe_warp* warp, const sfe_masks* masks, void* stream) {
  if (!dst_bf16) return fail(SFE_ERR_BAD_ARG, "sfe_normalize_bf16: NULL destination");
  return normalize_impl(src, reinterpret_cast<float*>(dst_bf16), feat_len, B, F, D, stats, stats_fixed, clamp, use_mean,
                        use_std, mean, std, per_utt, nullptr, nullptr, src_has_pad, pad_value, warp, masks, stream, nullptr,
                        dst_bf16);
}

extern "C" int sfe_normalize(const float* src, float* dst, const int64_t* feat_len, int32_t B, int64_t F, int32_t D,
                             const void* stats, int32_t stats_fixed, float clamp, int32_t use_mean, int32_t use_std,
                             const float* mean,
                             const float* std, int32_t per_utt, float* mean_out, float* std_out, int32_t src_has_pad,
                             float pad_value, const sfe_warp* warp, const sfe_masks* masks, void* stream) {
  return normalize_impl(src, dst, feat_len, B, F, D, stats, stats_fixed, clamp, use_mean, use_std, mean, std, per_utt,
                        mean_out, std_out, src_has_pad, pad_value, warp, masks, stream, nullptr);
}

static int normalize_impl(const float* src, float* dst, const int64_t* feat_len, int32_t B, int64_t F, int32_t D,
                          const void* stats, int32_t stats_fixed, float clamp, int32_t use_mean, int32_t use_std,
                          const float* mean, const float* std, int32_t per_utt, float* mean_out, float* std_out,
                          int32_t src_has_pad, float pad_value, const sfe_warp* warp, const sfe_masks* masks,
                          void* stream, int64_t* zero_sums, uint16_t* dst16) {
  if (!src || !dst || !feat_len) return fail(SFE_ERR_BAD_ARG, "sfe_normalize: NULL argument");
  if (B <= 0 || F <= 0 || D <= 0 || D > 4096) return fail(SFE_ERR_BAD_ARG, "sfe_normalize: bad shape");
  const bool warp_on = warp && warp->enabled;
  if (warp_on) {
    if (!dst16 && src == dst) return fail(SFE_ERR_BAD_ARG, "sfe_normalize: the time warp cannot run in place");
    if (!(warp->center > 0 && warp->center < warp->minlen && warp->pos > 0 && warp->pos < warp->minlen &&
          warp->minlen <= F))
      return fail(SFE_ERR_BAD_ARG, "sfe_normalize: inconsistent warp descriptor");
  }
  NormParams P;
  memset(&P, 0, sizeof(P));
  P.src = src;
  P.dst = dst16 ? nullptr : dst;
  P.dst16 = dst16;
  P.flen = reinterpret_cast<const long long*>(feat_len);
  P.B = B;
  P.F = F;
  P.D = D;
  P.stats = stats;
  P.stats_fixed = stats_fixed;
  P.clamp = clamp;
  P.use_mean = use_mean;
  P.use_std = use_std;
  P.mean = mean;
  P.sd = std;
  P.per_utt = per_utt;
  P.mean_out = mean_out;
  P.std_out = std_out;
  P.src_has_pad = src_has_pad;
  P.pad_value = pad_value;
  P.zero_sums = reinterpret_cast<long long*>(zero_sums);
  if (warp_on) {
    P.warp_on = (warp->enabled == 2) ? 2 : 1;  // 1 bicubic, 2 bilinear
    P.warp_center = warp->center;
    P.warp_pos = warp->pos;
    P.warp_minlen = warp->minlen;
  }
  if (int rc = fill_masks(masks, B, &P.mk, &P.fill, &P.fill_dev)) return rc;
  // A CTA pays a fixed price (statistics of its utterance in fp64, per-column constants) before it streams:
  // size the frame blocks for about SFE_NORM_CTAS_PER_SM (default 4) CTAs per SM, at least ~32 KB of features each.
  static const int per_sm = [] {  // tuning knob, read once
    int v = 4;
    if (const char* ev = getenv("SFE_NORM_CTAS_PER_SM")) {
      const int x = atoi(ev);
      if (x >= 1 && x <= 64) v = x;
    }
    return v;
  }();
  const int n_sm = cached_sm_count();
#ifdef SFE_NORM_CEIL
  long long blocks_per_utt = (static_cast<long long>(n_sm) * per_sm + B - 1) / B;
#else
  // (rounded DOWN: the grid must fit the resident CTAs — one wave; a second, partly filled one costs a quarter)
  // (the instance with the time-warp gather runs one CTA per SM less)
  long long blocks_per_utt = static_cast<long long>(n_sm) * (warp_on && per_sm > 1 ? per_sm - 1 : per_sm) / B;
#endif
  if (blocks_per_utt < 1) blocks_per_utt = 1;
  long long fpc64 = (F + blocks_per_utt - 1) / blocks_per_utt;
  const long long fmin = 8192 / D > 1 ? 8192 / D : 1;
  if (fpc64 < fmin) fpc64 = fmin;
  const int fpc = static_cast<int>(fpc64 > 0x3fffffff ? 0x3fffffff : fpc64);
  P.frames_per_cta = fpc;
  dim3 grid(static_cast<unsigned>((F + fpc - 1) / fpc), static_cast<unsigned>(B));
  SFE_CUDA(launch_pdl<NormParams>(warp_on ? normalize_kernel<true> : normalize_kernel<false>, grid, dim3(256),
                                  3 * D * sizeof(float), static_cast<cudaStream_t>(stream), P));
  SFE_CUDA(cudaGetLastError());
  return SFE_OK;
}

// In-place time warp of rows [0, minlen) + time masks (warp_inplace_kernel): the tail of sfe_chain_forward's warp
// schedules.  warp_tail_cols() says whether it applies (Mel bins per CTA, or 0) — to be asked BEFORE the launches in
// front of it are issued.
static int warp_tail_cols(int B, int D, const float* x, const sfe_warp* warp, const sfe_masks* masks) {
  if (D % 4 != 0 || (reinterpret_cast<uintptr_t>(x) & 15)) return 0;
  if (masks && (masks->fill != 0.f || masks->fill_dev)) return 0;  // (interpolating a masked bin must give the fill)
  // about two CTAs per SM, and no more than 200 KB of shared memory to stage minlen rows
  const int per_utt = (2 * cached_sm_count() + B - 1) / B;
  int cols = ((D + per_utt - 1) / per_utt + 3) / 4 * 4;
  const int fit = static_cast<int>((200 * 1024) / (static_cast<long long>(warp->minlen) * 4)) / 4 * 4;
  if (cols > fit) cols = fit;
  if (cols > D) cols = D;
  return cols < 4 ? 0 : cols;
}

static int warp_tail(float* x, int32_t B, int64_t F, int D, int cols, const sfe_warp* warp, const sfe_masks* masks,
                     void* stream) {
  const size_t smem = static_cast<size_t>(warp->minlen) * cols * sizeof(float);
  static std::mutex mu;
  static bool raised = false;
  if (smem > 48 * 1024) {
    std::lock_guard<std::mutex> lk(mu);
    if (!raised) {
      SFE_CUDA(cudaFuncSetAttribute(warp_inplace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      raised = true;
    }
  }
  WarpInPlaceParams P;
  memset(&P, 0, sizeof(P));
  P.x = x;
  P.B = B;
  P.F = F;
  P.D = D;
  P.cols = cols;
  P.mode = warp->enabled == 2 ? 2 : 1;
  P.center = warp->center;
  P.pos = warp->pos;
  P.minlen = warp->minlen;
  if (masks && masks->n_time > 0) {
    sfe_masks tm = *masks;  // time masks only
    tm.n_freq = 0;
    tm.freq_pos = tm.freq_len = nullptr;
    float fill;
    const float* fill_dev;
    if (int rc = fill_masks(&tm, B, &P.mk, &fill, &fill_dev)) return rc;
    P.fill = fill;
  }
  dim3 grid(static_cast<unsigned>((D + cols - 1) / cols), static_cast<unsigned>(B));
  SFE_CUDA(launch_pdl<WarpInPlaceParams>(warp_inplace_kernel, grid, dim3(256), smem, static_cast<cudaStream_t>(stream), P));
  SFE_CUDA(cudaGetLastError());
  return SFE_OK;
}

// frequency masks only (the launch in front of warp_tail applies them to every row)
static const sfe_masks* freq_only(const sfe_masks* masks, sfe_masks* tmp) {
  if (!masks || masks->n_freq <= 0) return nullptr;
  *tmp = *masks;
  tmp->n_time = 0;
  tmp->time_pos = tmp->time_len = nullptr;
  return tmp;
}

extern "C" int sfe_chain_forward(const sfe_plan* pl, const float* wav, int64_t wav_pitch, const int64_t* wav_len,
                                 int32_t B, int64_t T_max, float* mel, int64_t* feat_len, int64_t* stats_ws,
                                 int32_t ws_half, int32_t norm_mode, const float* mean, const float* std, float clamp,
                                 int32_t use_mean, int32_t use_std, float* mean_out, float* std_out,
                                 const sfe_masks* masks, void* stream, const int64_t* wav_len_host,
                                 const sfe_warp* warp, float* warp_out) {
  if (!pl) return fail(SFE_ERR_BAD_ARG, "sfe_chain_forward: NULL plan");
  if (!mel || !feat_len) return fail(SFE_ERR_BAD_ARG, "sfe_chain_forward: mel / feat_len is NULL");
  const bool warp_on = warp && warp->enabled;
  if (warp_on && (!warp_out || warp_out == mel))
    return fail(SFE_ERR_BAD_ARG, "sfe_chain_forward: a time warp needs warp_out, a second (B, F, n_mels) buffer");
  const int64_t F = sfe_plan_num_frames(pl, T_max);
  if (norm_mode == 0 || norm_mode == 1) {
    // masks ride in the frontend kernel's epilogue — unless a time warp comes first (SpecAugment warps, then masks).
    // Then: the kernel writes every row normalised and FREQUENCY-masked into warp_out, and a small kernel warps and
    // time-masks rows [0, minlen) of it in place (warp_chain_in_place); where that is not possible (a fill value
    // other than 0, a minlen too long to stage) the kernel writes `mel` and ONE more pass over the whole tensor
    // gathers the warp and applies the masks.
    if (warp_on) {
      if (!(warp->center > 0 && warp->center < warp->minlen && warp->pos > 0 && warp->pos < warp->minlen && warp->minlen <= F))
        return fail(SFE_ERR_BAD_ARG, "sfe_chain_forward: inconsistent warp descriptor");
      if (const int cols = warp_tail_cols(B, pl->cfg.n_mels, warp_out, warp, masks)) {
        sfe_masks fm;
        if (int rc = sfe_frontend_forward(pl, wav, wav_pitch, wav_len, B, T_max, warp_out, nullptr, nullptr, feat_len,
                                          nullptr, (norm_mode == 1 && use_mean) ? mean : nullptr,
                                          (norm_mode == 1 && use_std) ? std : nullptr, freq_only(masks, &fm), 0, stream,
                                          wav_len_host))
          return rc;
        return warp_tail(warp_out, B, F, pl->cfg.n_mels, cols, warp, masks, stream);
      }
    }
    if (int rc = sfe_frontend_forward(pl, wav, wav_pitch, wav_len, B, T_max, mel, nullptr, nullptr, feat_len, nullptr,
                                      (norm_mode == 1 && use_mean) ? mean : nullptr,
                                      (norm_mode == 1 && use_std) ? std : nullptr, warp_on ? nullptr : masks, 0, stream,
                                      wav_len_host))
      return rc;
    if (!warp_on) return SFE_OK;
    return normalize_impl(mel, warp_out, feat_len, B, F, pl->cfg.n_mels, nullptr, 0, clamp, 0, 0, nullptr, nullptr, 0,
                          nullptr, nullptr, 1, 0.f, warp, masks, stream, nullptr);
  }
  if (norm_mode != 2 && norm_mode != 3) return fail(SFE_ERR_BAD_ARG, "sfe_chain_forward: norm_mode must be 0, 1, 2 or 3");
  if (!stats_ws) return fail(SFE_ERR_BAD_ARG, "sfe_chain_forward: norm_mode 2 / 3 needs stats_ws");
  if (!pl->cfg.mel_logging) return fail(SFE_ERR_BAD_ARG, "sfe_chain_forward: norm_mode 2 / 3 needs a log-mel plan");
  if (norm_mode == 3) {
    // batch / global statistics that are still being averaged: raw rows + sums into one half of stats_ws (the
    // protocol of norm_mode 2), then the packed [sum of means | sum of stds | B] of sfe_stats_finalize into mean_out
    // by a launch that also clears the other half.  The caller all-reduces those 2 n_mels + 1 floats when it runs
    // data-parallel and finishes with sfe_running_normalize.
    if (!mean_out) return fail(SFE_ERR_BAD_ARG, "sfe_chain_forward: norm_mode 3 writes the packed sums to mean_out");
    const size_t half3 = static_cast<size_t>(B) * pl->cfg.n_mels * 2;
    int64_t* cur3 = stats_ws + (ws_half ? half3 : 0);
    int64_t* other3 = stats_ws + (ws_half ? 0 : half3);
    if (int rc = frontend_forward_impl(pl, wav, wav_pitch, wav_len, B, T_max, mel, nullptr, nullptr, feat_len, cur3,
                                       nullptr, nullptr, nullptr, 1, stream, false, wav_len_host))
      return rc;
    return stats_finalize_impl(cur3, 1, feat_len, B, pl->cfg.n_mels, clamp, nullptr, nullptr, mean_out, stream, other3,
                               true);
  }
  // stats_ws holds TWO (B, n_mels, 2) buffers, both zero before the first call: this call accumulates into half
  // `ws_half` and its normalise launch clears the other half for the next call (which must pass 1 - ws_half),
  // so no memset launch is needed
  const size_t half = static_cast<size_t>(B) * pl->cfg.n_mels * 2;
  int64_t* cur = stats_ws + (ws_half ? half : 0);
  int64_t* other = stats_ws + (ws_half ? 0 : half);
  if (int rc = frontend_forward_impl(pl, wav, wav_pitch, wav_len, B, T_max, mel, nullptr, nullptr, feat_len, cur,
                                     nullptr, nullptr, nullptr, 1, stream, false, wav_len_host))
    return rc;
  if (warp_on) {
    if (!(warp->center > 0 && warp->center < warp->minlen && warp->pos > 0 && warp->pos < warp->minlen && warp->minlen <= F))
      return fail(SFE_ERR_BAD_ARG, "sfe_chain_forward: inconsistent warp descriptor");
    if (const int cols = warp_tail_cols(B, pl->cfg.n_mels, warp_out, warp, masks)) {
      // normalise + frequency masks into warp_out (every row), then warp + time masks on its rows [0, minlen) in place
      sfe_masks fm;
      if (int rc = normalize_impl(mel, warp_out, feat_len, B, F, pl->cfg.n_mels, cur, 1, clamp, use_mean, use_std, nullptr,
                                  nullptr, 0, mean_out, std_out, 0, pl->proto.pad_mel, nullptr, freq_only(masks, &fm), stream,
                                  other))
        return rc;
      return warp_tail(warp_out, B, F, pl->cfg.n_mels, cols, warp, masks, stream);
    }
  }
  // in place, or out of place into warp_out with the time-warp gather
  if (int rc = normalize_impl(mel, warp_on ? warp_out : mel, feat_len, B, F, pl->cfg.n_mels, cur, 1, clamp,
                              use_mean, use_std, nullptr, nullptr, 0, mean_out, std_out, 0, pl->proto.pad_mel,
                              warp_on ? warp : nullptr, masks, stream, other))
    return rc;
  return SFE_OK;
}

extern "C" int sfe_running_normalize(const float* packed, int32_t D, float weight, int32_t first, float* run_mean,
                                     float* run_std, int64_t* counter, int64_t count, const float* feat,
                                     const int64_t* feat_len, int32_t B, int64_t F, int32_t src_has_pad, float pad_value,
                                     const sfe_masks* masks, void* stream, const sfe_warp* warp, float* out) {
  if (!packed || D <= 0 || (!run_mean && !run_std))
    return fail(SFE_ERR_BAD_ARG, "sfe_running_normalize: bad argument");
  if (!out) return fail(SFE_ERR_BAD_ARG, "sfe_running_normalize: NULL destination");
  const bool warp_on = warp && warp->enabled;
  if (warp_on && out == feat)
    return fail(SFE_ERR_BAD_ARG, "sfe_running_normalize: the time warp cannot run in place");
  // A plain launch on purpose.  With programmatic serialisation here as well, the whole chain frontend -> finalise ->
  // this -> normalise launches early, and the normalise CTAs take the SMs the frontend kernel frees FIRST, four per
  // SM: a grid that does not fill every slot (C4: 512 CTAs for 592) then leaves SMs idle for the whole pass
  // (C4 step 606 us against 574 with the chain broken here; C2-sized steps lose 1 us to the launch gap).
  running_update_kernel<<<(D + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      packed, D, weight, first, 1, run_mean, run_std, nullptr, nullptr, reinterpret_cast<long long*>(counter), count);
  SFE_CUDA(cudaGetLastError());
  if (warp_on) {
    if (!(warp->center > 0 && warp->center < warp->minlen && warp->pos > 0 && warp->pos < warp->minlen && warp->minlen <= F))
      return fail(SFE_ERR_BAD_ARG, "sfe_running_normalize: inconsistent warp descriptor");
    if (const int cols = warp_tail_cols(B, D, out, warp, masks)) {
      // normalise + frequency masks into `out` (every row), then warp + time masks on its rows [0, minlen) in place
      sfe_masks fm;
      if (int rc = normalize_impl(feat, out, feat_len, B, F, D, nullptr, 0, 0.f, 0, 0, run_mean, run_std, 0, nullptr,
                                  nullptr, src_has_pad, pad_value, nullptr, freq_only(masks, &fm), stream, nullptr))
        return rc;
      return warp_tail(out, B, F, D, cols, warp, masks, stream);
    }
  }
  return normalize_impl(feat, out, feat_len, B, F, D, nullptr, 0, 0.f, 0, 0, run_mean, run_std, 0, nullptr, nullptr,
                        src_has_pad, pad_value, warp_on ? warp : nullptr, masks, stream, nullptr);
}

// Row groups of the ragged upload: [b, e) travels as one pitched copy of width w (see include/sfe.h).
static bool upload_grouping() {
  static const bool group = [] {
    const char* ev = getenv("SFE_UPLOAD_GROUP");  // SFE_UPLOAD_GROUP=0: one copy per row
    return !(ev && ev[0] == '0');
  }();
  return group;
}
static int upload_next_group(const int64_t* len_host, int B, int b, int64_t* width) {
  static const int den = [] {  // a group's padding stays below 1/den of its bytes (tuning knob; measured on C2: 1/3 0.57 M, 1/5 0.69 M, 1/8 0.73 M, 1/12 0.754 M, 1/24 0.755 M audio-s/s, one copy per row 0.64 M)
    const char* ev = getenv("SFE_UPLOAD_WASTE_DEN");
    const int v = ev ? atoi(ev) : 24;
    return v >= 2 && v <= 1024 ? v : 24;
  }();
  int64_t w = len_host[b], sum = len_host[b];
  int e = b + 1;
  while (upload_grouping() && e < B && e - b < 64) {
    const int64_t w2 = w > len_host[e] ? w : len_host[e], s2 = sum + len_host[e];
    const int64_t area = w2 * (e - b + 1);
    if (area - s2 > area / den) break;
    w = w2;
    sum = s2;
    ++e;
  }
  *width = w;
  return e;
}

extern "C" int64_t sfe_upload_bytes(const int64_t* len_host, int32_t B) {
  if (!len_host || B <= 0) return 0;
  int64_t total = 0;
  for (int b = 0; b < B;) {
    int64_t w;
    const int e = upload_next_group(len_host, B, b, &w);
    total += w * (e - b) * static_cast<int64_t>(sizeof(float));
    b = e;
  }
  return total;
}

extern "C" int sfe_upload_waveforms(const float* wav_host, int64_t host_pitch, const int64_t* len_host, int32_t B,
                                    float* wav_dev, int64_t dev_pitch, void* stream) {
  if (!wav_host || !len_host || !wav_dev || B <= 0) return fail(SFE_ERR_BAD_ARG, "sfe_upload_waveforms: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int b = 0; b < B; ++b)
    if (len_host[b] < 0 || len_host[b] > host_pitch || len_host[b] > dev_pitch)
      return fail(SFE_ERR_BAD_ARG, "sfe_upload_waveforms: length exceeds the row pitch");
  // a DMA descriptor per row costs more than the few padding samples a pitched copy of a row group drags along
  for (int b = 0; b < B;) {
    int64_t w;
    const int e = upload_next_group(len_host, B, b, &w);
    if (w > 0) {
      if (e - b == 1)
        SFE_CUDA(cudaMemcpyAsync(wav_dev + b * dev_pitch, wav_host + b * host_pitch, static_cast<size_t>(w) * sizeof(float),
                                 cudaMemcpyHostToDevice, st));
      else
        SFE_CUDA(cudaMemcpy2DAsync(wav_dev + b * dev_pitch, static_cast<size_t>(dev_pitch) * sizeof(float),
                                   wav_host + b * host_pitch, static_cast<size_t>(host_pitch) * sizeof(float),
                                   static_cast<size_t>(w) * sizeof(float), static_cast<size_t>(e - b),
                                   cudaMemcpyHostToDevice, st));
    }
    b = e;
  }
  return SFE_OK;
}

extern "C" int sfe_prestft_norm(const float* wav, int64_t wav_pitch, const int64_t* wav_len, int32_t B, int64_t T,
                                int32_t has_preemphasis, float preemphasis, int32_t mode, double* scratch, float* out,
                                int64_t out_pitch, void* stream) {
  if (!wav || !wav_len || !scratch || !out || B <= 0 || T <= 0 || wav_pitch < T || out_pitch < T)
    return fail(SFE_ERR_BAD_ARG, "sfe_prestft_norm: bad argument");
  if (mode != 1 && mode != 2) return fail(SFE_ERR_BAD_ARG, "sfe_prestft_norm: mode must be 1 (mean_std) or 2 (min_max)");
  if (mode == 1 && B != 1)
    return fail(SFE_ERR_BAD_ARG, "pre_stft_norm='mean_std' only broadcasts for batch size 1 (speech2linear.py:179 has no keepdim)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  prenorm_stats_kernel<<<static_cast<unsigned>(B), 1024, 0, st>>>(wav, wav_pitch, reinterpret_cast<const long long*>(wav_len),
                                                                 T, has_preemphasis, preemphasis, mode, scratch);
  SFE_CUDA(cudaGetLastError());
  int64_t chunks = (T + 256 * 8 - 1) / (256 * 8);
  const int64_t cap = std::max<int64_t>(1, static_cast<int64_t>(cached_sm_count()) * 8 / B);
  if (chunks > cap) chunks = cap;
  prenorm_apply_kernel<<<dim3(static_cast<unsigned>(chunks), static_cast<unsigned>(B)), 256, 0, st>>>(
      wav, wav_pitch, reinterpret_cast<const long long*>(wav_len), T, has_preemphasis, preemphasis, mode, scratch, out, out_pitch);
  SFE_CUDA(cudaGetLastError());
  return SFE_OK;
}

extern "C" int64_t sfe_upload_i16_stage_bytes(int32_t B, int64_t T) {
  if (B <= 0 || T <= 0) return 0;
  const int64_t pitch = (T + 7) / 8 * 8;  // rows start on 16 bytes
  return (static_cast<int64_t>(B) * 8 + 255) / 256 * 256 + static_cast<int64_t>(B) * pitch * 2;
}

extern "C" int sfe_upload_waveforms_i16(const int16_t* wav_host, int64_t host_pitch, const int64_t* len_host, int32_t B,
                                        void* stage_dev, int64_t stage_bytes, float* wav_dev, int64_t dev_pitch,
                                        void* stream) {
  if (!wav_host || !len_host || !stage_dev || !wav_dev || B <= 0)
    return fail(SFE_ERR_BAD_ARG, "sfe_upload_waveforms_i16: bad argument");
  int64_t T = 0;
  for (int b = 0; b < B; ++b) {
    if (len_host[b] < 0 || len_host[b] > host_pitch || len_host[b] > dev_pitch)
      return fail(SFE_ERR_BAD_ARG, "sfe_upload_waveforms_i16: length exceeds the row pitch");
    T = std::max<int64_t>(T, len_host[b]);
  }
  if (T == 0) return SFE_OK;
  if (stage_bytes < sfe_upload_i16_stage_bytes(B, T)) return fail(SFE_ERR_BAD_ARG, "sfe_upload_waveforms_i16: staging buffer too small");
  if (reinterpret_cast<uintptr_t>(stage_dev) & 15) return fail(SFE_ERR_BAD_ARG, "sfe_upload_waveforms_i16: staging buffer must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t pitch = (T + 7) / 8 * 8;
  long long* len_dev = static_cast<long long*>(stage_dev);
  short* rows = reinterpret_cast<short*>(static_cast<unsigned char*>(stage_dev) + (static_cast<int64_t>(B) * 8 + 255) / 256 * 256);
  SFE_CUDA(cudaMemcpyAsync(len_dev, len_host, sizeof(int64_t) * B, cudaMemcpyHostToDevice, st));
  for (int b = 0; b < B;) {  // same row groups as the fp32 upload
    int64_t w;
    const int e = upload_next_group(len_host, B, b, &w);
    if (w > 0) {
      if (e - b == 1)
        SFE_CUDA(cudaMemcpyAsync(rows + b * pitch, wav_host + b * host_pitch, static_cast<size_t>(w) * 2, cudaMemcpyHostToDevice, st));
      else
        SFE_CUDA(cudaMemcpy2DAsync(rows + b * pitch, static_cast<size_t>(pitch) * 2, wav_host + b * host_pitch,
                                   static_cast<size_t>(host_pitch) * 2, static_cast<size_t>(w) * 2,
                                   static_cast<size_t>(e - b), cudaMemcpyHostToDevice, st));
    }
    b = e;
  }
  const int vec_ok = !(reinterpret_cast<uintptr_t>(wav_dev) & 15) && (dev_pitch % 4 == 0);
  int64_t chunks = (T + 256 * 8 - 1) / (256 * 8);
  const int64_t cap = std::max<int64_t>(1, static_cast<int64_t>(cached_sm_count()) * 8 / B);
  if (chunks > cap) chunks = cap;
  pcm16_to_float_kernel<<<dim3(static_cast<unsigned>(chunks), static_cast<unsigned>(B)), 256, 0, st>>>(
      len_dev, rows, pitch, wav_dev, dev_pitch, vec_ok);
  SFE_CUDA(cudaGetLastError());
  return SFE_OK;
}

extern "C" int sfe_delta(const float* feat, int32_t B, int64_t F, int32_t D, int32_t order, int32_t N, float* out,
                         void* stream) {
  if (!feat || !out) return fail(SFE_ERR_BAD_ARG, "sfe_delta: NULL argument");
  if (B <= 0 || F <= 0 || D <= 0 || (order != 1 && order != 2) || N < 1) return fail(SFE_ERR_BAD_ARG, "sfe_delta: bad argument");
  int denom = 0;
  for (int i = 1; i <= N; ++i) denom += i * i;
  const long long total = static_cast<long long>(B) * F * D;
  long long blocks = (total + 255) / 256;
  if (blocks > cached_sm_count() * 16LL) blocks = cached_sm_count() * 16LL;
  delta_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      feat, B, F, D, order, N, static_cast<float>(2 * denom), out);
  SFE_CUDA(cudaGetLastError());
  return SFE_OK;
}

extern "C" int sfe_twosided(const float* power, const int64_t* feat_len, int32_t B, int64_t F, int32_t K, int32_t n_fft,
                            int32_t mag_spec, int32_t logging, float clamp, float log_div, float* out, float* energy,
                            void* stream) {
  if (!power || !feat_len || !out) return fail(SFE_ERR_BAD_ARG, "sfe_twosided: NULL argument");
  if (B <= 0 || F <= 0 || n_fft < 2 || K != n_fft / 2 + 1) return fail(SFE_ERR_BAD_ARG, "sfe_twosided: bad shape");
  if (logging && !(clamp >= 1.17549435e-38f)) return fail(SFE_ERR_BAD_ARG, "sfe_twosided: clamp must be a normal positive number");
  const long long rows = static_cast<long long>(B) * F;
  long long blocks = (rows + 7) / 8;
  if (blocks > cached_sm_count() * 16LL) blocks = cached_sm_count() * 16LL;
  twosided_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      power, reinterpret_cast<const long long*>(feat_len), rows, F, K, n_fft, mag_spec, logging, clamp,
      0.69314718055994530942f / log_div, out, energy);
  SFE_CUDA(cudaGetLastError());
  return SFE_OK;
}

// scratch accumulator for sfe_mean_all lives with the library (one word per device context)
extern "C" int sfe_mean_all(const float* x, int64_t n, float* out, void* stream) {
  if (!x || !out || n <= 0) return fail(SFE_ERR_BAD_ARG, "sfe_mean_all: bad argument");
  static thread_local double* acc = nullptr;
  static thread_local int acc_dev = -1;
  int dev = 0;
  SFE_CUDA(cudaGetDevice(&dev));
  if (!acc || acc_dev != dev) {
    SFE_CUDA(cudaMalloc(&acc, sizeof(double)));
    SFE_CUDA(cudaMemset(acc, 0, sizeof(double)));
    acc_dev = dev;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > cached_sm_count() * 8LL) blocks = cached_sm_count() * 8LL;
  // the word is cleared on THIS call's stream: an aborted earlier call (or another stream driven by the same
  // host thread) cannot leave it dirty
  SFE_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
  sum_all_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, n, acc);
  SFE_CUDA(cudaGetLastError());
  mean_finish_kernel<<<1, 1, 0, st>>>(acc, n, out);
  SFE_CUDA(cudaGetLastError());
  return SFE_OK;
}

extern "C" int sfe_mel2linear(const float* mel, const int64_t* feat_len, int32_t B, int64_t F, int32_t D, int32_t K,
                              const float* pinv_t, int32_t log_mode, float log_base, int64_t max_len, float clamp,
                              float* lin, void* stream) {
  if (!mel || !feat_len || !pinv_t || !lin || B < 0 || F < 0 || D <= 0 || K <= 0 || log_mode < 0 || log_mode > 2)
    return fail(SFE_ERR_BAD_ARG, "sfe_mel2linear: bad argument");
  if (D > 32 * MAX_MEL_ROUNDS) return fail(SFE_ERR_BAD_ARG, "sfe_mel2linear: n_mels must be <= " + std::to_string(32 * MAX_MEL_ROUNDS));
  const long long n_rows = static_cast<long long>(B) * F;
  if (n_rows == 0) return SFE_OK;
  const int smem = D * M2L_ROWS * static_cast<int>(sizeof(float));
  if (smem > 48 * 1024) SFE_CUDA(raise_dyn_smem(reinterpret_cast<const void*>(&mel2linear_kernel), smem));
  const long long grid = (n_rows + M2L_ROWS - 1) / M2L_ROWS;
  if (grid > 0x7fffffffLL) return fail(SFE_ERR_BAD_ARG, "sfe_mel2linear: too many rows");
  mel2linear_kernel<<<static_cast<unsigned>(grid), M2L_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      mel, reinterpret_cast<const long long*>(feat_len), n_rows, F, D, K, pinv_t, log_mode, log_base, max_len, clamp, lin);
  SFE_CUDA(cudaGetLastError());
  return SFE_OK;
}

// ---------------------------------------------------------------------------------------------
// Griffin-Lim
// ---------------------------------------------------------------------------------------------
static size_t al256(size_t x) { return (x + 255) / 256 * 256; }

extern "C" int64_t sfe_griffinlim_workspace(const sfe_plan* pl, int32_t B, int64_t F) {
  if (!pl || B < 0 || F < 0) return -1;
  const size_t n = static_cast<size_t>(B) * F * (pl->cfg.n_fft / 2 + 1);
  const size_t L = F > 0 ? static_cast<size_t>(pl->cfg.hop_length) * (F - 1) : 0;
  return static_cast<int64_t>(al256(n * 4) + 2 * al256(n * 8) + al256(static_cast<size_t>(B) * F * pl->cfg.n_fft * 4) +
                              al256(static_cast<size_t>(B) * L * 4));
}

extern "C" int sfe_griffinlim(sfe_plan* pl, const float* spec_in, const int64_t* feat_len, int32_t B, int64_t F,
                              int32_t n_iter, float momentum, const float* init_angles, int32_t inv_preemph_taps,
                              void* work, int64_t work_bytes, float* wav, int64_t* wav_len, void* stream) {
  if (!pl || !spec_in || !feat_len || !work || !wav || B <= 0 || n_iter < 0 || !(momentum >= 0.f && momentum < 1.f) ||
      inv_preemph_taps < 0)
    return fail(SFE_ERR_BAD_ARG, "sfe_griffinlim: bad argument");
  const int n_fft = pl->cfg.n_fft, hop = pl->cfg.hop_length, win = pl->cfg.win_length, K = n_fft / 2 + 1;
  if (!pl->cfg.center) return fail(SFE_ERR_BAD_CONFIG, "sfe_griffinlim: needs center=True (torch.istft trims n_fft/2 per side)");
  const long long L = static_cast<long long>(hop) * (F - 1);
  if (F < 2 || L <= n_fft / 2)
    return fail(SFE_ERR_BAD_ARG, "sfe_griffinlim: too few frames (reflect padding needs hop*(F-1) > n_fft/2)");
  if (work_bytes < sfe_griffinlim_workspace(pl, B, F)) return fail(SFE_ERR_BAD_ARG, "sfe_griffinlim: workspace too small");
  const FftEntry* fe = nullptr;
  for (const auto& e : registry())
    if (e.n_fft == n_fft && e.win == 0) fe = &e;
  if (!fe || !fe->gl_stft || fe->R1 != pl->fft->R1 || fe->R2 != pl->fft->R2 || fe->R3 != pl->fft->R3 || fe->MP != pl->fft->MP)
    return fail(SFE_ERR_UNSUPPORTED_NFFT, "sfe_griffinlim: no generic-window plan for this n_fft");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool preemph = pl->cfg.has_preemphasis && inv_preemph_taps > 0;

  // ---- constant tables, once per plan (and per tap count) ----
  const int ones_floats = round_up(n_fft + n_fft / 2 + 8, 4);
  if (!pl->d_gl || pl->gl_taps != inv_preemph_taps) {
    std::vector<float> t(ones_floats, 1.f);
    pl->gl_o_synth = static_cast<int>(t.size());
    const int lp = (n_fft - win) / 2;
    for (int j = 0; j < n_fft; ++j) t.push_back((j >= lp && j < lp + win) ? pl->h_window[j - lp] / static_cast<float>(n_fft) : 0.f);
    pl->gl_o_wsq = static_cast<int>(t.size());
    for (int j = 0; j < n_fft; ++j) t.push_back((j >= lp && j < lp + win) ? pl->h_window[j - lp] * pl->h_window[j - lp] : 0.f);
    pl->gl_o_pw = static_cast<int>(t.size());
    for (int j = 0; j < inv_preemph_taps; ++j) t.push_back(powf(pl->cfg.preemphasis, static_cast<float>(j)));
    while (t.size() % 4) t.push_back(0.f);
    pl->gl_o_awin = static_cast<int>(t.size());  // analysis window x 1/2 at its centred position (phase matters here)
    for (int j = 0; j < n_fft; ++j) t.push_back((j >= lp && j < lp + win) ? 0.5f * pl->h_window[j - lp] : 0.f);
    if (pl->d_gl) {
      SFE_CUDA(cudaStreamSynchronize(st));
      SFE_CUDA(cudaFree(pl->d_gl));
      pl->d_gl = nullptr;
    }
    SFE_CUDA(cudaMalloc(&pl->d_gl, t.size() * sizeof(float)));
    SFE_CUDA(cudaMemcpy(pl->d_gl, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
    pl->gl_taps = inv_preemph_taps;
  }

  // ---- workspace carve-up ----
  const size_t n = static_cast<size_t>(B) * F * K;
  unsigned char* w = static_cast<unsigned char*>(work);
  float* mag = reinterpret_cast<float*>(w); w += al256(n * 4);
  float2* tprev = reinterpret_cast<float2*>(w); w += al256(n * 8);
  float2* spec = reinterpret_cast<float2*>(w); w += al256(n * 8);
  float* frames = reinterpret_cast<float*>(w); w += al256(static_cast<size_t>(B) * F * n_fft * 4);
  float* tmp = reinterpret_cast<float*>(w);

  GlParams P;
  memset(&P, 0, sizeof(P));
  P.B = B; P.F = static_cast<int>(F); P.K = K; P.n_fft = n_fft; P.win = win; P.hop = hop;
  P.L = L;
  P.n_frames = static_cast<long long>(B) * F;
  P.tables = pl->d_tables;
  P.tab_floats = pl->proto.o_mtab;  // window, Stockham twiddles, unpack twiddles (everything before the mel tables)
  P.o_win = pl->proto.o_win; P.o_tw2 = pl->proto.o_tw2; P.o_tw3 = pl->proto.o_tw3; P.o_wpost = pl->proto.o_wpost;
  P.gl_tab = pl->d_gl;
  P.o_ones = 0; P.o_synth = pl->gl_o_synth; P.o_wsq = pl->gl_o_wsq; P.o_awin = pl->gl_o_awin;
  P.wav = tmp; P.mag = mag; P.tprev = tprev; P.spec = spec; P.frames = frames;
  P.momentum = momentum / (1.f + momentum);
  P.stage_floats = ones_floats;
  P.sm_warp = round_up(P.tab_floats * 4, 16);
  P.sm_stage_off = round_up(fe->C * fe->MP * 8, 16);
  P.sm_warp_bytes = P.sm_stage_off + round_up(fe->C * P.stage_floats * 4, 16);
  const int smem = P.sm_warp + GL_WARPS * P.sm_warp_bytes;
  SFE_CUDA(raise_dyn_smem(reinterpret_cast<const void*>(fe->gl_stft), smem));
  SFE_CUDA(raise_dyn_smem(reinterpret_cast<const void*>(fe->gl_istft), smem));
  const long long groups = (P.n_frames + fe->C - 1) / fe->C;
  long long gridf = (groups + GL_WARPS - 1) / GL_WARPS;
  if (gridf > pl->n_sm * 8LL) gridf = pl->n_sm * 8LL;
  const unsigned grid_n = static_cast<unsigned>((n + 255) / 256);
  const long long BL = static_cast<long long>(B) * L;
  const unsigned grid_w = static_cast<unsigned>(((BL > B ? BL : B) + 255) / 256);

  gl_init_kernel<<<grid_n, 256, 0, st>>>(spec_in, reinterpret_cast<const float2*>(init_angles), static_cast<long long>(n),
                                         pl->cfg.mag_spec ? 0 : 1, mag, spec);
  for (int it = 0; it <= n_iter; ++it) {
    fe->gl_istft<<<static_cast<unsigned>(gridf), GL_THREADS, smem, st>>>(P);
    gl_ola_kernel<<<grid_w, 256, 0, st>>>(frames, pl->d_gl + pl->gl_o_wsq, B, static_cast<int>(F), n_fft, hop, L, tmp);
    if (it == n_iter) break;
    P.first = (it == 0) ? 1 : 0;
    fe->gl_stft<<<static_cast<unsigned>(gridf), GL_THREADS, smem, st>>>(P);
  }
  gl_finish_kernel<<<grid_w, 256, 0, st>>>(tmp, preemph ? pl->d_gl + pl->gl_o_pw : nullptr, inv_preemph_taps,
                                           reinterpret_cast<const long long*>(feat_len), hop, B, L, wav,
                                           reinterpret_cast<long long*>(wav_len));
  SFE_CUDA(cudaGetLastError());
  return SFE_OK;
}
