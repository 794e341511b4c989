// One FFT plan's kernel instances per translation unit: compile with -DSFE_PLAN_TU=<index in SFE_PLAN_LIST>.
#include "sfe_registry.h"

#ifndef SFE_PLAN_TU
#error "compile with -DSFE_PLAN_TU=<plan index>"
#endif

namespace sfe {

template <int I>
struct PlanOf;
#define SFE_DEF_PLAN(i, M, G, R1, R2, R3, Q, W) \
  template <>                                   \
  struct PlanOf<i> {                            \
    typedef FftPlan<M, G, R1, R2, R3, Q, W> type; \
  };
SFE_PLAN_LIST(SFE_DEF_PLAN)
#undef SFE_DEF_PLAN

#define SFE_CAT_(a, b) a##b
#define SFE_CAT(a, b) SFE_CAT_(a, b)

template <class PL>
static FftEntry make_entry() {
  FftEntry e{2 * PL::M, PL::WIN, PL::M, PL::G, PL::R1, PL::R2, PL::R3, PL::MP, PL::C, PL::MAXW,
             {&fused_frontend_kernel<PL, false>, &fused_frontend_kernel<PL, true>}, nullptr, nullptr,
             {nullptr, nullptr, nullptr}, 0, 0, 0, 0, 0};
  if constexpr (TileCfg<PL>::SUPPORTED) {
    e.t2[0] = &tile_frontend_kernel<PL, false, false>;
    e.t2[1] = &tile_frontend_kernel<PL, true, false>;
    e.t2[2] = &tile_frontend_kernel<PL, false, true>;
    e.t2_ts = TileCfg<PL>::TS;
    e.t2_nw = TileCfg<PL>::NW;
    e.t2_bank = TileCfg<PL>::BANK;
    e.t2_nf = TileCfg<PL>::NF;
    e.t2_npb = TileCfg<PL>::NPB;
  }
  if constexpr (PL::WIN == 0) {
    e.gl_stft = &gl_stft_kernel<PL>;
    e.gl_istft = &gl_istft_kernel<PL>;
  }
  return e;
}

FftEntry SFE_CAT(sfe_plan_entry_, SFE_PLAN_TU)() { return make_entry<PlanOf<SFE_PLAN_TU>::type>(); }

}  // namespace sfe

#if defined(SFE_T2_TRACE)
// development only: copy the trace of the tile kernel out (tools/trace_tile.py)
extern "C" int sfe_debug_trace(unsigned long long* t, unsigned int* tag, int cap) {
  unsigned int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, sfe::g_t2_trace_n, sizeof(n));
  if (n > (1u << 15)) n = 1u << 15;
  if (static_cast<int>(n) > cap) n = cap;
  cudaMemcpyFromSymbol(t, sfe::g_t2_trace_t, sizeof(unsigned long long) * n);
  cudaMemcpyFromSymbol(tag, sfe::g_t2_trace_tag, sizeof(unsigned int) * n);
  const unsigned int zero = 0;
  cudaMemcpyToSymbol(sfe::g_t2_trace_n, &zero, sizeof(zero));
  return static_cast<int>(n);
}
#endif

#if defined(SFE_T2_TRACE)
extern "C" int sfe_debug_span(unsigned long long* out) {  // development only: [cta][warp][entry, exit] ns
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, sfe::g_t2_span, sizeof(unsigned long long) * 2 * 32 * 256);
  return 0;
}
#endif

#if defined(SFE_T2_PROF)
extern "C" int sfe_debug_prof(long long* out) {  // development only (tools/prof_tile.py)
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, sfe::g_t2_prof, sizeof(long long) * 32 * 16);
  return 0;
}
#endif
