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
             {&fused_frontend_kernel<PL, false>, &fused_frontend_kernel<PL, true>}, nullptr, nullptr};
  if constexpr (PL::WIN == 0) {
    e.gl_stft = &gl_stft_kernel<PL>;
    e.gl_istft = &gl_istft_kernel<PL>;
  }
  return e;
}

FftEntry SFE_CAT(sfe_plan_entry_, SFE_PLAN_TU)() { return make_entry<PlanOf<SFE_PLAN_TU>::type>(); }

}  // namespace sfe
