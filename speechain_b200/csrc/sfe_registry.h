// FFT plan registry: n_fft (and optionally a compile-time window length) -> kernel instances.
// Every plan is its own translation unit (sfe_plans.cu compiled with -DSFE_PLAN_TU=i), so the library
// builds in parallel; sfe_lib.cu only sees the table of entry getters below.
#pragma once
#include "sfe_fused.cuh"
#include "sfe_gl.cuh"
#include "sfe_tile.cuh"

namespace sfe {

typedef void (*fused_fn)(const FwdParams);
typedef void (*gl_fn)(const GlParams);

struct FftEntry {
  int n_fft, win;  // win == 0: any window length (run-time tests)
  int M, G, R1, R2, R3, MP, C, MAXW;
  fused_fn fn[2];  // [0] no pre-emphasis, [1] pre-emphasis applied while loading
  gl_fn gl_stft, gl_istft;  // Griffin-Lim transforms (generic-window entries only)
  // tile kernel (sfe_tile.cuh; nullptr where the plan has none): [0] plain, [1] pre-emphasis, [2] energy output
  fused_fn t2[3];
  int t2_ts, t2_nw, t2_bank, t2_nf, t2_npb;  // pairs per tile, warps per CTA, compiled Mel program (sfe_melprog.h), FFT warps, power-spectrum buffers
};

// (M, G, R1, R2, R3, PADQ, WIN) of every compiled plan, in registry order
#define SFE_PLAN_LIST(X)                                                                              \
  X(0, 64, 16, 8, 8, 1, 8, 0)        /* n_fft 128 */                                                  \
  X(1, 100, 16, 4, 5, 5, 4, 0)       /* n_fft 200 */                                                  \
  X(2, 128, 16, 8, 16, 1, 8, 0)      /* n_fft 256 */                                                  \
  X(3, 160, 16, 8, 5, 4, 8, 0)       /* n_fft 320  (20 ms @ 16 kHz) */                                \
  X(4, 200, 16, 8, 5, 5, 8, 0)       /* n_fft 400  (25 ms @ 16 kHz: the ASR recipes' default) */      \
  X(5, 240, 16, 16, 5, 3, 16, 0)     /* n_fft 480 */                                                  \
  X(6, 256, 16, 16, 16, 1, 16, 0)    /* n_fft 512 */                                                  \
  X(7, 320, 32, 16, 5, 4, 16, 0)     /* n_fft 640 */                                                  \
  X(8, 400, 32, 16, 5, 5, 16, 0)     /* n_fft 800 */                                                  \
  X(9, 512, 32, 4, 8, 16, 16, 0)     /* n_fft 1024 (TTS recipes): last radix-16 pass in registers */  \
  X(10, 1024, 32, 16, 8, 8, 16, 0)   /* n_fft 2048 */                                                 \
  /* window length fixed at compile time for the (n_fft, win_length) pairs of the shipped recipes */  \
  X(11, 200, 16, 8, 5, 5, 8, 400)    /* 25 ms window @16 kHz, n_fft = win (ASR recipe default) */     \
  X(12, 256, 16, 16, 16, 1, 16, 400) /* 25 ms window @16 kHz in a 512-point FFT (BASELINE C1/C2) */   \
  X(13, 512, 32, 4, 8, 16, 16, 1024) /* 1024/1024 (TTS recipes, BASELINE C3) */

#define SFE_DECL_TU(i, M, G, R1, R2, R3, Q, W) FftEntry sfe_plan_entry_##i();
SFE_PLAN_LIST(SFE_DECL_TU)
#undef SFE_DECL_TU

typedef FftEntry (*plan_entry_fn)();
#define SFE_TU_PTR(i, M, G, R1, R2, R3, Q, W) &sfe_plan_entry_##i,
static const plan_entry_fn kPlanTus[] = {SFE_PLAN_LIST(SFE_TU_PTR)};
#undef SFE_TU_PTR
constexpr int kNumPlanTus = sizeof(kPlanTus) / sizeof(kPlanTus[0]);

}  // namespace sfe
