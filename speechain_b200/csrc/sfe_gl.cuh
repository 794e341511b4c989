// Griffin-Lim phase recovery (Speech2LinearSpec.recover, speech2linear.py:252-351; the iteration itself is
// torchaudio.functional.griffinlim, restated in oracle/frontend_np.py:griffinlim).
//
// One iteration = inverse STFT of (magnitude x current phase) followed by a forward STFT of the result and the
// momentum phase update.  Both transforms run on the frontend's own Stockham passes:
//   gl_istft_kernel  spectrum (B, F, K) complex -> synthesis-windowed frames (B, F, n_fft)
//                    (inverse real FFT as a forward half-size FFT of the conjugated, re-packed spectrum)
//   gl_ola_kernel    frames -> waveform (B, hop (F-1)): overlap-add as a gather, divided by the window envelope
//   gl_stft_kernel   waveform -> rebuilt spectrum; angles = rebuilt - m * previous; next spectrum = magnitude *
//                    angles / (|angles| + 1e-16), all in the epilogue of the FFT (nothing complex is re-read)
// A warp owns PL::C frames at a time (the lane grouping of the frontend kernel); scratch is warp-private.
#pragma once
#include "sfe_fused.cuh"

namespace sfe {

constexpr int GL_WARPS = 4;
constexpr int GL_THREADS = GL_WARPS * 32;

struct GlParams {
  int B, F, K, n_fft, win, hop;
  long long L;         // samples per waveform row = hop * (F - 1)
  long long n_frames;  // B * F
  // frontend tables (window x 1/2, Stockham twiddles, unpack twiddles) and their float offsets
  const float* tables;
  int tab_floats;
  int o_win, o_tw2, o_tw3, o_wpost;
  // Griffin-Lim tables: [ones (n_fft * 3 / 2 + 8) | synthesis window / n_fft (n_fft) | window^2 (n_fft) |
  //                       inverse pre-emphasis taps | analysis window x 1/2 centred in n_fft (n_fft)]
  const float* gl_tab;
  int o_ones, o_synth, o_wsq, o_awin;
  // data
  const float* wav;     // (B, L)
  const float* mag;     // (B, F, K)
  float2* tprev;        // (B, F, K) previous rebuilt spectrum
  float2* spec;         // (B, F, K) magnitude x phase estimate
  float* frames;        // (B, F, n_fft)
  float momentum;       // already m / (1 + m)
  int first;            // 1: no previous iterate yet (tprev is write-only)
  // shared-memory carve-up (bytes)
  int sm_warp, sm_warp_bytes, sm_stage_off, stage_floats;
};

template <class PL>
__device__ __forceinline__ void gl_setup(const GlParams& p, unsigned char* smem, int& g, int& l, cpx*& Zw, float*& stage_w,
                                         float*& stage_g) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* tab = reinterpret_cast<float*>(smem);
  for (int i = tid; i < p.tab_floats; i += GL_THREADS) tab[i] = __ldg(p.tables + i);
  __syncthreads();
  g = lane / PL::G;
  l = lane % PL::G;
  unsigned char* wbase = smem + p.sm_warp + warp * p.sm_warp_bytes;
  Zw = reinterpret_cast<cpx*>(wbase) + g * PL::MP;
  stage_w = reinterpret_cast<float*>(wbase + p.sm_stage_off);
  stage_g = stage_w + g * p.stage_floats;
}

// Runs the Stockham passes over `fr` and hands the M outputs of this lane's frame to `emit(k, value)`.
template <class PL, class Emit>
__device__ __forceinline__ void gl_fft_all(cpx* Zw, const cpx* tw2, const cpx* tw3, int l, const float* fr,
                                           const float* wtab, int win, Emit emit) {
  constexpr int M = PL::M, G = PL::G, Q = PL::PADQ;
  fft_frame<PL, true, false>(Zw, tw2, tw3, l, fr, wtab, win, 0.f);
  if constexpr (PL::LASTREG) {
    constexpr int RL = PL::RL;
    cpx u[RL];
    stockham_last_reg<PL, RL, M / RL>(Zw, (PL::R3 > 1) ? tw3 : tw2, l, u);
#pragma unroll
    for (int i = 0; i < RL; ++i) emit(l + G * i, u[i]);
  } else {
    for (int k = l; k < M; k += G) emit(k, Zw[padidx<Q>(k)]);
  }
}

// ---------------------------------------------------------------------------------------------
// spectrum -> windowed frames
// ---------------------------------------------------------------------------------------------
template <class PL>
__global__ void __launch_bounds__(GL_THREADS) gl_istft_kernel(const __grid_constant__ GlParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int M = PL::M, C = PL::C;
  int g, l;
  cpx* Zw;
  float *stage_w, *stage_g;
  gl_setup<PL>(p, smem, g, l, Zw, stage_w, stage_g);
  const float* tab = reinterpret_cast<const float*>(smem);
  const cpx* tw2 = reinterpret_cast<const cpx*>(tab + p.o_tw2);
  const cpx* tw3 = reinterpret_cast<const cpx*>(tab + p.o_tw3);
  const cpx* wpost = reinterpret_cast<const cpx*>(tab + p.o_wpost);
  const float* ones = p.gl_tab + p.o_ones;
  const float* synth = p.gl_tab + p.o_synth;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (long long f0 = (static_cast<long long>(blockIdx.x) * GL_WARPS + warp) * C; f0 < p.n_frames;
       f0 += static_cast<long long>(gridDim.x) * GL_WARPS * C) {
    // Re-pack the one-sided spectrum X[0..M] of a real signal into the half-size complex spectrum
    // Z2[k] = (X[k] + conj X[M-k]) + i (X[k] - conj X[M-k]) conj(W^k)  (twice the textbook Z: the 1/2 joins the
    // 1/n_fft in the synthesis window), stored conjugated: an inverse FFT is conj(FFT(conj(.))).
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const long long fi = f0 + c;
      cpx* st = reinterpret_cast<cpx*>(stage_w + c * p.stage_floats);
      if (fi < p.n_frames) {
        const float2* X = p.spec + fi * p.K;
        for (int k = lane; k <= M / 2; k += 32) {
          const int m = M - k;
          cpx A = X[k], Bm = X[m];
          if (k == 0) { A.y = 0.f; Bm.y = 0.f; }  // a real signal's DC and Nyquist bins are real (C2R ignores the rest)
          const cpx E2 = make_float2(A.x + Bm.x, A.y - Bm.y);
          const cpx Dd = make_float2(A.x - Bm.x, A.y + Bm.y);
          const cpx w = wpost[k];
          const cpx O2 = make_float2(Dd.x * w.x + Dd.y * w.y, Dd.y * w.x - Dd.x * w.y);  // Dd * conj(w)
          st[k] = make_float2(E2.x - O2.y, -(E2.y + O2.x));
          if (k != 0 && k != m) st[m] = make_float2(E2.x + O2.y, -(O2.x - E2.y));
        }
      } else {
        for (int k = lane; k < M; k += 32) st[k] = make_float2(0.f, 0.f);
      }
    }
    __syncwarp();
    const long long fi = f0 + g;
    float* out = p.frames + (fi < p.n_frames ? fi : 0) * p.n_fft;
    const bool live = fi < p.n_frames;
    gl_fft_all<PL>(Zw, tw2, tw3, l, stage_g, ones, p.n_fft, [&](int k, cpx y) {
      // x[2k] = Re conj(y), x[2k+1] = Im conj(y); synthesis window carries 1 / n_fft
      const float2 s = *reinterpret_cast<const float2*>(synth + 2 * k);
      if (live) *reinterpret_cast<float2*>(out + 2 * k) = make_float2(y.x * s.x, -y.y * s.y);
    });
    __syncwarp();
  }
}

#ifndef SFE_PLAN_TU  // not templated on the plan: defined once, in sfe_lib.cu
// ---------------------------------------------------------------------------------------------
// frames -> waveform (overlap-add as a gather; torch.istft's window-envelope normalisation and centre trim)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gl_ola_kernel(const float* __restrict__ frames, const float* __restrict__ wsq, int B,
                                                     int F, int n_fft, int hop, long long L, float* __restrict__ wav) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * L) return;
  const long long b = i / L, n = i - b * L;
  const long long pos = n + n_fft / 2;
  long long t_hi = pos / hop;
  if (t_hi > F - 1) t_hi = F - 1;
  long long t_lo = (pos - n_fft + hop) / hop;  // ceil((pos - n_fft + 1) / hop)
  if (pos - n_fft + 1 <= 0) t_lo = 0;
  float sum = 0.f, env = 0.f;
  for (long long t = t_lo; t <= t_hi; ++t) {
    const int j = static_cast<int>(pos - t * hop);
    sum += __ldg(frames + (b * F + t) * n_fft + j);
    env += __ldg(wsq + j);
  }
  wav[i] = (env > 1e-11f) ? sum / env : sum;
}
#endif

// ---------------------------------------------------------------------------------------------
// waveform -> rebuilt spectrum -> phase update -> next spectrum
// ---------------------------------------------------------------------------------------------
template <class PL>
__global__ void __launch_bounds__(GL_THREADS) gl_stft_kernel(const __grid_constant__ GlParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int M = PL::M, G = PL::G, Q = PL::PADQ, C = PL::C;
  int g, l;
  cpx* Zw;
  float *stage_w, *stage_g;
  gl_setup<PL>(p, smem, g, l, Zw, stage_w, stage_g);
  const float* tab = reinterpret_cast<const float*>(smem);
  // Phase matters here (unlike the frontend, whose window sits at [0, win) of the FFT buffer): the window is
  // centred in n_fft like torch.stft does, so the frame is n_fft samples starting at t*hop - n_fft/2.
  const float* wtab = p.gl_tab + p.o_awin;
  const cpx* tw2 = reinterpret_cast<const cpx*>(tab + p.o_tw2);
  const cpx* tw3 = reinterpret_cast<const cpx*>(tab + p.o_tw3);
  const cpx* wpost = reinterpret_cast<const cpx*>(tab + p.o_wpost);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lp = (p.n_fft - p.win) / 2;

  for (long long f0 = (static_cast<long long>(blockIdx.x) * GL_WARPS + warp) * C; f0 < p.n_frames;
       f0 += static_cast<long long>(gridDim.x) * GL_WARPS * C) {
    // gather the window's samples of each frame (reflect padding of n_fft/2 at both ends of the row)
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const long long fi = f0 + c;
      float* st = stage_w + c * p.stage_floats;
      if (fi < p.n_frames) {
        const long long b = fi / p.F, t = fi - b * p.F;
        const float* row = p.wav + b * p.L;
        const long long base = t * p.hop - p.n_fft / 2;
        for (int j = lane; j < p.n_fft; j += 32) {
          long long o = base + j;
          if (o < 0) o = -o;
          if (o >= p.L) o = 2 * (p.L - 1) - o;
          st[j] = (j >= lp && j < lp + p.win) ? __ldg(row + o) : 0.f;
        }
      } else {
        for (int j = lane; j < p.n_fft; j += 32) st[j] = 0.f;
      }
    }
    __syncwarp();
    fft_frame<PL, true, false>(Zw, tw2, tw3, l, stage_g, wtab, p.n_fft, 0.f);

    // real-FFT unpack exactly as in the frontend kernel: X[k] = E + W^k O, X[M-k] = conj(E - W^k O)
    constexpr int RL = PL::RL;
    constexpr int NI = PL::LASTREG ? (RL / 2 + 1) : ((M / 2) / G + 1);
    cpx za[NI], zb[NI];
    if constexpr (PL::LASTREG) {
      cpx u[RL];
      stockham_last_reg<PL, RL, M / RL>(Zw, (PL::R3 > 1) ? tw3 : tw2, l, u);
      const int src = (G - l) & (G - 1);
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        za[i] = u[i];
        const float bx = __shfl_sync(0xffffffffu, u[RL - 1 - i].x, src, G);
        const float by = __shfl_sync(0xffffffffu, u[RL - 1 - i].y, src, G);
        zb[i] = (l == 0) ? u[(RL - i) % RL] : make_float2(bx, by);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int k = l + i * G;
        if (i < NI - 1 || k <= M / 2) {
          za[i] = Zw[padidx<Q>(k)];
          zb[i] = Zw[padidx<Q>(k == 0 ? 0 : M - k)];
        }
      }
    }
    const long long fi = f0 + g;
    if (fi < p.n_frames) {
      const long long q0 = fi * p.K;
      auto update = [&](int bin, cpx r) {
        cpx a = r;
        if (!p.first) {
          const cpx tp = p.tprev[q0 + bin];
          a = make_float2(r.x - p.momentum * tp.x, r.y - p.momentum * tp.y);
        }
        const float nrm = sqrtf(a.x * a.x + a.y * a.y) + 1e-16f;
        const float mg = __ldg(p.mag + q0 + bin);
        p.tprev[q0 + bin] = r;
        p.spec[q0 + bin] = make_float2(a.x / nrm * mg, a.y / nrm * mg);
      };
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
          update(k, X);
          if (2 * k != M) update(M - k, make_float2(Y.x, -Y.y));
        }
      }
    }
    __syncwarp();
  }
}

#ifndef SFE_PLAN_TU  // not templated on the plan: defined once, in sfe_lib.cu
// ---------------------------------------------------------------------------------------------
// small elementwise kernels
// ---------------------------------------------------------------------------------------------
// magnitude = spec^(1/power); first spectrum = magnitude x initial angles (complex, not unit length: torch.rand)
__global__ void __launch_bounds__(256) gl_init_kernel(const float* __restrict__ spec_in, const float2* __restrict__ angles,
                                                      long long n, int power2, float* __restrict__ mag,
                                                      float2* __restrict__ spec) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float m = spec_in[i];
  if (power2) m = sqrtf(m);
  mag[i] = m;
  const float2 a = angles ? angles[i] : make_float2(1.f, 0.f);
  spec[i] = make_float2(a.x * m, a.y * m);
}
#endif

#ifndef SFE_PLAN_TU  // not templated on the plan: defined once, in sfe_lib.cu
// inverse pre-emphasis over a sliding window (speech2linear.py:297-344): y[n] = sum_j a^j x[n-j], j < taps;
// then silence beyond wav_len[b] = (feat_len[b] - 1) * hop (:346-348).  pw[j] = a^j (nullptr: no filter).
__global__ void __launch_bounds__(256) gl_finish_kernel(const float* __restrict__ x, const float* __restrict__ pw, int taps,
                                                        const long long* __restrict__ feat_len, int hop, int B, long long L,
                                                        float* __restrict__ y, long long* __restrict__ wav_len) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < B && wav_len) wav_len[i] = (feat_len[i] - 1) * hop;
  if (i >= static_cast<long long>(B) * L) return;
  const long long b = i / L, n = i - b * L;
  const long long len = (feat_len[b] - 1) * hop;
  float v = 0.f;
  if (n < len) {
    if (pw) {
      // same summation order as the reference's Conv1d: oldest sample first
      const int jmax = (n < taps - 1) ? static_cast<int>(n) : taps - 1;
      for (int j = jmax; j >= 0; --j) v = fmaf(pw[j], x[i - j], v);
    } else {
      v = x[i];
    }
  }
  y[i] = v;
}
#endif

}  // namespace sfe
