/*
 * sfe.h — C ABI of the B200-native speech frontend ("sfe") library, libsfe.so.
 *
 * This is the drop-in boundary for the on-the-fly acoustic frontend of bagustris/speechain
 * (SURVEY.md §8b).  The reference is pure Python: its "FFI" for this path is the module API
 * of speechain/module/{frontend,norm,augment}; each entry point below names the reference
 * code whose arithmetic it replaces.  Plain C types and raw device pointers only — no torch
 * types cross this boundary, and the library owns no per-call memory: every buffer is
 * allocated by the caller (PyTorch's caching allocator in the shipped host layer).
 *
 * Conventions
 *   - All array pointers are DEVICE pointers unless the name ends in _host.
 *   - All calls are asynchronous on `stream` (a cudaStream_t passed as void*); nothing
 *     synchronises the host.  A plan is immutable after creation and may be used from any
 *     thread that has the plan's device current.
 *   - Return value 0 = success; non-zero = sfe_status.  No C++ exception crosses the ABI;
 *     sfe_last_error() returns a thread-local message for the last failing call.
 *   - Features are fp32, lengths int64, mask descriptors int32.
 */
#ifndef SFE_H_
#define SFE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFE_ABI_VERSION 1

typedef enum sfe_status {
  SFE_OK = 0,
  SFE_ERR_BAD_CONFIG = 1,       /* inconsistent sfe_config                                  */
  SFE_ERR_UNSUPPORTED_NFFT = 2, /* no FFT plan compiled for this n_fft                      */
  SFE_ERR_BAD_ARG = 3,          /* NULL / misaligned pointer, negative size, D too large …  */
  SFE_ERR_CUDA = 4              /* a CUDA runtime call or kernel launch failed              */
} sfe_status;

/* Constructor arguments of Speech2LinearSpec + LinearSpec2MelSpec, already resolved by the host
 * layer (seconds -> samples, n_fft default, fmax default): speech2linear.py:85-140,
 * linear2mel.py:117-161, speech2mel.py:107-151. */
typedef struct sfe_config {
  int32_t n_fft;          /* even; must have a compiled plan (see sfe_supported_nfft)          */
  int32_t win_length;     /* 1..n_fft                                                          */
  int32_t hop_length;     /* >= 1                                                              */
  int32_t center;         /* torch.stft(center=...): batch-level reflect padding by n_fft/2    */
  int32_t has_preemphasis;
  float   preemphasis;    /* y[n] = x[n] - a*x[n-1], y[0]=x[0], y[>=len]=0  (s2l.py:167-173)   */
  int32_t mag_spec;       /* sqrt of the power spectrum before the mel bank (s2l.py:236-237)   */
  /* linear-spectrogram output transform (Speech2LinearSpec standalone, s2l.py:240-245) */
  int32_t lin_logging;
  float   lin_clamp;
  float   lin_log_div;    /* ln(log_base), or 1.0f for natural log                             */
  /* mel stage (linear2mel.py:177-187); n_mels == 0 => plan without a mel stage */
  int32_t n_mels;
  int32_t mel_logging;
  float   mel_clamp;
  float   mel_log_div;
  const float* window_host;     /* win_length floats (already * n_fft^-0.5 if normalized=True)  */
  const float* mel_weight_host; /* (n_mels, n_fft/2+1) row-major, rows banded; NULL if n_mels==0 */
} sfe_config;

typedef struct sfe_plan sfe_plan;

/* SpecAugment masks drawn by the host with the reference's torch.randint sequence
 * (specaug.py:145-199): int32 arrays of shape (B, n_freq) / (B, n_time).
 * on_host = 0: the four pointers are DEVICE pointers.
 * on_host = 1: they are HOST pointers and the descriptors travel inside the kernel parameter
 *              block (no H2D copy, nothing to keep alive after the call returns); allowed while
 *              2*(n_freq+n_time)*B <= SFE_MASK_INLINE_MAX, else SFE_ERR_BAD_ARG. */
#define SFE_MASK_INLINE_MAX 2304
typedef struct sfe_masks {
  const int32_t* freq_pos;
  const int32_t* freq_len;
  int32_t n_freq;
  const int32_t* time_pos;
  const int32_t* time_len;
  int32_t n_time;
  float fill;              /* 0.0 when feat_norm=True (specaug.py:203)                         */
  const float* fill_dev;   /* if non-NULL, the fill value is read from device (feat.mean())    */
  int32_t on_host;
} sfe_masks;

/* Batch-shared SpecAugment time warp (specaug.py:96-124): rows [0,center) are resampled to
 * [0,pos) and [center,minlen) to [pos,minlen) by cubic convolution (A=-0.75, align_corners). */
typedef struct sfe_warp {
  int32_t enabled;  /* 0 off, 1 bicubic (the reference's default time_warp_mode), 2 bilinear; both align_corners=True */
  int32_t center;
  int32_t pos;
  int32_t minlen;
} sfe_warp;

int  sfe_abi_version(void);
const char* sfe_last_error(void);
/* writes up to cap supported n_fft values, returns how many exist */
int  sfe_supported_nfft(int32_t* out, int cap);

int  sfe_plan_create(const sfe_config* cfg, sfe_plan** out);
void sfe_plan_destroy(sfe_plan* plan);
/* frames torch.stft yields for a padded batch of T_max samples (s2l.py:198) */
int64_t sfe_plan_num_frames(const sfe_plan* plan, int64_t T_max);
/* dynamic shared memory per CTA and CTAs per SM of the fused kernel (for reports) */
int  sfe_plan_query(const sfe_plan* plan, int32_t* smem_bytes, int32_t* ctas_per_sm, int32_t* n_sm);
/* Host only (no CUDA call): which compiled straight-line Mel program of the tile kernel (sfe_melprog.h) has
 * exactly the band structure of the filterbank `mel_weight_host` (n_mels, n_bins) — linear2mel.py:135-155 builds
 * the matrix, :177 multiplies by it — and the weights in the layout that program reads.
 *   meta[10] = {bank id (0: none, the gather kernel runs), n_bins, n_mels, groups per stream, first bin of
 *               stream B, d_split, b_off, c_split, row floats, tail emits}
 *   table    = groups * 16 floats: per 4-bin group {A bin 0 (even filter, odd filter), A bin 1 | A bin 2, A bin 3 |
 *              B bin 0, B bin 1 | B bin 2, B bin 3}
 * Returns the number of table floats (written if table_cap is large enough), 0 if no program matches, < 0 on a
 * bad argument.  Exists so that program + weights can be checked against the filterbank without a GPU
 * (tests/test_mel_schedule.py executes the generated program text). */
int  sfe_mel_schedule(const float* mel_weight_host, int32_t n_mels, int32_t n_bins, int32_t* meta, float* table,
                      int32_t table_cap);

/* Fused waveform -> (log-)mel / linear spectrogram / energy (+ per-utterance sums).
 * Replaces Speech2LinearSpec.forward + LinearSpec2MelSpec.forward (= Speech2MelSpec.forward),
 * speech2linear.py:142-250, linear2mel.py:163-189.
 *   wav        (B, T_max) fp32, row pitch `wav_pitch` samples, zero beyond each length
 *   wav_len    (B,) int64 device
 *   mel        (B, F, n_mels) or NULL        F = sfe_plan_num_frames(T_max)
 *   lin        (B, F, n_fft/2+1) or NULL
 *   energy     (B, F) or NULL                (s2l.py:216-219,230-233)
 *   feat_len   (B,) int64 or NULL: len//hop+1 (s2l.py:201-212); wav_len is NOT mutated
 *   stats      (B, n_mels, 2) int64 or NULL: FIXED-POINT sums of `mel` over valid frames, [..][0] =
 *              sum x * 2^32, [..][1] = sum x^2 * 2^24 (exact, order independent -> bit-reproducible;
 *              log-mel plans only; zeroed by this call); pass to sfe_stats_finalize /
 *              sfe_normalize with stats_fixed = 1
 *   norm_mean/norm_std (n_mels,) or NULL: fused global normalisation (x-mean)/std in the
 *              epilogue (feat_norm.py:327-336, frozen/eval statistics)
 *   masks      NULL or SpecAugment masks applied in the epilogue (needs min over feat_len
 *              already folded into the drawn positions)
 *   skip_pad   if non-zero, mel rows t >= feat_len are left unwritten (a following
 *              sfe_normalize call synthesises them)
 *   wav_len_host  NULL, or a HOST copy of wav_len (the collated batch's lengths, which the data
 *              loader holds on the host anyway: model/abs.py:329-352 copies them to the device).
 *              With it the library cuts the batch into per-SM ranges on the host and the kernel
 *              starts without its length-load -> scan -> search prologue (~2 us per launch).  The
 *              values must equal wav_len's; results do not depend on whether it is passed. */
int sfe_frontend_forward(const sfe_plan* plan, const float* wav, int64_t wav_pitch,
                         const int64_t* wav_len, int32_t B, int64_t T_max, float* mel, float* lin,
                         float* energy, int64_t* feat_len, int64_t* stats, const float* norm_mean,
                         const float* norm_std, const sfe_masks* masks, int32_t skip_pad,
                         void* stream, const int64_t* wav_len_host);

/* The frontend -> normalise -> SpecAugment-mask chain of ASREncoder.forward (encoder/asr.py:103-119)
 * in one call and at most two launches (no time warp; see sfe_normalize for that):
 *   norm_mode 0  no normalisation: one launch, masks applied in the epilogue
 *   norm_mode 1  statistics known up front (mean/std vectors (n_mels,), either may be NULL): one
 *                launch, (x-mean)/std and masks in the epilogue         (feat_norm.py:327-336)
 *   norm_mode 2  utterance statistics: launch 1 writes the valid rows and the fixed-point sums into
 *                one half of stats_ws, launch 2 normalises in place, synthesises the padded rows,
 *                applies the masks and clears the other half for the next call (feat_norm.py:137-139).
 *                stats_ws is a (2, B, n_mels, 2) int64 scratch, ALL ZERO before its first use and kept
 *                by the caller between calls; ws_half (0/1) names the half this call accumulates into
 *                and must alternate from call to call on the same scratch; after a failed call zero
 *                the scratch again.
 *   norm_mode 3  batch / global statistics that are still being averaged (training, epoch <= max_epoch_num;
 *                feat_norm.py:232-266, 455-508): launch 1 as in mode 2 (same stats_ws / ws_half protocol; the two
 *                modes may share one scratch), launch 2 reduces the per-utterance statistics to the packed
 *                [sum_i mean_i | sum_i std_i | B] of sfe_stats_finalize, written to `mean_out` (2 n_mels + 1 floats),
 *                and clears the other half.
 *                `mel` holds the RAW valid rows afterwards; the caller all-reduces the packed sums when it runs
 *                data-parallel (one collective) and finishes with sfe_running_normalize.  masks / warp are ignored.
 * mean_out/std_out (B, n_mels), optional: the per-utterance statistics of norm_mode 2.
 * wav_len_host: as in sfe_frontend_forward (optional host copy of wav_len).
 * warp / warp_out: NULL, or the SpecAugment time warp (specaug.py:89-127: the warp comes BEFORE the masks) and a second
 *                (B, F, n_mels) buffer: the last pass then runs out of place — `mel` holds the unwarped (norm_mode 0/1:
 *                normalised, norm_mode 2: raw) features afterwards, `warp_out` the result.  The recipes' SpecAugment
 *                blocks leave time_warp at its default (on), so this is their training step in one call. */
int sfe_chain_forward(const sfe_plan* plan, const float* wav, int64_t wav_pitch,
                      const int64_t* wav_len, int32_t B, int64_t T_max, float* mel,
                      int64_t* feat_len, int64_t* stats_ws, int32_t ws_half, int32_t norm_mode,
                      const float* mean,
                      const float* std, float clamp, int32_t use_mean, int32_t use_std,
                      float* mean_out, float* std_out, const sfe_masks* masks, void* stream,
                      const int64_t* wav_len_host, const sfe_warp* warp, float* warp_out);

/* Second half of the averaging step (feat_norm.py:250,265,455-508, 327-336, then specaug.py:73-206): the running
 * average of register_mean_std_batch from the (all-reduced) packed sums — run = first ? batch : weight*batch +
 * (1-weight)*run, weight = 1/count in fp32 — written to run_mean / run_std (n_mels,; either may be NULL: mean_norm /
 * std_norm off), `count` stored into the `{prefix}_batch` buffer `counter` (int64 on the device, or NULL), and then
 * out = (feat - run_mean) / run_std over (B, F, D) — `out` may be `feat` itself (in place) — with the rows at or past
 * feat_len synthesised from pad_value when src_has_pad == 0, the SpecAugment masks, and — with `warp` — the time warp
 * in front of the masks (then `out` must be a second buffer, as in sfe_chain_forward).  Two launches (three with a
 * warp). */
int sfe_running_normalize(const float* packed, int32_t D, float weight, int32_t first, float* run_mean,
                          float* run_std, int64_t* counter, int64_t count, const float* feat,
                          const int64_t* feat_len, int32_t B, int64_t F, int32_t src_has_pad,
                          float pad_value, const sfe_masks* masks, void* stream, const sfe_warp* warp,
                          float* out);

/* LinearSpec2MelSpec.forward on an existing linear spectrogram (linear2mel.py:163-189).
 * lin (B*F, n_fft/2+1) -> mel (B*F, n_mels). */
int sfe_linear2mel_forward(const sfe_plan* plan, const float* lin, int64_t n_rows, float* mel,
                           void* stream);

/* Sum / sum of squares over valid frames of an arbitrary (B, F, D) tensor
 * (feat_norm.py:115-133 first half).  stats (B, D, 2) fp64 is zeroed by this call. */
int sfe_stats_accumulate(const float* feat, const int64_t* feat_len, int32_t B, int64_t F,
                         int32_t D, double* stats, void* stream);

/* Per-utterance mean and unbiased std (clamped) from the sums (feat_norm.py:115-133), and
 * the packed local sums for batch/global statistics (feat_norm.py:239-266):
 *   stats       (B, D, 2): fp64 from sfe_stats_accumulate (stats_fixed = 0) or the fixed-point
 *               sums of sfe_frontend_forward (stats_fixed = 1)
 *   mean, std   (B, D) fp32 (either may be NULL)
 *   packed      (2*D+1,) fp32 = [sum_i mean_i | sum_i std_i | B] or NULL */
int sfe_stats_finalize(const void* stats, int32_t stats_fixed, const int64_t* feat_len, int32_t B,
                       int32_t D, float clamp, float* mean, float* std, float* packed, void* stream);

/* Batch statistics from (all-reduced) packed sums and the running-average update of
 * register_mean_std_batch (feat_norm.py:250,265,483-508):
 *   batch_mean = packed[0:D]/packed[2D], batch_std = packed[D:2D]/packed[2D]
 *   if update: run = first ? batch : w*batch + (1-w)*run          (w = 1/global_batch)
 *   batch_mean_out / batch_std_out (D,) optional copies of the batch statistics. */
int sfe_running_update(const float* packed, int32_t D, float weight, int32_t first,
                       int32_t update, float* run_mean, float* run_std, float* batch_mean_out,
                       float* batch_std_out, void* stream);

/* (x - mean)/std, optional time warp, optional SpecAugment masking, one read + one write.
 * Replaces feat_norm.py:137-139 / 283-299 / 327-336 and specaug.py:111-124,202-204.
 *   src, dst   (B, F, D); may alias when warp is NULL/disabled
 *   stats      if non-NULL: per-utterance statistics are derived from these sums (encoding per
 *              stats_fixed, see sfe_stats_finalize)
 *              (utterance mode) and written to mean_out/std_out (B, D) when non-NULL
 *   mean, std  if stats is NULL: vectors of shape (D,) when per_utt == 0 (batch/global
 *              mode) or (B, D) when per_utt != 0; either may be NULL (mean_norm/std_norm off)
 *   use_mean/use_std: the mean_norm / std_norm switches when stats != NULL
 *   pad_value  value of unwritten rows t >= feat_len when src_has_pad == 0 */
int sfe_normalize(const float* src, float* dst, const int64_t* feat_len, int32_t B, int64_t F,
                  int32_t D, const void* stats, int32_t stats_fixed, float clamp, int32_t use_mean,
                  int32_t use_std,
                  const float* mean, const float* std, int32_t per_utt, float* mean_out,
                  float* std_out, int32_t src_has_pad, float pad_value, const sfe_warp* warp,
                  const sfe_masks* masks, void* stream);

/* Waveform pre-normalisation before the STFT (the `pre_stft_norm` key, speech2linear.py:165-187): pre-emphasis
 * y[n] = x[n] - a x[n-1] with samples at or beyond each length forced to zero, then over the WHOLE padded row
 * (padding included, like the reference) mode 1 "mean_std": (y - mean) / std (unbiased; B must be 1 — the reference
 * only broadcasts for one utterance), mode 2 "min_max": (y - min) / (max - min) * 2 - 1.  `out` (B, out_pitch)
 * receives full rows (the padding is no longer silence afterwards: the STFT that follows reads whole rows);
 * `scratch` holds 2 * B doubles. */
int sfe_prestft_norm(const float* wav, int64_t wav_pitch, const int64_t* wav_len, int32_t B, int64_t T,
                     int32_t has_preemphasis, float preemphasis, int32_t mode, double* scratch, float* out,
                     int64_t out_pitch, void* stream);

/* sfe_normalize writing bfloat16 (round to nearest even, what `.to(torch.bfloat16)` does): the layout and dtype the
 * first consumer takes under AMP — Conv2dPrenet.forward, prenet/conv2d.py:220-276, unsqueezes (B, F, D) and runs its
 * Conv2d under the model's autocast — so the cast pass over the feature tensor disappears and the last write of the
 * path is two bytes per element.  Never in place (src stays fp32). */
int sfe_normalize_bf16(const float* src, uint16_t* dst_bf16, const int64_t* feat_len, int32_t B, int64_t F, int32_t D,
                       const void* stats, int32_t stats_fixed, float clamp, int32_t use_mean, int32_t use_std,
                       const float* mean, const float* std, int32_t per_utt, int32_t src_has_pad, float pad_value,
                       const sfe_warp* warp, const sfe_masks* masks, void* stream);

/* Ragged host->device staging of a padded waveform batch (the step before the path:
 * Model.batch_to_cuda, model/abs.py:329-352, copies the zero padding too).  Copies the first
 * len_host[b] samples of every row from (pinned) host memory; neighbouring rows of similar length
 * (batches are sorted by length) travel as one pitched copy as wide as the longest of them while the
 * padding dragged along stays below 1/24 of the group's bytes.  Device samples beyond a row's length
 * may or may not be written — the frontend kernel never reads them. */
int sfe_upload_waveforms(const float* wav_host, int64_t host_pitch, const int64_t* len_host, int32_t B,
                         float* wav_dev, int64_t dev_pitch, void* stream);
/* Bytes sfe_upload_waveforms moves for these lengths (valid samples + the padding inside row groups). */
int64_t sfe_upload_bytes(const int64_t* len_host, int32_t B);
/* The same staging for 16-bit PCM, the sample type speech corpora are stored in: half the bytes cross PCIe and
 * the GPU does the conversion soundfile does on the CPU (utilbox/data_loading_util.py:71-78, sf.read(dtype=
 * "float32"): x / 32768, exact in fp32).  The valid prefix of every int16 row is copied into `stage_dev`
 * (sfe_upload_i16_stage_bytes(B, T) bytes: the lengths, then the rows) and one kernel writes the fp32 rows
 * of `wav_dev` (B, dev_pitch); device samples beyond a row's length are not written. */
int64_t sfe_upload_i16_stage_bytes(int32_t B, int64_t T);
int sfe_upload_waveforms_i16(const int16_t* wav_host, int64_t host_pitch, const int64_t* len_host, int32_t B,
                             void* stage_dev, int64_t stage_bytes, float* wav_dev, int64_t dev_pitch, void* stream);

/* Speech2LinearSpec with onesided=False (speech2linear.py:20-37,197-246: torch.stft returns all n_fft bins): `power`
 * (B, F, n_fft/2+1) is the un-logged POWER spectrogram of sfe_frontend_forward (a plan without mag_spec / lin_logging;
 * rows at or past feat_len are zero); `out` (B, F, n_fft) receives its mirror image |X[n_fft-k]| = |X[k]| after the
 * reference's point-wise tail (sqrt for mag_spec, clamp + log / log_div), `energy` (B, F) or NULL the frame energy
 * sqrt(max(sum over ALL n_fft bins, 1e-10)), zero at or past feat_len. */
int sfe_twosided(const float* power, const int64_t* feat_len, int32_t B, int64_t F, int32_t K, int32_t n_fft,
                 int32_t mag_spec, int32_t logging, float clamp, float log_div, float* out, float* energy,
                 void* stream);

/* DeltaFeature.forward (delta_feat.py:57-81): (B, F, D) -> (B, F, D*(order+1)). */
int sfe_delta(const float* feat, int32_t B, int64_t F, int32_t D, int32_t order, int32_t N,
              float* out, void* stream);

/* mean over all elements of a tensor into out[0] (fill value of specaug.py:203 when
 * feat_norm=False). */
int sfe_mean_all(const float* x, int64_t n, float* out, void* stream);

/* LinearSpec2MelSpec.recover (linear2mel.py:191-224): (log-)mel (B, F, D) -> linear spectrogram (B, F, K).
 * Undoes the logarithm (pow(base, x); log_mode 0 none, 1 natural, 2 base `log_base`), applies the
 * minimum-norm least-squares solution of  W x = mel  (what torch.linalg.lstsq(driver="gels") returns for the
 * full-rank n_mels x K filterbank W) as the product with `pinv_t` = pinv(W)^T laid out (D, K), zeroes rows
 * f >= feat_len[b] of every utterance shorter than `max_len` (the reference's rule, :215-218) and clamps to
 * `clamp` (1e-10, :221).  The caller computes pinv(W) once per filterbank on the host. */
int sfe_mel2linear(const float* mel, const int64_t* feat_len, int32_t B, int64_t F, int32_t D, int32_t K,
                   const float* pinv_t, int32_t log_mode, float log_base, int64_t max_len, float clamp,
                   float* lin, void* stream);

/* Speech2LinearSpec.recover (speech2linear.py:252-351): linear spectrogram -> waveform by Griffin-Lim
 * (torchaudio.transforms.GriffinLim as configured at :281-292: n_iter 32, momentum 0.99, random initial
 * phase), inverse pre-emphasis over a sliding window of `inv_preemph_taps` samples (:297-344, skipped when the
 * plan has no pre-emphasis or taps == 0) and silence beyond wav_len[b] = (feat_len[b] - 1) * hop (:346-348).
 *   spec        (B, F, K) power (or magnitude when the plan has mag_spec) spectrogram, F = padded batch length
 *   init_angles (B, F, K, 2) initial complex phase estimate — the caller draws it (torch.rand of the
 *               reference's (B, K, F) shape, transposed) so that a seeded run reproduces the reference; NULL: ones
 *   work        sfe_griffinlim_workspace(plan, B, F) bytes of device scratch
 *   wav         (B, hop * (F - 1)) output;  wav_len (B,) int64 output or NULL
 * Needs center=True and hop * (F - 1) > n_fft / 2 (torch's reflect padding has the same requirement). */
int64_t sfe_griffinlim_workspace(const sfe_plan* plan, int32_t B, int64_t F);
int sfe_griffinlim(sfe_plan* plan, const float* spec, const int64_t* feat_len, int32_t B, int64_t F,
                   int32_t n_iter, float momentum, const float* init_angles, int32_t inv_preemph_taps,
                   void* work, int64_t work_bytes, float* wav, int64_t* wav_len, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SFE_H_ */
