/*
 * avzoom.h -- C ABI of libavzoom.so: the B200 (sm_100a) mask-driven 2-mic MVDR hot path.
 *
 * The reference (Senpai-sama06/real-time-audio-visual-zooming, pure Python) has no FFI of its own;
 * each entry point below replaces one inline numpy/scipy block of it and says which (paths are
 * relative to the reference root).  A reference-side binding is a ctypes stub.
 *
 * Conventions
 *   - every pointer is a caller-owned DEVICE buffer unless the name ends in _host; nothing is
 *     allocated per call (per-(device, n_fft) constant tables are created once, see avz_init);
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and re-entrant;
 *   - return value: 0 on success, a negative AVZ_E* code otherwise; avz_last_error() gives the
 *     thread-local message.  Nothing throws or exits across this boundary;
 *   - layouts follow the reference: spectra (M, F, T) with T contiguous, masks (F, T), batch
 *     dimension prepended; complex numbers are interleaved float pairs (re, im);
 *   - F = n_fft/2 + 1, T = avz_num_frames(L, n_fft, hop) = ceil(L/hop) + 1 for hop | n_fft,
 *     iSTFT length = (T - 1) * hop  (scipy.signal.stft/istft, boundary='zeros', padded=True).
 *   - supported: n_fft in {256, 512, 1024}, hop with n_fft % hop == 0 and 2 <= n_fft/hop <= 8, L >= n_fft, B <= 65535
 *     for the fused passes.  L < n_fft returns AVZ_EINVAL: scipy.signal.stft would shrink nperseg to L there (or raise
 *     "noverlap must be less than nperseg" when L <= n_fft - hop), which changes the number of bins, and every reference
 *     call site then indexes out of range (oracle_debug.py:60 loops over N_FFT//2 + 1 bins) - the reference has no
 *     behaviour to reproduce for such inputs.
 *     n_fft 512 with hop 128 or 256 (every BASELINE shape) runs on the register-resident fast path, and so does
 *     n_fft 1024 with hop 512 (the learned pipelines' shape: features, mask covariance, apply).
 */
#ifndef AVZOOM_H_
#define AVZOOM_H_

#include <stdint.h>

#if defined(__GNUC__)
#define AVZ_API __attribute__((visibility("default")))
#else
#define AVZ_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define AVZ_VERSION 100

#define AVZ_OK 0
#define AVZ_EINVAL (-1)   /* bad argument (shape, n_fft, hop, null pointer) */
#define AVZ_ECUDA (-2)    /* CUDA runtime error; message has cudaGetErrorString */
#define AVZ_ENOGPU (-3)   /* no sm_100 device / kernels not loadable on this device */

/* post-filter applied to the beamformed spectrum */
#define AVZ_POST_NONE 0            /* masked_mvdr.py:124-127                                  */
#define AVZ_POST_ONE_MINUS_NOISE 1 /* oracle_debug.py:84-90     S * (1 - noise_mask)           */
#define AVZ_POST_FLOOR 2           /* full_audio.../inference.py:116   S * max(mask, floor)    */
#define AVZ_POST_MASK 3            /* Final_pipeline/src/inference.py:219   S * mask           */

/* what bins below the high-pass do (SURVEY 8-A2 "bins below the high-pass") */
#define AVZ_HP_NONE 0 /* tf_lite_version/inference.py batch_mvdr: no high-pass   */
#define AVZ_HP_ZERO 1 /* oracle_debug.py:69  `continue` -> bin stays 0           */
#define AVZ_HP_MIC0 2 /* Final_pipeline/src/inference.py:51-53  pass mic 0       */

/* feature layouts */
#define AVZ_FEAT_LOGMAG_IPD 0         /* (2,F,T): ln(|Y0|+1e-7), angle(Y0)-angle(Y1); full_audio.../inference.py:91-94 */
#define AVZ_FEAT_LOGMAG_IPD_WRAPPED 1 /* same, IPD wrapped to (-pi, pi]; notebook cell7:130-137                          */
#define AVZ_FEAT_PHYSICS_NHWC 2       /* (F,T,4): logmag, sin ipd, cos ipd, k/(F-1); Final_pipeline/src/inference.py:117-128 */

/* Every constant in which the reference's call sites differ (SURVEY 8-A2 variant table). */
typedef struct AvzMvdrCfg {
  float sigma;      /* diagonal loading:            oracle_debug.py:24,70 (1), masked_mvdr.py:16 (1e-7), .../inference.py:25 (1e-5) */
  float norm_eps;   /* sum(mask) + norm_eps:        oracle_debug.py:64 (1e-6)                                                        */
  float sqrt_eps;   /* sqrt(mask + sqrt_eps):       tf_lite_version/inference.py:111 (1e-10), 0 elsewhere                            */
  float w_eps;      /* d^H u + w_eps:               oracle_debug.py:77 (1e-10)                                                       */
  int32_t hp_bins;  /* number of leading bins with f < hp_hz (4 for 100 Hz at 512/16k)                                               */
  int32_t hp_mode;  /* AVZ_HP_*                                                                                                      */
  int32_t post_mode;/* AVZ_POST_*                                                                                                    */
  float post_floor; /* 0.05 at full_audio.../inference.py:116                                                                        */
} AvzMvdrCfg;

AVZ_API int avz_version(void);
AVZ_API const char* avz_last_error(void);
/* How this binary was built: target architecture, compiler, "release" (reads no environment variable) or "experiment"
 * (-DAVZ_EXPERIMENT: tuning knobs from the environment, tools/build_exp.sh).  Static string; bench.py records it. */
AVZ_API const char* avz_build_info(void);

/* Create the per-device constant tables (window, twiddles) for n_fft ahead of time (optional;
 * otherwise done on first use).  Do this before capturing calls into a CUDA graph. */
AVZ_API int avz_init(int n_fft);

/* Per-kernel timing of the fused fast path, for benchmarks: while enabled, CUDA events are recorded on the launching
 * stream around each kernel; avz_profile_get() waits for them and returns the device time (ms) of the LAST launch of
 * each slot: 0 k512_ibm, 1 k512_ibm_fixup, 2 k512_cov, 3 k_cov_finalize, 4 k_mvdr_weights, 5 k512_apply,
 * 6 k_peak_normalise (-1 if that kernel has not run since enabling).  ms_host: host array of >= 7 floats.
 * Not thread-safe; meant for single-stream measurement runs. */
AVZ_API int avz_profile_enable(int on);
AVZ_API int avz_profile_get(float* ms_host, int n);

/* scipy.signal.stft frame count (boundary='zeros', padded=True). */
AVZ_API int64_t avz_num_frames(int64_t L, int n_fft, int hop);

/* ---- STFT: replaces scipy.signal.stft(x, fs, nperseg=n_fft, noverlap=n_fft-hop) at
 * oracle_debug.py:42-44, masked_mvdr.py:76, full_audio.../inference.py:90.
 * x [B, C, L] f32  ->  Y [B, C, F, T] complex64. */
AVZ_API int avz_stft_f32(const float* x, int B, int C, int64_t L, int n_fft, int hop, float* Y, void* stream);

/* ---- iSTFT + overlap-add: replaces scipy.signal.istft at oracle_debug.py:93, masked_mvdr.py:127.
 * S [B, F, T] complex64 -> x [B, (T-1)*hop] f32.  If peak != NULL, peak[b] receives max|x[b,:]|
 * (peak must be zeroed by the caller). */
AVZ_API int avz_istft_f32(const float* S, int B, int T, int n_fft, int hop, float* x, float* peak, void* stream);

/* ---- x[b,:] /= (peak[b] + peak_eps): oracle_debug.py:94 (eps 0), masked_mvdr.py:128 (1e-6). */
AVZ_API int avz_peak_normalise_f32(float* x, int B, int64_t n, const float* peak, float peak_eps, void* stream);

/* ---- fused pass A of the oracle path: STFT(tgt), STFT(int), STFT(mix) -> IBM -> masked covariance,
 * without materialising any spectrum.  Replaces oracle_debug.py:42-64.
 * mix [B,2,L], tgt [B,L], itf [B,L] f32.
 * ibm_bits [B, T, ceil(F/32)] u32: bit (k & 31) of word (k >> 5) of frame t = (|S_int| > |S_tgt|) at bin k.
 * R [B, F, 4] f32 = (R00, R11, Re R01, Im R01) of  sum_t m y y^H / (sum_t m + norm_eps);  msum [B, F] = sum_t m.
 * ws: workspace of avz_ibm_cov_ws_bytes() bytes (partial sums of frame chunks). */
AVZ_API int64_t avz_ibm_cov_ws_bytes(int B, int64_t L, int n_fft, int hop);
AVZ_API int avz_ibm_cov_f32(const float* mix, const float* tgt, const float* itf, int B, int64_t L, int n_fft, int hop,
                    float norm_eps, uint32_t* ibm_bits, float* R, float* msum, void* ws, void* stream);

/* ---- the IBM of oracle_debug.py:42-53 with every bin decided in float64 (direct DFT): slow reference form of the
 * mask avz_ibm_cov_f32 produces, for checking it at sizes a CPU cannot reach.  n_fft = 512 only.
 * ibm_bits [B, T, 9] u32 as above; ws16: 16 bytes of device scratch. */
AVZ_API int avz_ibm_exact_f32(const float* tgt, const float* itf, int B, int64_t L, int n_fft, int hop, uint32_t* ibm_bits,
                      void* ws16, void* stream);

/* ---- masked covariance from a waveform and a given (target-probability) mask, learned-mask path:
 * replaces full_audio.../inference.py:90,102-108 / tf_lite_version/inference.py:97-127.
 * mask [B, F, T] f32 is the TARGET probability; the noise weight is (1 - mask). */
AVZ_API int avz_wave_mask_cov_f32(const float* mix, const float* mask, int B, int64_t L, int n_fft, int hop,
                          float sqrt_eps, float norm_eps, float* R, float* msum, void* ws, void* stream);

/* ---- masked covariance from a given spectrum: generic form of oracle_debug.py:56-64.
 * Y [B,2,F,T] complex64, noise_w [B,F,T] f32 (the weight itself, not 1 - mask). */
AVZ_API int avz_spec_mask_cov_f32(const float* Y, const float* noise_w, int B, int F, int T, float sqrt_eps, float norm_eps,
                          float* R, float* msum, void* stream);

/* ---- closed-form 2x2 MVDR weights: replaces oracle_debug.py:68-79 (np.linalg.solve per bin).
 * R [B,F,4], dvec [F,2] complex64 (steering vectors), -> w [B,F,2] complex64.
 * bins k < cfg->hp_bins get w = 0 (AVZ_HP_ZERO) or [1,0] (AVZ_HP_MIC0); det == 0 -> [1,0]. */
AVZ_API int avz_mvdr_weights_f32(const float* R, const float* dvec, int B, int F, const AvzMvdrCfg* cfg, float* w, void* stream);

/* ---- hybrid hard-null weights: replaces Final_pipeline/src/inference.py:56-94 (eigh + cond + solve per bin) with a
 * closed form in float64.  R [B,F,4] is the interference covariance sum m y y^H / (sum m + 1e-6), m = 1 - mask
 * (avz_spec_mask_cov_f32 / avz_wave_mask_cov_f32 with sqrt_eps = 0); dvec [F,2] the un-normalised steering vectors;
 * bins k < bypass_bins (f < 200 Hz there) get w = [1, 0] (mic 0 passes).  Condition number > 10 -> w = v_tgt / 2.
 * A bin whose principal eigenvector has a zero first component (an exactly zero or diagonal covariance with R00 <= R11)
 * gets delay-and-sum, or - zero_cov_nan != 0 - NaN, which is what the reference's v_int / (v_int[0] / (|v_int[0]| + 1e-10))
 * evaluates to there (Final_pipeline/src/inference.py:66-68).  w [B,F,2] complex64,
 * usable by avz_beamform_f32 and by avz_mvdr_apply_f32 (any weights are "beamformer weights" to pass B). */
AVZ_API int avz_hybrid_null_weights_f32(const float* R, const float* dvec, int B, int F, int bypass_bins, int zero_cov_nan,
                                float* w, void* stream);

/* ---- beamform a given spectrum: oracle_debug.py:80.  w [B,F,2], Y [B,2,F,T] -> S [B,F,T] complex64. */
AVZ_API int avz_beamform_f32(const float* w, const float* Y, int B, int F, int T, float* S, void* stream);

/* ---- fused pass B: STFT(mix) -> w^H y -> post-filter -> iSTFT/overlap-add (+ per-utterance peak).
 * Replaces oracle_debug.py:80-93 / full_audio.../inference.py:114-117.
 * Exactly one of ibm_bits / mask may be non-NULL (both NULL only with AVZ_POST_NONE).
 * out [B, (T-1)*hop] f32; peak [B] (zeroed by caller) or NULL. */
AVZ_API int avz_mvdr_apply_f32(const float* mix, const float* w, const uint32_t* ibm_bits, const float* mask, int B, int64_t L,
                       int n_fft, int hop, const AvzMvdrCfg* cfg, float* out, float* peak, void* stream);

/* ---- "kept spectrum" variants of the two fused passes (n_fft 512, hop 128 / 256).
 * The path is fp32-issue-bound, not HBM-bound, so pass A can keep the packed two-mic spectrum of
 * every frame (4096 B per frame in `spec`, avz_spec_ws_bytes() bytes in total; 0 = not supported for this shape) and
 * pass B then skips its forward transform: ~25 % fewer instructions for 64 B/sample of otherwise idle HBM bandwidth.
 * Results are bit-identical to the recomputing variants (same transform, same arithmetic order).
 * The buffer also has room for per-utterance completion counters and for a transposed copy (B, T, 264) of the mask:
 * the mask variants re-lay the caller's (B, F, T) mask there first, so that a frame's 257 weights are contiguous.
 * avz_mvdr_apply_kept_f32 with a float-mask post-filter (AVZ_POST_FLOOR / AVZ_POST_MASK) and mask == NULL uses that
 * transposed copy as it stands - valid right after avz_wave_mask_cov_keep_f32 on the same `spec`, and it saves the second
 * transposition (n_fft 512; with a mask pointer the mask is re-laid again, which is always safe).
 * n_fft 1024 / hop 512: the learned-mask pair (avz_wave_mask_cov_keep_f32 -> avz_mvdr_apply_kept_f32, float masks)
 * keeps both one-sided spectra, 8320 B per frame = 16 B per sample; the IBM variant and the fused normalisation are
 * n_fft 512 only (AVZ_EINVAL otherwise). */
AVZ_API int64_t avz_spec_ws_bytes(int B, int64_t L, int n_fft, int hop);
AVZ_API int avz_ibm_cov_keep_f32(const float* mix, const float* tgt, const float* itf, int B, int64_t L, int n_fft, int hop,
                         float norm_eps, uint32_t* ibm_bits, float* R, float* msum, void* ws, void* spec, void* stream);
AVZ_API int avz_wave_mask_cov_keep_f32(const float* mix, const float* mask, int B, int64_t L, int n_fft, int hop,
                               float sqrt_eps, float norm_eps, float* R, float* msum, void* ws, void* spec, void* stream);
AVZ_API int avz_mvdr_apply_kept_f32(const void* spec, const float* w, const uint32_t* ibm_bits, const float* mask, int B,
                            int64_t L, int n_fft, int hop, const AvzMvdrCfg* cfg, float* out, float* peak, void* stream);

/* Sparse kept spectrum for the oracle path (post-filter AVZ_POST_ONE_MINUS_NOISE, oracle_debug.py:84-90): pass B
 * multiplies every TF bin whose noise bit is set by zero, so pass A keeps only the bins with a clear bit, compacted to
 * the front of each frame's block in an order both passes derive from the same ibm_bits.  On speech-like mixtures two
 * thirds of the bins are noise-dominated: two thirds of the kept-spectrum traffic - what bounds pass A and pass B -
 * disappears; the waveform is bit-identical to the dense pair.  `spec` is sized by avz_spec_ws_bytes() as before.  A
 * sparse spectrum read by avz_mvdr_apply_kept_f32 (or a dense one read here) is detected on the device (a header word
 * in `spec`) and poisons the output with NaN instead of producing garbage. */
AVZ_API int avz_ibm_cov_keep_sparse_f32(const float* mix, const float* tgt, const float* itf, int B, int64_t L, int n_fft,
                                int hop, float norm_eps, uint32_t* ibm_bits, float* R, float* msum, void* ws, void* spec,
                                void* stream);
AVZ_API int avz_mvdr_apply_kept_sparse_f32(const void* spec, const float* w, const uint32_t* ibm_bits, int B, int64_t L,
                                   int n_fft, int hop, const AvzMvdrCfg* cfg, float* out, float* peak, void* stream);

/* Pass A for a pass B that applies AVZ_POST_ONE_MINUS_NOISE with these same ibm_bits (the oracle path): dense layout
 * as avz_ibm_cov_keep_f32, but a 32-byte sector of the kept spectrum (the slots of two adjacent bins) whose two bins are
 * both noise-dominated is not written - pass B multiplies whatever sits there by the gain 0.  `spec` must therefore hold
 * finite values before its first use (zero it ONCE after allocating it; later calls find the previous call's finite
 * spectra there) - a NaN / Inf left in it would surface as NaN in the output, loudly, never as a wrong number.  Read back
 * with avz_mvdr_apply_kept_f32 (post-filter AVZ_POST_ONE_MINUS_NOISE only: any other post-filter on such a buffer is
 * detected on the device and yields NaN).  Same bits out; about 45 % of pass A's store stream - which is what bounds it -
 * is never sent to HBM. */
AVZ_API int avz_ibm_cov_keep_postmask_f32(const float* mix, const float* tgt, const float* itf, int B, int64_t L, int n_fft,
                                  int hop, float norm_eps, uint32_t* ibm_bits, float* R, float* msum, void* ws, void* spec,
                                  void* stream);

/* Same as avz_mvdr_apply_kept_f32 plus the peak normalisation of oracle_debug.py:94 fused in: the thread blocks of an
 * utterance run as one cluster, agree on max|x| through distributed shared memory and divide their own output range
 * by (peak + peak_eps) while it is still in L2 (more than 8 blocks per utterance: a separate pass follows instead).
 * Bit-identical to apply followed by avz_peak_normalise_f32; slower than that pair at the BASELINE shapes
 * (profiles/README.md), offered for callers that want one launch.  peak [B] zeroed by the caller. */
AVZ_API int avz_mvdr_apply_kept_norm_f32(const void* spec, const float* w, const uint32_t* ibm_bits, const float* mask, int B,
                                 int64_t L, int n_fft, int hop, const AvzMvdrCfg* cfg, float peak_eps, float* out,
                                 float* peak, void* stream);

/* ---- pass A with its per-utterance tail folded in (n_fft 512, hop 128 / 256): as avz_ibm_cov_keep_f32 (spec may be
 * NULL: no kept spectrum), and the block that finishes an utterance's last frame chunk also sums the chunk partials
 * (oracle_debug.py:60-64) and solves the 257 2x2 systems (oracle_debug.py:68-79) - R, msum and w come out of the one
 * launch, bit-identical to avz_ibm_cov_keep_f32 + avz_mvdr_weights_f32.  Two launches fewer per step (measured: no gain,
 * profiles/README.md).  dvec [257,2] complex64; sparse: 0 dense kept spectrum, 1 as avz_ibm_cov_keep_sparse_f32,
 * 2 as avz_ibm_cov_keep_postmask_f32. */
AVZ_API int avz_ibm_cov_weights_keep_f32(const float* mix, const float* tgt, const float* itf, int B, int64_t L, int n_fft,
                                 int hop, const AvzMvdrCfg* cfg, const float* dvec, uint32_t* ibm_bits, float* R,
                                 float* msum, float* w, void* ws, void* spec, int sparse, void* stream);

/* ---- the whole oracle path of oracle_debug.py:42-94 as ONE call (n_fft 512, hop 128 / 256): k512_ibm + k512_ibm_fixup,
 * then a single persistent kernel whose tasks are pass A (masked covariance, spectrum kept), the per-utterance
 * finalize + 2x2 solve, pass B (beamform, post-filter, iSTFT / overlap-add) and the peak normalisation.  Pass B of an
 * utterance runs microseconds after its pass A and the kept spectrum lives in a ring of utterance slots that stays in
 * L2, so it never travels to HBM (the separate kernels move 4.2 GB of it per 1024 x 4 s and are bound by that stream).
 * Same arithmetic, same device functions: ibm_bits, R, msum, w, out and peak equal what avz_ibm_cov_keep_f32 ->
 * avz_mvdr_weights_f32 -> avz_mvdr_apply_kept_f32 -> avz_peak_normalise_f32 produce (bit for bit when both cut an
 * utterance into the same frame chunks, otherwise to float32 summation order in R).
 * dvec [257,2] complex64; cfg->post_mode AVZ_POST_ONE_MINUS_NOISE or AVZ_POST_NONE; peak_eps < 0: no normalisation;
 * peak [B] is zeroed by the call; ws: avz_oracle_fused_ws_bytes() bytes. */
AVZ_API int64_t avz_oracle_fused_ws_bytes(int B, int64_t L, int n_fft, int hop);
AVZ_API int avz_oracle_fused_f32(const float* mix, const float* tgt, const float* itf, int B, int64_t L, int n_fft, int hop,
                         const AvzMvdrCfg* cfg, float peak_eps, const float* dvec, uint32_t* ibm_bits, float* R, float* msum,
                         float* w, float* out, float* peak, void* ws, void* stream);

/* ---- streaming mode (BASELINE config 4; NOT in the reference - defined by this project, parity unpinned):
 *   R_t = lambda R_{t-1} + (1-lambda) m_t y_t y_t^H,  n_t = lambda n_{t-1} + (1-lambda) m_t,
 *   w_t = mvdr(R_t/(n_t+norm_eps) + sigma I),  S_t = w_t^H y_t,  same 512/128 framing and overlap-add as the batch path.
 * One call consumes one hop (128 new samples per mic) of n_streams independent streams and emits 128 output samples
 * per stream.  Call number h (0, 1, ...) carries t = h - 1 (the scipy frame completed by this hop) and returns the
 * output samples [(h-3)*128, (h-2)*128): ignore the first three returns; after the last input hop feed three zero hops
 * with t_end = number of frames (= input hops + 1) to flush.  t_end = INT_MAX while the stream is open.
 * state: avz_stream_state_bytes(n_streams) bytes, zeroed by the caller before the first hop.
 * hop_in [S,2,128] f32, noise_w [S,257] f32 (noise weight m_t of this frame) or NULL (= 1), dvec [257,2] complex64. */
AVZ_API int64_t avz_stream_state_bytes(int n_streams);
AVZ_API int avz_stream_step_f32(float* state, const float* hop_in, const float* noise_w, const float* dvec, int n_streams,
                        int t, int t_end, float lambda, const AvzMvdrCfg* cfg, float* hop_out, void* stream);

/* ---- chunk drivers: main_deploy / process_audio_file / enhance_audio cut a recording into
 * WIN_SIZE = 32000-sample windows at stride 16000, zero-pad the tail, enhance every window and overlap-add the outputs
 * with a per-sample count (full_audio.../inference.py:127-156, resnet_model_mvdr/inference.py:226-265,
 * tf_lite_version/inference.py:267-375, Final_pipeline/src/inference.py:172-235).  Here R planar recordings
 * rec [R][2][rec_len] are read IN PLACE: window i of recording r is utterance b = r * n_windows + i of the fused
 * kernels (samples past rec_len read as zero), so no gathered copy of the windows exists, and one gather kernel does
 * the count-averaged overlap-add.  n_fft 1024 / hop 512 (the STFT shape of all those drivers); R * n_windows <= 65535. */
typedef struct AvzChunkView {
  int64_t rec_len;    /* samples per channel of a recording                                   */
  int32_t n_windows;  /* windows per recording: ceil(rec_len / stride)                          */
  int32_t stride;     /* samples between window starts (WIN_SIZE / 2 = 16000 in the reference)  */
} AvzChunkView;
/* features of every window: X [R*n_windows, ...] in the layout of `mode` (process_chunk, full_audio.../inference.py:90-94) */
AVZ_API int avz_chunk_features_f32(const float* rec, int R, const AvzChunkView* cv, int64_t win, int n_fft, int hop, int mode,
                           float* X, void* stream);
/* learned-mask pass A per window (full_audio.../inference.py:102-108): mask [R*n_windows, F, T] target probabilities ->
 * Rcov [R*n_windows, F, 4], msum; ws as avz_ibm_cov_ws_bytes(R*n_windows, win, ...); spec (may be NULL) keeps the spectra
 * (avz_spec_ws_bytes(R*n_windows, win, ...)) for avz_chunk_mvdr_apply_f32. */
AVZ_API int avz_chunk_mask_cov_f32(const float* rec, const float* mask, int R, const AvzChunkView* cv, int64_t win, int n_fft,
                           int hop, float sqrt_eps, float norm_eps, float* Rcov, float* msum, void* ws, void* spec,
                           void* stream);
/* pass B per window (full_audio.../inference.py:114-117): from `spec` if not NULL, else from the waveform (rec).
 * out [R*n_windows, (T-1)*hop]; peak [R*n_windows] zeroed by the caller, or NULL. */
AVZ_API int avz_chunk_mvdr_apply_f32(const float* rec, const void* spec, const float* w, const float* mask, int R,
                             const AvzChunkView* cv, int64_t win, int n_fft, int hop, const AvzMvdrCfg* cfg, float* out,
                             float* peak, void* stream);
/* count-averaged overlap-add of the window outputs: final[r][s] = sum_i outs[r*n_windows+i][s - i*stride] / max(count, 1)
 * over the windows whose first `use_len` output samples cover s (use_len = min(olen, WIN_SIZE) at
 * full_audio.../inference.py:151-153, = olen at Final_pipeline/src/inference.py:225-227 where the buffer ends at rec_len),
 * summed in window order like the reference's loop.  final [R][rec_len]; peak [R] (zeroed by the caller) receives
 * max|final[r]| for the closing `final / (max|final| + 1e-9)` (avz_peak_normalise_f32), or NULL. */
AVZ_API int avz_chunk_ola_f32(const float* outs, int R, int n_windows, int64_t olen, int64_t rec_len, int stride,
                      int64_t use_len, float* final_, float* peak, void* stream);
/* WAV frames -> planar float32: pcm [R][n][C] interleaved int16 (soundfile.read's (frames, channels)) -> out [R][C][n]
 * = pcm / 32768 (oracle_debug.py:35-39). */
AVZ_API int avz_pcm16_frames_to_planar_f32(const int16_t* pcm, int R, int64_t n, int C, float* out, void* stream);

/* ---- float64 forms of the unfused operators.  The reference computes in float64 / complex128 whenever its inputs are
 * float64 (scipy.signal.stft follows the input dtype; R, w, S are `dtype=complex`, oracle_debug.py:57,67).  Two call
 * sites are ill-conditioned enough for a float32 STFT to show in the output - masked_mvdr.main (sigma = 1e-7,
 * masked_mvdr.py:76-128) and hybrid_hard_null_bf (eigenvector + cond threshold + solve, Final_pipeline/src/inference.py:
 * 28-98) - and run on these.  Same layouts as the _f32 entry points with double / complex128 elements; R is
 * [B,F,4] doubles, dvec [F,2] complex128, w [B,F,2] complex128.  Not on the throughput path. */
AVZ_API int avz_stft_f64(const double* x, int B, int C, int64_t L, int n_fft, int hop, double* Y, void* stream);
AVZ_API int64_t avz_istft_f64_ws_bytes(int B, int T, int n_fft);
AVZ_API int avz_istft_f64(const double* S, int B, int T, int n_fft, int hop, double* x, void* ws, void* stream);
/* x[b,:] /= max|x[b,:]| + peak_eps (masked_mvdr.py:128); peak [B] receives the maxima if not NULL. */
AVZ_API int avz_peak_normalise_f64(double* x, int B, int64_t n, double peak_eps, double* peak, void* stream);
AVZ_API int avz_geometric_mask_f64(const double* Y, int B, int F, int T, double* mask, void* stream);
AVZ_API int avz_spec_mask_cov_f64(const double* Y, const double* noise_w, int B, int F, int T, double sqrt_eps,
                          double norm_eps, double* R, double* msum, void* stream);
/* Learned-mask covariance in float64 straight from the float32 waveform and target-probability mask (weight 1 - mask):
 * the ill-conditioned consumers (hybrid hard-null) get full_audio.../inference.py:90,102-108 at the reference's precision.
 * ws: avz_wave_mask_cov_f64_ws_bytes() bytes (the complex128 spectrum). */
AVZ_API int64_t avz_wave_mask_cov_f64_ws_bytes(int B, int64_t L, int n_fft, int hop);
AVZ_API int avz_wave_mask_cov_f64(const float* mix, const float* mask, int B, int64_t L, int n_fft, int hop, double sqrt_eps,
                          double norm_eps, double* R, double* msum, void* ws, void* stream);
/* the same for the windows of planar recordings read in place (AvzChunkView above); ws for R * n_windows utterances */
AVZ_API int avz_chunk_mask_cov_f64(const float* rec, const float* mask, int R, const AvzChunkView* cv, int64_t win, int n_fft,
                           int hop, double sqrt_eps, double norm_eps, double* Rcov, double* msum, void* ws, void* stream);
AVZ_API int avz_mvdr_weights_f64(const double* R, const double* dvec, int B, int F, double sigma, double w_eps, int hp_bins,
                         int hp_mode, double* w, void* stream);
AVZ_API int avz_hybrid_null_weights_f64(const double* R, const double* dvec, int B, int F, int bypass_bins, int zero_cov_nan,
                                double* w, void* stream);
/* float64 arithmetic, weights rounded once to complex64 for the float32 pass B (avz_mvdr_apply_f32 / avz_chunk_mvdr_apply_f32) */
AVZ_API int avz_hybrid_null_weights_f64_w32(const double* R, const double* dvec, int B, int F, int bypass_bins, int zero_cov_nan,
                                    float* w, void* stream);
AVZ_API int avz_beamform_f64(const double* w, const double* Y, int B, int F, int T, double* S, void* stream);

/* ---- IBM from given spectra: oracle_debug.py:49-53 (noise polarity) / model_training.py:90 (target polarity).
 * a, b [n] complex64 -> out [n] f32 = (|a| > |b|) ? 1 : 0, compared exactly (float64 squares). */
AVZ_API int avz_mag_greater_f32(const float* a, const float* b, int64_t n, float* out, void* stream);

/* ---- soft (ideal-ratio) post-filter mask: oracle_reverb.py:143-147.
 * s_tgt, s_int [n] complex64 -> out [n] f32 = sqrt(|s_tgt|^2 / (|s_tgt|^2 + |s_int|^2 + 1e-10)). */
AVZ_API int avz_irm_f32(const float* s_tgt, const float* s_int, int64_t n, float* out, void* stream);

/* ---- geometric phase mask: masked_mvdr.py:37-46.  Y [B,2,F,T] -> mask [B,F,T] in {0.01, 1}. */
AVZ_API int avz_geometric_mask_f32(const float* Y, int B, int F, int T, float* mask, void* stream);

/* ---- unpack ibm_bits [B,T,ceil(F/32)] to a float mask [B,F,T] of {0,1}. */
AVZ_API int avz_ibm_unpack_f32(const uint32_t* ibm_bits, int B, int F, int T, float* mask, void* stream);

/* ---- features from a spectrum: full_audio.../inference.py:91-94.  Y [B,2,F,T] -> X (mode layout). */
AVZ_API int avz_features_f32(const float* Y, int B, int F, int T, int mode, float* X, void* stream);

/* ---- features straight from the waveform (STFT fused, no spectrum written): mix [B,2,L] -> X. */
AVZ_API int avz_wave_features_f32(const float* mix, int B, int64_t L, int n_fft, int hop, int mode, float* X, void* stream);

/* ---- projection scores: Final_pipeline/src/metrics.py:102-123 and scripts/run_metrics.py:6-36.
 * est [B, n_est], tgt [B, n_ref], itf [B, n_ref]; uses the first min(n_est, n_ref) samples.
 * scores [B,4] f32 = (OSINR, OSIR, SDR_unit_output, SIR_unit_output) in dB. */
AVZ_API int avz_sir_f32(const float* est, const float* tgt, const float* itf, int B, int64_t n_est, int64_t n_ref,
                float* scores, void* stream);

/* ---- far-field 2-mic mixer: rt_av_zoom/core/tf_lite_version/world_building.py:46-52 (apply_frac_delay: whole-signal
 * rFFT, phase ramp exp(-2 pi i f tau), irfft) and :61-93 (mix_and_save: per-source delays to both mics, references =
 * the mic-1 images, everything divided by max|mix| + peak_eps).
 * src [B,S,L] (source 0 is the target, the others interferers); delays_host [S][2] seconds (mic 1, mic 2) is a HOST
 * pointer read during the call.  mix [B,2,L], tgt [B,L], itf [B,L].  peak_eps < 0 skips the division (plain delays).
 * Any L (as pocketfft in the reference): L = N1*N2 with N1 the largest power of two <= 512 dividing L and N2 <= 1024
 * (64000 = 512*125, 80000 = 128*625) runs the two-factor transform directly; every other length up to 262144 samples
 * (primes, odd lengths, ...) runs the same passes as a chirp-z (Bluestein) convolution of length M >= 2L-1 (one plan
 * per (device, L), built on the first call, which synchronises `stream` once).  Longer non-factorable lengths return
 * AVZ_EINVAL.  1 <= S <= 8.  ws: avz_farfield_mix_ws_bytes(B,S,L) bytes of device scratch (0 = unsupported shape).
 * S <= 4 and L = 4 * M with M <= 25600 a product of 2, 3, 5, 7 (64000, 80000, 16000, 4096 ...): ONE kernel, one 8-CTA
 * thread-block cluster per utterance - the two complex planes stay in the cluster's shared memory from the load of the
 * sources to the store of the normalised signals (in-place mixed-radix FFT whose first radix-4 stage runs across the
 * CTAs through distributed shared memory); `ws` is not touched.  avz_farfield_mix_passes_f32 is the same operator
 * always on the multi-pass path (what S > 4 and all other lengths use): same arguments, results equal to float32
 * rounding of two different transform factorisations. */
AVZ_API int64_t avz_farfield_mix_ws_bytes(int B, int S, int64_t L);
AVZ_API int avz_farfield_mix_f32(const float* src, const double* delays_host, int B, int S, int64_t L, double fs,
                         float peak_eps, float* mix, float* tgt, float* itf, void* ws, void* stream);
AVZ_API int avz_farfield_mix_passes_f32(const float* src, const double* delays_host, int B, int S, int64_t L, double fs,
                                float peak_eps, float* mix, float* tgt, float* itf, void* ws, void* stream);

/* ---- PCM16 wire format (soundfile semantics, oracle_debug.py:35-39,96): read = int16 / 32768 (exact in float32);
 * write = libsndfile's default PCM_16 conversion, round-half-even(x * 32767) (0x7FFF), clipped to [-32768, 32767].
 * n elements, any layout; the device-side halves of a PCM16 host<->device transfer (half the PCIe bytes of float32). */
AVZ_API int avz_pcm16_to_f32(const int16_t* pcm, int64_t n, float* out, void* stream);
AVZ_API int avz_f32_to_pcm16(const float* x, int64_t n, int16_t* pcm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVZOOM_H_ */
