/*
 * ftb200.h -- C ABI of the B200-native ForwardTacotron / FastPitch inference
 * path and the STFT->log-mel feature extractor.
 *
 * The reference (tarepan/ForwardTacotron) is pure Python: it has no FFI of its
 * own, so the drop-in boundary is its Python surface
 *     ForwardTacotron.from_config / from_checkpoint / generate      models/forward_tacotron.py:244-268,338-350
 *     FastPitch.from_config / from_checkpoint / generate            models/fast_pitch.py:286-303,342-354
 *     DSP.from_config / DSP.wav_to_mel                              utils/dsp.py:59-61,71-87
 * which forwardtacotron_b200/{models,utils}/ mirrors.  This header is the
 * boundary UNDER that surface: what the mirrored classes bind with ctypes and
 * what a maintainer of the reference would bind.
 *
 * Conventions
 *   - every entry point returns 0 (FTB_OK) or a negative ftb_status; the text of
 *     the last failure on the calling thread is ftb_last_error();
 *   - all data pointers are DEVICE pointers owned by the caller unless the name
 *     says host_; nothing is allocated behind the caller's back except the
 *     packed weight copy held by a model handle;
 *   - every launch goes to the cudaStream_t passed as `stream` (void*);
 *   - activations are channel-last: (B, S, C) row-major; mel outputs are the
 *     reference's (B, n_mels, L);
 *   - a handle is bound to one device and is not thread-safe.
 */
#ifndef FTB200_H_
#define FTB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FTB_ABI_VERSION 1

typedef enum ftb_status {
  FTB_OK = 0,
  FTB_ERR_INVALID = -1,     /* bad argument / unsupported shape                 */
  FTB_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed              */
  FTB_ERR_MISSING = -3,     /* a state_dict entry is missing or has wrong shape */
  FTB_ERR_WORKSPACE = -4,   /* caller workspace too small                       */
  FTB_ERR_UNSUPPORTED = -5  /* config outside what the kernels are built for    */
} ftb_status;

typedef enum ftb_dtype { FTB_F32 = 0, FTB_I64 = 1, FTB_BF16 = 2, FTB_I32 = 3 } ftb_dtype;

/* One state_dict entry: the name is the reference's parameter / buffer name
 * (e.g. "prenet.conv1d_bank.3.conv.weight"), data is a device pointer in the
 * reference's own layout and dtype. */
typedef struct ftb_tensor {
  const char* name;
  const void* data;
  int32_t dtype; /* ftb_dtype */
  int32_t ndim;
  int64_t shape[4];
} ftb_tensor;

const char* ftb_last_error(void);
int ftb_abi_version(void);
/* sizeof() of the ABI structs as compiled (0 ftb_tensor, 1 ftb_conv_desc, 2 ftb_mel_config,
 * 3 ftb_ft_config, 4 ftb_fp_config); bindings use it to verify their struct layout. */
int ftb_struct_size(int which);
/* Kernel launches issued by this library since it was loaded (bench bookkeeping). */
long long ftb_launch_count(void);
/* Optional per-kernel-family timing: CUDA events recorded on the launching stream around every launch
 * while enabled (enable(1) also clears earlier records).  collect() synchronises the events and fills
 * arrays of ftb_profile_families() entries: milliseconds, algorithmic FLOPs, algorithmic bytes, launches. */
int ftb_profile_families(void);
const char* ftb_profile_family_name(int family);
int ftb_profile_enable(int on);
int ftb_profile_collect(double* ms, double* flops, double* bytes, long long* launches);
/* Enables kernels on `device` to load / store memory resident on `peer` (one node, NVLink / NVSwitch).  The output
 * pointers of every entry point may then point into a peer GPU: a sharded run lets the last epilogue store its result
 * directly into the collecting rank's buffer (no separate gather pass). */
int ftb_enable_peer_access(int device, int peer);
/* Peer-mapped result window of a sharded run: the collecting rank allocates and exports a CUDA IPC handle (64 bytes,
 * to be sent to the other processes of the node by any means), the others import it with their own device current and
 * pass pointers into it as output arguments.  ftb_ipc_release: owner != 0 frees, otherwise un-maps. */
int ftb_ipc_alloc(int64_t bytes, int device, void** dev_ptr, unsigned char* handle64);
int ftb_ipc_open(const unsigned char* handle64, int device, void** dev_ptr);
int ftb_ipc_release(void* dev_ptr, int owner);
/* Process-wide tuning.  FTB_TUNE_LSTM_MIN_CHUNK (8..32, default 8): smallest number of utterances a decoder-LSTM
 * cluster takes.  8 = lowest latency of one call (as many clusters as fit in a wave); 32 = throughput mode for
 * several generate() calls in flight on different streams (fewer SMs pinned by the latency-bound recurrence). */
#define FTB_TUNE_LSTM_MIN_CHUNK 1
/* FTB_TUNE_GRU_MIN_CHUNK (8..32, default 8): the same for the CBHG GRU clusters (4 CTAs each): 8 utterances per
 * cluster = 16 clusters / 64 SMs for a batch of 64; 16 halves the SMs held for a ~40 % longer step. */
#define FTB_TUNE_GRU_MIN_CHUNK 2
int ftb_tune(int key, int value);
/* Number of SMs / compute capability of `device`; fails on anything but sm_100. */
int ftb_device_check(int device, int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------- *
 * Operator level (each replaces one torch call site of the reference)
 * ------------------------------------------------------------------------- */

/* LengthRegulator.forward, models/common_layers.py:12-19.
 * plan: dur (B,T) f32 is clamped at 0 IN PLACE (line 13); reps = trunc(dur+0.5)
 * (line 16); cum (B,T) int32 receives the INCLUSIVE prefix sum of reps per row,
 * total (B) int32 the per-row frame count.  The caller reads `total` back (the
 * one D2H of generate) to size L = max(total).
 * expand: out (B,L,C) <- x (B,T,C) rows repeated, zero padded (line 18).
 * elem_bytes is 2 (bf16) or 4 (f32); C*elem_bytes must be a multiple of 16. */
int ftb_length_plan(float* dur, int32_t* cum, int32_t* total, int B, int T, void* stream);
int ftb_length_expand(const void* x, const int32_t* cum, void* out, int B, int T, int L, int C,
                      int elem_bytes, void* stream);
/* The same expansion as an index: idx (B,L) int32 <- b*T + t for the phoneme t whose frames hold j
 * (cum[b,t-1] <= j < cum[b,t]); pad_row for the zero-padded tail (line 18).  out[b,j] == x_rows[idx[b,j]] when row
 * pad_row of x_rows is zero.  Lets a per-row-linear consumer (the decoder LSTM's input projection,
 * models/forward_tacotron.py:317-321) run once per phoneme instead of once per frame. */
int ftb_length_index(const int32_t* cum, int32_t* idx, int B, int T, int L, int pad_row, void* stream);

/* Ragged batches (the batched replacement of gen_forward.py:106-118, which runs one sentence per generate() call):
 * zero the rows t >= lens[b] of a (B, S, row_bytes) tensor (lens: (B) int32, device).  With zeros beyond a row's end
 * every conv sees the zero padding of the solo run; together with the length-aware recurrences (ftb_rnn_bidir_packed)
 * a padded batch reproduces the per-sentence outputs. */
int ftb_zero_tail_rows(void* x, int B, int S, int64_t row_bytes, const int32_t* lens, void* stream);
/* The duration fallback of models/forward_tacotron.py:254-255 decided per ROW over its lens[b] valid positions (what the
 * reference's per-sentence loop computes). */
int ftb_duration_fallback_rows(float* dur, const int32_t* lens, int B, int T, void* stream);

/* Duration fallback, models/forward_tacotron.py:254-255: if the batch-global sum
 * of trunc-toward-zero(dur) is <= 0, fill dur with 2.0.  scratch: 8 bytes. */
int ftb_duration_fallback(float* dur, int64_t n, void* scratch8, void* stream);

/* Conv1d (stride 1, zero padded, truncated to S) as implicit GEMM on
 * channel-last activations, models/common_layers.py:45-52:
 *   y[b,t,n] = sum_{j<k} sum_{c<Cin} w[n, j*Cin + c] * x[b, t + j - pad_left, c]
 * followed by the fused epilogue, in this order:
 *   (+bias[n]) -> (ReLU) -> (*scale[n] + shift[n]) -> (+residual[b,t,n]) -> *out_scale
 * Outputs (any subset): out_f32 (B,S,ldo) / out_bf16 (B,S,ldo) / out_t (B,N,S) f32.
 * A linear layer is k = 1.  Weights must be pre-packed (N, k*Cin) K-major. */
typedef struct ftb_conv_desc {
  int32_t B, S, Cin, N, ktaps, pad_left;
  int32_t lda;  /* row stride of x in elements (>= Cin)               */
  int32_t ldo;  /* row stride of out_f32 / out_bf16 (>= n_offset + N)  */
  int32_t n_offset; /* column offset inside the output row (conv bank concat) */
  int32_t relu;
  const float* bias;
  const float* scale;
  const float* shift;
  const float* residual_f32; /* (B,S,ldr) or NULL */
  const void* residual_bf16; /* (B,S,ldr) or NULL */
  int32_t ldr;
  float out_scale;
  float* out_f32;
  void* out_bf16;
  float* out_t;
} ftb_conv_desc;

/* fp32 SIMT kernel (fp32-accurate: used for the duration predictor). */
int ftb_conv_gemm_f32(const float* x, const float* w_packed, const ftb_conv_desc* d, void* stream);
/* tcgen05 / TMA kernel: x (B,S,lda) bf16, w bf16 (N, k*Cin) with Cin % 64 == 0. */
int ftb_conv_gemm_bf16(const void* x, const void* w_packed, const ftb_conv_desc* d, void* stream);
/* CBHG conv bank, models/common_layers.py:92-100: n_convs (<= 16) convs of the SAME input, written side by
 * side into one (B,S,ldo) tensor at their n_offset (the reference's torch.cat), optionally followed by
 * MaxPool1d(kernel 2, stride 1, padding 1)[:S] along t -- out[t] = max(y[t-1], y[t]), out[0] = y[0] -- fused
 * into the epilogue.  One persistent tcgen05 launch.  descs[i] share B, S, Cin, lda, ldo and the output
 * pointers; no residual / transposed output.  All N must select the same tile width (equal N is enough). */
int ftb_conv_bank_bf16(const void* x, const void* const* w_packed, const ftb_conv_desc* descs, int n_convs,
                       int maxpool, void* stream);
/* Every mbarrier wait of the pipelined kernels (tcgen05 GEMM, LSTM / GRU clusters) is bounded so a protocol bug cannot
 * hang the GPU; a wait that expires counts itself and TRAPS the kernel, so the launch fails loudly (the next CUDA call
 * returns an error) instead of continuing on data that has not arrived.  Returns the number of expired waits since the
 * library was loaded: 0 in a healthy run, -1 when the context is already dead. */
int ftb_tc_timeout_count(void);
/* Pack a reference-layout conv weight (N, Cin, k) f32 into (Npad, k*Cin_pad) K-major
 * f32 (out_bf16 = 0), bf16 (1) or IEEE half (2), zero padded.  Mode 3: (Npad, 6, k, Cin_pad) bf16, the K axis of the
 * split-precision GEMM (three bf16 parts per weight, arranged for the six part products).  Modes 4 (bf16) / 5 (IEEE
 * half): (Npad, 3, k, Cin_pad) = [hi | hi | lo], the K axis of a GEMM over a two-part activation hi | lo
 * (ftb_rnn_bidir_rows lo_off): hi.hi + lo.hi + hi.lo, the operand pair of the output heads. */
int ftb_pack_conv_weight(const float* w, void* out, int N, int Cin, int k, int Npad, int Cin_pad,
                         int out_bf16, void* stream);

/* Linear layer over a two-part 16-bit activation (the output heads lin / post_proj, models/forward_tacotron.py:322,326):
 * x_pair (B,S,2*Cin) = [hi | lo] as written by ftb_rnn_bidir_rows with lo_off = Cin, w_packed from
 * ftb_pack_conv_weight mode 4 (bf16) / 5 (IEEE half); out_f32 (B,S,ldo) = (hi + lo) . (w_hi + w_lo)^T + bias up to the
 * lo.lo term, fp32 accumulation on tcgen05.  Cin % 64 == 0. */
int ftb_linear_pair(const void* x_pair, const void* w_packed, int B, int S, int Cin, int N, const float* bias,
                    float* out_f32, int ldo, int fp16, void* stream);

/* Bidirectional single-layer GRU / LSTM recurrence (torch.nn.GRU / nn.LSTM,
 * batch_first, zero initial state; models/common_layers.py:84, models/forward_tacotron.py:39,165).
 * xg (B,S,2,G*H) f32 holds W_ih x + b_ih (+ b_hh for the gates where it can be
 * folded: all LSTM gates, GRU r and z); w_hh (2,G*H,H) f32 in torch gate order;
 * b_hn (2,H) f32 is the GRU n-gate hidden bias (NULL for LSTM).
 * out (B,S,2H): f32 when out_bf16 == 0, bf16 when 1, IEEE half when 2.
 * H in {64,128}: one CTA per (row, direction), one thread per gate row with its W_hh row in registers, exact fp32.
 * H = 256 (GRU): cluster of 4 CTAs, W_hh resident in registers as 16-bit mma.sync fragments.
 * H = 512 (LSTM): cluster of 16 CTAs, W_hh resident in shared memory as the tcgen05 A operand.
 * Both exchange the hidden state through distributed shared memory.
 * Any other H that is a multiple of 4 up to 2048 (checkpoints trained with non-default sizes): generic fp32 kernel,
 * one CTA per (row, direction) streaming W_hh from L2 every step -- correct, not fast.  Other H: FTB_ERR_UNSUPPORTED. */
int ftb_rnn_bidir(const float* xg, const float* w_hh, const float* b_hn, void* out, int B, int S, int H,
                  int is_lstm, int out_bf16, void* stream);
/* As ftb_rnn_bidir with two extensions the models use (nn.LSTM after LengthRegulator, forward_tacotron.py:317-321):
 *   xrow (H = 512 LSTM; NULL = off): (B,S) int32, frame (b,t) reads row xrow[b,t] of xg (rows of 2*G*H floats) --
 *     the input projection is computed once per PHONEME and gathered by the recurrence (ftb_length_index);
 *   ldo / lo_off (H >= 256): out rows are ldo elements apart (0 = 2H); with a 16-bit out_kind and lo_off > 0 the
 *     rounding remainder h - hi is stored lo_off elements after hi in the same 16-bit type, so the consumer GEMM can
 *     read hi + lo (the output heads need more than an 11-bit activation at trained mel magnitude). */
int ftb_rnn_bidir_rows(const float* xg, const int32_t* xrow, const float* w_hh, const float* b_hn, void* out,
                       int B, int S, int H, int is_lstm, int out_kind, int ldo, int lo_off, void* stream);
/* Packed-sequence semantics (pack_padded_sequence / pad_packed_sequence, models/forward_tacotron.py:224-231): row b is
 * a sequence of lens[b] <= S steps ((B) int32, device); the reverse direction starts at its last valid step with a zero
 * state, and the output beyond it is pad_value (H = 512 LSTM; 0 for the GRUs, where pad_value must be 0). */
int ftb_rnn_bidir_packed(const float* xg, const int32_t* lens, float pad_value, const float* w_hh, const float* b_hn,
                         void* out, int B, int S, int H, int is_lstm, int out_kind, void* stream);

/* DSP.wav_to_mel, utils/dsp.py:71-87,105-107, for a batch of clips packed back to
 * back: audio f32, clip_offsets (n_clips+1) int64 sample offsets, frame_offsets
 * (n_clips+1) int64 with frames_i = 1 + N_i / hop.  out: (n_mels, total_frames)
 * is written per clip as an (n_mels, frames_i) row-major block starting at
 * out + n_mels * frame_offsets[i].  Requires n_fft == win_length == 1024. */
typedef struct ftb_mel_config {
  int32_t sample_rate, n_fft, hop_length, win_length, num_mels;
  float fmin, fmax;
} ftb_mel_config;
typedef struct ftb_mel_handle ftb_mel_handle;
int ftb_mel_create(const ftb_mel_config* cfg, int device, ftb_mel_handle** out);
void ftb_mel_destroy(ftb_mel_handle* h);
int ftb_mel_run(ftb_mel_handle* h, const float* audio, const int64_t* clip_offsets, const int64_t* frame_offsets,
                int n_clips, int64_t total_frames, float* out, int normalize, void* stream);
/* Copies the (num_mels, 1 + n_fft/2) f32 filterbank the handle uses to HOST memory. */
int ftb_mel_filterbank(ftb_mel_handle* h, float* host_out);

/* DSP.griffinlim, utils/dsp.py:89-103 (the step after generate in gen_forward.py:132-134), in its two library calls:
 * ftb_mel_to_stft = librosa.feature.inverse.mel_to_stft(exp(mel), power=1): mel (n_mels, n_frames) f32 (log-mel when
 *   denormalize != 0) -> S_out (1 + n_fft/2, n_frames) f32 >= 0 minimising ||A S - M||^2 (non-negative least squares from
 *   the clipped least-squares start, `iters` accelerated projected-gradient steps; 0 = 64).  The minimiser is not unique;
 *   the objective reached is at or below the reference's L-BFGS-B result.
 * ftb_griffinlim = librosa.griffinlim(S, n_iter, hop, win) with momentum (0.99 upstream) and init='random': the
 *   reference draws the initial phases from an unseeded RNG, so the caller passes them: phase_u (1 + n_fft/2, n_frames)
 *   f32 uniform in [0, 1), angles0 = exp(2 pi i u).  wav_out: hop * (n_frames - 1) f32 samples (istft with center=True,
 *   length=None). */
int ftb_mel_to_stft(ftb_mel_handle* h, const float* mel, int n_frames, int denormalize, int iters, float* S_out,
                    void* stream);
int64_t ftb_griffinlim_workspace_bytes(int n_frames);
int ftb_griffinlim(ftb_mel_handle* h, const float* S, const float* phase_u, int n_frames, int n_iter, float momentum,
                   float* wav_out, void* workspace, int64_t workspace_bytes, void* stream);
/* DSP.trim_silence = librosa.effects.trim(wav, top_db, frame_length=2048, hop_length=512), utils/dsp.py:112-113, for
 * clips packed back to back (clip_offsets as in ftb_mel_run; max_clip_samples = the longest clip): bounds (n_clips, 2)
 * int64 receives [start, end) of the non-silent region of every clip in samples relative to the clip (0, 0 when every
 * frame is below the threshold).  workspace: n_clips * (1 + max_clip_samples / hop_length) floats. */
int ftb_trim_silence(const float* audio, const int64_t* clip_offsets, int n_clips, int max_clip_samples, float top_db,
                     int frame_length, int hop_length, int64_t* bounds, void* workspace, int64_t workspace_bytes,
                     void* stream);

/* ------------------------------------------------------------------------- *
 * Model level: ForwardTacotron.generate, models/forward_tacotron.py:244-330
 * ------------------------------------------------------------------------- */
typedef struct ftb_ft_config { /* keys of config.yaml forward_tacotron.model + num_chars, n_mels */
  int32_t num_chars, embed_dims, series_embed_dims;
  int32_t durpred_conv_dims, durpred_rnn_dims;
  int32_t pitch_conv_dims, pitch_rnn_dims;
  int32_t energy_conv_dims, energy_rnn_dims;
  int32_t rnn_dims;
  int32_t prenet_dims, prenet_k, prenet_num_highways;
  int32_t postnet_dims, postnet_k, postnet_num_highways;
  int32_t n_mels;
  float pitch_strength, energy_strength;
  int32_t gemm_mode; /* 0 (default): IEEE-half operands, fp32 accumulation on tcgen05 (11-bit significand: holds the
                        absolute mel tolerance at trained-checkpoint magnitude; duration predictor fp32-grade, output
                        heads on two-part operands); 1: all GEMMs fp32 SIMT; 2: bf16 operands (same kernels and rate) */
} ftb_ft_config;

typedef struct ftb_ft_handle ftb_ft_handle;

/* Packs the weights once (GEMM layouts, BN -> scale/shift, bias folding).
 * `tensors` is the model's state_dict (load_state_dict contract: every name the
 * reference's strict load expects must be present with the reference's shape). */
int ftb_ft_create(const ftb_ft_config* cfg, const ftb_tensor* tensors, int n_tensors, int device,
                  ftb_ft_handle** out);
void ftb_ft_destroy(ftb_ft_handle* h);

/* Bytes of caller workspace needed by predict / synthesize for these sizes. */
int64_t ftb_ft_workspace_bytes(const ftb_ft_handle* h, int B, int T, int L);

/* Stage A (generate lines 251-262): three SeriesPredictors + duration fallback.
 * tokens (B,T) int64; dur (B,T), pitch (B,T), energy (B,T) f32 out
 * ((B,1,T) of the reference is the same memory).  The three predictors are independent: they run on three
 * internal streams forked from / joined into `stream` with events, so the call is ordered like any other
 * launch on `stream`. */
int ftb_ft_predict(ftb_ft_handle* h, const int64_t* tokens, int B, int T, float alpha, float* dur, float* pitch,
                   float* energy, void* workspace, int64_t workspace_bytes, void* stream);

/* Handle options.  FTB_OPT_OVERLAP_PRENET (default 0): ftb_ft_predict also starts the prenet CBHG of stage B -- it
 * depends on the tokens only -- on an internal stream into handle-owned memory, and the NEXT ftb_ft_synthesize call
 * with the same tokens pointer, B and T picks it up instead of recomputing it.  The caller must not modify the
 * token buffer between the two calls (generate() does not); one-shot, dropped by any other call sequence. */
#define FTB_OPT_OVERLAP_PRENET 1
/* FTB_OPT_SERIALIZE (default 0): run every launch on the caller's stream, one after the other (no side streams, no
 * prenet prefetch).  For profiling: per-launch CUDA-event timings are kernel durations only when nothing overlaps. */
#define FTB_OPT_SERIALIZE 2
/* FTB_OPT_DUR_SIMT (default 0; env FTB_DUR_SIMT): in the 16-bit modes the duration predictor -- whose rounded output
 * must be bit-exact -- runs its convs / GRU input projection as split-precision tensor-core GEMMs: every fp32 operand
 * is carried as three bf16 parts (hi + mid + lo = the fp32 value), six part products accumulate in fp32.  1 = use the
 * fp32 SIMT GEMM instead (the all-fp32 gemm_mode 1 always does). */
#define FTB_OPT_DUR_SIMT 3
/* FTB_OPT_UNFUSED_TAIL (default 0; env FTB_UNFUSED_TAIL): 1 = run pre_highway, the highway layers and the GRU input
 * projection of each CBHG (models/common_layers.py:113-118) as one tcgen05 launch per layer instead of the fused
 * persistent kernel that keeps the activations in shared memory across all of them (csrc/cbhg_tail.cu).  Both paths
 * produce the same bits; the option exists for that comparison and for profiling. */
#define FTB_OPT_UNFUSED_TAIL 4
/* FTB_OPT_LSTM_MIN_CHUNK (default 0 = the process-wide ftb_tune(FTB_TUNE_LSTM_MIN_CHUNK) value): smallest number of
 * utterances per decoder-LSTM cluster for launches of THIS handle.  32 = throughput setting for several batches in
 * flight (4 instead of 6 clusters hold 64 instead of 96 SMs for the whole recurrence; one call alone +12 %). */
#define FTB_OPT_LSTM_MIN_CHUNK 5
int ftb_ft_set_option(ftb_ft_handle* h, int option, int value);

/* Between the stages the Python callbacks pitch_function / energy_function run
 * (generate lines 259, 263), then ftb_length_plan + the D2H of `total`. */

/* Stage B (_generate_mel, lines 289-330).  dur must already be clamped/planned:
 * cum is the output of ftb_length_plan.  mel / mel_post: (B, n_mels, L) f32. */
int ftb_ft_synthesize(ftb_ft_handle* h, const int64_t* tokens, const int32_t* cum, const float* pitch,
                      const float* energy, int B, int T, int L, float* mel, float* mel_post, void* workspace,
                      int64_t workspace_bytes, void* stream);

/* Stage B of the teacher-forced forward() in eval mode (models/forward_tacotron.py:203-242; the GTA feature dump of
 * train_forward.py:33-52): as ftb_ft_synthesize, but the decoder LSTM runs over PACKED sequences -- row b stops at
 * mel_lens[b] (int32, device; pack_padded_sequence), the reverse direction starts at its last valid frame, and the
 * LSTM output beyond it is pad_value (pad_packed_sequence, padding_value = -11.5129).  lin and the postnet then run
 * over the padded rows exactly as the reference does.  L = max(mel_lens) <= the expanded length. */
int ftb_ft_synthesize_packed(ftb_ft_handle* h, const int64_t* tokens, const int32_t* cum, const float* pitch,
                             const float* energy, const int32_t* mel_lens, float pad_value, int B, int T, int L,
                             float* mel, float* mel_post, void* workspace, int64_t workspace_bytes, void* stream);

/* Ragged batch = the reference's per-sentence loop (gen_forward.py:106-118) in one call: row b holds tok_lens[b] tokens
 * (padded with anything), and every output of row b equals what generate() returns for that sentence alone: convs see
 * zero padding beyond the row's end, recurrences run over the row's own length, the duration fallback is decided per
 * row.  predict_ragged writes 0 to dur / pitch / energy beyond tok_lens[b]; after the callbacks the caller zeroes those
 * positions again (ftb_zero_tail_rows), plans the lengths, and passes total (B) as mel_lens to synthesize_ragged.  Frames
 * of row b beyond mel_lens[b] in mel / mel_post are padding without meaning. */
int ftb_ft_predict_ragged(ftb_ft_handle* h, const int64_t* tokens, const int32_t* tok_lens, int B, int T, float alpha,
                          float* dur, float* pitch, float* energy, void* workspace, int64_t workspace_bytes, void* stream);
int ftb_ft_synthesize_ragged(ftb_ft_handle* h, const int64_t* tokens, const int32_t* tok_lens, const int32_t* cum,
                             const float* pitch, const float* energy, const int32_t* mel_lens, int B, int T, int L,
                             float* mel, float* mel_post, void* workspace, int64_t workspace_bytes, void* stream);

/* Sub-module entry points (row a3 / a4 of the scope table; used by the mirrored
 * SeriesPredictor / CBHG modules and their parity tests).
 * which: 0 dur_pred, 1 pitch_pred, 2 energy_pred.  out (B,T) f32 (no fallback). */
int ftb_ft_series_predictor(ftb_ft_handle* h, int which, const int64_t* tokens, int B, int T, float alpha,
                            float* out, void* workspace, int64_t workspace_bytes, void* stream);
/* which: 0 prenet, 1 postnet.  x (B,S,Cin) f32 channel-last -> out (B,S,2*dims) f32. */
int ftb_ft_cbhg(ftb_ft_handle* h, int which, const float* x, int B, int S, float* out, void* workspace,
                int64_t workspace_bytes, void* stream);
/* Number of kernels the last predict / synthesize call launched (bench bookkeeping). */
int ftb_ft_last_launch_count(const ftb_ft_handle* h);

/* ------------------------------------------------------------------------- *
 * Model level: FastPitch.generate, models/fast_pitch.py:286-340
 * ------------------------------------------------------------------------- */
/* The 16-bit multi-head attention core of FastPitch's FFT blocks on its own (nn.MultiheadAttention inside
 * models/fast_pitch.py:64,80-82): ctx (B,S,E) = softmax(q k^T / sqrt(hd) + key_padding_mask) v on the packed projection
 * qkv (B,S,3E) = [q | k | v], heads split along E (hd = E / heads in {64, 128}).  tokens: (B,S) int64 ids, keys with id
 * 0 are masked; or NULL.  fp16: the 16-bit type is IEEE half (else bfloat16).  impl 0 = the tcgen05 / TMEM kernel the
 * models use, 1 = the mma.sync kernel it replaced (kept for comparison). */
int ftb_attention_16(const void* qkv, const int64_t* tokens, void* ctx, int B, int S, int E, int heads, int fp16, int impl,
                     void* stream);

typedef struct ftb_fp_config { /* keys of config.yaml fast_pitch.model + num_chars, n_mels */
  int32_t num_chars, n_mels;
  int32_t durpred_d_model, durpred_n_heads, durpred_layers, durpred_d_fft;
  int32_t pitch_d_model, pitch_n_heads, pitch_layers, pitch_d_fft;
  int32_t energy_d_model, energy_n_heads, energy_layers, energy_d_fft;
  int32_t d_model, conv1_kernel, conv2_kernel;
  int32_t prenet_layers, prenet_heads, prenet_fft;
  int32_t postnet_layers, postnet_heads, postnet_fft;
  float pitch_strength, energy_strength;
  int32_t gemm_mode; /* 0: IEEE-half tcgen05 GEMMs + tensor-core attention (duration predictor fp32); 1: all fp32; 2: bf16 */
} ftb_fp_config;

typedef struct ftb_fp_handle ftb_fp_handle;
int ftb_fp_create(const ftb_fp_config* cfg, const ftb_tensor* tensors, int n_tensors, int device,
                  ftb_fp_handle** out);
void ftb_fp_destroy(ftb_fp_handle* h);
int64_t ftb_fp_workspace_bytes(const ftb_fp_handle* h, int B, int T, int L);
int ftb_fp_predict(ftb_fp_handle* h, const int64_t* tokens, int B, int T, float alpha, float* dur, float* pitch,
                   float* energy, void* workspace, int64_t workspace_bytes, void* stream);
/* mel (B, n_mels, L) f32; the reference returns the same tensor as 'mel' and 'mel_post'. */
int ftb_fp_synthesize(ftb_fp_handle* h, const int64_t* tokens, const int32_t* cum, const float* pitch,
                      const float* energy, int B, int T, int L, float* mel, void* workspace,
                      int64_t workspace_bytes, void* stream);

/* FastPitch.forward in eval mode (models/fast_pitch.py:243-283; the teacher-forced pass): the three predictors with the
 * token padding mask (keys with id 0 ignored) and no fallback -> dur_hat / pitch_hat / energy_hat (B,T); then stage B
 * with the batch's durations (cum from ftb_length_plan), pitch and energy, the prenet masked by the tokens and the
 * postnet by frame_mask: (B,L) int64, 0 = padded frame (t >= mel_len[b], make_mel_len_mask :47-51).  mel (B,n_mels,L). */
int ftb_fp_forward_eval(ftb_fp_handle* h, const int64_t* tokens, const int32_t* cum, const float* pitch,
                        const float* energy, const int64_t* frame_mask, int B, int T, int L, float* dur_hat,
                        float* pitch_hat, float* energy_hat, float* mel, void* workspace, int64_t workspace_bytes,
                        void* stream);
int ftb_fp_last_launch_count(const ftb_fp_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* FTB200_H_ */
