/*
 * bsed.h -- C ABI of libbsed.so: the B200 (sm_100a) implementation of the sound-event-detection
 * hot path of fumchin/bird-sound-event-detecion.
 *
 * The reference is 100% Python and has no FFI of its own; the boundary it offers is a set of Python
 * call signatures.  Each entry point below names the reference interface it stands behind
 * (file:line relative to the reference tree).  The Python package `bird-sound-event-detecion_b200`
 * keeps those Python signatures and binds this library with ctypes.
 *
 * Conventions
 *   - every function returns int: 0 = BSED_OK, negative = BSED_E_*; never throws, never exits.
 *     `bsed_last_error()` gives a thread-local message for the last failure.
 *   - all tensor arguments are DEVICE pointers to contiguous fp32 (unless stated), owned by the
 *     caller; the library never allocates or frees caller-visible memory.  Scratch space is passed
 *     in as `workspace` (size from the matching *_workspace_bytes call).
 *   - all work is enqueued on the caller's stream (`void* stream` is a cudaStream_t); no hidden
 *     synchronisation, no default-stream use => CUDA-graph capturable.
 *   - activations inside the CRNN are channels-last (B, T, F, C); the public tensors
 *     (B,1,T,128) input and (B,313,256)/(B,313,20)/(B,20) outputs are layout-identical to the
 *     reference's.
 *   - there is no CPU fallback: on a machine without a usable GPU every compute call fails with
 *     BSED_E_CUDA.
 */
#ifndef BSED_H_
#define BSED_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSED_ABI_VERSION 4

#define BSED_OK 0
#define BSED_E_INVALID (-1) /* bad argument / shape / alignment */
#define BSED_E_CUDA (-2)    /* CUDA runtime error (message has the detail) */
#define BSED_E_WORKSPACE (-3) /* workspace too small */
#define BSED_E_STATE (-4)   /* call order violation (e.g. backward without a saved forward) */

typedef struct bsed_context* bsed_handle;

int bsed_version(void);
const char* bsed_last_error(void);

/* One handle per (process, device).  Builds the constant tables (Hamming window, FFT twiddles,
 * Slaney mel filterbank) on the device.  Constants follow src/data/config.py:47-57. */
int bsed_create(int device, bsed_handle* out);
int bsed_destroy(bsed_handle h);

/* ------------------------------------------------------------------------------------------
 * Frontend.
 * ------------------------------------------------------------------------------------------ */

/* 1 + n_samples / 255   (librosa.stft, center=True) */
int bsed_frontend_n_frames(int n_samples);

/* preprocess(audio, compute_log=False)            src/data/preprocess.py:18-45
 * audio [B][n_samples] -> mel [B][n_frames][128] amplitude-mel (|STFT| x Slaney filterbank).
 * n_samples >= 1025 (reflect padding). */
int bsed_melspec(bsed_handle h, const float* audio, int B, int n_samples, float* mel, void* stream);

/* get_transforms(frames, scaler, add_axis=0, noise_dict_params) applied to cached amplitude-mel:
 *   [AugmentGaussianNoise(snr)] -> ApplyLog -> PadOrTrunc(frames) -> ToTensor -> [Normalize]
 *                                                 src/data/Transforms.py:74-139,155-197,304-322
 * mel [B][t_in][128]; unit_noise [B][t_in][128] standard-normal draws or NULL (no noise);
 * scaler_mean / scaler_std [128] or NULL; out [B][frames][128] (rows >= t_in are 0, truncation if
 * t_in > frames).  workspace: bsed_amp_to_db_workspace_bytes(B). */
size_t bsed_amp_to_db_workspace_bytes(int B);
int bsed_amp_to_db(bsed_handle h, const float* mel, const float* unit_noise, float snr_db, int B,
                   int t_in, int frames, const float* scaler_mean, const float* scaler_std,
                   float* out, void* workspace, size_t workspace_bytes, void* stream);

/* preprocess(audio, compute_log=True) -- or preprocess + the load-time ApplyLog -> PadOrTrunc -> [Normalize] of a clean
 * clip -- in one call: the STFT + mel kernel also takes the per-clip maximum the 80 dB clamp needs, so the dB pass reads
 * the amplitude-mel once.  mel [B][n_frames][128] receives the amplitude-mel (the cache format), out [B][frames][128] the
 * log-mel; bit-identical to bsed_melspec followed by bsed_amp_to_db without noise.  workspace as bsed_amp_to_db. */
int bsed_logmel(bsed_handle h, const float* audio, int B, int n_samples, int frames, const float* scaler_mean,
                const float* scaler_std, float* mel, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Post-processing.         src/evaluation_measures.py:188-209, src/utilities/ManyHotEncoder.py:148-164
 * strong [B][T][C] probabilities -> events, class-major then time, per clip.
 *   b = p >= threshold ; median over `win` frames (scipy 'reflect') ; maximal runs [on, off).
 * events [B][max_events][3] int32 (class, onset_frame, offset_frame); n_events [B] int32 (the true
 * count; entries beyond max_events are dropped).  T <= 1024, C <= 32.
 * ------------------------------------------------------------------------------------------ */
int bsed_median_decode(bsed_handle h, const float* strong, int B, int T, int C, float threshold,
                       int win, int32_t* events, int max_events, int32_t* n_events, void* stream);

/* ------------------------------------------------------------------------------------------
 * CRNN + Predictor engine.   src/models/CNN.py:33-84, RNN.py:7-16, CRNN.py:178-240,548-577
 *
 * Parameters live in ONE flat fp32 buffer in the reference's named_parameters() order
 * (cnn.conv0.weight, cnn.conv0.bias, cnn.batchnorm0.weight, cnn.batchnorm0.bias,
 *  cnn.glu0.linear.weight, cnn.glu0.linear.bias, ... x7, rnn.rnn.weight_ih_l0, weight_hh_l0,
 *  bias_ih_l0, bias_hh_l0, *_reverse, *_l1, *_l1_reverse), each tensor in the reference's own shape.
 * The Predictor has its own flat buffer (see bsed_predictor_*); hosts normally place it right after
 * the CRNN parameters so one optimiser call covers both.
 * BatchNorm buffers live in a second flat fp32 buffer (running_mean_i, running_var_i per block) and
 * an int64 array num_batches_tracked[7].
 * ------------------------------------------------------------------------------------------ */
#define BSED_MAX_CNN_LAYERS 8

/* fpn = 1 selects CRNN_fpn / CNN_FPN (src/models/CRNN.py:243-337, src/models/CNN_FPN.py:33-100): after the CNN trunk
 * the shared-weight stage cnn_fcn (Conv3x3 128->128) -> bn_fcn -> glu -> dropout -> AvgPool[2,1] is applied twice
 * (313 -> 156 -> 78 frames); rnn / rnn_2 / rnn_4 run on the three time scales; after dropout the scales are merged
 * top-down: x_2 = conv1x1_2(cat(x_2, up(x_4))), x = conv1x1_4(cat(x, up(x_2))) with bilinear align_corners=True
 * upsampling along time.  The flat parameter buffer follows CRNN_fpn.named_parameters(): cnn.cnn.* (7 blocks),
 * cnn.cnn_fcn.{weight,bias}, cnn.glu.linear.{weight,bias}, cnn.bn_fcn.{weight,bias}, cnn.conv1x1.{weight,bias}
 * (registered but unused by the reference's forward: its gradient stays zero), rnn.rnn.*, rnn_2.rnn.*, rnn_4.rnn.*,
 * conv1x1_2.{weight,bias}, conv1x1_4.{weight,bias}.  BatchNorm buffers: the 7 blocks then bn_fcn
 * (num_batches_tracked[8]; bn_fcn is updated twice per train-mode forward, as in the reference). */

typedef struct {
  int n_frames;                        /* 1255 */
  int n_mels;                          /* 128  */
  int n_cnn;                           /* 7    */
  int filters[BSED_MAX_CNN_LAYERS];    /* 16,32,64,128,128,128,128  (multiples of 16, <= 128) */
  int pool_t[BSED_MAX_CNN_LAYERS];     /* 2,2,1,1,1,1,1 */
  int pool_f[BSED_MAX_CNN_LAYERS];     /* 2,2,2,2,2,2,2 */
  int rnn_hidden;                      /* 128 (fixed by the recurrence kernel) */
  int rnn_layers;                      /* 2 */
  int n_class;                         /* 20 (<= 20) */
  float dropout;                       /* 0.5 */
  float bn_eps;                        /* 1e-3 */
  float bn_momentum;                   /* 0.99 */
  int fpn;                             /* 0 = CRNN (src/models/CRNN.py:178-240); 1 = CRNN_fpn (:243-337), see below */
} bsed_crnn_cfg;

typedef struct bsed_crnn_plan* bsed_plan;

/* Build a plan for batches of up to max_clips clips. */
int bsed_plan_create(bsed_handle h, const bsed_crnn_cfg* cfg, int max_clips, bsed_plan* out);
int bsed_plan_destroy(bsed_plan p);

/* Arithmetic of the dense contractions (3x3 convolutions, GLU / GRU-input linears and their gradients):
 *   BSED_PRECISION_TF32X3 (default) error-compensated 3xTF32 on the same tcgen05 kernels: the activation tile is split
 *                                  in shared memory into its tf32 high part and fp32 remainder, the weights are
 *                                  pre-split, and every k-step issues a*w_hi + a*w_lo + a_lo*w_hi with fp32
 *                                  accumulation in TMEM -- fp32-grade products (relative error ~2^-21), the parity mode
 *                                  against the reference's fp32 arithmetic (src/models/CNN.py:9-16 nn.Linear is fp32
 *                                  on a GPU as well).  Forward and data-gradient contractions; the weight-gradient
 *                                  reductions stay single-pass tf32 (what cuDNN gives the reference on a GPU).
 *   BSED_PRECISION_TF32            single-pass tcgen05.mma kind::tf32 everywhere -- what the reference gets from cuDNN
 *                                  convolutions on a GPU (torch.backends.cudnn.allow_tf32 defaults to True); stated
 *                                  tolerance 5e-3 on train-mode probabilities, not an inference parity mode;
 *   BSED_PRECISION_FP32            fp32 FMA on the CUDA cores (cross-check of the two above).
 * Everything else (BatchNorm, gates, pooling, GRU recurrence, head, losses, optimiser) is fp32 in all three. */
#define BSED_PRECISION_FP32 0
#define BSED_PRECISION_TF32 1
#define BSED_PRECISION_TF32X3 2
int bsed_plan_set_precision(bsed_plan p, int precision);
int bsed_plan_get_precision(bsed_plan p);

int64_t bsed_plan_param_count(bsed_plan p);       /* floats in the flat parameter buffer  */
int64_t bsed_plan_bn_buffer_count(bsed_plan p);   /* floats in the flat BN running-stat buffer */
int bsed_plan_out_frames(bsed_plan p);            /* 313 */
/* offsets (in floats) of the tensors of the flat parameter buffer, in order; returns the count.
 * Lets the host verify its view of the layout. */
int bsed_plan_param_offsets(bsed_plan p, int64_t* offsets, int max_n);
size_t bsed_plan_workspace_bytes(bsed_plan p);

/* A forward "group" is one reference model call: BatchNorm batch statistics (train mode) are taken
 * over the clips of one group only; groups are processed together in the same launches.
 * flags */
#define BSED_F_TRAIN 1      /* BN batch statistics + running-stat update + dropout (model.train()) */
#define BSED_F_SAVE 2       /* keep what backward needs (requires BSED_F_TRAIN)                     */

typedef struct {
  const float* params;        /* flat parameter buffer used by this group                          */
  float* bn_buffers;          /* flat running stats (updated in train mode), may alias across groups */
  int64_t* num_batches_tracked; /* [n_cnn] or NULL                                                  */
  int first_clip, n_clips;    /* clips [first_clip, first_clip + n_clips) of x                     */
} bsed_group;

/* CRNN.forward for n_groups groups over x [B][n_frames][n_mels] (== (B,1,T,128)).
 * enc [B][313][256] receives the encoder output after the final dropout (what CRNN.forward returns
 * twice, as `x` and `d_input`).   src/models/CRNN.py:211-240
 * dropout_seed/step key the stateless dropout hash (oracle/crnn.py: mix_key / keep_mask). */
int bsed_crnn_forward(bsed_plan p, const bsed_group* groups, int n_groups, const float* x, int B,
                      int flags, uint64_t dropout_seed, uint64_t dropout_step, float* enc,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the last BSED_F_SAVE forward for the groups whose bit is set in `group_mask`, all of
 * which must be adjacent and share one parameter buffer.  d_enc [B][313][256]: gradient w.r.t. the
 * encoder output (rows of unmasked clips are ignored).  grads: flat buffer, same layout as params;
 * accumulate != 0 adds to it, else it is overwritten.  x of the forward must still be valid. */
int bsed_crnn_backward(bsed_plan p, uint32_t group_mask, const float* d_enc, float* grads,
                       int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* Predictor.forward / backward                                   src/models/CRNN.py:548-577
 * Parameters: flat fp32 buffer dense.weight [C][256], dense.bias [C], dense_softmax.weight [C][256],
 * dense_softmax.bias [C].  enc [n][313][256]; logits [n][313][bsed_predictor_ldl()] is written by the
 * forward and consumed by the backward (cols 0..C-1 dense, C..2C-1 dense_softmax pre-activations);
 * strong [n][313][C], weak [n][C].  inference != 0: strong *= (weak > 0.5)   (CRNN.py:570-574).
 * Backward: d_strong / d_weak may be NULL (= zero); writes d_enc [n][313][256] and the parameter
 * gradients (same layout as the parameters).  workspace: bsed_predictor_workspace_bytes(p, n). */
int64_t bsed_predictor_param_count(bsed_plan p);
int bsed_predictor_param_offsets(bsed_plan p, int64_t* offsets, int max_n);
int bsed_predictor_ldl(void);
size_t bsed_predictor_workspace_bytes(bsed_plan p, int n_clips);
int bsed_predictor_forward(bsed_plan p, const float* pred_params, const float* enc, int n_clips,
                           int inference, float* logits, float* strong, float* weak, void* workspace,
                           size_t workspace_bytes, void* stream);
int bsed_predictor_backward(bsed_plan p, const float* pred_params, const float* enc, const float* logits,
                            const float* strong, const float* weak, const float* d_strong,
                            const float* d_weak, int n_clips, float* d_enc, float* grads, int accumulate,
                            void* workspace, size_t workspace_bytes, void* stream);

/* Pointer to a named intermediate of the last forward/backward inside the workspace (debug and
 * tests): "xhat<i>", "lin<i>", "pool<i>", "gru<l>", "dxn"  */
int bsed_plan_debug_tensor(bsed_plan p, void* workspace, const char* name, float** ptr,
                           int64_t* numel);

/* ------------------------------------------------------------------------------------------
 * Losses of the mean-teacher step.                      src/main.py:376,405,434,439-449,474-477
 * Clips [syn_first, syn_first+syn_n) are the synthetic (strongly labelled) batch: BCE(strong, target)
 * + BCE(weak, max_t target); clips [real_first, real_first+real_n) are the real batch: cons_w *
 * (MSE(strong, strong_ema) + MSE(weak, weak_ema)), teacher tensors indexed from 0.
 * Writes losses[4] = {strong_bce, weak_bce, cons_strong, cons_weak} (device), and the gradients
 * d_strong / d_weak (same shapes as strong / weak, zero outside the two ranges).
 * ------------------------------------------------------------------------------------------ */
int bsed_mt_loss(bsed_handle h, const float* strong, const float* weak, int B, int T, int C,
                 int syn_first, int syn_n, const float* syn_target, int real_first, int real_n,
                 const float* strong_ema, const float* weak_ema, float cons_w, float* losses,
                 float* d_strong, float* d_weak, void* stream);

/* ------------------------------------------------------------------------------------------
 * Shift-consistency training (the ISP / SCT branch).           src/main_baseline.py:229-277,372-529
 * bsed_roll_clips: out[b] = torch.roll(torch.roll(x[b], shift_t[b], time), shift_f[b], frequency) for clips
 * x [B][T][F] (shift arrays on the device, either may be NULL = no shift along that axis); out must not alias x.
 * bsed_loss_terms: a list of BCELoss / MSELoss terms over clip ranges of strong [B][T][C] / weak [B][C]:
 *   kind BCE_STRONG / MSE_STRONG: pred = strong[pred_first .. +n_clips), ref [n_clips][T][C], optionally rolled per clip
 *        along time (ref'[k] = torch.roll(ref[k], roll[k], 0), roll on the device) -- the rolled targets and rolled
 *        (detached) predictions of the reference;
 *   kind BCE_WEAK / MSE_WEAK: pred = weak[pred_first .. +n_clips), ref [n_clips][C], or with ref_is_strong a strong
 *        target [n_clips][T][C] whose maximum over time is the weak target (syn_target.max(-2)[0]).
 * Every term is a mean over its own elements (torch reduction='mean'); losses[slot] += weight * mean and
 * d_strong / d_weak += grad_weight * d(mean)/d(pred) (grad_weight 0: value only, e.g. terms the reference only logs).
 * `terms` is a HOST array; losses [n_slots], d_strong [B][T][C], d_weak [B][C] are device buffers, zeroed here first.
 * Terms run as consecutive launches in array order, so the accumulation order is fixed.
 * ------------------------------------------------------------------------------------------ */
#define BSED_LOSS_BCE_STRONG 0
#define BSED_LOSS_BCE_WEAK 1
#define BSED_LOSS_MSE_STRONG 2
#define BSED_LOSS_MSE_WEAK 3
typedef struct {
  int kind;
  int pred_first, n_clips;
  const float* ref;
  const int32_t* roll;
  int ref_is_strong;
  float weight;
  float grad_weight;
  int slot;
} bsed_loss_term;
int bsed_loss_terms(bsed_handle h, const float* strong, const float* weak, int B, int T, int C,
                    const bsed_loss_term* terms, int n_terms, float* losses, int n_slots, float* d_strong,
                    float* d_weak, void* stream);
int bsed_roll_clips(bsed_handle h, const float* x, const int32_t* shift_t, const int32_t* shift_f, float* out, int B,
                    int T, int F, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser + EMA in one pass over the flat buffers.
 *   Adam   torch.optim.Adam(lr, betas=(.9,.999), eps=1e-8, weight_decay=0)   src/main.py:823-828
 *   SGD    torch.optim.SGD(momentum=.9, weight_decay=1e-4, nesterov=True)
 *                                                    src/main_scmt_ada_weak_seperate.py:858-870
 *   EMA    ema = a*ema + (1-a)*param, a = min(1 - 1/(step+1), alpha)   src/main.py:86-100
 * grad is multiplied by grad_scale first (1/world_size after a sum all-reduce).
 * ema may be NULL (no teacher).  `step` is the 1-based optimiser step (Adam bias correction);
 * `ema_step` is the global_step passed to update_ema_variables.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int kind;            /* 0 = Adam, 1 = SGD-Nesterov */
  float lr, beta1, beta2, eps, weight_decay, momentum;
  float grad_scale;
  float ema_alpha;     /* 0.999 */
  int64_t step;
  int64_t ema_step;
} bsed_opt_cfg;

int bsed_opt_ema_step(bsed_handle h, float* params, const float* grads, float* m, float* v,
                      float* ema, int64_t n, const bsed_opt_cfg* cfg, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel step in one kernel: gradient reduce-scatter over NVLink peer memory + optimiser + EMA + all-gather.
 * The reference trains on one GPU; data-parallel replicas exchange ONE flat gradient buffer per step.
 * Each rank exports its flat gradient, parameter and EMA buffers and a zero-initialised int32[64] flag block
 * (bsed_ipc_export -> 64-byte handle + byte offset inside the allocation, exchanged by the host), maps its peers'
 * (bsed_ipc_open), and every step calls bsed_dp_opt_ema_step with the per-rank device pointers (entry `rank` = its own
 * buffers; peer_ema NULL = no teacher) and epoch = 1, 2, 3, ...  Rank r owns slice r of the flat buffers: the kernel
 * waits until every peer's gradients are complete, sums the slice in rank order straight out of peer memory, applies
 * bsed_opt_ema_step's update with cfg->grad_scale (1 / world) to its own copy, stores the new parameter / EMA values
 * into every peer's buffers, and leaves only when every peer has done the same (all replicas bit-identical; m and v are
 * only touched inside the owner's slice).  world <= 8.  A peer that does not arrive within BSED_DP_TIMEOUT_S (default
 * 60 s) is fatal for the run: the rank raises the sticky flag[33], applies nothing further and never signals depart
 * (its peers time out as well); hosts must poll flag[33].  Buffers must come from cudaMalloc-backed allocations (the
 * default torch allocator).
 * ------------------------------------------------------------------------------------------ */
#define BSED_IPC_HANDLE_BYTES 64
int bsed_ipc_export(bsed_handle h, const void* dev_ptr, unsigned char* handle, uint64_t* offset);
int bsed_ipc_open(bsed_handle h, const unsigned char* handle, uint64_t offset, void** mapped);
int bsed_ipc_close(bsed_handle h, void* mapped, uint64_t offset);
int bsed_dp_opt_ema_step(bsed_handle h, int rank, int world, const float* const* peer_grads,
                         float* const* peer_params, float* const* peer_ema, int32_t* const* peer_flags,
                         int64_t epoch, float* m, float* v, int64_t n, const bsed_opt_cfg* cfg, void* stream);

/* ------------------------------------------------------------------------------------------
 * Device-resident step state: what makes one training iteration CUDA-graph capturable.
 * Everything that changes from one iteration to the next on the host side of the reference loop -- the dropout draw,
 * the Adam bias corrections (src/main.py:823-828 torch.optim.Adam), the EMA coefficient min(1 - 1/(step+1), alpha)
 * (src/main.py:86-100), the consistency weight max_consistency_cost * exp_rampup(step) (src/main.py:474-477,
 * src/utilities/ramps.py:11-18), the epoch of the data-parallel exchange -- lives in one small device struct.
 * bsed_step_state_advance enqueues a one-thread kernel that increments the counters and recomputes the derived
 * scalars (double precision, the host formulas); while a state is installed with bsed_set_step_state, the
 * step-dependent entry points ignore their by-value step arguments and read the struct instead:
 *   bsed_crnn_forward (dropout_seed / dropout_step -> keys[]), bsed_mt_loss (cons_w), bsed_opt_ema_step and
 *   bsed_dp_opt_ema_step (cfg->step, cfg->ema_step, cfg->lr, epoch), bsed_ema_buffers (ema_step).
 * A captured graph of [advance, forward, loss, backward, optimiser + EMA] can then be replayed unchanged.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t global_step;   /* index of the current iteration (dropout draw, rampup); the host presets start - 1 */
  int64_t opt_step;      /* optimiser updates applied so far, counting the current one (Adam's t)          */
  int64_t dp_epoch;      /* bsed_dp_opt_ema_step epoch of the current iteration                             */
  uint32_t keys[16];     /* dropout keys of the current iteration, one per stream (oracle/crnn.py: mix_key)  */
  float lr;              /* learning rate: written by the HOST whenever the schedule changes it              */
  float cons_w;          /* consistency weight of the current iteration                                      */
  float step_size;       /* lr / (1 - beta1^t)                                                               */
  float bc2_sqrt;        /* sqrt(1 - beta2^t)                                                                */
  float ema_a, ema_b;    /* EMA coefficient of this iteration and (float)(1 - a)                             */
  int32_t first_step;    /* SGD: the momentum buffer starts as the gradient                                  */
  int32_t pad;
} bsed_step_state;

typedef struct {
  uint64_t dropout_seed;
  int64_t key_mul, key_add;     /* dropout step of iteration g = key_mul * g + key_add (1, 0 for src/main.py)   */
  float beta1, beta2;           /* Adam                                                                          */
  float ema_alpha;              /* 0.999                                                                         */
  float max_consistency_cost;   /* src/data/config.py:84                                                         */
  int64_t rampup_length;        /* n_epoch_rampup * len(loader); 0 = weight 1                                    */
} bsed_step_cfg;

int bsed_step_state_advance(bsed_handle h, bsed_step_state* state, const bsed_step_cfg* cfg, void* stream);
/* Install (device pointer) or remove (NULL) the step state the entry points above read.  Per handle, not per stream. */
int bsed_set_step_state(bsed_handle h, const bsed_step_state* state);

/* State-dict flavour of update_ema_variables for the non-parameter entries: BN running stats
 * (fp32) and num_batches_tracked (int64, blended in fp32 and truncated, as load_state_dict does). */
int bsed_ema_buffers(bsed_handle h, const float* bn_buffers, float* ema_bn_buffers, int64_t n,
                     const int64_t* nbt, int64_t* ema_nbt, int n_nbt, float ema_alpha,
                     int64_t ema_step, void* stream);

/* ------------------------------------------------------------------------------------------
 * Adversarial domain-adaptation branch.      src/models/CRNN_GRL.py:16-52, src/DA/cdan_frame.py:89-119
 * Clip_Discriminator: d_input [B][313][256] -> 5 x [conv3x3 stride 2 -> BatchNorm2d -> LeakyReLU(.2)]
 * (channels 128, 64, 32, 16, 8) -> AdaptiveAvgPool2d((2,1)) -> Linear(16,1) -> sigmoid -> prob [B].
 * Parameters: one flat fp32 buffer in the reference's named_parameters() order (conv_1..5 weight/bias, dense_d
 * weight/bias, bn_1..5 weight/bias; bsed_disc_param_count floats); BatchNorm running stats: flat fp32
 * (running_mean_l, running_var_l per layer) + int64 num_batches_tracked[5].
 * bsed_disc_backward undoes the last train-mode bsed_disc_forward on the same workspace; d_dinput (may be NULL) is
 * the gradient handed to the gradient-reversal layer.  bsed_disc_bce: mean BCE against the domain labels + gradient.
 * ------------------------------------------------------------------------------------------ */
int bsed_disc_set_precision(bsed_handle h, int precision);   /* BSED_PRECISION_FP32 (default) | _TF32 | _TF32X3 */
/* dst = alpha * src (n floats; dst may alias src).  The gradient-reversal layer's backward, -coeff * grad
 * (src/DA/grl.py:19-31), between bsed_disc_backward and bsed_crnn_backward. */
int bsed_scale_f32(bsed_handle h, float* dst, const float* src, int64_t n, float alpha, void* stream);
int64_t bsed_disc_param_count(void);
int64_t bsed_disc_bn_buffer_count(void);
size_t bsed_disc_workspace_bytes(int B);
int bsed_disc_forward(bsed_handle h, const float* params, float* bn_buffers, int64_t* num_batches_tracked,
                      const float* d_input, int B, int train, float* prob, void* workspace, size_t workspace_bytes,
                      void* stream);
int bsed_disc_backward(bsed_handle h, const float* params, const float* prob, const float* d_prob, int B, float* grads,
                       int accumulate, float* d_dinput, void* workspace, size_t workspace_bytes, void* stream);
int bsed_disc_bce(bsed_handle h, const float* prob, const float* label, int B, float* loss, float* d_prob, void* stream);

/* ------------------------------------------------------------------------------------------
 * ResNet-18 weak tagger, inference path.    src/audio_tagging_system_cnn.py:50-64 (Net_resnet = torchvision resnet18,
 * conv1 -> Conv2d(1,64,7,2,3,bias=False), fc -> Linear(512,20), sigmoid), src/audio_tagging_inference.py:123-133,289-316
 * In eval mode each BatchNorm folds into the preceding convolution (host, at load time); a stage then is
 *   bsed_im2col_nhwc -> bsed_gemm_nt_tc / bsed_gemm_nn (+ bias) -> bsed_add_relu (with the residual for a block's
 * second convolution), plus the stem's max-pool, the global average pool and the sigmoid.  Tensors are channels-last.
 *   im2col: col[(b*Ho+ho)*Wo+wo][(ky*kw+kx)*Cin+ci] = x[b][ho*sh-ph+ky][wo*sw-pw+kx][ci] (0 outside), columns up to
 *           Kpad (a multiple of 4, >= kh*kw*Cin) zero-filled; Ho = (H+2ph-kh)/sh+1, Wo likewise.
 *   add_relu: y = relu(y + residual) in place, residual may be NULL, n % 4 == 0.
 *   maxpool: nn.MaxPool2d(k, s, p); avgpool: AdaptiveAvgPool2d(1) over the HW pixels; sigmoid_rows: out[r][c] =
 *           sigmoid(logits[r*ld + c]), c < C.
 * ------------------------------------------------------------------------------------------ */
int bsed_im2col_nhwc(bsed_handle h, const float* x, float* col, int B, int H, int W, int Cin, int kh, int kw, int sh,
                     int sw, int ph, int pw, int Ho, int Wo, int Kpad, void* stream);
int bsed_add_relu(bsed_handle h, float* y, const float* residual, int64_t n, void* stream);
int bsed_maxpool_nhwc(bsed_handle h, const float* x, float* y, int B, int H, int W, int C, int k, int s, int p, int Ho,
                      int Wo, void* stream);
int bsed_avgpool_nhwc(bsed_handle h, const float* x, float* y, int B, int HW, int C, void* stream);
int bsed_sigmoid_rows(bsed_handle h, const float* logits, int ld, float* out, int rows, int C, void* stream);

/* Training path of the tagger (src/audio_tagging_system_cnn.py:199-416): every convolution is bias-free and followed by a
 * train-mode BatchNorm2d (eps 1e-5, momentum 0.1) over the rows [M][C] of the GEMM output.
 *   bsed_bn_rows_train: batch statistics -> running statistics (unbiased variance) and num_batches_tracked -> x becomes
 *       xhat in place, y = gamma*xhat + beta [+ residual] [ReLU]; mean_rstd [2][C] is kept for the backward.
 *   bsed_bn_rows_backward: dy (in: gradient w.r.t. y; out: gradient w.r.t. the convolution output); y != NULL applies the
 *       ReLU mask first; d_residual (may be NULL) receives the masked gradient (the identity branch); d_gamma / d_beta +=.
 *   bsed_col2im_nhwc: transpose of bsed_im2col_nhwc (gather, fixed summation order), dx = or += (accumulate).
 *   bsed_maxpool_nhwc_backward: gradient to the first maximum of each window (ATen's rule).
 *   bsed_sigmoid_rows_backward: d_logits [rows][ld] = d_p * p (1-p), padding columns zero.
 * workspace: bsed_bn_rows_workspace_bytes(C).  Weight gradients and data gradients of the convolutions are GEMMs
 * (bsed_gemm_tn on the recomputed im2col matrix, bsed_gemm_nn with the packed weights).  C % 4 == 0, C <= 1024. */
size_t bsed_bn_rows_workspace_bytes(int C);
int bsed_bn_rows_train(bsed_handle h, float* x, int64_t M, int C, const float* gamma, const float* beta, float eps,
                       float momentum, float* run_mean, float* run_var, int64_t* nbt, const float* residual, int relu,
                       float* y, float* mean_rstd, void* workspace, size_t workspace_bytes, void* stream);
int bsed_bn_rows_backward(bsed_handle h, float* dy, const float* y, const float* xhat, int64_t M, int C,
                          const float* gamma, const float* mean_rstd, float* d_gamma, float* d_beta, float* d_residual,
                          void* workspace, size_t workspace_bytes, void* stream);
int bsed_col2im_nhwc(bsed_handle h, const float* dcol, float* dx, int B, int H, int W, int Cin, int kh, int kw, int sh,
                     int sw, int ph, int pw, int Ho, int Wo, int Kpad, int accumulate, void* stream);
int bsed_maxpool_nhwc_backward(bsed_handle h, const float* x, const float* dy, float* dx, int B, int H, int W, int C, int k,
                               int s, int p, int Ho, int Wo, void* stream);
int bsed_avgpool_nhwc_backward(bsed_handle h, const float* dy, float* dx, int B, int HW, int C, void* stream);
int bsed_sigmoid_rows_backward(bsed_handle h, const float* p, const float* dp, float* d_logits, int rows, int C, int ld,
                               void* stream);

/* ------------------------------------------------------------------------------------------
 * Measurement hooks (bench.py).
 *   bsed_launch_count: kernels this library has launched in this process.
 *   bsed_profile_begin(cls) .. bsed_profile_end: CUDA-event time, summed over the launches of one
 *   kernel class on their launching stream, with the algorithmic flops / bytes of those launches.
 *   classes: 1 implicit-GEMM conv (fwd + dgrad), 2 conv weight gradient, 3 plain GEMM, 4 split-K
 *   reduction GEMM, 5 log-mel frontend (STFT+mel), 6 GRU recurrence.
 * ------------------------------------------------------------------------------------------ */
uint64_t bsed_launch_count(void);
int bsed_profile_begin(int kernel_class);
int bsed_profile_end(double* total_ms, double* total_flops, double* total_bytes, int* n_launches);

/* ------------------------------------------------------------------------------------------
 * Generic kernels exported for unit tests (row-major fp32).
 *   gemm_nn: C[M][N] (ldc) = A[M][K] (lda) * Bm[K][N] (ldb) (+ bias[N]) (+ C if accumulate)
 *            K % 16 == 0, N % 16 == 0
 *   gemm_tn: C[M][N] += sum_k A[k][M] * Bm[k][N]   (C must be initialised; split-K atomics)
 *            M % 16 == 0, N % 16 == 0
 *   conv3x3: channels-last 3x3 / stride 1 / pad 1, weight in the reference's (Cout,Cin,3,3) layout,
 *            y [B][T][F][Cout]; wpack scratch of 9*Cin*Cout floats.
 * ------------------------------------------------------------------------------------------ */
int bsed_gemm_nn(bsed_handle h, const float* A, int lda, const float* Bm, int ldb, float* C, int ldc,
                 int M, int N, int K, const float* bias, int accumulate, void* stream);
int bsed_gemm_tn(bsed_handle h, const float* A, int lda, const float* Bm, int ldb, float* C, int ldc,
                 int M, int N, int K, void* stream);
int bsed_conv3x3(bsed_handle h, const float* x, const float* weight, const float* bias, float* y,
                 int B, int T, int F, int Cin, int Cout, float* wpack, void* stream);
/* tcgen05 (kind::tf32) + TMA variants: conv3x3_tc needs F dividing 128 and Cout <= 128;
 * gemm_nt_tc: C[M][N] (+)= A[M][K] * Bk[N][K]^T (+ bias), N <= 128, K % 16 == 0. */
int bsed_conv3x3_tc(bsed_handle h, const float* x, const float* weight, const float* bias, float* y,
                    int B, int T, int F, int Cin, int Cout, float* wpack, void* stream);
/* dw (Cout,Cin,3,3) += sum_pixels dy[p][co] * x[p + tap][ci]; tensor_cores != 0 needs F dividing 64,
 * Cout % 32 == 0 and a workspace of bsed_conv3x3_wgrad_workspace_bytes(h). */
size_t bsed_conv3x3_wgrad_workspace_bytes(bsed_handle h);
int bsed_conv3x3_wgrad(bsed_handle h, const float* x, const float* dy, float* dw, int B, int T, int F, int Cin,
                       int Cout, int tensor_cores, float* workspace, size_t workspace_bytes, void* stream);
int bsed_gemm_nt_tc(bsed_handle h, const float* A, int lda, const float* Bk, int ldb, float* C, int ldc,
                    int M, int N, int K, const float* bias, int accumulate, void* stream);
/* gemm_tn_tc: C[M][N] += sum_k A[k][M] * Bm[k][N] on tcgen05 (the weight-gradient kernel, rows = k); M % 32 == 0,
 * N % 32 == 0 and, from 128 up, N % 128 == 0; workspace of bsed_conv3x3_wgrad_workspace_bytes(h). */
int bsed_gemm_tn_tc(bsed_handle h, const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int M, int N,
                    int64_t K, float* workspace, size_t workspace_bytes, void* stream);
/* Error-compensated 3xTF32 variants of the two above (BSED_PRECISION_TF32X3: fp32-grade products on the tensor cores).
 * conv3x3_tc3: wpack scratch of 2 * 9*Cin*Cout floats; gemm_nt_tc3: split_ws scratch of 2 * N * ldb floats (the
 * tf32-rounded copy of Bk and its remainders). */
int bsed_conv3x3_tc3(bsed_handle h, const float* x, const float* weight, const float* bias, float* y,
                     int B, int T, int F, int Cin, int Cout, float* wpack, void* stream);
int bsed_gemm_nt_tc3(bsed_handle h, const float* A, int lda, const float* Bk, int ldb, float* C, int ldc,
                     int M, int N, int K, const float* bias, int accumulate, float* split_ws, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BSED_H_ */
