// launch.h -- host-side launcher prototypes shared between the translation units of libbsed.
#pragma once
#include "common.cuh"

namespace bsed {

// per-device one-time kernel attributes (one handle per device; cudaFuncSetAttribute is a per-device setting)
constexpr int kMaxDevices = 64;
static inline bool first_use_on_device(bool (&done)[kMaxDevices]) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) return true;
  const bool first = !done[dev];
  done[dev] = true;
  return first;
}


constexpr int kMaxGroups = 4;

// clip ranges of the forward groups (one group == one reference model call)
struct Groups {
  int n;
  int first[kMaxGroups];
  int count[kMaxGroups];
};

// per-group pointers of one BatchNorm layer
struct BNPtrs {
  const float* gamma[kMaxGroups];
  const float* beta[kMaxGroups];
  float* mean[kMaxGroups];   // [C] saved batch (or running) mean
  float* rstd[kMaxGroups];   // [C] 1/sqrt(var + eps)
};

struct FloatPtrs {
  const float* p[kMaxGroups];
};

// ---- gemm.cu
int gemm_nn(const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int M, int N, int K,
            const float* bias, int accumulate, cudaStream_t st);
int conv3x3_nn(const float* X, const float* Wp, float* Y, int B, int T, int F, int Cin, int Cout,
               const float* bias, int accumulate, cudaStream_t st);
int gemm_tn(const float* A, int lda, const float* Bm, int ldb, float* C, long long rs, long long cs, int M,
            int N, long long K, int target_ctas, cudaStream_t st);
int conv3x3_wgrad(const float* X, const float* dY, float* dW, int B, int T, int F, int Cin, int Cout,
                  int target_ctas, cudaStream_t st);
int gru_whh_grad(const float* dG, int ldg, const float* H, int ldh, int dt, float* dW, int T, long long BT,
                 int target_ctas, cudaStream_t st);

// ---- cnn_ops.cu
int conv0_fwd(const float* x, const Groups& g, const FloatPtrs& w, const FloatPtrs& bias, float* y, int T,
              int F, int Cout, double* stats, int num_sms, cudaStream_t st);
int conv0_wgrad(const float* x, const float* dY, float* dW, int first_clip, int n_clips, int T, int F,
                int Cout, int num_sms, cudaStream_t st);
// mode 0: (sum a, sum a^2); mode 1: (sum a, sum a*b); out: double [group][C][2], accumulated
int col_stats(const float* a, const float* b, int mode, const Groups& g, long long rows_per_clip, int C,
              double* out, int num_sms, cudaStream_t st);
int bn_finalize_train(const double* stats, const Groups& g, long long rows_per_clip, int C, float eps,
                      float momentum, const BNPtrs& bn, float* const* run_mean, float* const* run_var,
                      int64_t* const* nbt, cudaStream_t st);
int bn_running_update2(const double* statsA, long long rowsA, const double* statsB, long long rowsB, const Groups& g, int C,
                       float momentum, float* const* run_mean, float* const* run_var, int64_t* const* nbt,
                       cudaStream_t st);
int bn_prepare_eval(const Groups& g, int C, float eps, const BNPtrs& bn, float* const* run_mean,
                    float* const* run_var, cudaStream_t st);
int bn_normalize(float* y, const Groups& g, long long rows_per_clip, int C, const BNPtrs& bn,
                 cudaStream_t st);
int glu_gate_pool_fwd(const float* xhat, const float* lin, float* pooled, const Groups& g, const BNPtrs& bn,
                      int T, int F, int C, int pt, int pf, DropKey key, uint32_t thresh, float inv_keep,
                      cudaStream_t st);
int glu_gate_pool_bwd(const float* xhat, float* lin_dlin, const float* dpooled, float* dxn, const Groups& g,
                      const BNPtrs& bn, int T, int F, int C, int pt, int pf, DropKey key, uint32_t thresh,
                      float inv_keep, cudaStream_t st);
int glu_gate_pool_bwd_sums(const float* xhat, float* lin_dlin, const float* dpooled, float* dxn, const Groups& g,
                           const BNPtrs& bn, int T, int F, int C, int pt, int pf, DropKey key, uint32_t thresh,
                           float inv_keep, double* sums, int num_sms, cudaStream_t st);
int bn_bwd_prepare(const double* sums, const float* G, int n_groups, int C, const Groups& g, long long rows_per_clip,
                   const BNPtrs& bn, const float* wg, const float* gamma, const float* beta, int pack, float* tab,
                   float* d_gamma, float* d_beta, float* d_wg, float* d_bg, cudaStream_t st);
int bn_bwd_apply(float* dxn_dy, const float* xhat, const double* stats2, const Groups& g,
                 long long rows_per_clip, int C, const BNPtrs& bn, cudaStream_t st);
// dgamma += sum_g s2, dbeta += sum_g s1; dWg += gamma[c] * G[c'][c] + beta[c] * dbg[c']; dbg_out += dbg
int bn_glu_param_grads(const double* stats2, int n_groups, int C, const float* gamma, const float* beta,
                       const float* G, const double* dbg, float* d_gamma, float* d_beta, float* d_wg,
                       float* d_bg, cudaStream_t st);
// out[c] += sum_rows a[row][c]  (double accumulators [C][2], second slot unused) -> float add
int col_sum_to(const float* a, long long rows, int C, float* out, double* scratch, int num_sms,
               cudaStream_t st);
int add_double_to_float(const double* src, int stride, float* dst, int n, cudaStream_t st);

// weight preparation (packing / transposes / BN folding), see prep.cu
struct PrepOp {
  int type;
  int d0, d1, d2, d3;
  const float* src;
  const float* aux0;
  const float* aux1;
  const float* aux2;
  float* dst;
  float* dst2;
};
enum { PREP_CONV_PACK = 0, PREP_CONV_PACK_FLIP = 1, PREP_GLU_FOLD = 2, PREP_TRANSPOSE = 3, PREP_COPY = 4, PREP_ZERO = 5, PREP_ZERO_COLS = 6, PREP_CONV_KMAJOR = 7, PREP_CONV_KMAJOR_FLIP = 8, PREP_TRANSPOSE_BD = 9, PREP_CONV_PAIR = 10, PREP_GATE_TAB = 11 };
constexpr int kMaxPrepOps = 64;
struct PrepTable {
  int n;
  PrepOp ops[kMaxPrepOps];
};
int run_prep(const PrepTable& table, cudaStream_t st);
// 3xTF32 weight split of packed GEMM operands: ranges [off, off + n) of `packed` become their tf32 roundings, the
// remainders go to packed + lo_offset
constexpr int kMaxSplitRanges = 64;
struct SplitTable {
  int n;
  long long off[kMaxSplitRanges], len[kMaxSplitRanges];
};
int run_split(const SplitTable& table, float* packed, long long lo_offset, cudaStream_t st);

// ---- gru.cu
// xg [B][T][768] input projections (both directions); whhT per group: [2][128][384]; bhh per group: [2][384]
// out [B][T][256]; enc (optional) = dropout(out); saved (optional) [B][T][2][4][128] = r, z, n, W_hn h + b_hn
int gru_forward(const float* xg, const Groups& g, const FloatPtrs& whhT, const FloatPtrs& bhh, float* out,
                float* enc, float* saved, int T, DropKey key, uint32_t thresh, float inv_keep,
                cudaStream_t st);
// whh: [2][384][128]; dxg / dgh [B][T][768]
int gru_backward(const float* dout, const float* saved, const float* out, const float* whh, float* dxg,
                 float* dgh, int T, int first_clip, int n_clips, cudaStream_t st);

// ---- head.cu
int head_forward(const float* logits, float* strong, float* weak, int B, int T, int C, int ldl,
                 int inference, cudaStream_t st);
int head_backward(const float* logits, const float* strong, const float* weak, const float* d_strong,
                  const float* d_weak, float* d_logits, int first_clip, int n_clips, int T, int C, int ldl,
                  cudaStream_t st);
int dropout_bwd_mask(float* d, const float* extra, long long first_elem, long long n, DropKey key,
                     uint32_t thresh, float inv_keep, cudaStream_t st);
int mt_loss(const float* strong, const float* weak, int B, int T, int C, int syn_first, int syn_n,
            const float* syn_target, int real_first, int real_n, const float* strong_ema, const float* weak_ema,
            float cons_w, const bsed_step_state* ss, float* losses, float* d_strong, float* d_weak, cudaStream_t st);
int loss_terms(const float* strong, const float* weak, int B, int T, int C, const bsed_loss_term* terms, int n_terms,
               float* losses, int n_slots, float* d_strong, float* d_weak, cudaStream_t st);
int roll_clips(const float* x, const int* shift_t, const int* shift_f, float* out, int B, int T, int F, cudaStream_t st);
int opt_ema_step(float* params, const float* grads, float* m, float* v, float* ema, long long n,
                 const bsed_opt_cfg* cfg, const bsed_step_state* ss, cudaStream_t st);
int step_state_advance(bsed_step_state* state, const bsed_step_cfg* cfg, cudaStream_t st);
int dp_opt_ema_step(int rank, int world, const float* const* peer_grads, float* const* peer_params, float* const* peer_ema,
                    int* const* peer_flags, long long epoch, float* m, float* v, long long n, const bsed_opt_cfg* cfg,
                    const bsed_step_state* ss, int num_sms, cudaStream_t st);
int ema_buffers(const float* bn_buffers, float* ema_bn_buffers, long long n, const int64_t* nbt, int64_t* ema_nbt,
                int n_nbt, float ema_alpha, int64_t ema_step, const bsed_step_state* ss, cudaStream_t st);
int add_f32(float* dst, const float* src, long long n, cudaStream_t st);
int scale_f32(float* dst, const float* src, long long n, float alpha, cudaStream_t st);
// feature-pyramid merge (src/models/CRNN.py:323-328): cat[b][t][0:256] = a[b][t], cat[b][t][256:512] = bilinear
// (align_corners=True) upsampling of b from Tb to Ta frames; backward splits dcat into da and the transposed upsampling db
int fpn_cat_upsample_fwd(const float* a, const float* b, float* cat, int B, int Ta, int Tb, cudaStream_t st);
int fpn_cat_upsample_bwd(const float* dcat, float* da, float* db, int B, int Ta, int Tb, cudaStream_t st);

// ---- tc_gemm.cu (tcgen05 / TMA)
// Wk_lo / Bk_lo != nullptr select the error-compensated 3xTF32 products (fp32-grade): the hi array holds the tf32-rounded
// weights, the lo array their fp32 remainders (split_hi_lo below); nullptr = plain kind::tf32.
int tc_conv3x3(const float* X, const float* Wk, const float* Wk_lo, float* Y, int B, int T, int F, int Cin, int Cout,
               const float* bias, int accumulate, int sms, cudaStream_t st);
// stats != nullptr: also accumulates the per-channel sum / sum of squares of the output per group
// (stats[(group * Cout + c) * 2 + {0,1}], group = last k with clip >= gfirst[k], clips relative to X)
int tc_conv3x3_stats(const float* X, const float* Wk, const float* Wk_lo, float* Y, int B, int T, int F, int Cin, int Cout,
                     const float* bias, int accumulate, double* stats, int stats_groups, const int* gfirst, int sms,
                     cudaStream_t st);
// column-tiled variant with shared-memory halo reuse (tc_conv.cu); no accumulate mode
bool tc_conv_col_supported(int F, int Cin, int Cout);
// stats_c: number of true channels the statistics fold onto (output column c counts for channel c % stats_c)
int tc_conv3x3_col(const float* X, const float* Wk, const float* Wk_lo, float* Y, int B, int T, int F, int Cin, int Cout,
                   const float* bias, double* stats, int stats_groups, const int* gfirst, int sms, cudaStream_t st,
                   int stats_c = 0);
int tc_gemm_nt(const float* A, int lda, const float* Bk, const float* Bk_lo, int ldb, float* C, int ldc, long long M, int N,
               int K, const float* bias, int accumulate, int sms, cudaStream_t st);
// in place: w[i] <- tf32(w[i]) (round to nearest), lo[i] <- w[i] - tf32(w[i])
int split_hi_lo(float* w, float* lo, long long n, cudaStream_t st);

int tc_glu_gate_fwd(const float* xhat, const float* Wk, const float* bias, const float* tab, float* lin, float* pooled,
                    int B, int T, int F, int C, int pack, int pt, int pf, uint32_t key, uint32_t thresh, float inv_keep,
                    uint32_t drop_base, int store_lin, int sms, cudaStream_t st);
int tc_gemm_nt_bnbwd(const float* A, const float* Bk, const float* Bk_lo, float* C, const float* xhat, long long M, int N,
                     int K, const float* tab, int groups, long long rows_per_clip, const int* gfirst, int sms, cudaStream_t st);
size_t tc_wgrad_workspace_bytes(int sms);
// [rows][ld] fp32 tensor, channels [c0, c0 + C) of every row take part
struct TcOperand {
  const float* p;
  int ld, c0, C;
};
int tc_wgrad_ex(TcOperand A, TcOperand Bm, int Bn, int T, int Fv, int mode, int dt0, float* dW, long long rs,
                long long cs, long long ts, float* part, size_t part_bytes, int sms, cudaStream_t st);
int tc_wgrad(const float* X, const float* dY, float* dW, long long rs, long long cs, long long ts, int B, int T, int F,
             int Cin, int Cout, int ntaps, float* part, size_t part_bytes, int sms, cudaStream_t st);

// ---- frontend.cu
int melspec(bsed_context* h, const float* audio, int B, int n_samples, float* mel, cudaStream_t st);
int logmel(bsed_context* h, const float* audio, int B, int n_samples, int frames, const float* sc_mean, const float* sc_std,
           float* mel, float* out, void* ws, size_t ws_bytes, cudaStream_t st);
int amp_to_db(const float* mel, const float* noise, float snr_db, int B, int t_in, int frames,
              const float* sc_mean, const float* sc_std, float* out, void* ws, size_t ws_bytes, cudaStream_t st);
int median_decode(const float* strong, int B, int T, int C, float threshold, int win, int32_t* events,
                  int max_events, int32_t* n_events, cudaStream_t st);

// channels-last im2col (float4 when Cin % 4 == 0) and its transpose in gather form (resnet.cu; also the discriminator's
// stride-2 convolutions): col[(b*Ho+ho)*Wo+wo][(ky*kw+kx)*Cin+ci], columns [kh*kw*Cin, Kpad) zero
int im2col_nhwc(const float* x, float* col, int B, int H, int W, int Cin, int kh, int kw, int sh, int sw, int ph, int pw,
                int Ho, int Wo, int Kpad, cudaStream_t st);
int col2im_nhwc(const float* dcol, float* dx, int B, int H, int W, int Cin, int kh, int kw, int sh, int sw, int ph, int pw,
                int Ho, int Wo, int Kpad, int accumulate, cudaStream_t st);

}  // namespace bsed
