// engine.cu -- CRNN + Predictor plan: parameter layout, workspace carving, forward / backward
// orchestration over the kernels of this library.
//
// Reference call structure: CRNN.forward (src/models/CRNN.py:211-240) = 7 x gated conv block
// (src/models/CNN.py:43-67) -> BiGRU x2 (src/models/RNN.py) -> dropout; Predictor.forward
// (src/models/CRNN.py:559-577).  One "group" is one reference model call (its own BatchNorm batch
// statistics); groups that share a parameter buffer are batched in the same GEMM launches.
#include <stdlib.h>

#include <string>
#include <vector>

#include "launch.h"

using namespace bsed;

namespace {

struct LayerGeom {
  int Cin, Cout, T, F, pt, pf, To, Fo;
  long long rows, prows;  // T*F and To*Fo per clip
};

// A "block" is one Conv3x3 -> BatchNorm -> GLU -> dropout -> AvgPool application: the n_cnn trunk blocks, then (fpn) the
// two applications of the shared-weight stage cnn_fcn / bn_fcn / glu (src/models/CNN_FPN.py:85-98), which are blocks
// n_cnn and n_cnn + 1 with identical parameter / packed-operand / running-stat offsets and their own activations.
constexpr int kMaxBlocks = BSED_MAX_CNN_LAYERS + 2;
// A "stack" is one BidirectionalGRU: rnn, and (fpn) rnn_2, rnn_4 on the 156- and 78-frame scales.
constexpr int kMaxStacks = 3;

struct ParamLayout {
  long long conv_w[kMaxBlocks], conv_b[kMaxBlocks], bn_w[kMaxBlocks], bn_b[kMaxBlocks], glu_w[kMaxBlocks], glu_b[kMaxBlocks];
  long long wih[kMaxStacks][4][2], whh[kMaxStacks][4][2], bih[kMaxStacks][4][2], bhh[kMaxStacks][4][2];
  long long c1_w, c1_b;        // cnn.conv1x1 of CNN_FPN: registered by the reference, never used in forward
  long long m_w[2], m_b[2];    // [0] conv1x1_2 (156-frame merge), [1] conv1x1_4 (313-frame merge): [256][512], [256]
  long long total;
  long long dense_w, dense_b, sm_w, sm_b, pred_total;  // Predictor: its own flat buffer
  std::vector<long long> pred_order;
  long long rm[kMaxBlocks], rv[kMaxBlocks], bn_total;
  std::vector<long long> order;
};

struct PackedLayout {
  long long wp[kMaxBlocks], wd[kMaxBlocks], glu_wT[kMaxBlocks], glu_bf[kMaxBlocks], glu_wgT[kMaxBlocks], wpair[kMaxBlocks],
      bpair[kMaxBlocks], gate_tab[kMaxBlocks];
  long long wihT[kMaxStacks][4], bih[kMaxStacks][4], whhT[kMaxStacks][4], whh[kMaxStacks][4], bhh[kMaxStacks][4],
      wih_cat[kMaxStacks][4];
  long long m_w[2], m_wT[2];   // merge convolutions: [256][512] copy and its [512][256] transpose
  long long total;
  long long wcatT, bcat, wcat, pred_total;  // Predictor operands (own region)
};

// pixels packed per GEMM row for the GLU linears of narrow blocks (tensor-core path): 16 / 32 channels are viewed as
// rows of 64 floats with block-diagonal weights -- 64-byte TMA rows move at a fraction of the 128-byte row rate
inline int glu_pack(int C) { return C == 16 ? 4 : C == 32 ? 2 : 1; }

// forward conv of a 16-input-channel block on the column-tiled kernel: two pixels per row (32 floats in, 2*Cout out)
inline bool conv_pair_ok(const LayerGeom& g) {
  return g.Cin == 16 && g.F % 2 == 0 && 2 * g.Cout <= 128 && tc_conv_col_supported(g.F / 2, 32, 2 * g.Cout);
}

// blocks whose GLU linear runs with the gate / dropout / pool fused into the GEMM epilogue (tc_glu_gate_fwd)
inline bool glu_fused(const LayerGeom& g) {
  const int pack = glu_pack(g.Cout), Fp = g.F / pack;
  return g.Cout * pack == 64 && g.F % pack == 0 && Fp <= 128 && 128 % Fp == 0 && (128 / Fp) % g.pt == 0 && g.F % g.pf == 0 &&
         getenv("BSED_GLU_FUSED");   // opt-in: with 4 epilogue warps the fused gate math makes the kernel epilogue-bound
                                     // (7.9 ms vs 6.8 ms per step on B200), so the separate gate kernel stays the default
}

// 3xTF32 covers every forward contraction (the parity grade of the probabilities) and the data gradients of the linear
// layers (GLU, GRU input projections, fpn merges: nn.Linear / 1x1 convolutions are fp32 on the reference's GPU path).
// The data gradient of the 3x3 convolutions stays single-pass tf32 -- what cuDNN gives the reference
// (torch.backends.cudnn.allow_tf32) -- unless BSED_X3_DGRAD=1: with 16 ... 64 output columns its three-fold MMA count
// costs 0.6 ms per step for gradient digits the weight-gradient reductions (single-pass as well) do not keep.
inline bool x3_conv_dgrad() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BSED_X3_DGRAD");
    v = e && atoi(e) ? 1 : 0;
  }
  return v == 1;
}

constexpr int kLdl = 48;  // padded logits row: [0,20) dense, [20,40) dense_softmax, rest zero

}  // namespace

struct bsed_crnn_plan {
  bsed_context* ctx;
  bsed_crnn_cfg cfg;
  int max_clips;
  int Tout;
  int n_blocks;                 // n_cnn (+ 2 with fpn)
  int bn_slot[kMaxBlocks];      // index of a block's BatchNorm in num_batches_tracked
  int n_stacks;                 // 1 (+ 2 with fpn)
  int stackT[kMaxStacks];       // frames of each GRU stack: 313, 156, 78
  int stack_src[kMaxStacks];    // block whose pooled output feeds the stack
  LayerGeom L[kMaxBlocks];
  ParamLayout pl;
  PackedLayout pk;
  // workspace offsets (bytes)
  size_t off_packed[2];
  size_t off_xhat[kMaxBlocks], off_lin[kMaxBlocks], off_pool[kMaxBlocks];
  size_t off_stats, off_stats2, off_meanrstd;
  // per BRANCH scratch: branch 0 runs on the caller's stream; with fpn, rnn and rnn_2 (forward and backward) run on two
  // plan-owned streams forked from / joined to it, so the three independent recurrences overlap
  size_t off_xg[kMaxStacks], off_gru_out[kMaxStacks][4], off_gru_saved[kMaxStacks][4], off_enc[kMaxStacks];
  size_t off_dxn, off_dpool[2], off_denc[kMaxStacks], off_dx1[kMaxStacks], off_dxg[kMaxStacks], off_dgh[kMaxStacks];
  size_t off_dx0[kMaxStacks];   // aux branches: gradient w.r.t. the stack input, added to the block gradient after the join
  int n_branches;
  cudaStream_t aux[2];
  cudaEvent_t ev_fork[2], ev_join[2];
  // conv weight gradients are off the backward critical path (only the optimiser needs them): with BSED_WGRAD_ASIDE=1
  // they run on a plan-owned side stream, reading a double-buffered dY, while the caller's stream continues with the
  // data gradient and the next block's gate / BatchNorm kernels.  Opt-in: both sides fill the GPU, the step gains 0.6 %
  // (CRNN) to 2 % (CRNN_fpn), and kernels timed on the caller's stream slow down under the sharing
  cudaStream_t side;
  cudaEvent_t ev_dy[2], ev_wg[2];
  size_t off_dxn2, off_wgpart_side;
  size_t off_cat[2], off_y2, off_dcat, off_dy2;   // fpn merge: cat[0] (B,156,512), cat[1] (B,313,512), y2 (B,156,256)
  size_t off_G, off_dscratch[kMaxStacks], off_wgpart[kMaxStacks], off_bsums, off_bntab;
  size_t wgpart_bytes;
  size_t ws_bytes;
  int precision;  // BSED_PRECISION_FP32 (SIMT fp32), BSED_PRECISION_TF32 (tcgen05 kind::tf32), BSED_PRECISION_TF32X3
  // state of the last forward
  bool saved_valid;
  int n_groups, B;
  Groups groups;
  const float* gparams[kMaxGroups];
  int gpset[kMaxGroups];
  const float* pset_params[2];
  int n_psets;
  const float* x_in;
  DropKey bkeys[kMaxBlocks];    // dropout keys of the blocks (by value, or references into the device step state)
  DropKey skeys[kMaxStacks];    // dropout keys of the stack outputs
  uint32_t thresh;              // encoder-output and trunk-block dropout (cfg.dropout)
  float inv_keep;
  uint32_t bthresh[kMaxBlocks]; // per block: the fpn stage drops with p = 0.5 whatever cfg.dropout is
  float binv[kMaxBlocks];       // (CNN_FPN.__init__: self.dropout = nn.Dropout(0.5), src/models/CNN_FPN.py:79)
};

namespace {

size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

int build_layouts(bsed_crnn_plan* p) {
  const bsed_crnn_cfg& c = p->cfg;
  BSED_REQUIRE(c.n_cnn >= 1 && c.n_cnn <= BSED_MAX_CNN_LAYERS, "plan: n_cnn=%d", c.n_cnn);
  BSED_REQUIRE(c.rnn_hidden == 128, "plan: rnn_hidden must be 128 (got %d)", c.rnn_hidden);
  BSED_REQUIRE(c.rnn_layers >= 1 && c.rnn_layers <= 4, "plan: rnn_layers=%d", c.rnn_layers);
  BSED_REQUIRE(c.n_class >= 1 && c.n_class <= 20, "plan: n_class=%d", c.n_class);
  BSED_REQUIRE(c.dropout >= 0.f && c.dropout < 1.f, "plan: dropout=%f", c.dropout);
  int T = c.n_frames, F = c.n_mels, Cin = 1;
  for (int i = 0; i < c.n_cnn; ++i) {
    LayerGeom& g = p->L[i];
    BSED_REQUIRE(c.filters[i] % 16 == 0 && c.filters[i] >= 16 && c.filters[i] <= 128, "plan: filters[%d]=%d", i, c.filters[i]);
    BSED_REQUIRE(c.pool_t[i] >= 1 && c.pool_f[i] >= 1, "plan: pooling[%d]", i);
    g.Cin = Cin;
    g.Cout = c.filters[i];
    g.T = T;
    g.F = F;
    g.pt = c.pool_t[i];
    g.pf = c.pool_f[i];
    g.To = T / g.pt;
    g.Fo = F / g.pf;
    BSED_REQUIRE(g.To >= 1 && g.Fo >= 1, "plan: layer %d pools to nothing", i);
    BSED_REQUIRE(((long long)T * F) % glu_pack(g.Cout) == 0, "plan: layer %d: T*F must be a multiple of %d", i, glu_pack(g.Cout));
    g.rows = (long long)T * F;
    g.prows = (long long)g.To * g.Fo;
    T = g.To;
    F = g.Fo;
    Cin = g.Cout;
  }
  BSED_REQUIRE(F == 1, "plan: frequency axis must pool to 1 (got %d)", F);
  BSED_REQUIRE(Cin == 128, "plan: last CNN block must have 128 channels (GRU input), got %d", Cin);
  BSED_REQUIRE(c.fpn == 0 || c.fpn == 1, "plan: fpn=%d", c.fpn);
  p->Tout = T;
  p->n_blocks = c.n_cnn;
  p->n_stacks = 1;
  p->stackT[0] = T;
  p->stack_src[0] = c.n_cnn - 1;
  for (int i = 0; i < c.n_cnn; ++i) p->bn_slot[i] = i;
  if (c.fpn) {
    // CNN_FPN.forward (src/models/CNN_FPN.py:82-100): two applications of cnn_fcn -> bn_fcn -> glu -> dropout -> pool_fcn
    BSED_REQUIRE(c.n_cnn + 2 <= kMaxBlocks, "plan: too many CNN blocks for fpn");
    BSED_REQUIRE(T >= 4, "plan: fpn needs at least 4 output frames (got %d)", T);
    for (int j = 0; j < 2; ++j) {
      LayerGeom& g = p->L[c.n_cnn + j];
      g.Cin = g.Cout = 128;
      g.T = T;
      g.F = 1;
      g.pt = 2;
      g.pf = 1;
      g.To = T / 2;
      g.Fo = 1;
      g.rows = T;
      g.prows = g.To;
      T = g.To;
      p->bn_slot[c.n_cnn + j] = c.n_cnn;
      p->stackT[1 + j] = T;
      p->stack_src[1 + j] = c.n_cnn + j;
    }
    p->n_blocks = c.n_cnn + 2;
    p->n_stacks = 3;
  }
  // flat parameter layout == reference named_parameters() order
  ParamLayout& pl = p->pl;
  long long o = 0;
  auto take = [&](long long n) {
    long long r = o;
    pl.order.push_back(r);
    o += n;
    return r;
  };
  for (int i = 0; i < c.n_cnn; ++i) {
    const LayerGeom& g = p->L[i];
    pl.conv_w[i] = take((long long)g.Cout * g.Cin * 9);
    pl.conv_b[i] = take(g.Cout);
    pl.bn_w[i] = take(g.Cout);
    pl.bn_b[i] = take(g.Cout);
    pl.glu_w[i] = take((long long)g.Cout * g.Cout);
    pl.glu_b[i] = take(g.Cout);
  }
  if (c.fpn) {
    const int i = c.n_cnn;   // registration order of CNN_FPN.__init__: cnn_fcn, glu, bn_fcn, conv1x1
    pl.conv_w[i] = take(128LL * 128 * 9);
    pl.conv_b[i] = take(128);
    pl.glu_w[i] = take(128LL * 128);
    pl.glu_b[i] = take(128);
    pl.bn_w[i] = take(128);
    pl.bn_b[i] = take(128);
    pl.c1_w = take(128LL * 256);
    pl.c1_b = take(128);
    pl.conv_w[i + 1] = pl.conv_w[i];
    pl.conv_b[i + 1] = pl.conv_b[i];
    pl.glu_w[i + 1] = pl.glu_w[i];
    pl.glu_b[i + 1] = pl.glu_b[i];
    pl.bn_w[i + 1] = pl.bn_w[i];
    pl.bn_b[i + 1] = pl.bn_b[i];
  }
  for (int s = 0; s < p->n_stacks; ++s)
    for (int l = 0; l < c.rnn_layers; ++l) {
      int In = l == 0 ? 128 : 256;
      for (int d = 0; d < 2; ++d) {
        pl.wih[s][l][d] = take(384LL * In);
        pl.whh[s][l][d] = take(384LL * 128);
        pl.bih[s][l][d] = take(384);
        pl.bhh[s][l][d] = take(384);
      }
    }
  if (c.fpn)
    for (int j = 0; j < 2; ++j) {
      pl.m_w[j] = take(256LL * 512);
      pl.m_b[j] = take(256);
    }
  pl.total = o;
  {
    long long po = 0;
    auto ptk = [&](long long n) {
      long long r = po;
      pl.pred_order.push_back(r);
      po += n;
      return r;
    };
    pl.dense_w = ptk((long long)c.n_class * 256);
    pl.dense_b = ptk(c.n_class);
    pl.sm_w = ptk((long long)c.n_class * 256);
    pl.sm_b = ptk(c.n_class);
    pl.pred_total = po;
  }
  long long bo = 0;
  for (int i = 0; i < c.n_cnn + (c.fpn ? 1 : 0); ++i) {
    pl.rm[i] = bo;
    bo += p->L[i].Cout;
    pl.rv[i] = bo;
    bo += p->L[i].Cout;
  }
  if (c.fpn) {
    pl.rm[c.n_cnn + 1] = pl.rm[c.n_cnn];
    pl.rv[c.n_cnn + 1] = pl.rv[c.n_cnn];
  }
  pl.bn_total = bo;
  // packed operand layout (per parameter set)
  PackedLayout& pk = p->pk;
  long long q = 0;
  auto ptake = [&](long long n) {
    long long r = q;
    q += (n + 3) / 4 * 4;
    return r;
  };
  for (int i = 0; i < c.n_cnn + (c.fpn ? 1 : 0); ++i) {
    const LayerGeom& g = p->L[i];
    pk.wp[i] = ptake((long long)g.Cout * g.Cin * 9);
    pk.wd[i] = ptake((long long)g.Cout * g.Cin * 9);
    pk.wpair[i] = ptake(g.Cin == 16 ? 36LL * g.Cin * g.Cout : 0);   // pixel-pair conv weights (16-channel input)
    pk.bpair[i] = ptake(g.Cin == 16 ? 2LL * g.Cout : 0);
    const long long cp = (long long)g.Cout * glu_pack(g.Cout);
    pk.glu_wT[i] = ptake(cp * cp);
    pk.glu_bf[i] = ptake(cp);
    pk.glu_wgT[i] = ptake(cp * cp);
    pk.gate_tab[i] = ptake(2 * cp);
  }
  if (c.fpn) {
    const int i = c.n_cnn;
    pk.wp[i + 1] = pk.wp[i];
    pk.wd[i + 1] = pk.wd[i];
    pk.wpair[i + 1] = pk.wpair[i];
    pk.bpair[i + 1] = pk.bpair[i];
    pk.glu_wT[i + 1] = pk.glu_wT[i];
    pk.glu_bf[i + 1] = pk.glu_bf[i];
    pk.glu_wgT[i + 1] = pk.glu_wgT[i];
    pk.gate_tab[i + 1] = pk.gate_tab[i];
  }
  for (int s = 0; s < p->n_stacks; ++s)
    for (int l = 0; l < c.rnn_layers; ++l) {
      int In = l == 0 ? 128 : 256;
      pk.wihT[s][l] = ptake(768LL * In);
      pk.bih[s][l] = ptake(768);
      pk.whhT[s][l] = ptake(2LL * 128 * 384);
      pk.whh[s][l] = ptake(2LL * 384 * 128);
      pk.bhh[s][l] = ptake(768);
      pk.wih_cat[s][l] = ptake(768LL * In);
    }
  if (c.fpn)
    for (int j = 0; j < 2; ++j) {
      pk.m_w[j] = ptake(256LL * 512);
      pk.m_wT[j] = ptake(512LL * 256);
    }
  pk.total = q;
  q = 0;
  pk.wcatT = ptake(256LL * kLdl);
  pk.bcat = ptake(kLdl);
  pk.wcat = ptake((long long)kLdl * 256);
  pk.pred_total = q;
  return BSED_OK;
}

void carve_workspace(bsed_crnn_plan* p) {
  const bsed_crnn_cfg& c = p->cfg;
  const long long Bm = p->max_clips;
  size_t o = 0;
  auto takeb = [&](size_t bytes) {
    size_t r = o;
    o = align_up(o + bytes);
    return r;
  };
  // packed operands of each parameter set, followed by the low parts of the GEMM operands (3xTF32) at + pk.total
  for (int s = 0; s < 2; ++s) p->off_packed[s] = takeb(sizeof(float) * 2 * p->pk.total);
  long long max_full = 0, max_pool = 0;
  for (int i = 0; i < p->n_blocks; ++i) {
    const LayerGeom& g = p->L[i];
    long long full = Bm * g.rows * g.Cout, pool = Bm * g.prows * g.Cout;
    p->off_xhat[i] = takeb(sizeof(float) * full);
    p->off_lin[i] = takeb(sizeof(float) * full);
    p->off_pool[i] = takeb(sizeof(float) * pool);
    if (full > max_full) max_full = full;
    if (pool > max_pool) max_pool = pool;
  }
  p->off_stats = takeb(sizeof(double) * kMaxBlocks * kMaxGroups * 128 * 2);
  p->off_stats2 = takeb(sizeof(double) * kMaxGroups * 128 * 2);
  p->off_meanrstd = takeb(sizeof(float) * kMaxBlocks * kMaxGroups * 128 * 2);
  const long long BT = Bm * p->Tout;   // the longest stack
  p->n_branches = p->n_stacks;
  for (int b = 0; b < p->n_branches; ++b) p->off_xg[b] = takeb(sizeof(float) * BT * 768);
  for (int s = 0; s < p->n_stacks; ++s) {
    const long long BTs = Bm * p->stackT[s];
    for (int l = 0; l < c.rnn_layers; ++l) {
      p->off_gru_out[s][l] = takeb(sizeof(float) * BTs * 256);
      p->off_gru_saved[s][l] = takeb(sizeof(float) * BTs * 2 * 4 * 128);
    }
    p->off_enc[s] = takeb(sizeof(float) * BTs * 256);
    p->off_denc[s] = takeb(sizeof(float) * BTs * 256);
  }
  p->off_dxn = takeb(sizeof(float) * max_full);
  p->off_dxn2 = takeb(sizeof(float) * max_full);
  p->off_dpool[0] = takeb(sizeof(float) * max_pool);
  p->off_dpool[1] = takeb(sizeof(float) * max_pool);
  for (int b = 0; b < p->n_branches; ++b) {
    p->off_dx1[b] = takeb(sizeof(float) * BT * 256);
    p->off_dxg[b] = takeb(sizeof(float) * BT * 768);
    p->off_dgh[b] = takeb(sizeof(float) * BT * 768);
    p->off_dx0[b] = b > 0 ? takeb(sizeof(float) * BT * 128) : 0;
  }
  if (c.fpn) {
    p->off_cat[0] = takeb(sizeof(float) * Bm * p->stackT[1] * 512);
    p->off_cat[1] = takeb(sizeof(float) * BT * 512);
    p->off_y2 = takeb(sizeof(float) * Bm * p->stackT[1] * 256);
    p->off_dcat = takeb(sizeof(float) * BT * 512);
    p->off_dy2 = takeb(sizeof(float) * Bm * p->stackT[1] * 256);
  }
  p->off_G = takeb(sizeof(float) * kMaxGroups * 128 * 128);
  p->off_bsums = takeb(sizeof(double) * kMaxGroups * 128 * 4);
  p->off_bntab = takeb(sizeof(float) * kMaxGroups * 3 * 128);
  p->wgpart_bytes = tc_wgrad_workspace_bytes(p->ctx->num_sms);
  for (int b = 0; b < p->n_branches; ++b) {
    p->off_dscratch[b] = takeb(sizeof(double) * 2 * 768);
    p->off_wgpart[b] = takeb(p->wgpart_bytes);
  }
  p->off_wgpart_side = takeb(p->wgpart_bytes);
  p->ws_bytes = o;
}

template <class T>
T* wsp(void* ws, size_t off) {
  return reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(ws) + off);
}

struct PrepAdder {
  PrepTable* tb;
  void operator()(int type, const float* src, float* dst, int d0, int d1 = 0, int d2 = 0, int d3 = 0,
                  const float* a0 = nullptr, const float* a1 = nullptr, const float* a2 = nullptr,
                  float* dst2 = nullptr) const {
    PrepOp& op = tb->ops[tb->n++];
    op.type = type;
    op.src = src;
    op.dst = dst;
    op.dst2 = dst2;
    op.aux0 = a0;
    op.aux1 = a1;
    op.aux2 = a2;
    op.d0 = d0;
    op.d1 = d1;
    op.d2 = d2;
    op.d3 = d3;
  }
};

// conv / BatchNorm / GLU operands of the trunk blocks and (fpn) of the shared stage, prepared once for both applications
// `sp` collects the packed GEMM-operand ranges the 3xTF32 mode splits into (hi, lo) after the table has run
void build_prep_table(const bsed_crnn_plan* p, const float* params, float* packed, bool need_bwd, PrepTable* tb,
                      SplitTable* sp) {
  const bsed_crnn_cfg& c = p->cfg;
  const ParamLayout& pl = p->pl;
  const PackedLayout& pk = p->pk;
  const bool tc = p->precision != BSED_PRECISION_FP32;
  tb->n = 0;
  PrepAdder add{tb};
  auto split = [&](long long off, long long n) {
    sp->off[sp->n] = off;
    sp->len[sp->n++] = n;
  };
  for (int i = 0; i < c.n_cnn + (c.fpn ? 1 : 0); ++i) {
    const LayerGeom& g = p->L[i];
    const long long wn = (long long)g.Cout * g.Cin * 9;
    if (i > 0) {
      if (tc && conv_pair_ok(g)) {
        add(PREP_CONV_PAIR, params + pl.conv_w[i], packed + pk.wpair[i], g.Cout, g.Cin, 0, 0, params + pl.conv_b[i], nullptr,
            nullptr, packed + pk.bpair[i]);
        split(pk.wpair[i], 36LL * g.Cin * g.Cout);
      }
      add(tc ? PREP_CONV_KMAJOR : PREP_CONV_PACK, params + pl.conv_w[i], packed + pk.wp[i], g.Cout, g.Cin);
      split(pk.wp[i], wn);
      if (need_bwd) {
        add(tc ? PREP_CONV_KMAJOR_FLIP : PREP_CONV_PACK_FLIP, params + pl.conv_w[i], packed + pk.wd[i], g.Cout, g.Cin);
        if (x3_conv_dgrad()) split(pk.wd[i], wn);
      }
    }
    // d1 != 0: K-major folded matrix [c'][c] for the tensor-core GEMM (B operand [N][K])
    const long long cp = (long long)g.Cout * glu_pack(g.Cout);
    add(PREP_GLU_FOLD, params + pl.glu_w[i], packed + pk.glu_wT[i], g.Cout, tc ? 1 : 0, glu_pack(g.Cout), 0,
        params + pl.bn_w[i], params + pl.bn_b[i], params + pl.glu_b[i], packed + pk.glu_bf[i]);
    split(pk.glu_wT[i], cp * cp);
    if (tc) add(PREP_GATE_TAB, params + pl.bn_w[i], packed + pk.gate_tab[i], g.Cout, glu_pack(g.Cout), 0, 0, params + pl.bn_b[i]);
    if (tc && need_bwd) {
      add(PREP_TRANSPOSE_BD, params + pl.glu_w[i], packed + pk.glu_wgT[i], g.Cout, glu_pack(g.Cout));
      split(pk.glu_wgT[i], cp * cp);
    }
  }
}

// recurrent operands of GRU stack s (and, with s == 0 and fpn, the two merge convolutions)
void build_prep_table_stack(const bsed_crnn_plan* p, int s, const float* params, float* packed, bool need_bwd, PrepTable* tb,
                            SplitTable* sp) {
  const bsed_crnn_cfg& c = p->cfg;
  const ParamLayout& pl = p->pl;
  const PackedLayout& pk = p->pk;
  const bool tc = p->precision != BSED_PRECISION_FP32;
  tb->n = 0;
  PrepAdder add{tb};
  auto split = [&](long long off, long long n) {
    sp->off[sp->n] = off;
    sp->len[sp->n++] = n;
  };
  for (int l = 0; l < c.rnn_layers; ++l) {
    int In = l == 0 ? 128 : 256;
    for (int d = 0; d < 2; ++d) {
      // W_ih [384][In] -> wihT [In][768] columns d*384..
      add(PREP_TRANSPOSE, params + pl.wih[s][l][d], packed + pk.wihT[s][l], 384, In, 768, d * 384);
      add(PREP_COPY, params + pl.bih[s][l][d], packed + pk.bih[s][l] + d * 384, 384);
      // W_hh [384][128] -> whhT [d][128][384]
      add(PREP_TRANSPOSE, params + pl.whh[s][l][d], packed + pk.whhT[s][l] + (long long)d * 128 * 384, 384, 128, 384, 0);
      add(PREP_COPY, params + pl.bhh[s][l][d], packed + pk.bhh[s][l] + d * 384, 384);
      if (need_bwd) add(PREP_COPY, params + pl.whh[s][l][d], packed + pk.whh[s][l] + (long long)d * 384 * 128, 384 * 128);
      if (need_bwd || tc) add(PREP_COPY, params + pl.wih[s][l][d], packed + pk.wih_cat[s][l] + (long long)d * 384 * In, 384 * In);
    }
    // tensor-core B operands: wih_cat (forward projection), wihT (its data gradient)
    split(pk.wih_cat[s][l], 768LL * In);
    split(pk.wihT[s][l], 768LL * In);
  }
  if (c.fpn && s == 0)
    for (int j = 0; j < 2; ++j) {
      add(PREP_COPY, params + pl.m_w[j], packed + pk.m_w[j], 256 * 512);
      add(PREP_TRANSPOSE, params + pl.m_w[j], packed + pk.m_wT[j], 256, 512, 256, 0);   // [256][512] -> [512][256]
      split(pk.m_w[j], 256LL * 512);
      split(pk.m_wT[j], 256LL * 512);
    }
}

void build_prep_table_head(const bsed_crnn_plan* p, const float* params, float* packed, PrepTable* tb) {
  const bsed_crnn_cfg& c = p->cfg;
  const ParamLayout& pl = p->pl;
  const PackedLayout& pk = p->pk;
  const int C = c.n_class;
  tb->n = 0;
  auto add = [&](int type, const float* src, float* dst, int d0, int d1, int d2, int d3) {
    PrepOp& op = tb->ops[tb->n++];
    memset(&op, 0, sizeof(op));
    op.type = type;
    op.src = src;
    op.dst = dst;
    op.d0 = d0;
    op.d1 = d1;
    op.d2 = d2;
    op.d3 = d3;
  };
  // zero only the padding (cols/rows 2C..47): the ops of one table run concurrently
  add(PREP_ZERO_COLS, nullptr, packed + pk.wcatT, 256, kLdl, 2 * C, kLdl - 2 * C);
  add(PREP_ZERO, nullptr, packed + pk.bcat + 2 * C, kLdl - 2 * C, 0, 0, 0);
  add(PREP_ZERO, nullptr, packed + pk.wcat + (long long)2 * C * 256, (kLdl - 2 * C) * 256, 0, 0, 0);
  // dense.weight [C][256] -> wcatT [256][48] cols 0..C-1 ; dense_softmax -> cols C..2C-1
  add(PREP_TRANSPOSE, params + pl.dense_w, packed + pk.wcatT, C, 256, kLdl, 0);
  add(PREP_TRANSPOSE, params + pl.sm_w, packed + pk.wcatT, C, 256, kLdl, C);
  add(PREP_COPY, params + pl.dense_b, packed + pk.bcat, C, 0, 0, 0);
  add(PREP_COPY, params + pl.sm_b, packed + pk.bcat + C, C, 0, 0, 0);
  add(PREP_COPY, params + pl.dense_w, packed + pk.wcat, C * 256, 0, 0, 0);
  add(PREP_COPY, params + pl.sm_w, packed + pk.wcat + (long long)C * 256, C * 256, 0, 0, 0);
}

BNPtrs make_bn_ptrs(const bsed_crnn_plan* p, void* ws, int layer, const int* group_ids, int n) {
  BNPtrs bn;
  float* mr = wsp<float>(ws, p->off_meanrstd);
  for (int i = 0; i < kMaxGroups; ++i) {
    int gi = i < n ? group_ids[i] : group_ids[0];
    const float* par = p->gparams[gi];
    bn.gamma[i] = par + p->pl.bn_w[layer];
    bn.beta[i] = par + p->pl.bn_b[layer];
    bn.mean[i] = mr + ((size_t)(layer * kMaxGroups + gi) * 2 + 0) * 128;
    bn.rstd[i] = mr + ((size_t)(layer * kMaxGroups + gi) * 2 + 1) * 128;
  }
  return bn;
}

}  // namespace

// =================================================================================================
// C ABI: plan
// =================================================================================================
extern "C" int bsed_plan_destroy(bsed_plan p);
extern "C" int bsed_plan_create(bsed_handle h, const bsed_crnn_cfg* cfg, int max_clips, bsed_plan* out) {
  BSED_REQUIRE(h && cfg && out, "plan_create: null argument");
  BSED_REQUIRE(max_clips >= 1 && max_clips <= 4096, "plan_create: max_clips=%d", max_clips);
  bsed_crnn_plan* p = new bsed_crnn_plan();
  p->ctx = h;
  p->cfg = *cfg;
  p->max_clips = max_clips;
  p->saved_valid = false;
  p->precision = BSED_PRECISION_TF32X3;
  int r = build_layouts(p);
  if (r != BSED_OK) {
    delete p;
    return r;
  }
  carve_workspace(p);
  for (int i = 0; i < 2; ++i) {
    p->aux[i] = nullptr;
    p->ev_fork[i] = p->ev_join[i] = p->ev_dy[i] = p->ev_wg[i] = nullptr;
  }
  p->side = nullptr;
  if (getenv("BSED_WGRAD_ASIDE")) {
    bool ok = cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i)
      ok = cudaEventCreateWithFlags(&p->ev_dy[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&p->ev_wg[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      bsed_set_error("plan_create: cannot create the weight-gradient side stream: %s", cudaGetErrorString(cudaGetLastError()));
      bsed_plan_destroy(p);
      return BSED_E_CUDA;
    }
  }
  if (p->n_branches > 1 && !getenv("BSED_FPN_SERIAL")) {
    for (int i = 0; i < 2; ++i) {
      if (cudaStreamCreateWithFlags(&p->aux[i], cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&p->ev_fork[i], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&p->ev_join[i], cudaEventDisableTiming) != cudaSuccess) {
        bsed_set_error("plan_create: cannot create the fpn branch streams: %s", cudaGetErrorString(cudaGetLastError()));
        bsed_plan_destroy(p);
        return BSED_E_CUDA;
      }
    }
  }
  *out = p;
  return BSED_OK;
}

extern "C" int bsed_plan_destroy(bsed_plan p) {
  if (!p) return BSED_OK;
  for (int i = 0; i < 2; ++i) {
    if (p->aux[i]) cudaStreamDestroy(p->aux[i]);
    if (p->ev_fork[i]) cudaEventDestroy(p->ev_fork[i]);
    if (p->ev_join[i]) cudaEventDestroy(p->ev_join[i]);
    if (p->ev_dy[i]) cudaEventDestroy(p->ev_dy[i]);
    if (p->ev_wg[i]) cudaEventDestroy(p->ev_wg[i]);
  }
  if (p->side) cudaStreamDestroy(p->side);
  delete p;
  return BSED_OK;
}

extern "C" int bsed_plan_set_precision(bsed_plan p, int precision) {
  BSED_REQUIRE(p, "plan_set_precision: null plan");
  BSED_REQUIRE(precision == BSED_PRECISION_FP32 || precision == BSED_PRECISION_TF32 || precision == BSED_PRECISION_TF32X3,
               "plan_set_precision: mode %d", precision);
  p->precision = precision;
  p->saved_valid = false;
  return BSED_OK;
}
extern "C" int bsed_plan_get_precision(bsed_plan p) { return p ? p->precision : -1; }

extern "C" int64_t bsed_plan_param_count(bsed_plan p) { return p ? p->pl.total : -1; }
extern "C" int64_t bsed_plan_bn_buffer_count(bsed_plan p) { return p ? p->pl.bn_total : -1; }
extern "C" int bsed_plan_out_frames(bsed_plan p) { return p ? p->Tout : -1; }
extern "C" size_t bsed_plan_workspace_bytes(bsed_plan p) { return p ? p->ws_bytes : 0; }
extern "C" int bsed_plan_param_offsets(bsed_plan p, int64_t* offsets, int max_n) {
  if (!p) return -1;
  int n = (int)p->pl.order.size();
  for (int i = 0; i < n && i < max_n; ++i) offsets[i] = p->pl.order[i];
  return n;
}

// =================================================================================================
// forward
// =================================================================================================
extern "C" int bsed_crnn_forward(bsed_plan p, const bsed_group* groups, int n_groups, const float* x, int B,
                                 int flags, uint64_t dropout_seed, uint64_t dropout_step, float* enc,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  BSED_REQUIRE(p && groups && x && workspace && enc, "crnn_forward: null argument");
  BSED_REQUIRE(n_groups >= 1 && n_groups <= kMaxGroups, "crnn_forward: n_groups=%d (max %d)", n_groups, kMaxGroups);
  BSED_REQUIRE(B >= 1 && B <= p->max_clips, "crnn_forward: B=%d exceeds plan max_clips=%d", B, p->max_clips);
  if (workspace_bytes < p->ws_bytes) {
    bsed_set_error("crnn_forward: workspace %zu < %zu", workspace_bytes, p->ws_bytes);
    return BSED_E_WORKSPACE;
  }
  const bool train = flags & BSED_F_TRAIN;
  const bool save = flags & BSED_F_SAVE;
  BSED_REQUIRE(!save || train, "crnn_forward: BSED_F_SAVE requires BSED_F_TRAIN");
  cudaStream_t st = as_stream(stream);
  const bsed_crnn_cfg& c = p->cfg;
  void* ws = workspace;
  const bool tc = p->precision != BSED_PRECISION_FP32;
  const bool x3 = p->precision == BSED_PRECISION_TF32X3;
  const long long lo_off = p->pk.total;   // low parts of the packed GEMM operands (3xTF32)
  auto LO = [&](const float* w) -> const float* { return x3 ? w + lo_off : nullptr; };
  const int sms = p->ctx->num_sms;

  // groups must tile [0, B)
  Groups g;
  g.n = n_groups;
  int next = 0;
  p->n_psets = 0;
  for (int i = 0; i < kMaxGroups; ++i) {
    g.first[i] = 0;
    g.count[i] = 0;
  }
  for (int i = 0; i < n_groups; ++i) {
    BSED_REQUIRE(groups[i].first_clip == next && groups[i].n_clips >= 1, "crnn_forward: groups must tile [0,B) in order");
    BSED_REQUIRE(groups[i].params && groups[i].bn_buffers, "crnn_forward: group %d has null buffers", i);
    g.first[i] = groups[i].first_clip;
    g.count[i] = groups[i].n_clips;
    next += groups[i].n_clips;
    p->gparams[i] = groups[i].params;
    int ps = -1;
    for (int s = 0; s < p->n_psets; ++s)
      if (p->pset_params[s] == groups[i].params) ps = s;
    if (ps < 0) {
      BSED_REQUIRE(p->n_psets < 2, "crnn_forward: at most 2 distinct parameter buffers per call");
      ps = p->n_psets++;
      p->pset_params[ps] = groups[i].params;
    }
    p->gpset[i] = ps;
    if (i > 0)
      BSED_REQUIRE(p->gpset[i] >= p->gpset[i - 1], "crnn_forward: groups sharing a parameter buffer must be adjacent");
  }
  BSED_REQUIRE(next == B, "crnn_forward: groups cover %d clips, B=%d", next, B);
  p->groups = g;
  p->n_groups = n_groups;
  p->B = B;
  p->x_in = x;
  p->saved_valid = false;
  p->thresh = train ? bsed_drop_thresh(c.dropout) : 0u;
  p->inv_keep = train && c.dropout > 0.f ? 1.0f / (1.0f - c.dropout) : 1.0f;
  for (int i = 0; i < p->n_blocks; ++i) {
    const bool stage = i >= c.n_cnn;
    p->bthresh[i] = stage ? (train ? bsed_drop_thresh(0.5f) : 0u) : p->thresh;
    p->binv[i] = stage ? (train ? 2.0f : 1.0f) : p->inv_keep;
  }
  // dropout streams (restated in oracle/crnn.py): trunk block i -> i, encoder output -> 7, fpn stage applications -> 8, 9,
  // rnn_2 / rnn_4 outputs -> 10, 11
  // with a device-resident step state installed (bsed_set_step_state) the keys are read from it at run time, so that a
  // captured graph of the iteration draws fresh masks on every replay
  const bsed_step_state* ss = p->ctx->step_state;
  auto drop_key = [&](int stream) {
    DropKey k;
    k.key = ss ? 0u : bsed_mix_key(dropout_seed, dropout_step, stream);
    k.dev = ss ? ss->keys + stream : nullptr;
    return k;
  };
  for (int i = 0; i < p->n_blocks; ++i) p->bkeys[i] = drop_key(i < c.n_cnn ? i : 8 + (i - c.n_cnn));
  for (int s = 0; s < p->n_stacks; ++s) p->skeys[s] = drop_key(s == 0 ? 7 : 9 + s);

  // runs of clips sharing a parameter set
  struct Run {
    int first, count, pset;
  } runs[2];
  int n_runs = 0;
  for (int i = 0; i < n_groups; ++i) {
    if (n_runs && runs[n_runs - 1].pset == p->gpset[i]) runs[n_runs - 1].count += g.count[i];
    else runs[n_runs++] = Run{g.first[i], g.count[i], p->gpset[i]};
  }

  // operand preparation
  for (int s = 0; s < p->n_psets; ++s) {
    PrepTable tb;
    SplitTable sp;
    sp.n = 0;
    float* packed = wsp<float>(ws, p->off_packed[s]);
    build_prep_table(p, p->pset_params[s], packed, save, &tb, &sp);
    BSED_TRY(run_prep(tb, st));
    for (int k = 0; k < p->n_stacks; ++k) {
      build_prep_table_stack(p, k, p->pset_params[s], packed, save, &tb, &sp);
      BSED_TRY(run_prep(tb, st));
    }
    if (x3) BSED_TRY(run_split(sp, packed, lo_off, st));
  }

  // GRU stacks: rnn on the trunk output, (fpn) rnn_2 / rnn_4 on the two coarser scales; each followed by dropout.
  // branch of a stack: its scratch set and its stream (0 = the caller's stream)
  const bool forked = p->n_stacks > 1 && p->aux[0] != nullptr;
  auto branch_of = [&](int s) { return p->n_stacks == 1 ? 0 : (s == p->n_stacks - 1 ? 0 : s + 1); };
  auto stream_of = [&](int b) { return b == 0 || !forked ? st : p->aux[b - 1]; };
  auto stack_forward = [&](int s) -> int {
    const int T = p->stackT[s];
    const int br = branch_of(s);
    cudaStream_t ss = stream_of(br);
    float* xg = wsp<float>(ws, p->off_xg[br]);
    for (int l = 0; l < c.rnn_layers; ++l) {
      const int In = l == 0 ? 128 : 256;
      const float* X = l == 0 ? wsp<float>(ws, p->off_pool[p->stack_src[s]]) : wsp<float>(ws, p->off_gru_out[s][l - 1]);
      for (int r = 0; r < n_runs; ++r) {
        const float* packed = wsp<float>(ws, p->off_packed[runs[r].pset]);
        const float* Xr = X + (size_t)runs[r].first * T * In;
        float* xgr = xg + (size_t)runs[r].first * T * 768;
        if (tc) {
          // B operand = [W_ih ; W_ih_reverse], K-major as stored; the six 128-column blocks of the 768 gate
          // pre-activations in one launch
          BSED_TRY(tc_gemm_nt(Xr, In, packed + p->pk.wih_cat[s][l], LO(packed + p->pk.wih_cat[s][l]), In, xgr, 768,
                              (long long)runs[r].count * T, 768, In, packed + p->pk.bih[s][l], 0, sms, ss));
        } else {
          BSED_TRY(gemm_nn(Xr, In, packed + p->pk.wihT[s][l], 768, xgr, 768, runs[r].count * T, 768, In,
                           packed + p->pk.bih[s][l], 0, ss));
        }
      }
      FloatPtrs whhT, bhh;
      for (int k = 0; k < kMaxGroups; ++k) {
        int gi = k < n_groups ? k : 0;
        const float* packed = wsp<float>(ws, p->off_packed[p->gpset[gi]]);
        whhT.p[k] = packed + p->pk.whhT[s][l];
        bhh.p[k] = packed + p->pk.bhh[s][l];
      }
      const bool last = l == c.rnn_layers - 1;
      BSED_TRY(gru_forward(xg, g, whhT, bhh, wsp<float>(ws, p->off_gru_out[s][l]), last ? wsp<float>(ws, p->off_enc[s]) : nullptr,
                           save ? wsp<float>(ws, p->off_gru_saved[s][l]) : nullptr, T, p->skeys[s], p->thresh,
                           p->inv_keep, ss));
    }
    return BSED_OK;
  };
  // fork: a stack may start as soon as the block that feeds it is done
  auto fork_stack = [&](int s) -> int {
    const int br = branch_of(s);
    if (forked && br > 0) {
      BSED_CHECK_CUDA(cudaEventRecord(p->ev_fork[br - 1], st));
      BSED_CHECK_CUDA(cudaStreamWaitEvent(p->aux[br - 1], p->ev_fork[br - 1], 0));
    }
    BSED_TRY(stack_forward(s));
    if (forked && br > 0) BSED_CHECK_CUDA(cudaEventRecord(p->ev_join[br - 1], p->aux[br - 1]));
    return BSED_OK;
  };

  int all_ids[kMaxGroups] = {0, 1, 2, 3};
  double* stats = wsp<double>(ws, p->off_stats);
  if (train) BSED_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * kMaxBlocks * kMaxGroups * 128 * 2, st));

  for (int i = 0; i < p->n_blocks; ++i) {
    const LayerGeom& L = p->L[i];
    float* y = wsp<float>(ws, p->off_xhat[i]);
    float* lin = wsp<float>(ws, p->off_lin[i]);
    float* pool = wsp<float>(ws, p->off_pool[i]);
    if (i == 0) {
      FloatPtrs w, bias;
      for (int k = 0; k < kMaxGroups; ++k) {
        int gi = k < n_groups ? k : 0;
        w.p[k] = p->gparams[gi] + p->pl.conv_w[0];
        bias.p[k] = p->gparams[gi] + p->pl.conv_b[0];
      }
      BSED_TRY(conv0_fwd(x, g, w, bias, y, L.T, L.F, L.Cout, train ? stats : nullptr, sms, st));   // + batch statistics
    } else {
      const float* xin = wsp<float>(ws, p->off_pool[i - 1]);
      for (int r = 0; r < n_runs; ++r) {
        const float* packed = wsp<float>(ws, p->off_packed[runs[r].pset]);
        const float* xr = xin + (size_t)runs[r].first * L.rows * L.Cin;
        float* yr = y + (size_t)runs[r].first * L.rows * L.Cout;
        const float* cb = p->pset_params[runs[r].pset] + p->pl.conv_b[i];
        if (tc) {
          // batch statistics of the conv output are accumulated by the GEMM epilogue (train mode)
          int gfirst_rel[kMaxGroups] = {0, 0, 0, 0}, gbase = -1, ng = 0;
          for (int k = 0; k < n_groups; ++k)
            if (p->gpset[k] == runs[r].pset) {
              if (gbase < 0) gbase = k;
              gfirst_rel[ng++] = g.first[k] - runs[r].first;
            }
          double* st_run = train ? stats + (size_t)i * kMaxGroups * 128 * 2 + (size_t)gbase * L.Cout * 2 : nullptr;
          if (conv_pair_ok(L))
            BSED_TRY(tc_conv3x3_col(xr, packed + p->pk.wpair[i], LO(packed + p->pk.wpair[i]), yr, runs[r].count, L.T, L.F / 2, 32,
                                    2 * L.Cout, packed + p->pk.bpair[i], st_run, ng, gfirst_rel, sms, st, L.Cout));
          else
            BSED_TRY(tc_conv3x3_stats(xr, packed + p->pk.wp[i], LO(packed + p->pk.wp[i]), yr, runs[r].count, L.T, L.F, L.Cin,
                                      L.Cout, cb, 0, st_run, ng, gfirst_rel, sms, st));
        } else {
          BSED_TRY(conv3x3_nn(xr, packed + p->pk.wp[i], yr, runs[r].count, L.T, L.F, L.Cin, L.Cout, cb, 0, st));
        }
      }
    }
    BNPtrs bn = make_bn_ptrs(p, ws, i, all_ids, n_groups);
    float* rmean[kMaxGroups];
    float* rvar[kMaxGroups];
    int64_t* nbt[kMaxGroups];
    for (int k = 0; k < kMaxGroups; ++k) {
      int gi = k < n_groups ? k : 0;
      rmean[k] = groups[gi].bn_buffers + p->pl.rm[i];
      rvar[k] = groups[gi].bn_buffers + p->pl.rv[i];
      nbt[k] = groups[gi].num_batches_tracked ? groups[gi].num_batches_tracked + p->bn_slot[i] : nullptr;
    }
    if (train) {
      double* st_i = stats + (size_t)i * kMaxGroups * 128 * 2;
      if (!tc && i > 0) BSED_TRY(col_stats(y, nullptr, 0, g, L.rows, L.Cout, st_i, p->ctx->num_sms, st));
      // stats rows are indexed [group][C]: col_stats uses stride C, finalize too
      if (i < c.n_cnn) {
        BSED_TRY(bn_finalize_train(st_i, g, L.rows, L.Cout, c.bn_eps, c.bn_momentum, bn, rmean, rvar, nbt, st));
      } else {
        // shared bn_fcn: batch mean / rstd per application now; its running statistics are updated after the second
        // application, group by group, in the reference's order (application 1 then 2 of each model call)
        float* no_f[kMaxGroups] = {nullptr, nullptr, nullptr, nullptr};
        int64_t* no_i[kMaxGroups] = {nullptr, nullptr, nullptr, nullptr};
        BSED_TRY(bn_finalize_train(st_i, g, L.rows, L.Cout, c.bn_eps, c.bn_momentum, bn, no_f, no_f, no_i, st));
        if (i == c.n_cnn + 1)
          BSED_TRY(bn_running_update2(stats + (size_t)c.n_cnn * kMaxGroups * 128 * 2, p->L[c.n_cnn].rows, st_i, L.rows, g,
                                      L.Cout, c.bn_momentum, rmean, rvar, nbt, st));
      }
    } else {
      BSED_TRY(bn_prepare_eval(g, L.Cout, c.bn_eps, bn, rmean, rvar, st));
    }
    BSED_TRY(bn_normalize(y, g, L.rows, L.Cout, bn, st));
    for (int r = 0; r < n_runs; ++r) {
      const float* packed = wsp<float>(ws, p->off_packed[runs[r].pset]);
      size_t off = (size_t)runs[r].first * L.rows * L.Cout;
      long long M = (long long)runs[r].count * L.rows;
      BSED_REQUIRE(M < (1ll << 31), "crnn_forward: too many pixels");
      if (tc) {
        const int pack = glu_pack(L.Cout), CP = L.Cout * pack;   // L.rows is a multiple of 4 (F is even twice over)
        if (!x3 && !ss && glu_fused(L)) {
          // GEMM + gate + dropout + average pool in one kernel; lin is stored only when backward will need it
          BSED_TRY(tc_glu_gate_fwd(y + off, packed + p->pk.glu_wT[i], packed + p->pk.glu_bf[i], packed + p->pk.gate_tab[i],
                                   lin + off, pool + (size_t)runs[r].first * L.prows * L.Cout, runs[r].count, L.T, L.F,
                                   L.Cout, pack, L.pt, L.pf, p->bkeys[i].key, p->bthresh[i], p->binv[i], (uint32_t)off, save ? 1 : 0,
                                   sms, st));
        } else {
          BSED_TRY(tc_gemm_nt(y + off, CP, packed + p->pk.glu_wT[i], LO(packed + p->pk.glu_wT[i]), CP, lin + off, CP, M / pack,
                              CP, CP, packed + p->pk.glu_bf[i], 0, sms, st));
        }
      }
      else
        BSED_TRY(gemm_nn(y + off, L.Cout, packed + p->pk.glu_wT[i], L.Cout, lin + off, L.Cout, (int)M, L.Cout, L.Cout,
                         packed + p->pk.glu_bf[i], 0, st));
    }
    if (!(tc && !x3 && !ss && glu_fused(L)))
      BSED_TRY(glu_gate_pool_fwd(y, lin, pool, g, bn, L.T, L.F, L.Cout, L.pt, L.pf, p->bkeys[i], p->bthresh[i],
                                 p->binv[i], st));
    for (int s = 0; s < p->n_stacks; ++s)
      if (p->stack_src[s] == i) BSED_TRY(fork_stack(s));
  }

  // join the forked stacks before anything reads their outputs
  if (forked)
    for (int i = 0; i < 2; ++i) BSED_CHECK_CUDA(cudaStreamWaitEvent(st, p->ev_join[i], 0));

  if (!c.fpn) {
    BSED_CHECK_CUDA(cudaMemcpyAsync(enc, wsp<float>(ws, p->off_enc[0]), sizeof(float) * (size_t)B * p->Tout * 256,
                                    cudaMemcpyDeviceToDevice, st));
  } else {
    // feature pyramid merge (src/models/CRNN.py:323-328):
    //   x_2 = conv1x1_2(cat(x_2, upsample_4(x_4)))  (156 frames) ; x = conv1x1_4(cat(x, upsample_2(x_2)))  (313 frames)
    const float* lo[2] = {wsp<float>(ws, p->off_enc[2]), wsp<float>(ws, p->off_y2)};
    const float* hi[2] = {wsp<float>(ws, p->off_enc[1]), wsp<float>(ws, p->off_enc[0])};
    float* outs[2] = {wsp<float>(ws, p->off_y2), enc};
    for (int j = 0; j < 2; ++j) {
      const int Ta = p->stackT[1 - j], Tb = p->stackT[2 - j];
      float* cat = wsp<float>(ws, p->off_cat[j]);
      BSED_TRY(fpn_cat_upsample_fwd(hi[j], lo[j], cat, B, Ta, Tb, st));
      for (int r = 0; r < n_runs; ++r) {
        const float* packed = wsp<float>(ws, p->off_packed[runs[r].pset]);
        const float* bias = p->pset_params[runs[r].pset] + p->pl.m_b[j];
        const float* A = cat + (size_t)runs[r].first * Ta * 512;
        float* Y = outs[j] + (size_t)runs[r].first * Ta * 256;
        const long long M = (long long)runs[r].count * Ta;
        if (tc) {   // B operand rows = output channels, K-major as the reference stores them
          BSED_TRY(tc_gemm_nt(A, 512, packed + p->pk.m_w[j], LO(packed + p->pk.m_w[j]), 512, Y, 256, M, 256, 512, bias, 0, sms, st));
        } else {
          BSED_TRY(gemm_nn(A, 512, packed + p->pk.m_wT[j], 256, Y, 256, (int)M, 256, 512, bias, 0, st));
        }
      }
    }
  }
  p->saved_valid = save;
  return BSED_OK;
}

// =================================================================================================
// backward
// =================================================================================================
extern "C" int bsed_crnn_backward(bsed_plan p, uint32_t group_mask, const float* d_enc, float* grads,
                                  int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  BSED_REQUIRE(p && grads && workspace && d_enc, "crnn_backward: null argument");
  if (!p->saved_valid) {
    bsed_set_error("crnn_backward: no saved forward (call bsed_crnn_forward with BSED_F_TRAIN|BSED_F_SAVE first)");
    return BSED_E_STATE;
  }
  if (workspace_bytes < p->ws_bytes) {
    bsed_set_error("crnn_backward: workspace %zu < %zu", workspace_bytes, p->ws_bytes);
    return BSED_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  const bsed_crnn_cfg& c = p->cfg;
  const ParamLayout& pl = p->pl;
  void* ws = workspace;
  // contiguous subset of groups sharing one parameter buffer
  int ids[kMaxGroups], n = 0;
  for (int i = 0; i < p->n_groups; ++i)
    if (group_mask & (1u << i)) ids[n++] = i;
  BSED_REQUIRE(n >= 1, "crnn_backward: empty group mask");
  for (int k = 1; k < n; ++k) {
    BSED_REQUIRE(ids[k] == ids[k - 1] + 1, "crnn_backward: masked groups must be adjacent");
    BSED_REQUIRE(p->gparams[ids[k]] == p->gparams[ids[0]], "crnn_backward: masked groups must share parameters");
  }
  Groups gb;
  gb.n = n;
  for (int k = 0; k < kMaxGroups; ++k) {
    gb.first[k] = k < n ? p->groups.first[ids[k]] : 0;
    gb.count[k] = k < n ? p->groups.count[ids[k]] : 0;
  }
  const int first = gb.first[0];
  int nb = 0;
  for (int k = 0; k < n; ++k) nb += gb.count[k];
  const float* params = p->gparams[ids[0]];
  const float* packed = wsp<float>(ws, p->off_packed[p->gpset[ids[0]]]);
  const int sms = p->ctx->num_sms;
  const int target = sms * 4;
  const bool tc = p->precision != BSED_PRECISION_FP32;
  const bool x3 = p->precision == BSED_PRECISION_TF32X3;
  const long long lo_off = p->pk.total;
  auto LO = [&](const float* w) -> const float* { return x3 ? w + lo_off : nullptr; };
  float* wgpart = wsp<float>(ws, p->off_wgpart[0]);
  double* dscr = wsp<double>(ws, p->off_dscratch[0]);

  if (!accumulate) BSED_CHECK_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * pl.total, st));

  // ---- gradient w.r.t. the dropout output of every GRU stack
  Groups one;
  one.n = 1;
  for (int k = 0; k < kMaxGroups; ++k) one.first[k] = 0, one.count[k] = k == 0 ? 1 : 0;
  if (!c.fpn) {
    const size_t ro = (size_t)first * p->Tout;  // row offset of the first masked clip in (B*T)-row matrices
    BSED_CHECK_CUDA(cudaMemcpyAsync(wsp<float>(ws, p->off_denc[0]) + ro * 256, d_enc + ro * 256,
                                    sizeof(float) * (size_t)nb * p->Tout * 256, cudaMemcpyDeviceToDevice, st));
  } else {
    // backward of the feature-pyramid merge, fine scale first (forward: src/models/CRNN.py:323-328)
    float* dcat = wsp<float>(ws, p->off_dcat);
    const float* dY[2] = {wsp<float>(ws, p->off_dy2), d_enc};                 // gradient of the merge output of level j
    float* d_hi[2] = {wsp<float>(ws, p->off_denc[1]), wsp<float>(ws, p->off_denc[0])};
    float* d_lo[2] = {wsp<float>(ws, p->off_denc[2]), wsp<float>(ws, p->off_dy2)};
    for (int j = 1; j >= 0; --j) {
      const int Ta = p->stackT[1 - j], Tb = p->stackT[2 - j];
      const size_t roa = (size_t)first * Ta, rob = (size_t)first * Tb;
      const long long M = (long long)nb * Ta;
      const float* dy = dY[j] + roa * 256;
      const float* cat = wsp<float>(ws, p->off_cat[j]) + roa * 512;
      // dW[co][k] += sum_rows dy[row][co] * cat[row][k] ; db += column sums of dy
      if (tc) {
        TcOperand Aw{dy, 256, 0, 256}, Bw{cat, 512, 0, 512};
        BSED_TRY(tc_wgrad_ex(Aw, Bw, nb, Ta, 1, 1, 0, grads + pl.m_w[j], 512, 1, 0, wgpart, p->wgpart_bytes, sms, st));
      } else {
        BSED_TRY(gemm_tn(dy, 256, cat, 512, grads + pl.m_w[j], 512, 1, 256, 512, M, target, st));
      }
      BSED_CHECK_CUDA(cudaMemsetAsync(dscr, 0, sizeof(double) * 2 * 768, st));
      BSED_TRY(col_stats(dy, nullptr, 2, one, M, 256, dscr, sms, st));
      BSED_TRY(add_double_to_float(dscr, 2, grads + pl.m_b[j], 256, st));
      // dcat = dy * W   ([M][256] x [256][512])
      if (tc) {
        BSED_TRY(tc_gemm_nt(dy, 256, packed + p->pk.m_wT[j], LO(packed + p->pk.m_wT[j]), 256, dcat + roa * 512, 512, M, 512, 256,
                            nullptr, 0, sms, st));
      } else {
        BSED_TRY(gemm_nn(dy, 256, packed + p->pk.m_w[j], 512, dcat + roa * 512, 512, (int)M, 512, 256, nullptr, 0, st));
      }
      BSED_TRY(fpn_cat_upsample_bwd(dcat + roa * 512, d_hi[j] + roa * 256, d_lo[j] + rob * 256, nb, Ta, Tb, st));
    }
  }

  // ---- one GRU stack: dropout mask, layers top-down; dX of layer 0 goes to `dx_target` (= or +=).  `br` selects the
  // scratch set, `ss` the stream (the caller's, or a plan-owned branch stream)
  int cur = 0;
  float* dpool_cur = wsp<float>(ws, p->off_dpool[cur]);
  const bool forked = p->n_stacks > 1 && p->aux[0] != nullptr;
  auto stack_backward = [&](int s, int br, cudaStream_t ss, float* dx_target, int accumulate_dx) -> int {
    const int T = p->stackT[s];
    const size_t ro = (size_t)first * T;
    const long long BTn = (long long)nb * T;
    float* denc = wsp<float>(ws, p->off_denc[s]);
    float* dxg = wsp<float>(ws, p->off_dxg[br]);
    float* dgh = wsp<float>(ws, p->off_dgh[br]);
    float* dx1 = wsp<float>(ws, p->off_dx1[br]);
    float* wgp = wsp<float>(ws, p->off_wgpart[br]);
    double* dsc = wsp<double>(ws, p->off_dscratch[br]);
    BSED_TRY(dropout_bwd_mask(denc + ro * 256, nullptr, (long long)ro * 256, BTn * 256, p->skeys[s], p->thresh,
                              p->inv_keep, ss));
    for (int l = c.rnn_layers - 1; l >= 0; --l) {
      const int In = l == 0 ? 128 : 256;
      const float* X = l == 0 ? wsp<float>(ws, p->off_pool[p->stack_src[s]]) : wsp<float>(ws, p->off_gru_out[s][l - 1]);
      const float* out_l = wsp<float>(ws, p->off_gru_out[s][l]);
      // gradient w.r.t. this layer's output: denc for the top layer, dx1 (ping-pong with denc) below
      float* dout = ((c.rnn_layers - 1 - l) % 2 == 0) ? denc : dx1;
      float* dxin = ((c.rnn_layers - 1 - l) % 2 == 0) ? dx1 : denc;
      BSED_TRY(gru_backward(dout, wsp<float>(ws, p->off_gru_saved[s][l]), out_l, packed + p->pk.whh[s][l], dxg, dgh, T, first,
                            nb, ss));
      for (int d = 0; d < 2; ++d) {
        if (tc) {
          // dW_hh[j][i] += sum_{b,t} dgh[b][t][j] * h[b][t -/+ 1][i] ; dW_ih[j][i] += sum_{b,t} dxg[b][t][j] * x[b][t][i]
          TcOperand Ah{dgh + ro * 768, 768, d * 384, 384}, Bh{out_l + ro * 256, 256, d * 128, 128};
          BSED_TRY(tc_wgrad_ex(Ah, Bh, nb, T, 1, 1, d == 0 ? -1 : 1, grads + pl.whh[s][l][d], 128, 1, 0, wgp,
                               p->wgpart_bytes, sms, ss));
          TcOperand Ai{dxg + ro * 768, 768, d * 384, 384}, Bi{X + ro * In, In, 0, In};
          BSED_TRY(tc_wgrad_ex(Ai, Bi, nb, T, 1, 1, 0, grads + pl.wih[s][l][d], In, 1, 0, wgp, p->wgpart_bytes, sms, ss));
        } else {
          BSED_TRY(gru_whh_grad(dgh + ro * 768 + d * 384, 768, out_l + ro * 256 + d * 128, 256, d == 0 ? -1 : 1,
                                grads + pl.whh[s][l][d], T, BTn, target, ss));
          BSED_TRY(gemm_tn(dxg + ro * 768 + d * 384, 768, X + ro * In, In, grads + pl.wih[s][l][d], In, 1, 384, In, BTn,
                           target, ss));
        }
      }
      BSED_CHECK_CUDA(cudaMemsetAsync(dsc, 0, sizeof(double) * 2 * 768, ss));
      BSED_TRY(col_stats(dxg + ro * 768, nullptr, 2, one, BTn, 768, dsc, sms, ss));
      BSED_TRY(add_double_to_float(dsc, 2, grads + pl.bih[s][l][0], 384, ss));
      BSED_TRY(add_double_to_float(dsc + 2 * 384, 2, grads + pl.bih[s][l][1], 384, ss));
      BSED_CHECK_CUDA(cudaMemsetAsync(dsc, 0, sizeof(double) * 2 * 768, ss));
      BSED_TRY(col_stats(dgh + ro * 768, nullptr, 2, one, BTn, 768, dsc, sms, ss));
      BSED_TRY(add_double_to_float(dsc, 2, grads + pl.bhh[s][l][0], 384, ss));
      BSED_TRY(add_double_to_float(dsc + 2 * 384, 2, grads + pl.bhh[s][l][1], 384, ss));
      float* dX = l == 0 ? dx_target + ro * 128 : dxin + ro * 256;
      const int acc = l == 0 ? accumulate_dx : 0;
      if (tc) {   // dX = dxg * [W_ih ; W_ih_reverse]: B operand rows = wihT [In][768]
        BSED_TRY(tc_gemm_nt(dxg + ro * 768, 768, packed + p->pk.wihT[s][l], LO(packed + p->pk.wihT[s][l]), 768, dX, In, BTn, In,
                            768, nullptr, acc, sms, ss));
      } else {
        BSED_TRY(gemm_nn(dxg + ro * 768, 768, packed + p->pk.wih_cat[s][l], In, dX, In, (int)BTn, In, 768, nullptr, acc, ss));
      }
    }
    return BSED_OK;
  };

  // ---- one conv block: consumes dpool_cur (gradient of its pooled output), leaves the gradient of its input there
  float* dxn_buf[2] = {wsp<float>(ws, p->off_dxn), wsp<float>(ws, p->off_dxn2)};
  float* wgpart_side = wsp<float>(ws, p->off_wgpart_side);
  bool side_pending[2] = {false, false};
  const bool use_side = p->side != nullptr;
  double* stats2 = wsp<double>(ws, p->off_stats2);
  float* G = wsp<float>(ws, p->off_G);
  auto block_backward = [&](int i) -> int {
    const LayerGeom& L = p->L[i];
    const int kb = use_side ? (i & 1) : 0;      // dY buffer of this block
    float* dxn = dxn_buf[kb];
    if (side_pending[kb]) {                     // the weight gradient that last read this buffer must be done
      BSED_CHECK_CUDA(cudaStreamWaitEvent(st, p->ev_wg[kb], 0));
      side_pending[kb] = false;
    }
    float* xhat = wsp<float>(ws, p->off_xhat[i]);
    float* lin = wsp<float>(ws, p->off_lin[i]);
    BNPtrs bn = make_bn_ptrs(p, ws, i, ids, n);
    const size_t off = (size_t)first * L.rows * L.Cout;
    const long long M = (long long)nb * L.rows;
    if (tc) {
      // ---- tensor-core path: every BatchNorm-backward statistic comes from column sums taken inside the gate
      // kernel plus the per-group GLU weight-gradient GEMM (cnn_ops.cu: glu_gate_pool_bwd_sums), and the BatchNorm
      // backward itself is the epilogue of the GEMM that adds the GLU linear path to dxn.
      double* bsums = wsp<double>(ws, p->off_bsums);
      float* tab = wsp<float>(ws, p->off_bntab);
      BSED_CHECK_CUDA(cudaMemsetAsync(bsums, 0, sizeof(double) * kMaxGroups * 128 * 4, st));
      BSED_TRY(glu_gate_pool_bwd_sums(xhat, lin, dpool_cur, dxn, gb, bn, L.T, L.F, L.Cout, L.pt, L.pf, p->bkeys[i],
                                      p->bthresh[i], p->binv[i], bsums, sms, st));
      BSED_CHECK_CUDA(cudaMemsetAsync(G, 0, sizeof(float) * n * L.Cout * L.Cout, st));
      int gfirst_rel[kMaxGroups] = {0, 0, 0, 0};
      for (int k = 0; k < n; ++k) {
        gfirst_rel[k] = gb.first[k] - first;
        const size_t goff = (size_t)gb.first[k] * L.rows * L.Cout;
        float* Gk = G + (size_t)k * L.Cout * L.Cout;
        if (L.Cout % 32 == 0 && L.F <= 64 && 64 % L.F == 0) {
          BSED_TRY(tc_wgrad(xhat + goff, lin + goff, Gk, L.Cout, 1, 0, gb.count[k], L.T, L.F, L.Cout, L.Cout, 1, wgpart,
                            p->wgpart_bytes, sms, st));
        } else if (L.Cout == 16 && L.F % 2 == 0 && L.F <= 128 && 64 % (L.F / 2) == 0) {
          TcOperand Ag{lin + goff, 32, 0, 32}, Bg{xhat + goff, 32, 0, 32};   // two pixels per row (32 floats)
          BSED_TRY(tc_wgrad_ex(Ag, Bg, gb.count[k], L.T, L.F / 2, 3, 0, Gk, L.Cout, 1, 0, wgpart, p->wgpart_bytes, sms, st));
        } else {
          BSED_TRY(gemm_tn(lin + goff, L.Cout, xhat + goff, L.Cout, Gk, L.Cout, 1, L.Cout, L.Cout,
                           (long long)gb.count[k] * L.rows, target, st));
        }
      }
      const int pack = glu_pack(L.Cout), CP = L.Cout * pack;
      BSED_TRY(bn_bwd_prepare(bsums, G, n, L.Cout, gb, L.rows, bn, params + pl.glu_w[i], params + pl.bn_w[i],
                              params + pl.bn_b[i], pack, tab, grads + pl.bn_w[i], grads + pl.bn_b[i], grads + pl.glu_w[i],
                              grads + pl.glu_b[i], st));
      BSED_TRY(tc_gemm_nt_bnbwd(lin + off, packed + p->pk.glu_wgT[i], LO(packed + p->pk.glu_wgT[i]), dxn + off, xhat + off,
                                M / pack, CP, CP, tab, n, L.rows / pack, gfirst_rel, sms, st));
    } else {
    // gate / dropout / pool backward: lin -> d_lin (in place), dxn <- direct gate path
    BSED_TRY(glu_gate_pool_bwd(xhat, lin, dpool_cur, dxn, gb, bn, L.T, L.F, L.Cout, L.pt, L.pf, p->bkeys[i], p->bthresh[i],
                               p->binv[i], st));
    // dxn += d_lin * Wg        (Wg [c'][c] is already K-major for this product)
    BSED_TRY(gemm_nn(lin + off, L.Cout, params + pl.glu_w[i], L.Cout, dxn + off, L.Cout, (int)M, L.Cout, L.Cout,
                     nullptr, 1, st));
    // G = d_lin^T xhat ; dbg = colsum(d_lin)
    BSED_CHECK_CUDA(cudaMemsetAsync(G, 0, sizeof(float) * L.Cout * L.Cout, st));
    if (tc && L.Cout % 32 == 0 && L.F <= 64 && 64 % L.F == 0) {
      BSED_TRY(tc_wgrad(xhat + off, lin + off, G, L.Cout, 1, 0, nb, L.T, L.F, L.Cout, L.Cout, 1, wgpart, p->wgpart_bytes,
                        sms, st));
    } else if (tc && L.Cout == 16 && L.F % 2 == 0 && L.F <= 128 && 64 % (L.F / 2) == 0) {
      // 16-channel block: both operands viewed two pixels per row (32 floats)
      TcOperand Ag{lin + off, 32, 0, 32}, Bg{xhat + off, 32, 0, 32};
      BSED_TRY(tc_wgrad_ex(Ag, Bg, nb, L.T, L.F / 2, 3, 0, G, L.Cout, 1, 0, wgpart, p->wgpart_bytes, sms, st));
    } else
      BSED_TRY(gemm_tn(lin + off, L.Cout, xhat + off, L.Cout, G, L.Cout, 1, L.Cout, L.Cout, M, target, st));
    Groups one;
    one.n = 1;
    one.first[0] = 0;
    one.count[0] = 1;
    BSED_CHECK_CUDA(cudaMemsetAsync(dscr, 0, sizeof(double) * 2 * 128, st));
    BSED_TRY(col_stats(lin + off, nullptr, 2, one, M, L.Cout, dscr, sms, st));
    // BN backward reductions per group
    BSED_CHECK_CUDA(cudaMemsetAsync(stats2, 0, sizeof(double) * kMaxGroups * 128 * 2, st));
    BSED_TRY(col_stats(dxn, xhat, 1, gb, L.rows, L.Cout, stats2, sms, st));
    BSED_TRY(bn_glu_param_grads(stats2, n, L.Cout, params + pl.bn_w[i], params + pl.bn_b[i], G, dscr,
                                grads + pl.bn_w[i], grads + pl.bn_b[i], grads + pl.glu_w[i], grads + pl.glu_b[i], st));
    BSED_TRY(bn_bwd_apply(dxn, xhat, stats2, gb, L.rows, L.Cout, bn, st));
    }
    // conv bias gradient: sum_p dY = gamma*rstd*(sum dxn - n*mean(dxn) - mean(dxn*xhat) * sum xhat) = 0 exactly behind a
    // train-mode BatchNorm (the reference computes rounding noise here); grads[conv_b] keeps its zero / accumulated value
    // the conv weight gradient of this block: on the side stream once dY is final
    cudaStream_t wst = use_side ? p->side : st;
    float* wgp = use_side ? wgpart_side : wgpart;
    if (use_side) {
      BSED_CHECK_CUDA(cudaEventRecord(p->ev_dy[kb], st));
      BSED_CHECK_CUDA(cudaStreamWaitEvent(p->side, p->ev_dy[kb], 0));
    }
    if (i > 0) {
      const float* xin = wsp<float>(ws, p->off_pool[i - 1]) + (size_t)first * L.rows * L.Cin;
      const bool tc_w32 = L.Cin % 32 == 0 && L.F <= 64 && 64 % L.F == 0;
      const bool tc_w16 = L.Cin == 16 && L.F % 2 == 0 && L.F <= 128 && 64 % (L.F / 2) == 0;
      if (tc && L.Cout % 32 == 0 && (tc_w32 || tc_w16))
        BSED_TRY(tc_wgrad(xin, dxn + off, grads + pl.conv_w[i], (long long)L.Cin * 9, 9, 1, nb, L.T, L.F, L.Cin, L.Cout, 9,
                          wgp, p->wgpart_bytes, sms, wst));
      else
        BSED_TRY(conv3x3_wgrad(xin, dxn + off, grads + pl.conv_w[i], nb, L.T, L.F, L.Cin, L.Cout, target, wst));
      if (use_side) {
        BSED_CHECK_CUDA(cudaEventRecord(p->ev_wg[kb], p->side));
        side_pending[kb] = true;
      }
      cur ^= 1;
      float* dnext = wsp<float>(ws, p->off_dpool[cur]);
      if (tc)
        BSED_TRY(tc_conv3x3(dxn + off, packed + p->pk.wd[i], x3_conv_dgrad() ? LO(packed + p->pk.wd[i]) : nullptr,
                            dnext + (size_t)first * L.rows * L.Cin, nb, L.T, L.F, L.Cout, L.Cin, nullptr, 0, sms, st));
      else
        BSED_TRY(conv3x3_nn(dxn + off, packed + p->pk.wd[i], dnext + (size_t)first * L.rows * L.Cin, nb, L.T, L.F, L.Cout,
                            L.Cin, nullptr, 0, st));
      dpool_cur = dnext;
    } else {
      BSED_TRY(conv0_wgrad(p->x_in, dxn, grads + pl.conv_w[0], first, nb, L.T, L.F, L.Cout, sms, wst));
      if (use_side) {
        BSED_CHECK_CUDA(cudaEventRecord(p->ev_wg[kb], p->side));
        side_pending[kb] = true;
      }
    }
    return BSED_OK;
  };

  // coarse scales first: rnn_4 -> second stage application -> (+ rnn_2) -> first application -> (+ rnn) -> trunk.
  // With branch streams, rnn and rnn_2 run their backward concurrently with that chain into their own dX buffers, which
  // are added to the block gradients after the join.
  if (p->n_stacks == 1) {
    BSED_TRY(stack_backward(0, 0, st, dpool_cur, 0));
  } else if (!forked) {
    for (int s = p->n_stacks - 1; s >= 1; --s) {
      BSED_TRY(stack_backward(s, 0, st, dpool_cur, s != p->n_stacks - 1));
      BSED_TRY(block_backward(p->stack_src[s]));
    }
    BSED_TRY(stack_backward(0, 0, st, dpool_cur, 1));
  } else {
    for (int s = 0; s < p->n_stacks - 1; ++s) {   // stack s on branch s + 1
      BSED_CHECK_CUDA(cudaEventRecord(p->ev_fork[s], st));
      BSED_CHECK_CUDA(cudaStreamWaitEvent(p->aux[s], p->ev_fork[s], 0));
      BSED_TRY(stack_backward(s, s + 1, p->aux[s], wsp<float>(ws, p->off_dx0[s + 1]), 0));
      BSED_CHECK_CUDA(cudaEventRecord(p->ev_join[s], p->aux[s]));
    }
    BSED_TRY(stack_backward(p->n_stacks - 1, 0, st, dpool_cur, 0));
    for (int s = p->n_stacks - 1; s >= 1; --s) {
      BSED_TRY(block_backward(p->stack_src[s]));                        // leaves the gradient of pool[stack_src[s - 1]]
      BSED_CHECK_CUDA(cudaStreamWaitEvent(st, p->ev_join[s - 1], 0));
      const size_t ro = (size_t)first * p->stackT[s - 1] * 128;
      BSED_TRY(add_f32(dpool_cur + ro, wsp<float>(ws, p->off_dx0[s]) + ro, (long long)nb * p->stackT[s - 1] * 128, st));
    }
  }
  for (int i = c.n_cnn - 1; i >= 0; --i) BSED_TRY(block_backward(i));
  for (int k = 0; k < 2; ++k)   // join: the gradient buffer is complete when the caller's stream gets here
    if (side_pending[k]) BSED_CHECK_CUDA(cudaStreamWaitEvent(st, p->ev_wg[k], 0));
  p->saved_valid = false;  // lin buffers now hold gradients
  return BSED_OK;
}

// =================================================================================================
// debug access
// =================================================================================================
extern "C" int bsed_plan_debug_tensor(bsed_plan p, void* workspace, const char* name, float** ptr, int64_t* numel) {
  BSED_REQUIRE(p && workspace && name && ptr && numel, "debug_tensor: null argument");
  std::string s(name);
  const long long Bm = p->max_clips;
  auto layer_of = [&](const char* prefix, int* idx) {
    size_t n = strlen(prefix);
    if (s.compare(0, n, prefix) != 0 || s.size() != n + 1) return false;
    *idx = s[n] - '0';
    return *idx >= 0;
  };
  int i;
  if (layer_of("xhat", &i) && i < p->n_blocks) {
    *ptr = wsp<float>(workspace, p->off_xhat[i]);
    *numel = Bm * p->L[i].rows * p->L[i].Cout;
    return BSED_OK;
  }
  if (layer_of("lin", &i) && i < p->n_blocks) {
    *ptr = wsp<float>(workspace, p->off_lin[i]);
    *numel = Bm * p->L[i].rows * p->L[i].Cout;
    return BSED_OK;
  }
  if (layer_of("pool", &i) && i < p->n_blocks) {
    *ptr = wsp<float>(workspace, p->off_pool[i]);
    *numel = Bm * p->L[i].prows * p->L[i].Cout;
    return BSED_OK;
  }
  if (layer_of("gru", &i) && i < p->cfg.rnn_layers) {
    *ptr = wsp<float>(workspace, p->off_gru_out[0][i]);
    *numel = Bm * p->Tout * 256;
    return BSED_OK;
  }
  struct Tap { const char* n; size_t off; long long numel; } taps[] = {
      {"denc", p->off_denc[0], Bm * p->Tout * 256}, {"dx1", p->off_dx1[0], Bm * p->Tout * 256},
      {"dxg", p->off_dxg[0], Bm * p->Tout * 768},   {"dgh", p->off_dgh[0], Bm * p->Tout * 768},
      {"xg", p->off_xg[0], Bm * p->Tout * 768},     {"enc", p->off_enc[0], Bm * p->Tout * 256},
      {"dpool0", p->off_dpool[0], Bm * p->L[0].prows * p->L[0].Cout},
      {"dpool1", p->off_dpool[1], Bm * p->L[0].prows * p->L[0].Cout},
      {"saved0", p->off_gru_saved[0][0], Bm * p->Tout * 1024}, {"saved1", p->off_gru_saved[0][1], Bm * p->Tout * 1024}};
  for (const Tap& t : taps)
    if (s == t.n) {
      *ptr = wsp<float>(workspace, t.off);
      *numel = t.numel;
      return BSED_OK;
    }
  if (p->cfg.fpn) {
    struct Tap { const char* n; size_t off; long long numel; } ftaps[] = {
        {"enc1", p->off_enc[1], Bm * p->stackT[1] * 256}, {"enc2", p->off_enc[2], Bm * p->stackT[2] * 256},
        {"y2", p->off_y2, Bm * p->stackT[1] * 256},       {"cat0", p->off_cat[0], Bm * p->stackT[1] * 512},
        {"cat1", p->off_cat[1], Bm * p->Tout * 512},      {"denc1", p->off_denc[1], Bm * p->stackT[1] * 256},
        {"denc2", p->off_denc[2], Bm * p->stackT[2] * 256}, {"dy2", p->off_dy2, Bm * p->stackT[1] * 256}};
    for (const Tap& t : ftaps)
      if (s == t.n) {
        *ptr = wsp<float>(workspace, t.off);
        *numel = t.numel;
        return BSED_OK;
      }
  }
  if (s == "dxn") {   // dY of block 0 (even blocks use the first buffer)
    *ptr = wsp<float>(workspace, p->off_dxn);
    *numel = Bm * p->L[0].rows * p->L[0].Cout;
    return BSED_OK;
  }
  bsed_set_error("debug_tensor: unknown tensor '%s'", name);
  return BSED_E_INVALID;
}

// =================================================================================================
// Predictor (attention pooling head)                                src/models/CRNN.py:548-577
// =================================================================================================
extern "C" int64_t bsed_predictor_param_count(bsed_plan p) { return p ? p->pl.pred_total : -1; }
extern "C" int bsed_predictor_param_offsets(bsed_plan p, int64_t* offsets, int max_n) {
  if (!p) return -1;
  int n = (int)p->pl.pred_order.size();
  for (int i = 0; i < n && i < max_n; ++i) offsets[i] = p->pl.pred_order[i];
  return n;
}
extern "C" int bsed_predictor_ldl(void) { return kLdl; }

namespace {
struct PredWs {
  size_t packed, dlog, tmpw, dscr, total;
};
PredWs pred_ws(const bsed_crnn_plan* p, int n_clips) {
  PredWs w;
  size_t o = 0;
  auto takeb = [&](size_t bytes) {
    size_t r = o;
    o = align_up(o + bytes);
    return r;
  };
  w.packed = takeb(sizeof(float) * p->pk.pred_total);
  w.dlog = takeb(sizeof(float) * (size_t)n_clips * p->Tout * kLdl);
  w.tmpw = takeb(sizeof(float) * kLdl * 256);
  w.dscr = takeb(sizeof(double) * 2 * kLdl);
  w.total = o;
  return w;
}
}  // namespace

extern "C" size_t bsed_predictor_workspace_bytes(bsed_plan p, int n_clips) {
  return p && n_clips > 0 ? pred_ws(p, n_clips).total : 0;
}

extern "C" int bsed_predictor_forward(bsed_plan p, const float* pred_params, const float* enc, int n_clips,
                                      int inference, float* logits, float* strong, float* weak, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  BSED_REQUIRE(p && pred_params && enc && logits && strong && weak && workspace, "predictor_forward: null argument");
  BSED_REQUIRE(n_clips >= 1, "predictor_forward: n_clips=%d", n_clips);
  PredWs w = pred_ws(p, n_clips);
  if (workspace_bytes < w.total) {
    bsed_set_error("predictor_forward: workspace %zu < %zu", workspace_bytes, w.total);
    return BSED_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  float* packed = wsp<float>(workspace, w.packed);
  PrepTable tb;
  build_prep_table_head(p, pred_params, packed, &tb);
  BSED_TRY(run_prep(tb, st));
  const int T = p->Tout;
  BSED_TRY(gemm_nn(enc, 256, packed + p->pk.wcatT, kLdl, logits, kLdl, n_clips * T, kLdl, 256, packed + p->pk.bcat, 0,
                   st));
  return head_forward(logits, strong, weak, n_clips, T, p->cfg.n_class, kLdl, inference ? 1 : 0, st);
}

extern "C" int bsed_predictor_backward(bsed_plan p, const float* pred_params, const float* enc, const float* logits,
                                       const float* strong, const float* weak, const float* d_strong,
                                       const float* d_weak, int n_clips, float* d_enc, float* grads, int accumulate,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  BSED_REQUIRE(p && pred_params && enc && logits && strong && weak && d_enc && grads && workspace,
               "predictor_backward: null argument");
  BSED_REQUIRE(n_clips >= 1, "predictor_backward: n_clips=%d", n_clips);
  PredWs w = pred_ws(p, n_clips);
  if (workspace_bytes < w.total) {
    bsed_set_error("predictor_backward: workspace %zu < %zu", workspace_bytes, w.total);
    return BSED_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  const ParamLayout& pl = p->pl;
  const int T = p->Tout, C = p->cfg.n_class;
  const long long BTn = (long long)n_clips * T;
  const int sms = p->ctx->num_sms;
  float* packed = wsp<float>(workspace, w.packed);
  float* dlog = wsp<float>(workspace, w.dlog);
  float* tmpw = wsp<float>(workspace, w.tmpw);
  double* dscr = wsp<double>(workspace, w.dscr);
  PrepTable tb;
  build_prep_table_head(p, pred_params, packed, &tb);
  BSED_TRY(run_prep(tb, st));
  if (!accumulate) BSED_CHECK_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * pl.pred_total, st));
  BSED_TRY(head_backward(logits, strong, weak, d_strong, d_weak, dlog, 0, n_clips, T, C, kLdl, st));
  BSED_CHECK_CUDA(cudaMemsetAsync(tmpw, 0, sizeof(float) * kLdl * 256, st));
  BSED_TRY(gemm_tn(dlog, kLdl, enc, 256, tmpw, 256, 1, kLdl, 256, BTn, sms * 4, st));
  BSED_TRY(add_f32(grads + pl.dense_w, tmpw, (long long)C * 256, st));
  BSED_TRY(add_f32(grads + pl.sm_w, tmpw + (size_t)C * 256, (long long)C * 256, st));
  Groups one;
  one.n = 1;
  one.first[0] = 0;
  one.count[0] = 1;
  BSED_CHECK_CUDA(cudaMemsetAsync(dscr, 0, sizeof(double) * 2 * kLdl, st));
  BSED_TRY(col_stats(dlog, nullptr, 2, one, BTn, kLdl, dscr, sms, st));
  BSED_TRY(add_double_to_float(dscr, 2, grads + pl.dense_b, C, st));
  BSED_TRY(add_double_to_float(dscr + 2 * C, 2, grads + pl.sm_b, C, st));
  return gemm_nn(dlog, kLdl, packed + p->pk.wcat, 256, d_enc, 256, (int)BTn, 256, kLdl, nullptr, 0, st);
}
