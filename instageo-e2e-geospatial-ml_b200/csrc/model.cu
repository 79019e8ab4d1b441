// PrithviSeg forward engine behind the C ABI: weight store + kernel schedule.
//
// Replaces instageo/model/model.py:392-419 (PrithviSeg.forward) and
// instageo/model/pritvhi.py:498-530 (PrithviViT.forward) for inference (eval semantics:
// Dropout = identity, BatchNorm running statistics).  The library owns only packed weights;
// every activation lives in the caller's workspace.
//
// Data layout in HBM (per batch of B chips, N = 1 + T*196 tokens):
//   patches bf16 [B*T*196, 1536]      tubelet rows (k = c*256 + kh*16 + kw, Conv3d weight order)
//   x       f32  [B*N, D]             residual stream (kept f32 across all blocks)
//   xn      bf16 [B*N, D]             LayerNorm output = A operand of qkv / fc1
//   qkv     bf16 [B*N, 3D]            timm layout, read in place by the attention kernel
//   att     bf16 [B*N, D]             attention output = A operand of proj
//   hid     bf16 [B*N, 4D]            GELU(fc1) = A operand of fc2
//   head    bf16 padded-flat NHWC     [guard + B*(H+2)*(W+2) + guard, C] per stage, zero border;
//                                     stage-0 channels stored t*D + d (weights permuted to match)
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "ig_gemm.cuh"
#include "ig_ops.cuh"

namespace {

typedef __nv_bfloat16 bf16;

struct LayerW {
  float *ln1g, *ln1b, *ln2g, *ln2b, *qkv_b, *proj_b, *fc1_b, *fc2_b;
  bf16 *qkv_w, *proj_w, *fc1_w, *fc2_w;
};
struct StageW {
  bf16 *ct_w, *cv_w;
  bf16* ct_ws;  // phase-stacked transposed-convolution weights [4*Cout][4*Cin], or null (wide stages)
  float *ct_b, *cv_b, *bn_g, *bn_b, *bn_m, *bn_v, *scale, *shift;
};

struct Buf {
  size_t off;       // byte offset in the workspace
  int Hp, Wp, C;    // padded geometry (head buffers)
  int guard;        // guard rows before/after
  int64_t rows;     // total rows incl. guards
};

struct Layout {
  size_t patches, x, xn, qkv, att, hid;
  Buf in0, t[4], a[3];
  size_t total;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct ig_fwd_plan;

struct ig_model {
  ig_model_cfg cfg;
  int D, L, heads, T, nc, g, ntok, K0;
  int dims[5];
  int device;
  bool finalized;
  std::vector<void*> allocs;
  std::vector<std::string> missing;  // required keys not loaded yet
  // encoder
  bf16* pe_w;
  float *pe_b, *pos, *cls, *norm_g, *norm_b;
  std::vector<LayerW> layers;
  StageW st[4];
  float *w1_raw, *b1_raw, *w1, *b1;
  // forward schedule cache (see "Forward schedule" below)
  std::vector<ig_fwd_plan*> plans;
  uint64_t use_ctr;
  char* ring_ws;      // (workspace, batch) whose head-buffer borders are currently zero
  int ring_batch;
  float* tap_buf;     // caller-owned [(L + 2), B, N, D] f32 or null (ig_model_set_tap_buffer)
  size_t tap_elems;
  int tap_batch;
  bool graphs_ok;
  cudaStream_t cap_stream;
  int last_path;      // 1 = the last forward was a graph launch
  int graph_kernels;  // kernel nodes of the most recently captured graph
  char graph_note[256];
};

namespace {

template <typename T>
int dev_alloc(ig_model* m, T** p, size_t n) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, n * sizeof(T) > 0 ? n * sizeof(T) : 16);
  if (e != cudaSuccess) {
    ig_set_error("cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
    return IG_ENOMEM;
  }
  m->allocs.push_back(q);
  *p = static_cast<T*>(q);
  return IG_OK;
}

Buf make_buf(size_t* cur, int B, int H, int C) {
  Buf b;
  b.Hp = H + 2;
  b.Wp = H + 2;
  b.C = C;
  b.guard = b.Wp + 8;
  b.rows = static_cast<int64_t>(b.guard) * 2 + static_cast<int64_t>(B) * b.Hp * b.Wp;
  b.off = *cur;
  *cur = align_up(*cur + static_cast<size_t>(b.rows) * C * sizeof(bf16), 1024);
  return b;
}

Layout make_layout(const ig_model* m, int B) {
  Layout l;
  size_t cur = 0;
  const size_t rows = static_cast<size_t>(B) * m->ntok;
  auto take = [&](size_t bytes) {
    size_t o = cur;
    cur = align_up(cur + bytes, 1024);
    return o;
  };
  l.patches = take(static_cast<size_t>(B) * m->T * m->g * m->g * m->K0 * sizeof(bf16));
  l.x = take(rows * m->D * sizeof(float));
  l.xn = take(rows * m->D * sizeof(bf16));
  l.qkv = take(rows * 3 * m->D * sizeof(bf16));
  l.att = take(rows * m->D * sizeof(bf16));
  l.hid = take(rows * 4 * m->D * sizeof(bf16));
  l.in0 = make_buf(&cur, B, m->g, m->dims[0]);
  int H = m->g;
  for (int i = 0; i < 4; ++i) {
    H *= 2;
    l.t[i] = make_buf(&cur, B, H, m->dims[i + 1]);
    if (i < 3) l.a[i] = make_buf(&cur, B, H, m->dims[i + 1]);
  }
  l.total = cur;
  return l;
}

bool starts_with(const char* s, const char* p) { return strncmp(s, p, strlen(p)) == 0; }

void mark_loaded(ig_model* m, const char* key) {
  for (size_t i = 0; i < m->missing.size(); ++i)
    if (m->missing[i] == key) {
      m->missing.erase(m->missing.begin() + i);
      return;
    }
}

int64_t numel(const int64_t* shape, int ndim) {
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  return n;
}

int copy_f32(float* dst, const float* src, int64_t n, int64_t expect, const char* key, cudaStream_t st) {
  IG_REQUIRE(n == expect, IG_ESHAPE, "weight %s has %lld elements, expected %lld", key,
             static_cast<long long>(n), static_cast<long long>(expect));
  IG_CUDA_OK(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return IG_OK;
}
int copy_bf16(bf16* dst, const float* src, int64_t n, int64_t expect, const char* key, cudaStream_t st) {
  IG_REQUIRE(n == expect, IG_ESHAPE, "weight %s has %lld elements, expected %lld", key,
             static_cast<long long>(n), static_cast<long long>(expect));
  return ops::cvt_bf16(src, dst, n, st);
}

}  // namespace

extern "C" int ig_model_create(const ig_model_cfg* cfg, ig_model** out) {
  IG_TRY(ig_check_device());
  IG_REQUIRE(cfg && out, IG_EINVAL, "ig_model_create: null pointer");
  IG_REQUIRE(cfg->patch_size == 16 && cfg->in_chans == 6, IG_ESHAPE,
             "only patch_size 16 / 6 input bands are supported (got %d / %d); the V2-600M variants "
             "(patch 14, head kernels 5/7) are out of scope", cfg->patch_size, cfg->in_chans);
  IG_REQUIRE(cfg->img_size >= 16 && cfg->img_size % 16 == 0 && cfg->img_size <= 224, IG_ESHAPE,
             "img_size %d unsupported", cfg->img_size);
  IG_REQUIRE(cfg->embed_dim % 128 == 0 && cfg->embed_dim >= 128 && cfg->embed_dim <= 1024, IG_ESHAPE,
             "embed_dim %d unsupported (multiple of 128, <= 1024)", cfg->embed_dim);
  IG_REQUIRE(cfg->num_heads * 64 == cfg->embed_dim, IG_ESHAPE, "head_dim must be 64 (D=%d heads=%d)",
             cfg->embed_dim, cfg->num_heads);
  IG_REQUIRE(cfg->depth >= 0 && cfg->temporal >= 1 && cfg->temporal <= 8, IG_ESHAPE, "bad depth/temporal");
  IG_REQUIRE(cfg->num_classes >= 1 && cfg->num_classes <= gemm::NCP, IG_ESHAPE, "num_classes %d unsupported (1..%d)",
             cfg->num_classes, gemm::NCP);
  IG_REQUIRE(cfg->head_dims[0] == cfg->embed_dim * cfg->temporal, IG_ESHAPE,
             "head_dims[0]=%d must equal embed_dim*temporal=%d", cfg->head_dims[0], cfg->embed_dim * cfg->temporal);
  for (int i = 0; i < 5; ++i)
    IG_REQUIRE(cfg->head_dims[i] >= 16 && cfg->head_dims[i] % 16 == 0 && gemm::pick_block_n(cfg->head_dims[i]) > 0,
               IG_ESHAPE, "head_dims[%d]=%d must be a multiple of 16", i, cfg->head_dims[i]);
  IG_REQUIRE(cfg->head_dims[4] <= gemm::MAX_BN, IG_ESHAPE, "head_dims[4]=%d must be <= %d for the fused 1x1 head",
             cfg->head_dims[4], gemm::MAX_BN);

  ig_model* m = new ig_model();
  m->cfg = *cfg;
  m->D = cfg->embed_dim;
  m->L = cfg->depth;
  m->heads = cfg->num_heads;
  m->T = cfg->temporal;
  m->nc = cfg->num_classes;
  m->g = cfg->img_size / 16;
  m->ntok = 1 + m->T * m->g * m->g;
  m->K0 = cfg->in_chans * 256;
  for (int i = 0; i < 5; ++i) m->dims[i] = cfg->head_dims[i];
  m->finalized = false;
  m->use_ctr = 0;
  m->ring_ws = nullptr;
  m->ring_batch = 0;
  m->tap_buf = nullptr;
  m->tap_elems = 0;
  m->tap_batch = 0;
  m->graphs_ok = true;
  m->cap_stream = nullptr;
  m->last_path = 0;
  m->graph_kernels = 0;
  m->graph_note[0] = 0;
  cudaGetDevice(&m->device);
  const int D = m->D;
  int rc = IG_OK;
#define A(ptr, n) if (rc == IG_OK) rc = dev_alloc(m, &(ptr), static_cast<size_t>(n))
  A(m->pe_w, static_cast<size_t>(D) * m->K0);
  A(m->pe_b, D);
  A(m->pos, static_cast<size_t>(m->ntok) * D);
  A(m->cls, D);
  A(m->norm_g, D);
  A(m->norm_b, D);
  m->layers.resize(m->L);
  for (int i = 0; i < m->L && rc == IG_OK; ++i) {
    LayerW& w = m->layers[i];
    A(w.ln1g, D); A(w.ln1b, D); A(w.ln2g, D); A(w.ln2b, D);
    A(w.qkv_b, 3 * D); A(w.proj_b, D); A(w.fc1_b, 4 * D); A(w.fc2_b, D);
    A(w.qkv_w, static_cast<size_t>(3) * D * D); A(w.proj_w, static_cast<size_t>(D) * D);
    A(w.fc1_w, static_cast<size_t>(4) * D * D); A(w.fc2_w, static_cast<size_t>(4) * D * D);
  }
  for (int i = 0; i < 4 && rc == IG_OK; ++i) {
    StageW& s = m->st[i];
    const size_t ci = m->dims[i], co = m->dims[i + 1];
    A(s.ct_w, 9 * ci * co); A(s.cv_w, 9 * co * co);
    // Narrow last stages (Cout <= 64, i.e. the T = 1 flood head's 96 -> 48): the four output-parity phases are
    // stacked along N (one GEMM, N = 4*Cout over the 2 x 2 input neighbourhood, 7 of 16 weight blocks zero) instead
    // of four GEMMs of N = Cout.  Tiles of 256 x 48 with one to four K = 96 taps carried 144-576 clk of tensor
    // work against ~3900 clk of per-tile epilogue: 165 TFLOP/s (ncu launch list, profiles/r02_launches_tile224.csv).
    s.ct_ws = nullptr;
    if (4 * co <= static_cast<size_t>(gemm::MAX_BN) && co % 8 == 0 && getenv("IG_NO_CONVT_STACK") == nullptr)
      A(s.ct_ws, 16 * ci * co);
    A(s.ct_b, co); A(s.cv_b, co); A(s.bn_g, co); A(s.bn_b, co); A(s.bn_m, co); A(s.bn_v, co);
    A(s.scale, co); A(s.shift, co);
  }
  A(m->w1_raw, static_cast<size_t>(m->nc) * m->dims[4]);
  A(m->b1_raw, m->nc);
  A(m->w1, static_cast<size_t>(m->dims[4]) * gemm::NCP);
  A(m->b1, gemm::NCP);
#undef A
  if (rc != IG_OK) {
    ig_model_destroy(m);
    return rc;
  }
  // required state_dict keys
  auto need = [&](const std::string& k) { m->missing.push_back(k); };
  const std::string e = "prithvi_encoder.";
  need(e + "cls_token"); need(e + "pos_embed");
  need(e + "patch_embed.proj.weight"); need(e + "patch_embed.proj.bias");
  need(e + "norm.weight"); need(e + "norm.bias");
  for (int i = 0; i < m->L; ++i) {
    const std::string b = e + "blocks." + std::to_string(i) + ".";
    for (const char* s : {"norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight",
                          "attn.proj.bias", "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias",
                          "mlp.fc2.weight", "mlp.fc2.bias"})
      need(b + s);
  }
  for (int i = 0; i < 4; ++i) {
    const std::string h = "segmentation_head." + std::to_string(i) + ".";
    for (const char* s : {"0.weight", "0.bias", "2.weight", "2.bias", "3.weight", "3.bias", "3.running_mean",
                          "3.running_var"})
      need(h + s);
  }
  need("segmentation_head.5.weight");
  need("segmentation_head.5.bias");
  *out = m;
  return IG_OK;
}

extern "C" int ig_model_destroy(ig_model* m) {
  if (!m) return IG_OK;
  ig_model_reset_cache(m);
  if (m->cap_stream) cudaStreamDestroy(m->cap_stream);
  for (void* p : m->allocs) cudaFree(p);
  delete m;
  return IG_OK;
}

extern "C" int ig_model_load_weight(ig_model* m, const char* key, const float* data, const int64_t* shape,
                                    int ndim, void* stream) {
  IG_REQUIRE(m && key && data && (shape || ndim == 0), IG_EINVAL, "ig_model_load_weight: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n = numel(shape, ndim);
  const int D = m->D;
  int rc = IG_OK;
  bool known = true;
  const char* e = "prithvi_encoder.";
  if (starts_with(key, e)) {
    const char* k = key + strlen(e);
    if (!strcmp(k, "cls_token")) rc = copy_f32(m->cls, data, n, D, key, st);
    else if (!strcmp(k, "pos_embed")) rc = copy_f32(m->pos, data, n, static_cast<int64_t>(m->ntok) * D, key, st);
    else if (!strcmp(k, "patch_embed.proj.weight")) rc = copy_bf16(m->pe_w, data, n, static_cast<int64_t>(D) * m->K0, key, st);
    else if (!strcmp(k, "patch_embed.proj.bias")) rc = copy_f32(m->pe_b, data, n, D, key, st);
    else if (!strcmp(k, "norm.weight")) rc = copy_f32(m->norm_g, data, n, D, key, st);
    else if (!strcmp(k, "norm.bias")) rc = copy_f32(m->norm_b, data, n, D, key, st);
    else if (starts_with(k, "blocks.")) {
      char* end = nullptr;
      const long li = strtol(k + 7, &end, 10);
      if (li < 0 || li >= m->L || !end || *end != '.') {
        known = false;  // blocks beyond `depth` are dropped like model.py:242-247
      } else {
        LayerW& w = m->layers[li];
        const char* s = end + 1;
        const int64_t DD = static_cast<int64_t>(D) * D;
        if (!strcmp(s, "norm1.weight")) rc = copy_f32(w.ln1g, data, n, D, key, st);
        else if (!strcmp(s, "norm1.bias")) rc = copy_f32(w.ln1b, data, n, D, key, st);
        else if (!strcmp(s, "norm2.weight")) rc = copy_f32(w.ln2g, data, n, D, key, st);
        else if (!strcmp(s, "norm2.bias")) rc = copy_f32(w.ln2b, data, n, D, key, st);
        else if (!strcmp(s, "attn.qkv.weight")) rc = copy_bf16(w.qkv_w, data, n, 3 * DD, key, st);
        else if (!strcmp(s, "attn.qkv.bias")) rc = copy_f32(w.qkv_b, data, n, 3 * D, key, st);
        else if (!strcmp(s, "attn.proj.weight")) rc = copy_bf16(w.proj_w, data, n, DD, key, st);
        else if (!strcmp(s, "attn.proj.bias")) rc = copy_f32(w.proj_b, data, n, D, key, st);
        else if (!strcmp(s, "mlp.fc1.weight")) rc = copy_bf16(w.fc1_w, data, n, 4 * DD, key, st);
        else if (!strcmp(s, "mlp.fc1.bias")) rc = copy_f32(w.fc1_b, data, n, 4 * D, key, st);
        else if (!strcmp(s, "mlp.fc2.weight")) rc = copy_bf16(w.fc2_w, data, n, 4 * DD, key, st);
        else if (!strcmp(s, "mlp.fc2.bias")) rc = copy_f32(w.fc2_b, data, n, D, key, st);
        else known = false;
      }
    } else {
      known = false;  // temporal_embed_enc.* / location_embed_enc.*: never read by forward
    }
  } else if (starts_with(key, "segmentation_head.")) {
    const char* k = key + strlen("segmentation_head.");
    if (!strcmp(k, "5.weight")) rc = copy_f32(m->w1_raw, data, n, static_cast<int64_t>(m->nc) * m->dims[4], key, st);
    else if (!strcmp(k, "5.bias")) rc = copy_f32(m->b1_raw, data, n, m->nc, key, st);
    else if (k[0] >= '0' && k[0] <= '3' && k[1] == '.') {
      const int i = k[0] - '0';
      StageW& s = m->st[i];
      const int ci = m->dims[i], co = m->dims[i + 1];
      const char* r = k + 2;
      if (!strcmp(r, "0.weight")) {
        IG_REQUIRE(ndim == 4 && shape[0] == ci && shape[1] == co && shape[2] == 3 && shape[3] == 3, IG_ESHAPE,
                   "%s: expected [%d,%d,3,3]", key, ci, co);
        rc = ops::repack_conv_weight(data, s.ct_w, ci, co, 1, i == 0 ? m->T : 1, st);
      } else if (!strcmp(r, "2.weight")) {
        IG_REQUIRE(ndim == 4 && shape[0] == co && shape[1] == co && shape[2] == 3 && shape[3] == 3, IG_ESHAPE,
                   "%s: expected [%d,%d,3,3] (head kernel sizes other than 3 are out of scope)", key, co, co);
        rc = ops::repack_conv_weight(data, s.cv_w, co, co, 0, 1, st);
      } else if (!strcmp(r, "0.bias")) rc = copy_f32(s.ct_b, data, n, co, key, st);
      else if (!strcmp(r, "2.bias")) rc = copy_f32(s.cv_b, data, n, co, key, st);
      else if (!strcmp(r, "3.weight")) rc = copy_f32(s.bn_g, data, n, co, key, st);
      else if (!strcmp(r, "3.bias")) rc = copy_f32(s.bn_b, data, n, co, key, st);
      else if (!strcmp(r, "3.running_mean")) rc = copy_f32(s.bn_m, data, n, co, key, st);
      else if (!strcmp(r, "3.running_var")) rc = copy_f32(s.bn_v, data, n, co, key, st);
      else known = false;
    } else {
      known = false;
    }
  } else {
    known = false;
  }
  if (rc != IG_OK) return rc;
  if (known) {
    mark_loaded(m, key);
    m->finalized = false;
  }
  return IG_OK;
}

extern "C" int ig_model_finalize(ig_model* m, void* stream) {
  IG_REQUIRE(m, IG_EINVAL, "ig_model_finalize: null model");
  if (!m->missing.empty()) {
    ig_set_error("ig_model_finalize: %zu weights not loaded, first missing: %s", m->missing.size(),
                 m->missing[0].c_str());
    return IG_ESTATE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < 4; ++i) {
    StageW& s = m->st[i];
    IG_TRY(ops::bn_fold(s.bn_g, s.bn_b, s.bn_m, s.bn_v, s.cv_b, s.scale, s.shift, m->dims[i + 1], st));
    if (s.ct_ws) IG_TRY(ops::stack_convt_weight(s.ct_w, s.ct_ws, m->dims[i], m->dims[i + 1], st));
  }
  IG_TRY(ops::repack_head1x1(m->w1_raw, m->b1_raw, m->w1, m->b1, m->nc, m->dims[4], gemm::NCP, st));
  m->finalized = true;
  return IG_OK;
}

extern "C" size_t ig_model_workspace_bytes(const ig_model* m, int batch) {
  if (!m || batch < 1) return 0;
  return make_layout(m, batch).total;
}

extern "C" int ig_model_launches_per_forward(const ig_model* m) {
  if (!m) return 0;
  // patchify (f32 entry only) + cls + patch-embed + 7 per block + final norm + 8 head GEMMs; the 5 border clears
  // run once per (batch, workspace) layout, not per forward
  return 3 + 7 * m->L + 1 + 8;
}

namespace {

int plan_conv(gemm::Plan* p, int epi, const ig_model* m, char* ws, const Buf& in, const void* w, int Cin, int Cout,
              int B, bool transposed, bool stacked = false) {
  gemm::Args& a = p->args;
  a = gemm::Args{};
  a.M = B * in.Hp * in.Wp;
  a.N = stacked ? 4 * Cout : Cout;
  a.block_n = gemm::pick_block_n(a.N);
  a.kc = Cin;
  a.num_m_tiles = (a.M + gemm::PAIR_M - 1) / gemm::PAIR_M;
  a.num_n_tiles = a.N / a.block_n;
  a.a_row_base = in.guard;
  a.Hp = in.Hp;
  a.Wp = in.Wp;
  a.ldo = Cout;
  a.a_box_rows = gemm::BM + gemm::A_HALO;
  if (!transposed) {
    // 3 A tiles (dy = -1, 0, 1), each shared by the 3 horizontal taps dx = -1, 0, 1 (row shifts 0, 1, 2)
    a.num_phases = 1;
    a.taps[0].n = 3;
    for (int ky = 0; ky < 3; ++ky) {
      gemm::TapGroup& g = a.taps[0].g[ky];
      g.a_off = (ky - 1) * in.Wp - 1;
      g.nsub = 3;
      for (int kx = 0; kx < 3; ++kx) {
        g.shift[kx] = kx;
        g.b_off[kx] = (ky * 3 + kx) * Cin;
      }
    }
  } else if (stacked) {
    // all four output parities in one accumulator row: N = 4*Cout, B = the stacked weights [4*Cout][4*Cin] whose
    // column block (iy*2 + ix) multiplies input pixel (y + iy, x + ix); one A tile per iy, shared by ix = 0, 1
    a.num_phases = 1;
    a.stack_cout = Cout;
    a.phase_a[0] = a.phase_b[0] = 0;
    a.taps[0].n = 2;
    for (int iy = 0; iy < 2; ++iy) {
      gemm::TapGroup& g = a.taps[0].g[iy];
      g.a_off = iy * in.Wp;
      g.nsub = 2;
      for (int ix = 0; ix < 2; ++ix) {
        g.shift[ix] = ix;
        g.b_off[ix] = (iy * 2 + ix) * Cin;
      }
    }
  } else {
    // out[2y+pa, 2x+pb]: pa=0 -> (ky=1, dy=0); pa=1 -> (ky=2, dy=0), (ky=0, dy=1); same along x.
    // One A tile per dy, shared by the dx = 0, 1 taps (row shifts 0, 1).
    a.num_phases = 4;
    for (int pa = 0; pa < 2; ++pa)
      for (int pb = 0; pb < 2; ++pb) {
        const int ph = pa * 2 + pb;
        a.phase_a[ph] = pa;
        a.phase_b[ph] = pb;
        gemm::Taps& t = a.taps[ph];
        t.n = pa ? 2 : 1;
        const int kys[2] = {pa ? 2 : 1, 0};
        const int kxs[2] = {pb ? 2 : 1, 0};
        for (int iy = 0; iy < t.n; ++iy) {
          gemm::TapGroup& g = t.g[iy];
          g.a_off = iy * in.Wp;
          g.nsub = pb ? 2 : 1;
          for (int ix = 0; ix < g.nsub; ++ix) {
            g.shift[ix] = ix;
            g.b_off[ix] = (kys[iy] * 3 + kxs[ix]) * Cin;
          }
        }
      }
  }
  gemm::finish_geometry(&a);
  p->epi = epi;
  const uint64_t wk = (stacked ? 4 : 9) * static_cast<uint64_t>(Cin);
  IG_TRY(gemm::make_maps(p, ws + in.off, in.rows, Cin, w, wk));
  (void)m;
  return IG_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// Forward schedule.  Everything that depends only on (batch, workspace address) -- ~30 + 4 L GEMM plans with their
// TMA descriptors, the attention descriptors -- is built ONCE per (batch, workspace) and kept in the model
// (FwdPlan); the production entry (bf16 tubelet rows in, no feature tap) is additionally captured into a CUDA
// graph per plan, so a steady-state forward is one cudaGraphLaunch.  Before this, every forward re-encoded ~120
// tensor maps on the host and issued ~108 launches: 17.02 ms per step for 16.52 ms of kernels (BENCH_r01).
// Caller pointers that may change between forwards (x, logits, argmax, prob) live in the parameters of exactly
// two kernel nodes (patch-embed GEMM, fused final conv); they are patched with cudaGraphExecKernelNodeSetParams.

namespace {

struct FwdIO {
  const void* x;
  int x_dtype;
  float* logits;
  int8_t* argmax;
  float* prob1;
  float* feats;
};

struct GraphEntry {
  unsigned flags;  // bit 0 logits, bit 1 argmax, bit 2 prob1
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  cudaGraphNode_t n_patch, n_final;
  FwdIO io;  // pointers currently baked into the executable graph
};

}  // namespace

struct ig_fwd_plan {
  int batch;
  char* ws;
  Layout l;
  gemm::Plan patch;        // A = ws + l.patches (f32 entry: rows written by patchify)
  gemm::Plan patch_ext;    // A = caller's tubelet rows (bf16 entry), re-encoded when the pointer changes
  const void* patch_ext_src;
  std::vector<gemm::Plan> enc;  // qkv, proj, fc1, fc2 per block
  CUtensorMap tmq, tmkv;
  gemm::Plan convt[4], conv[3], fin;
  std::vector<GraphEntry> graphs;
  uint64_t last_use;
};

namespace {

const int MAX_PLANS = 8;

void destroy_plan(ig_fwd_plan* fp) {
  for (GraphEntry& g : fp->graphs) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
  }
  delete fp;
}

int build_plan(ig_model* m, int batch, char* ws, ig_fwd_plan** out) {
  ig_fwd_plan* fp = new ig_fwd_plan();
  fp->batch = batch;
  fp->ws = ws;
  fp->l = make_layout(m, batch);
  fp->patch_ext_src = nullptr;
  fp->last_use = 0;
  const Layout& l = fp->l;
  const int B = batch, D = m->D, T = m->T, g = m->g, N = m->ntok;
  const int M = B * N, MP = B * T * g * g;
  float* xres = reinterpret_cast<float*>(ws + l.x);
  bf16* xn = reinterpret_cast<bf16*>(ws + l.xn);
  bf16* qkv = reinterpret_cast<bf16*>(ws + l.qkv);
  bf16* att = reinterpret_cast<bf16*>(ws + l.att);
  bf16* hid = reinterpret_cast<bf16*>(ws + l.hid);
  int rc = IG_OK;
#define PLAN_TRY(expr) do { if (rc == IG_OK) rc = (expr); } while (0)
  PLAN_TRY(gemm::plan_linear(&fp->patch, gemm::EPI_PATCH, ws + l.patches, m->K0, m->pe_w, MP, D, m->K0));
  if (rc == IG_OK) {
    gemm::Args& a = fp->patch.args;
    a.bias = m->pe_b;
    a.pos = m->pos;
    a.tok_per_img = T * g * g;
    a.ntok = N;
    a.out = xres;
    fp->patch_ext = fp->patch;
  }
  fp->enc.resize(static_cast<size_t>(4) * m->L);
  for (int i = 0; i < m->L && rc == IG_OK; ++i) {
    const LayerW& w = m->layers[i];
    gemm::Plan* p = &fp->enc[static_cast<size_t>(4) * i];
    PLAN_TRY(gemm::plan_linear(&p[0], gemm::EPI_BF16, xn, D, w.qkv_w, M, 3 * D, D));
    p[0].args.bias = w.qkv_b;
    p[0].args.out = qkv;
    PLAN_TRY(gemm::plan_linear(&p[1], gemm::EPI_RESID, att, D, w.proj_w, M, D, D));
    p[1].args.bias = w.proj_b;
    PLAN_TRY(gemm::set_residual_inplace(&p[1], xres));
    PLAN_TRY(gemm::plan_linear(&p[2], gemm::EPI_BF16, xn, D, w.fc1_w, M, 4 * D, D));
    p[2].args.bias = w.fc1_b;
    p[2].args.act = 1;
    p[2].args.out = hid;
    PLAN_TRY(gemm::plan_linear(&p[3], gemm::EPI_RESID, hid, 4 * D, w.fc2_w, M, D, 4 * D));
    p[3].args.bias = w.fc2_b;
    PLAN_TRY(gemm::set_residual_inplace(&p[3], xres));
  }
  if (m->L > 0) PLAN_TRY(ops::attention_maps(qkv, B, N, m->heads, &fp->tmq, &fp->tmkv));
  const Buf* in = &l.in0;
  for (int i = 0; i < 4 && rc == IG_OK; ++i) {
    const StageW& s = m->st[i];
    const Buf& tb = l.t[i];
    PLAN_TRY(plan_conv(&fp->convt[i], gemm::EPI_CONVT, m, ws, *in, s.ct_ws ? s.ct_ws : s.ct_w, m->dims[i],
                       m->dims[i + 1], B, true, s.ct_ws != nullptr));
    fp->convt[i].args.bias = s.ct_b;
    fp->convt[i].args.out = ws + tb.off;
    fp->convt[i].args.out_guard = tb.guard;
    if (i < 3) {
      const Buf& ab = l.a[i];
      PLAN_TRY(plan_conv(&fp->conv[i], gemm::EPI_CONV, m, ws, tb, s.cv_w, m->dims[i + 1], m->dims[i + 1], B, false));
      fp->conv[i].args.bias = s.scale;
      fp->conv[i].args.shift = s.shift;
      fp->conv[i].args.out = ws + ab.off;
      fp->conv[i].args.out_guard = ab.guard;
      in = &ab;
    } else {
      PLAN_TRY(plan_conv(&fp->fin, gemm::EPI_FINAL, m, ws, tb, s.cv_w, m->dims[4], m->dims[4], B, false));
      gemm::Args& a = fp->fin.args;
      a.bias = s.scale;
      a.shift = s.shift;
      a.w1 = m->w1;
      a.b1 = m->b1;
      a.nc = m->nc;
    }
  }
#undef PLAN_TRY
  if (rc != IG_OK) {
    destroy_plan(fp);
    return rc;
  }
  *out = fp;
  return IG_OK;
}

int get_plan(ig_model* m, int batch, char* ws, ig_fwd_plan** out) {
  for (ig_fwd_plan* fp : m->plans)
    if (fp->batch == batch && fp->ws == ws) {
      fp->last_use = ++m->use_ctr;
      *out = fp;
      return IG_OK;
    }
  if (static_cast<int>(m->plans.size()) >= MAX_PLANS) {  // evict the least recently used plan
    size_t k = 0;
    for (size_t i = 1; i < m->plans.size(); ++i)
      if (m->plans[i]->last_use < m->plans[k]->last_use) k = i;
    // an executable graph may still be in flight on the caller's stream
    cudaDeviceSynchronize();
    destroy_plan(m->plans[k]);
    m->plans.erase(m->plans.begin() + k);
  }
  ig_fwd_plan* fp = nullptr;
  IG_TRY(build_plan(m, batch, ws, &fp));
  fp->last_use = ++m->use_ctr;
  m->plans.push_back(fp);
  *out = fp;
  return IG_OK;
}

// Zero borders of the padded-flat head buffers that no kernel of the forward ever writes (the final LayerNorm and
// the transposed convolutions store interior pixels only): cleared when this (batch, workspace) layout takes the
// workspace over, not once per forward (5 launches per step before).
int prepare_rings(ig_model* m, ig_fwd_plan& fp, cudaStream_t st) {
  if (m->ring_ws == fp.ws && m->ring_batch == fp.batch) return IG_OK;
  const Layout& l = fp.l;
  IG_TRY(ops::zero_ring(fp.ws + l.in0.off + static_cast<size_t>(l.in0.guard) * l.in0.C * 2, fp.batch, l.in0.Hp, l.in0.Wp,
                        l.in0.C, st));
  for (int i = 0; i < 4; ++i) {
    const Buf& tb = l.t[i];
    IG_TRY(ops::zero_ring(fp.ws + tb.off + static_cast<size_t>(tb.guard) * tb.C * 2, fp.batch, tb.Hp, tb.Wp, tb.C, st));
  }
  m->ring_ws = fp.ws;
  m->ring_batch = fp.batch;
  return IG_OK;
}

// Enqueue the kernels of one forward on `st`.  *n_kernels = kernels launched; idx_patch / idx_final = position of the
// two launches that carry caller pointers (for the graph path).
int enqueue(ig_model* m, ig_fwd_plan& fp, const FwdIO& io, cudaStream_t st, bool taps, int* n_kernels, int* idx_patch,
            int* idx_final) {
  const Layout& l = fp.l;
  char* ws = fp.ws;
  const int B = fp.batch, D = m->D, T = m->T, g = m->g, N = m->ntok;
  const int M = B * N;
  float* xres = reinterpret_cast<float*>(ws + l.x);
  bf16* xn = reinterpret_cast<bf16*>(ws + l.xn);
  bf16* att = reinterpret_cast<bf16*>(ws + l.att);
  const size_t tap_stride = static_cast<size_t>(M) * D;
  int nk = 0;
  auto tap_x = [&](int slot) -> int {
    if (!taps) return IG_OK;
    IG_CUDA_OK(cudaMemcpyAsync(m->tap_buf + slot * tap_stride, xres, tap_stride * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
    return IG_OK;
  };

  // ---- kernel 2: tubelet patch embedding (im2col-free: rows are written once, in GEMM order)
  gemm::Plan* pp = &fp.patch;
  if (io.x_dtype == IG_F32) {
    IG_TRY(ops::patchify(static_cast<const float*>(io.x), ws + l.patches, B, m->cfg.in_chans, T, m->cfg.img_size, st));
    ++nk;
  } else {
    if (fp.patch_ext_src != io.x) {
      IG_TRY(ig_make_tmap_bf16(&fp.patch_ext.tmA, io.x, static_cast<uint64_t>(B) * T * g * g, m->K0, m->K0,
                               fp.patch_ext.args.a_box_rows, gemm::BK));
      fp.patch_ext.tmAr = fp.patch_ext.tmA;   // K0 = 1536: no narrow K block
      fp.patch_ext_src = io.x;
    }
    pp = &fp.patch_ext;
  }
  IG_TRY(ops::init_cls(xres, m->cls, m->pos, B, N, D, st));
  ++nk;
  *idx_patch = nk;
  IG_TRY(gemm::launch(*pp, st));
  ++nk;
  IG_TRY(tap_x(0));

  // ---- kernel 3: encoder blocks
  for (int i = 0; i < m->L; ++i) {
    const LayerW& w = m->layers[i];
    const gemm::Plan* p = &fp.enc[static_cast<size_t>(4) * i];
    IG_TRY(ops::layernorm(xres, w.ln1g, w.ln1b, xn, M, D, 0, 0, 0, 0, 0, st));
    IG_TRY(gemm::launch(p[0], st));
    IG_TRY(ops::attention_planned(fp.tmq, fp.tmkv, att, B, N, m->heads, st));
    IG_TRY(gemm::launch(p[1], st));
    IG_TRY(ops::layernorm(xres, w.ln2g, w.ln2b, xn, M, D, 0, 0, 0, 0, 0, st));
    IG_TRY(gemm::launch(p[2], st));
    IG_TRY(gemm::launch(p[3], st));
    nk += 7;
    IG_TRY(tap_x(1 + i));
  }

  // ---- final norm, written straight into the head's padded-flat input (cls dropped)
  IG_TRY(ops::layernorm(xres, m->norm_g, m->norm_b, ws + l.in0.off, M, D, 1, N, T, g, l.in0.guard, st));
  ++nk;
  if (taps) {  // the same rows in token order, cls included (oracle tap "tokens")
    IG_TRY(ops::layernorm(xres, m->norm_g, m->norm_b, xn, M, D, 0, 0, 0, 0, 0, st));
    IG_TRY(ops::cvt_f32(xn, m->tap_buf + (1 + m->L) * tap_stride, static_cast<int64_t>(tap_stride), st));
  }
  if (io.feats) {
    IG_TRY(ops::unpad_to_nchw(ws + l.in0.off, io.feats, B, l.in0.Hp, l.in0.Wp, l.in0.C, l.in0.guard, T, st));
    ++nk;
  }

  // ---- kernel 4: segmentation head
  for (int i = 0; i < 4; ++i) {
    IG_TRY(gemm::launch(fp.convt[i], st));
    ++nk;
    if (i < 3) {
      IG_TRY(gemm::launch(fp.conv[i], st));
      ++nk;
    } else {
      gemm::Args& a = fp.fin.args;
      a.logits = io.logits;
      a.argmax = (m->nc > 1) ? io.argmax : nullptr;
      a.prob1 = io.prob1;
      *idx_final = nk;
      IG_TRY(gemm::launch(fp.fin, st));
      ++nk;
    }
  }
  *n_kernels = nk;
  return IG_OK;
}

bool graphs_disabled_by_env() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("IG_NO_GRAPH");
    v = (e && e[0] && e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

// Capture one forward into a graph.  Any failure of the graph API leaves the model on the eager path for good
// (graphs_ok = false, message kept for ig_model_graph_status) -- never an error of the forward itself.
int capture_graph(ig_model* m, ig_fwd_plan& fp, const FwdIO& io, unsigned flags, GraphEntry** out) {
  *out = nullptr;
  auto give_up = [&](const char* what, cudaError_t e) {
    snprintf(m->graph_note, sizeof(m->graph_note), "%s: %s", what, cudaGetErrorString(e));
    m->graphs_ok = false;
    cudaGetLastError();
    return IG_OK;
  };
  cudaError_t e;
  if (!m->cap_stream) {
    e = cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return give_up("cudaStreamCreateWithFlags", e);
  }
  // Nothing executes on cap_stream: it only records the launches (the caller's stream may be the legacy default
  // stream, which cannot be captured).
  e = cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeRelaxed);
  if (e != cudaSuccess) return give_up("cudaStreamBeginCapture", e);
  int nk = 0, ip = -1, ifin = -1;
  const int rc = enqueue(m, fp, io, m->cap_stream, false, &nk, &ip, &ifin);
  cudaGraph_t graph = nullptr;
  e = cudaStreamEndCapture(m->cap_stream, &graph);
  if (rc != IG_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;  // a real planning / launch-configuration error: report it
  }
  if (e != cudaSuccess || !graph) return give_up("cudaStreamEndCapture", e);
  // the capture of a single stream is a chain: walk it from the root to number the kernel nodes
  std::vector<cudaGraphNode_t> chain;
  {
    size_t nroot = 0;
    e = cudaGraphGetRootNodes(graph, nullptr, &nroot);
    cudaGraphNode_t cur = nullptr;
    if (e == cudaSuccess && nroot == 1) {
      size_t one = 1;
      e = cudaGraphGetRootNodes(graph, &cur, &one);
    } else if (e == cudaSuccess) {
      e = cudaErrorUnknown;
    }
    while (e == cudaSuccess && cur) {
      cudaGraphNodeType ty;
      e = cudaGraphNodeGetType(cur, &ty);
      if (e != cudaSuccess) break;
      if (ty != cudaGraphNodeTypeKernel) { e = cudaErrorUnknown; break; }
      chain.push_back(cur);
      // _v2: the edges of programmatic dependent launches carry edge data (the plain query refuses them)
      size_t nd = 0;
      e = cudaGraphNodeGetDependentNodes_v2(cur, nullptr, nullptr, &nd);
      if (e != cudaSuccess) break;
      if (nd == 0) break;
      if (nd != 1) { e = cudaErrorUnknown; break; }
      cudaGraphNode_t nxt = nullptr;
      cudaGraphEdgeData ed;
      size_t one = 1;
      e = cudaGraphNodeGetDependentNodes_v2(cur, &nxt, &ed, &one);
      cur = nxt;
    }
  }
  if (e != cudaSuccess || static_cast<int>(chain.size()) != nk || ip < 0 || ifin < 0) {
    cudaGraphDestroy(graph);
    return give_up("captured graph is not the expected kernel chain", e == cudaSuccess ? cudaErrorUnknown : e);
  }
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  if (e != cudaSuccess) {
    cudaGraphDestroy(graph);
    return give_up("cudaGraphInstantiate", e);
  }
  GraphEntry ge;
  ge.flags = flags;
  ge.graph = graph;
  ge.exec = exec;
  ge.n_patch = chain[ip];
  ge.n_final = chain[ifin];
  ge.io = io;
  fp.graphs.push_back(ge);
  *out = &fp.graphs.back();
  m->graph_kernels = nk;
  return IG_OK;
}

// Re-point the two kernel nodes that carry caller pointers.  func / grid / block / shared memory are read back from
// the captured node; the argument list (tmA, tmB, tmAr, tmBr, tmO, Args) is gemm_kernel's.
int update_graph(ig_model* m, ig_fwd_plan& fp, GraphEntry& ge, const FwdIO& io) {
  cudaError_t e = cudaSuccess;
  if (io.x != ge.io.x) {
    if (fp.patch_ext_src != io.x) {
      IG_TRY(ig_make_tmap_bf16(&fp.patch_ext.tmA, io.x, static_cast<uint64_t>(fp.batch) * m->T * m->g * m->g, m->K0,
                               m->K0, fp.patch_ext.args.a_box_rows, gemm::BK));
      fp.patch_ext.tmAr = fp.patch_ext.tmA;
      fp.patch_ext_src = io.x;
    }
    cudaKernelNodeParams kp;
    e = cudaGraphKernelNodeGetParams(ge.n_patch, &kp);
    if (e == cudaSuccess) {
      void* args[6] = {&fp.patch_ext.tmA, &fp.patch_ext.tmB, &fp.patch_ext.tmAr, &fp.patch_ext.tmBr, &fp.patch_ext.tmO,
                       &fp.patch_ext.args};
      kp.kernelParams = args;
      kp.extra = nullptr;
      e = cudaGraphExecKernelNodeSetParams(ge.exec, ge.n_patch, &kp);
    }
  }
  if (e == cudaSuccess && (io.logits != ge.io.logits || io.argmax != ge.io.argmax || io.prob1 != ge.io.prob1)) {
    gemm::Args& a = fp.fin.args;
    a.logits = io.logits;
    a.argmax = (m->nc > 1) ? io.argmax : nullptr;
    a.prob1 = io.prob1;
    cudaKernelNodeParams kp;
    e = cudaGraphKernelNodeGetParams(ge.n_final, &kp);
    if (e == cudaSuccess) {
      void* args[6] = {&fp.fin.tmA, &fp.fin.tmB, &fp.fin.tmAr, &fp.fin.tmBr, &fp.fin.tmO, &fp.fin.args};
      kp.kernelParams = args;
      kp.extra = nullptr;
      e = cudaGraphExecKernelNodeSetParams(ge.exec, ge.n_final, &kp);
    }
  }
  if (e != cudaSuccess) {
    snprintf(m->graph_note, sizeof(m->graph_note), "graph node update: %s", cudaGetErrorString(e));
    m->graphs_ok = false;
    cudaGetLastError();
    return IG_ESTATE;  // caller falls back to the eager path for this call
  }
  ge.io = io;
  return IG_OK;
}

}  // namespace

static int forward_impl(ig_model* m, const void* x, int x_dtype, int batch, float* logits, int8_t* argmax,
                        float* prob1, float* feats, void* workspace, size_t workspace_bytes, void* stream) {
  IG_TRY(ig_check_device());
  IG_REQUIRE(m && x && workspace, IG_EINVAL, "ig_model_forward: null pointer");
  IG_REQUIRE(m->finalized, IG_ESTATE, "ig_model_forward: call ig_model_finalize after loading weights");
  IG_REQUIRE(batch >= 1, IG_ESHAPE, "ig_model_forward: batch %d", batch);
  IG_REQUIRE(x_dtype == IG_F32 || x_dtype == IG_BF16, IG_EINVAL, "ig_model_forward: x_dtype must be IG_F32 or IG_BF16");
  IG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, IG_EINVAL, "workspace must be 1024-byte aligned");
  IG_REQUIRE(static_cast<int64_t>(batch) * (226 * 226) < (1ll << 31) - 4096, IG_ESHAPE, "batch %d too large", batch);
  {
    int dev = -1;
    IG_CUDA_OK(cudaGetDevice(&dev));
    IG_REQUIRE(dev == m->device, IG_ESTATE, "ig_model_forward: model lives on device %d, current device is %d", m->device, dev);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  ig_fwd_plan* fp = nullptr;
  IG_TRY(get_plan(m, batch, ws, &fp));
  IG_REQUIRE(workspace_bytes >= fp->l.total, IG_ENOMEM, "workspace too small: %zu < %zu bytes", workspace_bytes, fp->l.total);
  const bool taps = m->tap_buf != nullptr;
  if (taps) {
    const size_t need = static_cast<size_t>(m->L + 2) * batch * m->ntok * m->D;
    IG_REQUIRE(m->tap_elems >= need, IG_ENOMEM, "tap buffer holds %zu floats, batch %d needs %zu", m->tap_elems, batch, need);
    m->tap_batch = batch;
  }
  IG_TRY(prepare_rings(m, *fp, st));
  const FwdIO io{x, x_dtype, logits, argmax, prob1, feats};
  int nk = 0, ip = -1, ifin = -1;

  bool use_graph = m->graphs_ok && !graphs_disabled_by_env() && !taps && !ig::prof_enabled() &&
                   x_dtype == IG_BF16 && feats == nullptr;
  if (use_graph) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      cudaGetLastError();
      use_graph = false;  // the caller is capturing: our launches become nodes of THEIR graph
    }
  }
  if (use_graph) {
    const unsigned flags = (logits ? 1u : 0u) | (argmax ? 2u : 0u) | (prob1 ? 4u : 0u);
    GraphEntry* ge = nullptr;
    for (GraphEntry& g : fp->graphs)
      if (g.flags == flags) ge = &g;
    if (!ge) IG_TRY(capture_graph(m, *fp, io, flags, &ge));
    if (ge) {
      int rc = IG_OK;
      if (io.x != ge->io.x || io.logits != ge->io.logits || io.argmax != ge->io.argmax || io.prob1 != ge->io.prob1)
        rc = update_graph(m, *fp, *ge, io);
      if (rc == IG_OK) {
        const cudaError_t e = cudaGraphLaunch(ge->exec, st);
        if (e == cudaSuccess) {
          m->last_path = 1;
          return IG_OK;
        }
        snprintf(m->graph_note, sizeof(m->graph_note), "cudaGraphLaunch: %s", cudaGetErrorString(e));
        m->graphs_ok = false;
        cudaGetLastError();
      }
    }
  }
  m->last_path = 0;
  return enqueue(m, *fp, io, st, taps, &nk, &ip, &ifin);
}

extern "C" int ig_model_forward(ig_model* m, const void* x, int x_dtype, int batch, float* logits, int8_t* argmax,
                                float* feats, void* workspace, size_t workspace_bytes, void* stream) {
  return forward_impl(m, x, x_dtype, batch, logits, argmax, nullptr, feats, workspace, workspace_bytes, stream);
}

extern "C" int ig_model_predict_proba(ig_model* m, const void* x, int x_dtype, int batch, float* prob_pos,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  IG_REQUIRE(m && prob_pos, IG_EINVAL, "ig_model_predict_proba: null pointer");
  IG_REQUIRE(m->nc >= 2, IG_ESHAPE, "ig_model_predict_proba: needs a classification head (num_classes >= 2)");
  return forward_impl(m, x, x_dtype, batch, nullptr, nullptr, prob_pos, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int ig_model_reset_cache(ig_model* m) {
  IG_REQUIRE(m, IG_EINVAL, "ig_model_reset_cache: null model");
  if (!m->plans.empty()) cudaDeviceSynchronize();
  for (ig_fwd_plan* fp : m->plans) destroy_plan(fp);
  m->plans.clear();
  m->ring_ws = nullptr;
  m->ring_batch = 0;
  return IG_OK;
}

extern "C" int ig_model_set_tap_buffer(ig_model* m, float* buf, size_t elems) {
  IG_REQUIRE(m, IG_EINVAL, "ig_model_set_tap_buffer: null model");
  m->tap_buf = buf;
  m->tap_elems = buf ? elems : 0;
  m->tap_batch = 0;
  return IG_OK;
}

extern "C" int ig_model_graph_status(const ig_model* m, int* last_forward_was_graph, int* kernels_in_graph,
                                     char* note, size_t note_bytes) {
  IG_REQUIRE(m, IG_EINVAL, "ig_model_graph_status: null model");
  if (last_forward_was_graph) *last_forward_was_graph = m->last_path;
  if (kernels_in_graph) *kernels_in_graph = m->graph_kernels;
  if (note && note_bytes) {
    strncpy(note, m->graph_note, note_bytes - 1);
    note[note_bytes - 1] = 0;
  }
  return m->graphs_ok ? 1 : 0;
}

extern "C" int ig_model_debug_tap(ig_model* m, const char* name, int batch, void* workspace, float* dst,
                                  size_t dst_elems, void* stream) {
  IG_REQUIRE(m && name && workspace && dst, IG_EINVAL, "ig_model_debug_tap: null pointer");
  const Layout l = make_layout(m, batch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  if (!strcmp(name, "x")) {  // residual stream after the last block, [B, N, D]
    const size_t n = static_cast<size_t>(batch) * m->ntok * m->D;
    IG_REQUIRE(dst_elems >= n, IG_ENOMEM, "tap x needs %zu elements", n);
    IG_CUDA_OK(cudaMemcpyAsync(dst, ws + l.x, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return IG_OK;
  }
  // encoder taps [B, N, D] recorded by the last forward into the buffer of ig_model_set_tap_buffer
  int slot = -1;
  if (!strcmp(name, "embed")) slot = 0;
  else if (!strcmp(name, "tokens")) slot = 1 + m->L;
  else if (starts_with(name, "block")) {
    char* end = nullptr;
    const long bi = strtol(name + 5, &end, 10);
    IG_REQUIRE(end && *end == 0 && end != name + 5 && bi >= 0 && bi < m->L, IG_EINVAL, "unknown tap '%s' (block0..block%d)", name, m->L - 1);
    slot = 1 + static_cast<int>(bi);
  }
  if (slot >= 0) {
    IG_REQUIRE(m->tap_buf != nullptr && m->tap_batch == batch, IG_ESTATE,
               "tap '%s': call ig_model_set_tap_buffer before the forward (recorded batch %d, asked %d)", name, m->tap_batch, batch);
    const size_t n = static_cast<size_t>(batch) * m->ntok * m->D;
    IG_REQUIRE(dst_elems >= n, IG_ENOMEM, "tap %s needs %zu elements", name, n);
    IG_CUDA_OK(cudaMemcpyAsync(dst, m->tap_buf + slot * n, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return IG_OK;
  }
  const Buf* b = nullptr;
  int permT = 1;
  if (!strcmp(name, "feat")) { b = &l.in0; permT = m->T; }
  else if (starts_with(name, "convt") && name[5] >= '0' && name[5] <= '3') b = &l.t[name[5] - '0'];
  else if (starts_with(name, "stage") && name[5] >= '0' && name[5] <= '2') b = &l.a[name[5] - '0'];
  IG_REQUIRE(b != nullptr, IG_EINVAL, "unknown tap '%s' (x, embed, block<i>, tokens, feat, convt0-3, stage0-2)", name);
  const size_t n = static_cast<size_t>(batch) * b->C * (b->Hp - 2) * (b->Wp - 2);
  IG_REQUIRE(dst_elems >= n, IG_ENOMEM, "tap %s needs %zu elements", name, n);
  return ops::unpad_to_nchw(ws + b->off, dst, batch, b->Hp, b->Wp, b->C, b->guard, permT, st);
}
