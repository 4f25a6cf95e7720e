// Native runtime: composes the kernels into the reference's four seam operations, batched.
//   md_vision_encode   <- MoondreamModel._vis_enc    (vision.py:64-74)
//   md_vision_project  <- reconstruct_from_crops + _vis_proj (image_crops.py:170-231, vision.py:77-89)
//   md_text_prefill    <- _prefill                   (text.py:128-160)
//   md_text_decode_step + md_lm_head_argmax <- _decode_one_tok (moondream.py:183-192)
// The model object stores pointers only (host memory); device memory belongs to the caller.
#include "engine.h"

#include <new>

namespace md {

static inline char* align_up(char* p, size_t a = 256) {
  return reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + a - 1) & ~(uintptr_t)(a - 1));
}
static inline long long pad256(long long b) { return (b + 255) & ~255LL; }

int model_num_weights(const md_dims& d) { return 3 + 12 * d.vis_layers + 6 + 1 + 10 * d.txt_layers + 4 + 14; }

int model_create(const md_dims& d, const void* const* w, int n, const void* lut, const float* rope,
                 Model** out) {
  if (n != model_num_weights(d)) return set_error("md_model_create: wrong number of weight tensors");
  if (d.vis_dim != d.vis_heads * 72) return set_error("md_model_create: vision head_dim must be 72");
  if (d.txt_dim != d.txt_heads * 64) return set_error("md_model_create: text head_dim must be 64");
  if (d.txt_kv_heads < 0 || (d.txt_kv_heads > 0 && d.txt_heads % d.txt_kv_heads))
    return set_error("md_model_create: txt_heads must be a multiple of txt_kv_heads");
  if (d.patch_k % 8 || d.vis_ff % 8 || d.vis_dim % 8 || d.txt_dim % 8 || d.txt_ff % 8 || d.vocab % 8 ||
      d.proj_inner % 8 || d.reg_inner % 8)
    return set_error("md_model_create: every GEMM dimension must be a multiple of 8 (pad when preparing)");
  if (d.grid * d.patch != d.crop || d.prefix_len != d.grid * d.grid + 1)
    return set_error("md_model_create: inconsistent crop geometry");
  for (int i = 0; i < n; ++i)
    if (!w[i] || (reinterpret_cast<uintptr_t>(w[i]) & 15))
      return set_error("md_model_create: weight pointers must be non-null and 16-byte aligned");
  Model* m = new (std::nothrow) Model();
  if (!m) return set_error("md_model_create: out of host memory");
  m->d = d;
  if (m->d.txt_kv_heads == 0) m->d.txt_kv_heads = d.txt_heads;      // 0 = multi-head attention
  m->lut = reinterpret_cast<const bf16*>(lut);
  m->rope = rope;
  int k = 0;
  auto next = [&]() { return reinterpret_cast<const bf16*>(w[k++]); };
  auto lin = [&]() { Lin l; l.w = next(); l.b = next(); l.ld = 0; return l; };
  m->pos_emb = next();
  m->patch_emb = lin();
  m->vis.resize(d.vis_layers);
  for (auto& b : m->vis) { b.ln1 = lin(); b.qkv = lin(); b.proj = lin(); b.ln2 = lin(); b.fc1 = lin(); b.fc2 = lin(); }
  m->vis_post_ln = lin();
  m->proj_fc1 = lin();
  m->proj_fc2 = lin();
  m->wte = next();
  m->txt.resize(d.txt_layers);
  for (auto& b : m->txt) { b.ln = lin(); b.qkv = lin(); b.proj = lin(); b.fc1 = lin(); b.fc2 = lin(); }
  m->txt_post_ln = lin();
  m->lm_head = lin();
  m->coord_features = next();
  m->size_features = next();
  m->coord_enc = lin();
  m->coord_dec1 = lin();
  m->coord_dec2 = lin();
  m->size_enc = lin();
  m->size_dec1 = lin();
  m->size_dec2 = lin();
  // natural row pitches (= fan-in) ...
  m->patch_emb.ld = d.patch_k;
  for (auto& b : m->vis) { b.qkv.ld = d.vis_dim; b.proj.ld = d.vis_dim; b.fc1.ld = d.vis_dim; b.fc2.ld = d.vis_ff; }
  m->proj_fc1.ld = 2 * d.vis_dim; m->proj_fc2.ld = d.proj_inner;
  m->lm_head.ld = d.txt_dim;
  m->coord_enc.ld = d.coord_feat; m->coord_dec1.ld = d.txt_dim; m->coord_dec2.ld = d.reg_inner;
  m->size_enc.ld = d.size_feat; m->size_dec1.ld = d.txt_dim; m->size_dec2.ld = d.reg_inner;
  const long long D = d.txt_dim, FF = d.txt_ff;
  const long long QKV = D + 2LL * m->d.txt_kv_heads * 64;            // rows of qkv.weight (text.py:36-38)
  for (auto& b : m->txt) {
    b.qkv.ld = D; b.fc1.ld = D;
    if (d.txt_fused) {
      // ... except the fused decode layout: W1 = [qkv ; fc1] rows, W2 = [proj | fc2] columns
      if (b.fc1.w != b.qkv.w + QKV * D || b.fc1.b != b.qkv.b + QKV || b.fc2.w != b.proj.w + D) {
        delete m;
        return set_error("md_model_create: txt_fused is set but the decoder weights are not views of "
                         "[qkv;fc1] / [proj|fc2] buffers");
      }
      b.proj.ld = D + FF; b.fc2.ld = D + FF;
    } else {
      b.proj.ld = D; b.fc2.ld = FF;
    }
  }
  *out = m;
  return 0;
}

// Quantised decoder blocks (layers.py:38-110): after this call for EVERY block, decode streams the packed bytes and
// prefill dequantises block i into the scratch the bf16 pointers alias (they must all name the same W1 / W2 buffers).
int model_set_quantized_block(Model& m, int layer, int bits, const void* w1q, const float* w1_scale, const float* w1_zero,
                              const void* w2q, const float* w2_scale, const float* w2_zero) {
  const md_dims& d = m.d;
  if (layer < 0 || layer >= d.txt_layers) return set_error("md_model_set_quantized_block: layer out of range");
  if (bits != 4 && bits != 8) return set_error("md_model_set_quantized_block: bits must be 4 or 8");
  if (!d.txt_fused) return set_error("md_model_set_quantized_block: needs the fused decode layout");
  if (d.txt_dim % 128 || d.txt_ff % 128) return set_error("md_model_set_quantized_block: txt_dim and txt_ff must be multiples of the group size 128");
  if (!w1q || !w1_scale || !w1_zero || !w2q || !w2_scale || !w2_zero) return set_error("md_model_set_quantized_block: null pointer");
  for (const void* p : {w1q, w2q, static_cast<const void*>(w1_scale), static_cast<const void*>(w1_zero),
                        static_cast<const void*>(w2_scale), static_cast<const void*>(w2_zero)})
    if (reinterpret_cast<uintptr_t>(p) & 15) return set_error("md_model_set_quantized_block: tensors must be 16-byte aligned");
  if (m.txt[layer].qkv.w != m.txt[0].qkv.w || m.txt[layer].proj.w != m.txt[0].proj.w)
    return set_error("md_model_set_quantized_block: the bf16 weight pointers of every decoder block must alias one scratch pair");
  if (m.tq.empty()) m.tq.resize(d.txt_layers);
  QuantBlock& q = m.tq[layer];
  q.bits = bits;
  q.w1q = reinterpret_cast<const uint8_t*>(w1q); q.w1s = w1_scale; q.w1z = w1_zero;
  q.w2q = reinterpret_cast<const uint8_t*>(w2q); q.w2s = w2_scale; q.w2z = w2_zero;
  return 0;
}

static int quant_ready(const Model& m) {
  for (const QuantBlock& q : m.tq)
    if (!q.bits) return set_error("quantised model: md_model_set_quantized_block has not been called for every block");
  return 0;
}

// prefill under quantised weights: rebuild block i's bf16 W1 / W2 in the shared scratch (0.1 GB written per block
// for the 2B, ~30 us; the row-form GEMMs that follow are compute-bound)
static int dequant_block(Model& m, int i, cudaStream_t st) {
  if (m.tq.empty()) return 0;
  const md_dims& d = m.d;
  const QuantBlock& q = m.tq[i];
  const int D = d.txt_dim, FF = d.txt_ff, QKV = D + 2 * d.txt_kv_heads * 64;
  const TxtBlock& b = m.txt[i];
  if (dequant_weights(q.bits, q.w1q, q.w1s, q.w1z, QKV + FF, D, const_cast<bf16*>(b.qkv.w), D, st)) return 1;
  return dequant_weights(q.bits, q.w2q, q.w2s, q.w2z, D, D + FF, const_cast<bf16*>(b.proj.w), D + FF, st);
}

// ------------------------------------------------------------------------------------------------
// vision encoder
// ------------------------------------------------------------------------------------------------
long long vision_encode_ws_bytes(const Model& m, int n_crops) {
  const md_dims& d = m.d;
  const long long T = static_cast<long long>(n_crops) * d.grid * d.grid;
  // patches | x | ln | qkv | attn | hidden(ff)
  return pad256(T * d.patch_k * 2) + 2 * pad256(T * d.vis_dim * 2) + pad256(T * 3 * d.vis_dim * 2) +
         pad256(T * d.vis_dim * 2) + pad256(T * d.vis_ff * 2) + 4096;
}

int vision_encode(Model& m, const uint8_t* crops, int n_crops, bf16* feats, void* ws, cudaStream_t st) {
  const md_dims& d = m.d;
  if (n_crops <= 0) return set_error("md_vision_encode: empty batch");
  const int tok = d.grid * d.grid;
  const int T = n_crops * tok;
  const int D = d.vis_dim;
  char* p = align_up(reinterpret_cast<char*>(ws));
  bf16* patches = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * d.patch_k * 2);
  bf16* x = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * D * 2);
  bf16* ln = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * D * 2);
  bf16* qkv = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * 3 * D * 2);
  bf16* att = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * D * 2);
  bf16* hid = reinterpret_cast<bf16*>(p);

  // x = bf16(bf16(patches W^T + b) + pos_emb[token])           vision.py:67-68
  if (!g_patch_embed_unfused && 3 * d.patch * d.patch <= 640) {
    // one im2col-fused tcgen05 GEMM straight from the uint8 crops (patch_embed.cu): no patch matrix in HBM
    if (patch_embed_fused(crops, n_crops, d.crop, d.patch, m.lut, m.patch_emb.w, d.patch_k, m.patch_emb.b, m.pos_emb, D, x,
                          st)) return 1;
  } else {
    if (patchify(crops, n_crops, d.crop, d.patch, d.patch_k, m.lut, patches, st)) return 1;
    if (gemm_rowform(patches, d.patch_k, m.patch_emb.w, d.patch_k, T, D, d.patch_k, EPI_BIAS_RESIDUAL,
                     m.patch_emb.b, m.pos_emb, D, tok, x, D, 0, 0, 0, st)) return 1;
  }
  for (int i = 0; i < d.vis_layers; ++i) {
    const VisBlock& b = m.vis[i];
    // x = x + attn(ln1(x))                                      vision.py:70, layers.py:155-166
    if (layernorm(x, D, b.ln1.w, b.ln1.b, ln, D, T, D, 1e-5f, st)) return 1;
    if (gemm_rowform(ln, D, b.qkv.w, D, T, 3 * D, D, EPI_BIAS, b.qkv.b, nullptr, 0, 0, qkv, 3 * D, 0, 0, 0, st)) return 1;
    if (g_attention_impl == 1 ? vit_attention(qkv, n_crops, tok, d.vis_heads, att, st)
                              : vit_attention_tc(qkv, n_crops, tok, d.vis_heads, att, st)) return 1;
    if (gemm_rowform(att, D, b.proj.w, D, T, D, D, EPI_BIAS_RESIDUAL, b.proj.b, x, D, 0, x, D, 0, 0, 0, st)) return 1;
    // x = x + fc2(gelu(fc1(ln2(x))))                            vision.py:71, layers.py:129-146
    if (layernorm(x, D, b.ln2.w, b.ln2.b, ln, D, T, D, 1e-5f, st)) return 1;
    if (gemm_rowform(ln, D, b.fc1.w, D, T, d.vis_ff, D, EPI_BIAS_GELU, b.fc1.b, nullptr, 0, 0, hid, d.vis_ff, 0, 0, 0, st)) return 1;
    if (gemm_rowform(hid, d.vis_ff, b.fc2.w, d.vis_ff, T, D, d.vis_ff, EPI_BIAS_RESIDUAL, b.fc2.b, x, D, 0, x, D, 0, 0, 0, st)) return 1;
  }
  return layernorm(x, D, m.vis_post_ln.w, m.vis_post_ln.b, feats, D, T, D, 1e-5f, st);
}

// ------------------------------------------------------------------------------------------------
// stitch + pool + projection MLP
// ------------------------------------------------------------------------------------------------
long long vision_project_ws_bytes(const Model& m, int n_images) {
  const md_dims& d = m.d;
  const long long T = static_cast<long long>(n_images) * d.grid * d.grid;
  return pad256(T * 2 * d.vis_dim * 2) + pad256(T * d.proj_inner * 2) + 4096;
}

int vision_project(Model& m, const bf16* feats, const int* crop_offsets, const int* tilings, int n_images,
                   bf16* embeds, int rows_per_image, void* ws, cudaStream_t st) {
  const md_dims& d = m.d;
  if (n_images <= 0) return set_error("md_vision_project: empty batch");
  const int tok = d.grid * d.grid;
  const int T = n_images * tok;
  char* p = align_up(reinterpret_cast<char*>(ws));
  bf16* cat = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * 2 * d.vis_dim * 2);
  bf16* hid = reinterpret_cast<bf16*>(p);
  if (stitch_pool_concat(feats, crop_offsets, tilings, n_images, d.grid, d.margin, d.vis_dim, cat, st)) return 1;
  if (gemm_rowform(cat, 2 * d.vis_dim, m.proj_fc1.w, 2 * d.vis_dim, T, d.proj_inner, 2 * d.vis_dim,
                   EPI_BIAS_GELU, m.proj_fc1.b, nullptr, 0, 0, hid, d.proj_inner, 0, 0, 0, st)) return 1;
  // rows of image i land at i*rows_per_image + 1 .. (row i*rows_per_image is the BOS embedding; rows after the
  // image tokens may hold the prompt embeddings when image and prompt are prefilled in one pass)
  if (rows_per_image < d.prefix_len) rows_per_image = d.prefix_len;
  return gemm_rowform(hid, d.proj_inner, m.proj_fc2.w, d.proj_inner, T, d.txt_dim, d.proj_inner, EPI_BIAS,
                      m.proj_fc2.b, nullptr, 0, 0, embeds, d.txt_dim, tok, rows_per_image, 1, st);
}

// `_vis_proj(g, r)` itself (vision.py:77-89) for ONE image whose local features are already stitched: the seam adapter's
// entry point (moondream_b200/seam.py); the batched product path is vision_project above.
int vision_project_stitched(Model& m, const bf16* global_feats, const bf16* stitched, int H, int W, bf16* out, void* ws,
                            cudaStream_t st) {
  const md_dims& d = m.d;
  const int T = d.grid * d.grid;
  char* p = align_up(reinterpret_cast<char*>(ws));
  bf16* cat = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * 2 * d.vis_dim * 2);
  bf16* hid = reinterpret_cast<bf16*>(p);
  if (pool_concat(global_feats, stitched, H, W, d.grid, d.vis_dim, cat, st)) return 1;
  if (gemm_rowform(cat, 2 * d.vis_dim, m.proj_fc1.w, 2 * d.vis_dim, T, d.proj_inner, 2 * d.vis_dim,
                   EPI_BIAS_GELU, m.proj_fc1.b, nullptr, 0, 0, hid, d.proj_inner, 0, 0, 0, st)) return 1;
  return gemm_rowform(hid, d.proj_inner, m.proj_fc2.w, d.proj_inner, T, d.txt_dim, d.proj_inner, EPI_BIAS,
                      m.proj_fc2.b, nullptr, 0, 0, out, d.txt_dim, 0, 0, 0, st);
}

// ------------------------------------------------------------------------------------------------
// text decoder: prefill
// ------------------------------------------------------------------------------------------------
long long text_prefill_ws_bytes(const Model& m, int T) {
  const md_dims& d = m.d;
  // ln | q | attn | tmp | hidden(ff)
  return pad256(1LL * T * d.txt_dim * 2) * 4 + pad256(1LL * T * d.txt_ff * 2) + 4096;
}

int text_prefill(Model& m, bf16* x, int T, const int* q_offsets, const int* start_pos, int n_seqs,
                 int max_q, int prefix_len, const md_kv& kv, void* ws, cudaStream_t st) {
  const md_dims& d = m.d;
  if (T <= 0 || n_seqs <= 0) return set_error("md_text_prefill: empty batch");
  if (prefix_len < 0) prefix_len = d.prefix_len;     // 0: pure causal mask (text-only query, moondream.py:565-574)
  const int D = d.txt_dim, H = d.txt_heads, KVH = d.txt_kv_heads;
  if (kv.n_kv_heads && kv.n_kv_heads != KVH) return set_error("md_text_prefill: the KV pool's head count differs from the model's");
  char* p = align_up(reinterpret_cast<char*>(ws));
  bf16* ln = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * D * 2);
  bf16* q = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * D * 2);
  bf16* att = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * D * 2);
  bf16* tmp = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * D * 2);
  bf16* hid = reinterpret_cast<bf16*>(p);
  bf16* pool = reinterpret_cast<bf16*>(kv.pool);
  if (!m.tq.empty() && quant_ready(m)) return 1;
  for (int i = 0; i < d.txt_layers; ++i) {
    const TxtBlock& b = m.txt[i];
    if (dequant_block(m, i, st)) return 1;
    // l = ln(x); x = x + attn(l) + mlp(l)                       text.py:145-158
    if (layernorm(x, D, b.ln.w, b.ln.b, ln, D, T, D, 1e-5f, st)) return 1;
    // QKV projection with bias, RoPE and the KV-page write in the GEMM epilogue (no qkv round trip through HBM)
    RopeEpilogue re{};
    re.D = D; re.n_heads = H; re.n_kv_heads = KVH; re.n_seqs = n_seqs; re.q_offsets = q_offsets; re.start_pos = start_pos;
    re.freqs = m.rope; re.q_out = q; re.kv_pool = pool; re.n_pages = kv.n_pages; re.block_tables = kv.block_tables;
    re.max_blocks = kv.max_blocks; re.layer = i;
    if (gemm_rowform_qkv_rope(ln, D, b.qkv.w, b.qkv.ld, T, D, b.qkv.b, re, st)) return 1;
    if (g_attention_impl == 1) {
      if (KVH != H) return set_error("md_text_prefill: the legacy mma.sync attention has no grouped-query path");
      if (prefill_attention(q, H, q_offsets, start_pos, n_seqs, max_q, prefix_len, pool, kv.n_pages,
                            kv.block_tables, kv.max_blocks, i, att, st)) return 1;
    } else {
      if (prefill_attention_tc(q, H, KVH, T, q_offsets, start_pos, n_seqs, max_q, prefix_len, pool, kv.n_pages,
                               kv.n_layers, kv.block_tables, kv.max_blocks, i, att, st)) return 1;
    }
    // tmp = bf16(x + bf16(proj(att)))  -- the reference adds l_attn first, then l_mlp (text.py:158)
    if (gemm_rowform(att, D, b.proj.w, b.proj.ld, T, D, D, EPI_BIAS_RESIDUAL, b.proj.b, x, D, 0, tmp, D, 0, 0, 0, st)) return 1;
    if (gemm_rowform(ln, D, b.fc1.w, b.fc1.ld, T, d.txt_ff, D, EPI_BIAS_GELU, b.fc1.b, nullptr, 0, 0, hid, d.txt_ff, 0, 0, 0, st)) return 1;
    if (gemm_rowform(hid, d.txt_ff, b.fc2.w, b.fc2.ld, T, D, d.txt_ff, EPI_BIAS_RESIDUAL, b.fc2.b, tmp, D, 0, x, D, 0, 0, 0, st)) return 1;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// text decoder with a LoRA variant (lora.py:55-79; text.py:31-32,54-56; layers.py:131-143): every Linear of a block
// gets `+ B (A x)` with the reference's roundings — the adapter product is a bf16 tensor added to the bf16 Linear
// output — and, as in the reference, the proj adapters are fed the block INPUT.  Composition of the existing
// kernels: two skinny row-form GEMMs per adapter (rank r in K, then in N), the main GEMM with the adapter product as
// its epilogue residual, RoPE + KV write as its own kernel (the fused QKV epilogue has no residual slot), and two
// small elementwise kernels where a block needs a second sum.  The same entry point serves decode under a variant
// (one row per sequence): variants are the rare path, the fused weight-stream step stays adapter-free.
// ------------------------------------------------------------------------------------------------
long long text_prefill_lora_ws_bytes(const Model& m, int T, int rank) {
  const md_dims& d = m.d;
  const long long wide = d.txt_ff > 3 * d.txt_dim ? d.txt_ff : 3 * d.txt_dim;
  // ln | q | att | attn_out | mlp_out | t | u | qkv | pre | hid
  return pad256(1LL * T * d.txt_dim * 2) * 5 + pad256(1LL * T * rank * 2) + pad256(1LL * T * wide * 2) +
         pad256(1LL * T * 3 * d.txt_dim * 2) + pad256(1LL * T * d.txt_ff * 2) * 2 + 4096;
}

int text_prefill_lora(Model& m, bf16* x, int T, const int* q_offsets, const int* start_pos, int n_seqs, int max_q,
                      int prefix_len, const md_kv& kv, const void* const* lora, int rank, void* ws, cudaStream_t st) {
  const md_dims& d = m.d;
  if (T <= 0 || n_seqs <= 0) return set_error("md_text_prefill_lora: empty batch");
  if (rank <= 0 || rank % 8) return set_error("md_text_prefill_lora: the adapter rank must be a positive multiple of 8");
  if (d.txt_kv_heads != d.txt_heads) return set_error("md_text_prefill_lora: grouped-query models are not supported with variants");
  if (prefix_len < 0) prefix_len = d.prefix_len;
  const int D = d.txt_dim, H = d.txt_heads, FF = d.txt_ff;
  const long long wide = FF > 3 * D ? FF : 3 * D;
  char* p = align_up(reinterpret_cast<char*>(ws));
  bf16* ln = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * D * 2);
  bf16* q = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * D * 2);
  bf16* att = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * D * 2);
  bf16* attn_out = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * D * 2);
  bf16* mlp_out = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * D * 2);
  bf16* t = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * rank * 2);
  bf16* u = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * wide * 2);
  bf16* qkv = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * 3 * D * 2);
  bf16* pre = reinterpret_cast<bf16*>(p); p += pad256(1LL * T * FF * 2);
  bf16* hid = reinterpret_cast<bf16*>(p);
  bf16* pool = reinterpret_cast<bf16*>(kv.pool);
  // u = bf16(B bf16(A in)):  in [T, K] -> t [T, rank] -> u [T, N]
  auto adapter = [&](const bf16* in, int K, const void* A, const void* B, int N) -> int {
    if (gemm_rowform(in, K, reinterpret_cast<const bf16*>(A), K, T, rank, K, EPI_BIAS, nullptr, nullptr, 0, 0, t, rank,
                     0, 0, 0, st)) return 1;
    return gemm_rowform(t, rank, reinterpret_cast<const bf16*>(B), rank, T, N, rank, EPI_BIAS, nullptr, nullptr, 0, 0, u, N,
                        0, 0, 0, st);
  };
  if (!m.tq.empty() && quant_ready(m)) return 1;
  for (int i = 0; i < d.txt_layers; ++i) {
    const TxtBlock& b = m.txt[i];
    if (dequant_block(m, i, st)) return 1;
    const void* const* L = lora + 8 * i;            // A/B of qkv, proj, fc1, fc2
    for (int j = 0; j < 8; ++j)
      if (!L[j] || (reinterpret_cast<uintptr_t>(L[j]) & 15)) return set_error("md_text_prefill_lora: null or unaligned adapter tensor");
    if (layernorm(x, D, b.ln.w, b.ln.b, ln, D, T, D, 1e-5f, st)) return 1;
    // qkv = bf16(bf16(qkv(l) + bias) + u)                                  text.py:30-32
    if (adapter(ln, D, L[0], L[1], 3 * D)) return 1;
    if (gemm_rowform(ln, D, b.qkv.w, b.qkv.ld, T, 3 * D, D, EPI_BIAS_RESIDUAL, b.qkv.b, u, 3 * D, 0, qkv, 3 * D, 0, 0, 0, st)) return 1;
    if (rope_kv_write(qkv, T, H, q_offsets, start_pos, n_seqs, m.rope, q, pool, kv.n_pages, kv.block_tables,
                      kv.max_blocks, i, st)) return 1;
    if (prefill_attention_tc(q, H, H, T, q_offsets, start_pos, n_seqs, max_q, prefix_len, pool, kv.n_pages, kv.n_layers,
                             kv.block_tables, kv.max_blocks, i, att, st)) return 1;
    // l_attn = bf16(bf16(proj(att) + bias) + bf16(B_p bf16(A_p l)))         text.py:53-56 (x there is the block input)
    if (adapter(ln, D, L[2], L[3], D)) return 1;
    if (gemm_rowform(att, D, b.proj.w, b.proj.ld, T, D, D, EPI_BIAS_RESIDUAL, b.proj.b, u, D, 0, attn_out, D, 0, 0, 0, st)) return 1;
    // mlp: x0 + x1, gelu, x0 + x1                                            layers.py:130-143
    if (adapter(ln, D, L[4], L[5], FF)) return 1;
    if (gemm_rowform(ln, D, b.fc1.w, b.fc1.ld, T, FF, D, EPI_BIAS_RESIDUAL, b.fc1.b, u, FF, 0, pre, FF, 0, 0, 0, st)) return 1;
    if (gelu_rows(pre, hid, 1LL * T * FF, st)) return 1;
    if (adapter(hid, FF, L[6], L[7], D)) return 1;
    if (gemm_rowform(hid, FF, b.fc2.w, b.fc2.ld, T, D, FF, EPI_BIAS_RESIDUAL, b.fc2.b, u, D, 0, mlp_out, D, 0, 0, 0, st)) return 1;
    // x = bf16(bf16(x + l_attn) + l_mlp)                                     text.py:158
    if (add3_rows(x, attn_out, mlp_out, x, 1LL * T * D, st)) return 1;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// text decoder: one decode step for `batch` sequences (fused layout, 5 launches per block)
// ------------------------------------------------------------------------------------------------
static long long smallbatch_ws_floats(const Model& m, int batch) {
  const md_dims& d = m.d;
  long long need = 0;
  auto upd = [&](int n_out, int K) {
    const long long f = 1LL * gemm_smallbatch_splits(n_out, K) * batch * n_out;
    if (f > need) need = f;
  };
  upd(d.txt_dim + 2 * d.txt_kv_heads * 64 + d.txt_ff, d.txt_dim);
  {
    const StreamPlan2 pl = plan_smallbatch_2seg(d.txt_dim, d.txt_dim + d.txt_ff, d.txt_dim);
    const long long f = 1LL * (pl.splits_a + pl.splits_b) * batch * d.txt_dim;
    if (f > need) need = f;
  }
  upd(d.vocab, d.txt_dim);
  upd(d.reg_inner, d.txt_dim); upd(d.coord_out, d.reg_inner); upd(d.size_out, d.reg_inner);
  upd(d.txt_dim, d.coord_feat); upd(d.txt_dim, d.size_feat);
  return need;
}

long long text_decode_ws_bytes(const Model& m, int batch) {
  const md_dims& d = m.d;
  return pad256(1LL * batch * d.txt_dim * 2) * 3 + pad256(1LL * batch * (d.txt_dim + d.txt_ff) * 2) +
         pad256(smallbatch_ws_floats(m, batch) * 4) + 4096;
}

static int small_linear(const bf16* x, long long ldx, const Lin& l, int batch, int n_out, int K, int mode,
                        const bf16* res, long long ldr, bf16* out, long long ldo, float* ws, cudaStream_t st) {
  const int used = gemm_smallbatch(l.w, l.ld, x, ldx, n_out, batch, K, gemm_smallbatch_splits(n_out, K), ws, st);
  if (used < 0) return 1;
  return splitk_epilogue(ws, used, batch, n_out, mode, l.b, res, ldr, out, ldo, st);
}

int g_debug_skip = 0;   // ablation timing only: bit0 GEMM1, bit1 epilogue1, bit2 attention, bit3 GEMM2, bit4 epilogue2

int text_decode_step(Model& m, bf16* x, const int* pos, int batch, const md_kv& kv, bf16* normed_out, void* ws,
                     cudaStream_t st) {
  const md_dims& d = m.d;
  if (batch <= 0) return set_error("md_text_decode_step: empty batch");
  if (!d.txt_fused) return set_error("md_text_decode_step: the model was created without the fused decode layout");
  const int D = d.txt_dim, FF = d.txt_ff, H = d.txt_heads, KVH = d.txt_kv_heads;
  const int QKV = D + 2 * KVH * 64;
  if (kv.n_kv_heads && kv.n_kv_heads != KVH) return set_error("md_text_decode_step: the KV pool's head count differs from the model's");
  char* p = align_up(reinterpret_cast<char*>(ws));
  bf16* ln = reinterpret_cast<bf16*>(p); p += pad256(1LL * batch * D * 2);
  bf16* ln_last = reinterpret_cast<bf16*>(p); p += pad256(1LL * batch * D * 2);
  bf16* qbuf = reinterpret_cast<bf16*>(p); p += pad256(1LL * batch * D * 2);          // grouped-query path only
  bf16* xcat = reinterpret_cast<bf16*>(p); p += pad256(1LL * batch * (D + FF) * 2);   // [att | gelu(fc1)]
  float* wsf = reinterpret_cast<float*>(p);
  const bool quant = !m.tq.empty();
  if (quant && quant_ready(m)) return 1;
  bf16* pool = reinterpret_cast<bf16*>(kv.pool);
  const StreamPlan2 pl2 = plan_smallbatch_2seg(D, D + FF, D);   // no split straddles proj | fc2
  const int proj_splits = pl2.splits_a;
  if (layernorm(x, D, m.txt[0].ln.w, m.txt[0].ln.b, ln, D, batch, D, 1e-5f, st)) return 1;
  for (int i = 0; i < d.txt_layers; ++i) {
    const TxtBlock& b = m.txt[i];
    // [qkv ; fc1] share the input l = ln(x) (text.py:145-157): one weight stream, then bias / RoPE /
    // KV-page write / GELU, which the attention kernel applies for its own (sequence, head).  (Fusing them into
    // the GEMM epilogue needs whole-K tiles, i.e. 112 streaming SMs instead of 148: measured slower, DESIGN.md.)
    int s1 = gemm_smallbatch_splits(QKV + FF, D);
    if (quant) s1 = gemm_smallbatch_quant(m.tq[i].bits, m.tq[i].w1q, m.tq[i].w1s, m.tq[i].w1z, ln, D, QKV + FF, batch, D, 0, wsf, st);
    else if (!(g_debug_skip & 1)) s1 = gemm_smallbatch(b.qkv.w, D, ln, D, QKV + FF, batch, D, 0, wsf, st);
    if (s1 < 0) return 1;
    if (KVH == H) {
      // bias / RoPE / KV-row write / GELU of that stream happen inside the attention kernel (one launch fewer)
      if (!(g_debug_skip & 4) &&
          decode_attention_fused(wsf, s1, D, FF, b.qkv.b, m.rope, xcat + D, D + FF, H, pos, batch, pool, kv.n_pages,
                                 kv.block_tables, kv.max_blocks, i, xcat, D + FF, st)) return 1;
    } else {
      // grouped-query attention: the G query heads of a group share one new K/V row, so the stream is finished by
      // its own small kernel and the plain paged kernel reads KV head h / G
      if (decode_qkv_finish(wsf, s1, batch, D, KVH, FF, b.qkv.b, m.rope, pos, qbuf, pool, kv.n_pages, kv.block_tables,
                            kv.max_blocks, i, xcat + D, D + FF, st)) return 1;
      if (decode_attention(qbuf, H, KVH, pos, batch, pool, kv.n_pages, kv.block_tables, kv.max_blocks, i, xcat,
                           D + FF, st)) return 1;
    }
    // proj(att) and fc2(hid) both land in the residual: one K-concatenated weight stream
    int s2 = pl2.splits_a + pl2.splits_b;
    const bool last = i + 1 == d.txt_layers;
    const Lin& nln = last ? m.txt_post_ln : m.txt[i + 1].ln;
    bf16* ln_dst = last ? (normed_out ? normed_out : ln_last) : ln;
    if (quant) s2 = gemm_smallbatch_quant(m.tq[i].bits, m.tq[i].w2q, m.tq[i].w2s, m.tq[i].w2z, xcat, D + FF, D, batch, D + FF, D, wsf, st);
    else if (!(g_debug_skip & 8)) s2 = gemm_smallbatch_2seg(b.proj.w, D + FF, xcat, D + FF, D, batch, D + FF, D, wsf, st);
    if (s2 < 0) return 1;
    if (!(g_debug_skip & 16) &&
        decode_residual_ln_epilogue(wsf, s2, proj_splits, batch, D, b.proj.b, b.fc2.b, x, nln.w, nln.b, ln_dst, st)) return 1;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// LM head + argmax
// ------------------------------------------------------------------------------------------------
long long lm_head_ws_bytes(const Model& m, int batch) {
  return pad256(1LL * batch * m.d.txt_dim * 2) + pad256(smallbatch_ws_floats(m, batch) * 4) +
         pad256(argmax_scratch_floats(batch) * 4) + 4096;
}

int lm_head_argmax(Model& m, const bf16* hidden, long long ldh, int prenormed, int batch, int mask_id, int mask_id2,
                   int* out_ids, long long out_stride, const int* out_index, float* out_margin,
                   bf16* out_logits, void* ws, cudaStream_t st) {
  const md_dims& d = m.d;
  if (batch <= 0) return set_error("md_lm_head_argmax: empty batch");
  char* p = align_up(reinterpret_cast<char*>(ws));
  bf16* ln = reinterpret_cast<bf16*>(p); p += pad256(1LL * batch * d.txt_dim * 2);
  float* wsf = reinterpret_cast<float*>(p); p += pad256(smallbatch_ws_floats(m, batch) * 4);
  float* scratch = reinterpret_cast<float*>(p);
  const bf16* normed = hidden;
  long long ldn = ldh;
  if (!prenormed) {
    if (layernorm(hidden, ldh, m.txt_post_ln.w, m.txt_post_ln.b, ln, d.txt_dim, batch, d.txt_dim, 1e-5f, st)) return 1;
    normed = ln;
    ldn = d.txt_dim;
  }
  const int used = gemm_smallbatch(m.lm_head.w, d.txt_dim, normed, ldn, d.vocab, batch, d.txt_dim,
                                gemm_smallbatch_splits(d.vocab, d.txt_dim), wsf, st);
  if (used < 0) return 1;
  return argmax_logits(wsf, used, batch, d.vocab, m.lm_head.b, 1, mask_id, mask_id2, out_ids, out_stride, out_index,
                       out_margin, out_logits, scratch, st);
}

// ------------------------------------------------------------------------------------------------
// region head
// ------------------------------------------------------------------------------------------------
long long region_ws_bytes(const Model& m, int batch) {
  const md_dims& d = m.d;
  const int feat = d.size_feat > d.coord_feat ? d.size_feat : d.coord_feat;
  return pad256(1LL * batch * d.reg_inner * 2) + pad256(1LL * batch * feat * 2) +
         pad256(smallbatch_ws_floats(m, batch) * 4) + 4096;
}

int region_decode(Model& m, int which, const bf16* hidden, long long ldh, int batch, int* out_bins,
                  void* ws, cudaStream_t st) {
  const md_dims& d = m.d;
  if (batch <= 0) return set_error("md_region_decode: empty batch");
  char* p = align_up(reinterpret_cast<char*>(ws));
  bf16* hid = reinterpret_cast<bf16*>(p); p += pad256(1LL * batch * d.reg_inner * 2);
  const int feat = d.size_feat > d.coord_feat ? d.size_feat : d.coord_feat;
  p += pad256(1LL * batch * feat * 2);
  float* wsf = reinterpret_cast<float*>(p);
  const Lin& l1 = which == 0 ? m.coord_dec1 : m.size_dec1;
  const Lin& l2 = which == 0 ? m.coord_dec2 : m.size_dec2;
  const int n_out = which == 0 ? d.coord_out : d.size_out;
  // mlp(hidden): fc1 + gelu, fc2 (region.py:46-57, 74-93)
  if (small_linear(hidden, ldh, l1, batch, d.reg_inner, d.txt_dim, EPI_BIAS_GELU, nullptr, 0, hid,
                   d.reg_inner, wsf, st)) return 1;
  const int used = gemm_smallbatch(l2.w, l2.ld, hid, d.reg_inner, n_out, batch, d.reg_inner,
                                gemm_smallbatch_splits(n_out, d.reg_inner), wsf, st);
  if (used < 0) return 1;
  if (which == 0)
    return argmax_logits(wsf, used, batch, n_out, l2.b, 1, -1, -1, out_bins, 1, nullptr, nullptr, nullptr, nullptr, st);
  // size logits are viewed as (2, -1): rows 2b (width bins) and 2b+1 (height bins)
  return argmax_logits(wsf, used, 2 * batch, n_out / 2, l2.b, 2, -1, -1, out_bins, 1, nullptr, nullptr, nullptr, nullptr, st);
}

int region_encode(Model& m, int which, const float* values, int batch, bf16* out, long long ldo, void* ws,
                  cudaStream_t st) {
  const md_dims& d = m.d;
  if (batch <= 0) return set_error("md_region_encode: empty batch");
  char* p = align_up(reinterpret_cast<char*>(ws));
  p += pad256(1LL * batch * d.reg_inner * 2);
  bf16* ff = reinterpret_cast<bf16*>(p);
  const int feat_max = d.size_feat > d.coord_feat ? d.size_feat : d.coord_feat;
  p += pad256(1LL * batch * feat_max * 2);
  float* wsf = reinterpret_cast<float*>(p);
  const int feat = which == 0 ? d.coord_feat : d.size_feat;
  const bf16* fw = which == 0 ? m.coord_features : m.size_features;
  const Lin& enc = which == 0 ? m.coord_enc : m.size_enc;
  if (fourier_features(values, batch, which == 0 ? 1 : 2, fw, feat / 2, ff, feat, st)) return 1;
  return small_linear(ff, feat, enc, batch, d.txt_dim, feat, EPI_BIAS, nullptr, 0, out, ldo, wsf, st);
}

}  // namespace md
