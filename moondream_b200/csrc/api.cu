// C-ABI of libmoondream_b200.so: plain pointers and sizes, no torch types (see include/moondream_b200.h
// for the contract and the reference call sites each entry point replaces).
#include <string.h>

#include <atomic>

#include "engine.h"

namespace md {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};   // engines on several host threads share the counter
int g_attention_impl = 0;
int g_pdl = 1;   // programmatic dependent launch: weight-streaming GEMMs prefetch their first ring of stages under the
                 // predecessor (decode attention / small epilogues trigger early); A/B in-run: 1.94 vs 2.23 ms per step

int set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "unknown error", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
  return 1;
}
const char* last_error() { return g_err; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
void reset_launch_count() { g_launches.store(0, std::memory_order_relaxed); }

}  // namespace md

using md::bf16;
#define BF(p) reinterpret_cast<const bf16*>(p)
#define BFM(p) reinterpret_cast<bf16*>(p)
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define NEED(p, what) \
  if (!(p)) return md::set_error(what ": null pointer")

extern "C" {

const char* md_last_error(void) { return md::last_error(); }
long long md_launch_count(void) { return md::launch_count(); }
void md_reset_launch_count(void) { md::reset_launch_count(); }
int md_abi_version(void) { return MD_ABI_VERSION; }
void md_profile_linear(int enable) { md::gemm_profile_enable(enable); }
void md_debug_force_cta_group(int cta_group) { md::gemm_force_cta_group(cta_group); }
int md_profile_linear_read(double* total_ms, double* total_flops, long long* launches) {
  NEED(total_ms && total_flops && launches, "md_profile_linear_read");
  return md::gemm_profile_read(total_ms, total_flops, launches);
}

// ---------------------------------------------------------------- operator level
int md_linear_bf16(const void* x, long long ldx, const void* w, long long ldw, int M, int N, int K,
                   int epilogue, const void* bias, const void* residual, long long ldr, int res_mod,
                   void* out, long long ldo, int remap_gin, int remap_gout, int remap_goff,
                   void* stream) {
  NEED(x && w && out, "md_linear_bf16");
  if (epilogue < 0 || epilogue > 2) return md::set_error("md_linear_bf16: bad epilogue");
  return md::gemm_rowform(BF(x), ldx, BF(w), ldw, M, N, K, epilogue, BF(bias), BF(residual), ldr,
                          res_mod, BFM(out), ldo, remap_gin, remap_gout, remap_goff, STREAM(stream));
}

int md_linear_small_batch_splits(int n_out, int K) { return md::gemm_smallbatch_splits(n_out, K); }

long long md_linear_small_batch_workspace_bytes(int n_out, int batch, int K) {
  return static_cast<long long>(md::gemm_smallbatch_splits(n_out, K)) * batch * n_out * 4;
}

int md_linear_small_batch_bf16(const void* x, long long ldx, const void* w, long long ldw, int batch,
                               int n_out, int K, int epilogue, const void* bias, const void* residual,
                               long long ldr, void* out, long long ldo, void* workspace,
                               void* stream) {
  NEED(x && w && out && workspace, "md_linear_small_batch_bf16");
  if (epilogue < 0 || epilogue > 2) return md::set_error("md_linear_small_batch_bf16: bad epilogue");
  const int want = md::gemm_smallbatch_splits(n_out, K);
  const int used = md::gemm_smallbatch(BF(w), ldw, BF(x), ldx, n_out, batch, K, want,
                                    reinterpret_cast<float*>(workspace), STREAM(stream));
  if (used < 0) return 1;
  return md::splitk_epilogue(reinterpret_cast<const float*>(workspace), used, batch, n_out, epilogue,
                             BF(bias), BF(residual), ldr, BFM(out), ldo, STREAM(stream));
}

int md_dequantize_weights(int bits, const void* wq, const float* scale, const float* zero, int N, int K, void* out,
                          long long ldo, void* stream) {
  NEED(wq && scale && zero && out, "md_dequantize_weights");
  return md::dequant_weights(bits, reinterpret_cast<const uint8_t*>(wq), scale, zero, N, K, BFM(out), ldo, STREAM(stream));
}

int md_linear_small_batch_quant(int bits, const void* x, long long ldx, const void* wq, const float* scale,
                                const float* zero, int batch, int n_out, int K, int epilogue, const void* bias,
                                const void* residual, long long ldr, void* out, long long ldo, void* workspace,
                                void* stream) {
  NEED(x && wq && scale && zero && out && workspace, "md_linear_small_batch_quant");
  if (epilogue < 0 || epilogue > 2) return md::set_error("md_linear_small_batch_quant: bad epilogue");
  const int used = md::gemm_smallbatch_quant(bits, reinterpret_cast<const uint8_t*>(wq), scale, zero, BF(x), ldx, n_out,
                                             batch, K, 0, reinterpret_cast<float*>(workspace), STREAM(stream));
  if (used < 0) return 1;
  return md::splitk_epilogue(reinterpret_cast<const float*>(workspace), used, batch, n_out, epilogue,
                             BF(bias), BF(residual), ldr, BFM(out), ldo, STREAM(stream));
}

int md_resample_u8(const uint8_t* src, int in_h, int in_w, int axis, const int* bounds, const int* coeffs, int ksize,
                   int out_size, uint8_t* dst, void* stream) {
  NEED(src && bounds && coeffs && dst, "md_resample_u8");
  return md::resample_u8(src, in_h, in_w, axis, bounds, coeffs, ksize, out_size, dst, STREAM(stream));
}

int md_extract_windows_u8(const uint8_t* canvas, int h, int w, int rows, int cols, int stride, int crop, uint8_t* crops,
                          void* stream) {
  NEED(canvas && crops, "md_extract_windows_u8");
  return md::extract_windows_u8(canvas, h, w, rows, cols, stride, crop, crops, STREAM(stream));
}

int md_patchify_u8(const uint8_t* crops, int n_crops, int crop, int patch, int k_pad, const void* pixel_lut, void* out,
                   void* stream) {
  NEED(crops && pixel_lut && out, "md_patchify_u8");
  if (n_crops <= 0 || patch <= 0 || crop % patch || k_pad < 3 * patch * patch || k_pad % 8)
    return md::set_error("md_patchify_u8: bad geometry");
  return md::patchify(crops, n_crops, crop, patch, k_pad, BF(pixel_lut), BFM(out), STREAM(stream));
}

int md_stitch_pool_concat_bf16(const void* feats, const int* crop_offsets, const int* tilings, int n_images, int grid,
                               int margin, int dim, void* out, void* stream) {
  NEED(feats && crop_offsets && tilings && out, "md_stitch_pool_concat_bf16");
  if (n_images <= 0 || grid <= 0 || dim % 8) return md::set_error("md_stitch_pool_concat_bf16: bad geometry");
  return md::stitch_pool_concat(BF(feats), crop_offsets, tilings, n_images, grid, margin, dim, BFM(out), STREAM(stream));
}

int md_layernorm_bf16(const void* x, long long ldx, const void* w, const void* b, void* y,
                      long long ldy, int rows, int dim, void* stream) {
  NEED(x && w && b && y, "md_layernorm_bf16");
  return md::layernorm(BF(x), ldx, BF(w), BF(b), BFM(y), ldy, rows, dim, 1e-5f, STREAM(stream));
}

int md_vit_attention_bf16(const void* qkv, int n_crops, int seq, int n_heads, void* out, void* stream) {
  NEED(qkv && out, "md_vit_attention_bf16");
  if (md::g_attention_impl == 1) return md::vit_attention(BF(qkv), n_crops, seq, n_heads, BFM(out), STREAM(stream));
  return md::vit_attention_tc(BF(qkv), n_crops, seq, n_heads, BFM(out), STREAM(stream));
}

int md_rope_kv_write_bf16(const void* qkv, int n_tokens, int n_heads, const int* q_offsets,
                          const int* start_pos, int n_seqs, const float* rope_table, void* q_out,
                          const md_kv* kv, int layer, void* stream) {
  NEED(qkv && start_pos && rope_table && q_out && kv && kv->pool && kv->block_tables, "md_rope_kv_write_bf16");
  if (kv->n_kv_heads && kv->n_kv_heads != n_heads) return md::set_error("md_rope_kv_write_bf16: multi-head layout only");
  return md::rope_kv_write(BF(qkv), n_tokens, n_heads, q_offsets, start_pos, n_seqs, rope_table,
                           BFM(q_out), BFM(kv->pool), kv->n_pages, kv->block_tables, kv->max_blocks,
                           layer, STREAM(stream));
}

int md_prefill_attention_bf16(const void* q, int n_heads, int total_tokens, const int* q_offsets,
                              const int* start_pos, int n_seqs, int max_q, int prefix_len,
                              const md_kv* kv, int layer, void* out, void* stream) {
  NEED(q && q_offsets && start_pos && kv && kv->pool && kv->block_tables && out, "md_prefill_attention_bf16");
  if (md::g_attention_impl == 1 && kv->n_kv_heads && kv->n_kv_heads != n_heads)
    return md::set_error("md_prefill_attention_bf16: the legacy mma.sync attention has no grouped-query path");
  if (md::g_attention_impl == 1)
    return md::prefill_attention(BF(q), n_heads, q_offsets, start_pos, n_seqs, max_q, prefix_len,
                                 BF(kv->pool), kv->n_pages, kv->block_tables, kv->max_blocks, layer,
                                 BFM(out), STREAM(stream));
  return md::prefill_attention_tc(BF(q), n_heads, kv->n_kv_heads, total_tokens, q_offsets, start_pos, n_seqs, max_q, prefix_len,
                                  BF(kv->pool), kv->n_pages, kv->n_layers, kv->block_tables, kv->max_blocks,
                                  layer, BFM(out), STREAM(stream));
}
void md_debug_attention_impl(int impl) { md::g_attention_impl = impl; }
void md_debug_set_pdl(int enable) { md::g_pdl = enable; }
void md_debug_skip_decode_kernels(int mask) { md::g_debug_skip = mask; }
void md_debug_gemm(int flags) { md::gemm_debug_flags(flags); }
void md_debug_gemm_sm_cap(int sms) { md::gemm_debug_sm_cap(sms); }
int md_debug_timeline(void* records, void* count, unsigned int capacity) {
  return md::timeline_install_all(static_cast<unsigned long long*>(records), static_cast<unsigned int*>(count), capacity);
}

int md_decode_attention_bf16(const void* q, int n_heads, const int* pos, int n_seqs, const md_kv* kv,
                             int layer, void* out, void* stream) {
  NEED(q && pos && kv && kv->pool && kv->block_tables && out, "md_decode_attention_bf16");
  return md::decode_attention(BF(q), n_heads, kv->n_kv_heads, pos, n_seqs, BF(kv->pool), kv->n_pages, kv->block_tables,
                              kv->max_blocks, layer, BFM(out), static_cast<long long>(n_heads) * 64,
                              STREAM(stream));
}

// ---------------------------------------------------------------- model level
int md_model_num_weights(const md_dims* dims) {
  if (!dims) return -1;
  return md::model_num_weights(*dims);
}

int md_model_create(const md_dims* dims, const void* const* weights, int n_weights,
                    const void* pixel_lut, const float* rope_table, md_model** out) {
  NEED(dims && weights && pixel_lut && rope_table && out, "md_model_create");
  md::Model* m = nullptr;
  if (md::model_create(*dims, weights, n_weights, pixel_lut, rope_table, &m)) return 1;
  *out = static_cast<md_model*>(m);
  return 0;
}

void md_model_destroy(md_model* model) { delete static_cast<md::Model*>(model); }

int md_model_set_quantized_block(md_model* model, int layer, int bits, const void* w1q, const float* w1_scale,
                                 const float* w1_zero, const void* w2q, const float* w2_scale, const float* w2_zero) {
  NEED(model, "md_model_set_quantized_block");
  return md::model_set_quantized_block(*model, layer, bits, w1q, w1_scale, w1_zero, w2q, w2_scale, w2_zero);
}

long long md_vision_encode_workspace_bytes(const md_model* model, int n_crops) {
  return model ? md::vision_encode_ws_bytes(*model, n_crops) : -1;
}
int md_vision_encode(md_model* model, const uint8_t* crops, int n_crops, void* feats, void* workspace,
                     void* stream) {
  NEED(model && crops && feats && workspace, "md_vision_encode");
  return md::vision_encode(*model, crops, n_crops, BFM(feats), workspace, STREAM(stream));
}

long long md_vision_project_workspace_bytes(const md_model* model, int n_images) {
  return model ? md::vision_project_ws_bytes(*model, n_images) : -1;
}
int md_vision_project(md_model* model, const void* feats, const int* crop_offsets, const int* tilings,
                      int n_images, void* embeds, int rows_per_image, void* workspace, void* stream) {
  NEED(model && feats && crop_offsets && tilings && embeds && workspace, "md_vision_project");
  return md::vision_project(*model, BF(feats), crop_offsets, tilings, n_images, BFM(embeds), rows_per_image,
                            workspace, STREAM(stream));
}

int md_vision_project_stitched(md_model* model, const void* global_feats, const void* stitched, int height, int width,
                               void* out, void* workspace, void* stream) {
  NEED(model && global_feats && stitched && out && workspace, "md_vision_project_stitched");
  return md::vision_project_stitched(*model, BF(global_feats), BF(stitched), height, width, BFM(out), workspace,
                                     STREAM(stream));
}

int md_embed_tokens(md_model* model, const int* ids, long long id_stride, int n, void* out,
                    long long ldo, void* stream) {
  NEED(model && ids && out, "md_embed_tokens");
  return md::embed_tokens(ids, id_stride, n, model->wte, model->d.txt_dim, model->d.vocab, BFM(out), ldo,
                          STREAM(stream));
}

long long md_text_prefill_workspace_bytes(const md_model* model, int total_tokens) {
  return model ? md::text_prefill_ws_bytes(*model, total_tokens) : -1;
}
int md_text_prefill(md_model* model, void* x, int total_tokens, const int* q_offsets,
                    const int* start_pos, int n_seqs, int max_q, int prefix_len, const md_kv* kv,
                    void* workspace, void* stream) {
  NEED(model && x && q_offsets && start_pos && kv && kv->pool && kv->block_tables && workspace, "md_text_prefill");
  return md::text_prefill(*model, BFM(x), total_tokens, q_offsets, start_pos, n_seqs, max_q, prefix_len, *kv,
                          workspace, STREAM(stream));
}

long long md_text_prefill_lora_workspace_bytes(const md_model* model, int total_tokens, int rank) {
  return model ? md::text_prefill_lora_ws_bytes(*model, total_tokens, rank) : -1;
}
int md_text_prefill_lora(md_model* model, void* x, int total_tokens, const int* q_offsets, const int* start_pos,
                         int n_seqs, int max_q, int prefix_len, const md_kv* kv, const void* const* lora, int rank,
                         void* workspace, void* stream) {
  NEED(model && x && q_offsets && start_pos && kv && kv->pool && kv->block_tables && lora && workspace, "md_text_prefill_lora");
  return md::text_prefill_lora(*model, BFM(x), total_tokens, q_offsets, start_pos, n_seqs, max_q, prefix_len, *kv, lora,
                               rank, workspace, STREAM(stream));
}

long long md_text_decode_workspace_bytes(const md_model* model, int batch) {
  return model ? md::text_decode_ws_bytes(*model, batch) : -1;
}
int md_text_decode_step(md_model* model, void* x, const int* pos, int batch, const md_kv* kv,
                        void* normed_out, void* workspace, void* stream) {
  NEED(model && x && pos && kv && kv->pool && kv->block_tables && workspace, "md_text_decode_step");
  return md::text_decode_step(*model, BFM(x), pos, batch, *kv, BFM(normed_out), workspace, STREAM(stream));
}

long long md_lm_head_workspace_bytes(const md_model* model, int batch) {
  return model ? md::lm_head_ws_bytes(*model, batch) : -1;
}
int md_lm_head_argmax(md_model* model, const void* hidden, long long ld_hidden, int prenormed, int batch,
                      int mask_id, int mask_id2, int* out_ids, long long out_stride, const int* out_index,
                      float* out_margin, void* out_logits, void* workspace, void* stream) {
  NEED(model && hidden && out_ids && workspace, "md_lm_head_argmax");
  return md::lm_head_argmax(*model, BF(hidden), ld_hidden, prenormed, batch, mask_id, mask_id2, out_ids, out_stride,
                            out_index, out_margin, BFM(out_logits), workspace, STREAM(stream));
}

int md_sample_top_p(const void* logits, int batch, int vocab, float temperature, float top_p,
                    const unsigned long long* seed, const int* step, const float* uniforms, void* scratch,
                    int keep_probs, int* out_ids, long long out_stride, int out_offset, void* stream) {
  NEED(logits && scratch && out_ids, "md_sample_top_p");
  return md::sample_top_p(BF(logits), batch, vocab, temperature, top_p, seed, step, uniforms, BFM(scratch), keep_probs,
                          out_ids, out_stride, out_offset, STREAM(stream));
}

int md_embed_tokens_select(md_model* model, const int* ids, long long id_stride, int n, int sel_id,
                           const void* alt, long long ld_alt, void* out, long long ldo, void* stream) {
  NEED(model && ids && out, "md_embed_tokens_select");
  return md::embed_select(ids, id_stride, n, model->wte, model->d.txt_dim, model->d.vocab, sel_id, BF(alt), ld_alt,
                          BFM(out), ldo, STREAM(stream));
}

int md_store_column_f32(const float* src, int n, float* dst, long long stride, const int* index, int offset,
                        void* stream) {
  NEED(src && dst, "md_store_column_f32");
  return md::store_column_f32(src, n, dst, stride, index, offset, STREAM(stream));
}

int md_decode_advance(int* cur_tok, int* pos, int* step, const int* preds, const int* forced,
                      long long stride, int batch, int eos_id, int* finished, void* stream) {
  NEED(cur_tok && pos && step && preds, "md_decode_advance");
  return md::decode_advance(cur_tok, pos, step, preds, forced, stride, batch, eos_id, finished,
                            STREAM(stream));
}

int md_gather_rows_bf16(const void* src, long long ld_src, const int* row_index, int n, int dim,
                        void* out, long long ldo, void* stream) {
  NEED(src && row_index && out, "md_gather_rows_bf16");
  return md::gather_rows(BF(src), ld_src, row_index, n, dim, BFM(out), ldo, STREAM(stream));
}

long long md_region_workspace_bytes(const md_model* model, int batch) {
  return model ? md::region_ws_bytes(*model, batch) : -1;
}
int md_region_decode(md_model* model, int which, const void* hidden, long long ld_hidden, int batch,
                     int* out_bins, void* workspace, void* stream) {
  NEED(model && hidden && out_bins && workspace, "md_region_decode");
  if (which != 0 && which != 1) return md::set_error("md_region_decode: which must be 0 or 1");
  return md::region_decode(*model, which, BF(hidden), ld_hidden, batch, out_bins, workspace, STREAM(stream));
}
int md_region_encode(md_model* model, int which, const float* values, int batch, void* out,
                     long long ldo, void* workspace, void* stream) {
  NEED(model && values && out && workspace, "md_region_encode");
  if (which != 0 && which != 1) return md::set_error("md_region_encode: which must be 0 or 1");
  return md::region_encode(*model, which, values, batch, BFM(out), ldo, workspace, STREAM(stream));
}
int md_region_bins_to_values(int which, const int* bins, int n, int n_bins, float* out, void* stream) {
  NEED(bins && out, "md_region_bins_to_values");
  return md::bins_to_values(which, bins, n, n_bins, out, STREAM(stream));
}

}  // extern "C"
