// Model object of the native runtime (host memory only; see engine.cu).
#pragma once
#include <vector>

#include "../../include/moondream_b200.h"
#include "kernels.cuh"

namespace md {

using bf16 = __nv_bfloat16;

struct Lin { const bf16* w; const bf16* b; long long ld; };   // Linear (weight [out,in], row pitch ld, bias) or LayerNorm (w, b)
struct VisBlock { Lin ln1, qkv, proj, ln2, fc1, fc2; };
struct TxtBlock { Lin ln, qkv, proj, fc1, fc2; };

// Weight-only quantised decoder block (gemm_quant.cu): packed stream tensors of W1 = [qkv ; fc1] and W2 = [proj | fc2];
// the bf16 pointers of every TxtBlock then alias ONE scratch pair that prefill fills block by block.
struct QuantBlock {
  int bits = 0;                                    // 0 = not set
  const uint8_t *w1q = nullptr, *w2q = nullptr;
  const float *w1s = nullptr, *w1z = nullptr, *w2s = nullptr, *w2z = nullptr;
};

struct Model {
  md_dims d;
  std::vector<QuantBlock> tq;                      // empty: bf16 decoder weights
  const bf16* lut;
  const float* rope;
  const bf16* pos_emb;
  Lin patch_emb;
  std::vector<VisBlock> vis;
  Lin vis_post_ln, proj_fc1, proj_fc2;
  const bf16* wte;
  std::vector<TxtBlock> txt;
  Lin txt_post_ln, lm_head;
  const bf16* coord_features;
  const bf16* size_features;
  Lin coord_enc, coord_dec1, coord_dec2, size_enc, size_dec1, size_dec2;
};

int model_num_weights(const md_dims& d);
int model_create(const md_dims& d, const void* const* w, int n, const void* lut, const float* rope, Model** out);
int model_set_quantized_block(Model& m, int layer, int bits, const void* w1q, const float* w1_scale, const float* w1_zero,
                              const void* w2q, const float* w2_scale, const float* w2_zero);

long long vision_encode_ws_bytes(const Model& m, int n_crops);
int vision_encode(Model& m, const uint8_t* crops, int n_crops, bf16* feats, void* ws, cudaStream_t st);
long long vision_project_ws_bytes(const Model& m, int n_images);
int vision_project(Model& m, const bf16* feats, const int* crop_offsets, const int* tilings, int n_images,
                   bf16* embeds, int rows_per_image, void* ws, cudaStream_t st);
int vision_project_stitched(Model& m, const bf16* global_feats, const bf16* stitched, int H, int W, bf16* out, void* ws,
                            cudaStream_t st);
long long text_prefill_ws_bytes(const Model& m, int T);
int text_prefill(Model& m, bf16* x, int T, const int* q_offsets, const int* start_pos, int n_seqs,
                 int max_q, int prefix_len, const md_kv& kv, void* ws, cudaStream_t st);
long long text_prefill_lora_ws_bytes(const Model& m, int T, int rank);
int text_prefill_lora(Model& m, bf16* x, int T, const int* q_offsets, const int* start_pos, int n_seqs, int max_q,
                      int prefix_len, const md_kv& kv, const void* const* lora, int rank, void* ws, cudaStream_t st);
extern int g_debug_skip;
long long text_decode_ws_bytes(const Model& m, int batch);
int text_decode_step(Model& m, bf16* x, const int* pos, int batch, const md_kv& kv, bf16* normed_out, void* ws,
                     cudaStream_t st);
long long lm_head_ws_bytes(const Model& m, int batch);
int lm_head_argmax(Model& m, const bf16* hidden, long long ldh, int prenormed, int batch, int mask_id, int mask_id2,
                   int* out_ids,
                   long long out_stride, const int* out_index, float* out_margin, bf16* out_logits,
                   void* ws, cudaStream_t st);
long long region_ws_bytes(const Model& m, int batch);
int region_decode(Model& m, int which, const bf16* hidden, long long ldh, int batch, int* out_bins,
                  void* ws, cudaStream_t st);
int region_encode(Model& m, int which, const float* values, int batch, bf16* out, long long ldo, void* ws,
                  cudaStream_t st);

}  // namespace md

struct md_model : md::Model {};
