/*
 * moondream_b200 — C-ABI of the B200-native moondream hot path (libmoondream_b200.so).
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes; device pointers unless marked HOST; bf16 = raw uint16 storage;
 *   - `stream` is a cudaStream_t passed as void*; entry points never synchronise and never
 *     allocate: the caller passes outputs and workspaces (see the *_workspace_bytes queries);
 *   - return 0 on success, non-zero on error; md_last_error() returns the message (thread-local);
 *   - leading dimensions (`ld*`) are in elements.
 *
 * The reference (vikhyat/moondream, /root/reference) has no FFI: its swap seam is the four bound
 * methods MoondreamModel._vis_enc/_vis_proj/_prefill/_decode_one_tok (moondream/torch/moondream.py
 * :168-192) that compile() rebinds (:194-204).  Each entry point below cites the reference lines
 * whose arithmetic it replaces; INTEGRATION.md shows the ctypes stubs that bind them.
 */
#ifndef MOONDREAM_B200_H
#define MOONDREAM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MD_ABI_VERSION 1

/* epilogue selectors for the linear entry points */
#define MD_EPI_BIAS 0          /* y = bf16(x W^T + b)                        layers.py:34-35   */
#define MD_EPI_BIAS_GELU 1     /* y = bf16(gelu_tanh(bf16(x W^T + b)))       layers.py:130,137 */
#define MD_EPI_BIAS_RESIDUAL 2 /* y = bf16(bf16(x W^T + b) + r)              vision.py:70-71, text.py:158 */

const char* md_last_error(void);
int md_abi_version(void);
/* number of kernels this library has launched since the last reset (bench.py's gpu_launches) */
long long md_launch_count(void);
void md_reset_launch_count(void);

/*
 * y[M,N] = epilogue(x[M,K] @ w[N,K]^T + bias[N])            (tcgen05 + TMA GEMM, bf16 in/out)
 * Replaces F.linear at layers.py:35 and its nn.Linear call sites vision.py:67 (patch_emb),
 * layers.py:159-165 (ViT qkv/proj), layers.py:130,139 (fc1/fc2), vision.py:89 (proj_mlp),
 * text.py:30,53 (decoder qkv/proj) for prefill-sized M.
 *   residual: [*, N] rows of stride ldr; when res_mod > 0 the residual row is (row % res_mod)
 *             (pos_emb broadcast over crops, vision.py:68).
 *   remap:    when remap_gin > 0 output row r is written at (r / gin) * gout + r % gin + goff
 *             (places the 729 projected image rows after the BOS row, moondream.py:254).
 * Requirements: N % 8 == 0, K % 8 == 0, 16-byte aligned pointers and row pitches.
 */
int md_linear_bf16(const void* x, long long ldx, const void* w, long long ldw, int M, int N, int K,
                   int epilogue, const void* bias, const void* residual, long long ldr, int res_mod,
                   void* out, long long ldo, int remap_gin, int remap_gout, int remap_goff,
                   void* stream);

/*
 * Same contraction for a small batch (decode step, region head): the weight matrix is the
 * M side of the MMA so every SM streams weights at HBM rate; K is split across CTAs and the fp32
 * partial sums are reduced in a fixed order (deterministic).  Replaces the M=1 F.linear calls the
 * reference issues per generated token (text.py:30,53; layers.py:130,139; region.py:43-93).
 */
int md_linear_small_batch_splits(int n_out, int K);
long long md_linear_small_batch_workspace_bytes(int n_out, int batch, int K);
int md_linear_small_batch_bf16(const void* x, long long ldx, const void* w, long long ldw, int batch,
                               int n_out, int K, int epilogue, const void* bias, const void* residual,
                               long long ldr, void* out, long long ldo, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOONDREAM_B200_H */
