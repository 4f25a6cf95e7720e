// C-ABI of libmoondream_b200.so: plain pointers and sizes, no torch types (see include/moondream_b200.h
// for the contract and the reference call sites each entry point replaces).
#include "../../include/moondream_b200.h"

#include <string.h>

#include "kernels.cuh"

namespace md {

static thread_local char g_err[512] = "";
static long long g_launches = 0;

int set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "unknown error", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
  return 1;
}
const char* last_error() { return g_err; }
void count_launch() { ++g_launches; }
long long launch_count() { return g_launches; }
void reset_launch_count() { g_launches = 0; }

}  // namespace md

using bf16 = __nv_bfloat16;
#define BF(p) reinterpret_cast<const bf16*>(p)
#define BFM(p) reinterpret_cast<bf16*>(p)
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

const char* md_last_error(void) { return md::last_error(); }
long long md_launch_count(void) { return md::launch_count(); }
void md_reset_launch_count(void) { md::reset_launch_count(); }
int md_abi_version(void) { return MD_ABI_VERSION; }

int md_linear_bf16(const void* x, long long ldx, const void* w, long long ldw, int M, int N, int K,
                   int epilogue, const void* bias, const void* residual, long long ldr, int res_mod,
                   void* out, long long ldo, int remap_gin, int remap_gout, int remap_goff,
                   void* stream) {
  return md::gemm_rowform(BF(x), ldx, BF(w), ldw, M, N, K, epilogue, BF(bias), BF(residual), ldr,
                          res_mod, BFM(out), ldo, remap_gin, remap_gout, remap_goff, STREAM(stream));
}

int md_linear_small_batch_splits(int n_out, int K) { return md::gemm_swapped_splits(n_out, K); }

long long md_linear_small_batch_workspace_bytes(int n_out, int batch, int K) {
  return static_cast<long long>(md::gemm_swapped_splits(n_out, K)) * batch * n_out * 4;
}

int md_linear_small_batch_bf16(const void* x, long long ldx, const void* w, long long ldw, int batch,
                               int n_out, int K, int epilogue, const void* bias, const void* residual,
                               long long ldr, void* out, long long ldo, void* workspace,
                               void* stream) {
  const int want = md::gemm_swapped_splits(n_out, K);
  const int used = md::gemm_swapped(BF(w), ldw, BF(x), ldx, n_out, batch, K, want,
                                    reinterpret_cast<float*>(workspace), STREAM(stream));
  if (used < 0) return 1;
  return md::splitk_epilogue(reinterpret_cast<const float*>(workspace), used, batch, n_out, epilogue,
                             BF(bias), BF(residual), ldr, BFM(out), ldo, STREAM(stream));
}

}  // extern "C"
