// Does programmatic dependent launch overlap kernels on this box, in a plain stream and in a captured graph?
// Kernel A triggers dependents at its start and then spins ~20 us; kernel B (launched with the PDL
// attribute) records %globaltimer when it starts and after griddepcontrol.wait.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void kA(unsigned long long* ts, int trigger) {
  if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  unsigned long long t0 = gtime();
  while (gtime() - t0 < 20000) {}
  if (threadIdx.x == 0 && blockIdx.x == 0) { ts[0] = t0; ts[1] = gtime(); }
}
__global__ void kB(unsigned long long* ts) {
  unsigned long long t0 = gtime();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  unsigned long long t1 = gtime();
  if (threadIdx.x == 0 && blockIdx.x == 0) { ts[2] = t0; ts[3] = t1; }
}
static void launchB(cudaStream_t s, unsigned long long* ts, int pdl) {
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.stream = s;
  cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; a[0].val.programmaticStreamSerializationAllowed = pdl;
  cfg.attrs = a; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kB, ts);
}
int main() {
  unsigned long long* ts; cudaMallocManaged(&ts, 64);
  cudaStream_t s; cudaStreamCreate(&s);
  for (int mode = 0; mode < 2; ++mode) for (int pdl = 0; pdl < 2; ++pdl) for (int trig = 0; trig < 2; ++trig) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) { kA<<<148, 128, 0, s>>>(ts, trig); launchB(s, ts, pdl); cudaStreamSynchronize(s); }
      else {
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
        kA<<<148, 128, 0, s>>>(ts, trig); launchB(s, ts, pdl);
        cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&ge, g, 0);
        cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
        cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
      }
    }
    printf("%s pdl_attr=%d trigger=%d : B starts %+6.1f us after A ends; B waited %5.1f us  (err %s)\n", mode ? "graph " : "stream", pdl, trig,
           ((double)ts[2] - (double)ts[1]) / 1000.0, ((double)ts[3] - (double)ts[2]) / 1000.0, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
