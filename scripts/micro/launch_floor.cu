// Floor of a dependent kernel chain inside a CUDA graph on this GPU: N empty kernels (one CTA each / 148 CTAs), with and without
// programmatic dependent launch, small and 1.7 KB parameter blocks.   nvcc -arch=sm_100a -O3 -o launch_floor launch_floor.cu
#include <cstdio>
#include <cuda_runtime.h>
struct Big { char pad[1700]; int* out; };
__global__ void k_small(int* out) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] += 1;
}
__global__ void k_big(const __grid_constant__ Big b) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0) b.out[0] += b.pad[5];
}
template <typename F> float time_graph(F launch, int n, cudaStream_t s) {
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n; ++i) launch(s);
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e0, s);
  for (int i = 0; i < 20; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms / 20.f * 1000.f / n;   // us per kernel
}
int main() {
  int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  const int n = 281;
  for (int grid : {1, 148}) for (int threads : {256, 320}) for (int pdl : {0, 1}) for (int big : {0, 1}) {
    auto launch = [&](cudaStream_t st) {
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.stream = st;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = pdl;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (big) { Big b{}; b.out = d; cudaLaunchKernelEx(&cfg, k_big, b); } else cudaLaunchKernelEx(&cfg, k_small, d);
    };
    printf("grid %3d threads %d pdl %d params %4zu B : %.2f us per dependent kernel\n", grid, threads, pdl, big ? sizeof(Big) : sizeof(int*), time_graph(launch, n, s));
  }
  return 0;
}
