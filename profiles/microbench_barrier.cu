// microbench_barrier.cu -- latency of grid-wide barrier flavours on B200 (design input for the
// ESDF wavefront kernel). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb profiles/microbench_barrier.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;

__device__ __forceinline__ void barA(unsigned* bar, unsigned& gen, unsigned n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    gen++;
    __threadfence();
    atomicAdd(bar, 1u);
    while (*(volatile unsigned*)bar < gen * n) {}
    __threadfence();
  }
  __syncthreads();
}
__device__ __forceinline__ void barB(unsigned* bar, unsigned& gen, unsigned n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    gen++;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < gen * n);
  }
  __syncthreads();
}
// relaxed arrive (no fence at all): lower bound, NOT a correct barrier for data
__device__ __forceinline__ void barR(unsigned* bar, unsigned& gen, unsigned n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    gen++;
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned v;
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < gen * n);
  }
  __syncthreads();
}
// fence + red + PIPELINED polling: kPoll loads in flight, issued ~RT/kPoll apart, so the release is seen
// ~one-way latency after it lands instead of on average half a round trip later
template <int kPoll>
__device__ __forceinline__ void barP(unsigned* bar, unsigned& gen, unsigned n, unsigned gap_ns) {
  __syncthreads();
  if (threadIdx.x == 0) {
    gen++;
    const unsigned target = gen * n;
    __threadfence();
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned v[kPoll];
#pragma unroll
    for (int i = 0; i < kPoll; i++) {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v[i]) : "l"(bar) : "memory");
      if (i + 1 < kPoll) __nanosleep(gap_ns);
    }
    bool done = false;
    while (!done) {
#pragma unroll
      for (int i = 0; i < kPoll; i++) {
        if (v[i] >= target) { done = true; break; }
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v[i]) : "l"(bar) : "memory");
      }
    }
    __threadfence();
  }
  __syncthreads();
}
template <int V>
__global__ void kbar(unsigned* bar, int iters, int work, float* sink) {
  unsigned gen = 0;
  float acc = 0;
  cg::grid_group g = cg::this_grid();
  for (int i = 0; i < iters; i++) {
    for (int w = 0; w < work; w++) acc = acc * 1.0001f + (float)threadIdx.x;  // stand-in for phase work
    if (V == 0) barA(bar, gen, gridDim.x);
    if (V == 1) barB(bar, gen, gridDim.x);
    if (V == 2) g.sync();
    if (V == 3) barR(bar, gen, gridDim.x);
    if (V == 4) barP<4>(bar, gen, gridDim.x, 120);
    if (V == 5) barP<2>(bar, gen, gridDim.x, 250);
    if (V == 6) barP<8>(bar, gen, gridDim.x, 60);
  }
  if (acc == 12345.f) *sink = acc;
}
__global__ void __cluster_dims__(8, 1, 1) kcluster(int iters, float* sink) {
  float acc = 0;
  for (int i = 0; i < iters; i++) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (acc == 1.f) *sink = acc;
}
template <int V>
float run(int grid, int threads, int iters) {
  unsigned* bar;
  float* sink;
  cudaMalloc(&bar, 256);
  cudaMalloc(&sink, 4);
  cudaMemset(bar, 0, 256);
  int work = 0;
  void* args[] = {&bar, &iters, &work, &sink};
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  cudaLaunchCooperativeKernel((void*)kbar<V>, dim3(grid), dim3(threads), args, 0, 0);  // warm
  cudaDeviceSynchronize();
  cudaMemset(bar, 0, 256);
  cudaEventRecord(a);
  cudaError_t e = cudaLaunchCooperativeKernel((void*)kbar<V>, dim3(grid), dim3(threads), args, 0, 0);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) return -1;
  cudaFree(bar), cudaFree(sink);
  return ms * 1000.f / iters;
}
int main() {
  const int iters = 2000;
  int grids[] = {148, 74, 32, 16, 8, 296};
  printf("grid threads  A(fence+atom+volatile)  B(red.release+ld.acquire)  C(cg grid.sync)  R(relaxed, lower bound)  P4 P2 P8 (pipelined polls)   [us per barrier]\n");
  for (int g : grids)
    for (int t : {256, 64}) {
      printf("%4d %4d   %8.3f   %8.3f   %8.3f   %8.3f   %8.3f %8.3f %8.3f\n", g, t, run<0>(g, t, iters), run<1>(g, t, iters),
             run<2>(g, t, iters), run<3>(g, t, iters), run<4>(g, t, iters), run<5>(g, t, iters), run<6>(g, t, iters));
    }
  float* sink;
  cudaMalloc(&sink, 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  kcluster<<<8, 256>>>(10, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  kcluster<<<8, 256>>>(iters, sink);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  printf("cluster(8) hardware barrier: %.3f us\n", ms * 1000.f / iters);
  return 0;
}
