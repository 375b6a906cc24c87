// microbench_xbarrier.cu -- latency of the barrier + count all-gather of the exchange-slab wavefront (flagBarrierScan) against the
// atomic-counter grid barrier + a separate read of the counts, for 1..148 participating CTAs of 512 threads.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -I include -I isaac_ros_nvblox_b200/csrc -o /tmp/mbx profiles/microbench_xbarrier.cu
#include <cstdio>
#include "../isaac_ros_nvblox_b200/csrc/nvb_esdf_wavex.cu"
using namespace nvb;

constexpr int kFlagStride = 16;  // 64-bit words between two CTAs' flags: one flag per 128-byte line

// variant 0: acquire polls
__device__ __forceinline__ void flagBarrierScan(XShared& xs, unsigned long long* flags, int nctas, int cta, int tid,
                                                unsigned int gen, int* K, int* M) {
  __syncthreads();
  if (tid == 0) {
    const unsigned long long v = ((unsigned long long)gen << 40) | ((unsigned long long)(unsigned)xs.ncand << 20) |
                                 (unsigned long long)(unsigned)xs.nchanged;
    xs.ncand = 0, xs.nchanged = 0;
    __threadfence();
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(flags + cta * kFlagStride), "l"(v) : "memory");
  }
  int inc = 0, chg = 0;
  if (tid < nctas) {
    unsigned long long v;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + tid * kFlagStride) : "memory");
    } while ((unsigned int)(v >> 40) != gen);
    inc = (int)((v >> 20) & 0xfffffu), chg = (int)(v & 0xfffffu);
  }
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) chg += __shfl_xor_sync(0xffffffffu, chg, o);
  if (warp < 8 && lane == 31) xs.warp_tot[0][warp] = inc;
  if (warp < 8 && lane == 0) xs.warp_tot[1][warp] = chg;
  __syncthreads();
  int base = 0, mtot = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    if (w < warp) base += xs.warp_tot[0][w];
    if (w * 32 < nctas) mtot += xs.warp_tot[1][w];
  }
  if (tid < nctas) xs.pre[tid + 1] = base + inc;
  if (tid == 0) xs.pre[0] = 0;
  __syncthreads();
  *K = xs.pre[nctas];
  *M = mtot;
}

// variant 1: relaxed polling + one acq_rel fence
__device__ __forceinline__ void flagBarrierScanRelaxed(XShared& xs, unsigned long long* flags, int nctas, int cta, int tid,
                                                       unsigned int gen, int* K, int* M) {
  __syncthreads();
  if (tid == 0) {
    const unsigned long long v = ((unsigned long long)gen << 40) | ((unsigned long long)(unsigned)xs.ncand << 20) |
                                 (unsigned long long)(unsigned)xs.nchanged;
    xs.ncand = 0, xs.nchanged = 0;
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(flags + cta * kFlagStride), "l"(v) : "memory");
  }
  int inc = 0, chg = 0;
  if (tid < nctas) {
    unsigned long long v;
    do {
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + tid * kFlagStride) : "memory");
    } while ((unsigned int)(v >> 40) != gen);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    inc = (int)((v >> 20) & 0xfffffu), chg = (int)(v & 0xfffffu);
  }
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) chg += __shfl_xor_sync(0xffffffffu, chg, o);
  if (warp < 8 && lane == 31) xs.warp_tot[0][warp] = inc;
  if (warp < 8 && lane == 0) xs.warp_tot[1][warp] = chg;
  __syncthreads();
  int base = 0, mtot = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    if (w < warp) base += xs.warp_tot[0][w];
    if (w * 32 < nctas) mtot += xs.warp_tot[1][w];
  }
  if (tid < nctas) xs.pre[tid + 1] = base + inc;
  if (tid == 0) xs.pre[0] = 0;
  __syncthreads();
  *K = xs.pre[nctas];
  *M = mtot;
}

template <int V>
__global__ void kbar(unsigned long long* flags, unsigned int* bar, int2* counts, int iters, int participants, long long* out) {
  __shared__ XShared xs;
  const int cta = blockIdx.x, tid = threadIdx.x;
  if (cta >= participants) return;
  if (tid == 0) xs.ncand = 0, xs.nchanged = 0;
  __syncthreads();
  unsigned int generation = 0;
  int K = 0, M = 0, acc = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (tid == 0) xs.ncand = cta & 3, xs.nchanged = 1;
    if (V == 0) flagBarrierScan(xs, flags + (i & 1) * kMaxCtas * kFlagStride, participants, cta, tid, 0x5000u + i, &K, &M);
    if (V == 1) flagBarrierScanRelaxed(xs, flags + (i & 1) * kMaxCtas * kFlagStride, participants, cta, tid, 0x5000u + i, &K, &M);
    if (V == 2) {  // atomic counter barrier, then the counts
      __syncthreads();
      if (tid == 0) counts[(i & 1) * kMaxCtas + cta] = make_int2(xs.ncand, xs.nchanged);
      gridBarrier(bar, generation, participants);
      int2 v = make_int2(0, 0);
      if (tid < participants) v = __ldcg(counts + (i & 1) * kMaxCtas + tid);
      K = __syncthreads_count(v.x > 0), M = v.y;
    }
    acc += K + M;
  }
  const long long t1 = clock64();
  if (tid == 0 && cta == 0) out[0] = (t1 - t0) / iters, out[1] = acc;
}

int main() {
  unsigned long long* flags;
  unsigned int* bar;
  int2* counts;
  long long* out;
  cudaMalloc(&flags, 2 * kMaxCtas * kFlagStride * 8);
  cudaMalloc(&bar, 64);
  cudaMalloc(&counts, 2 * kMaxCtas * sizeof(int2));
  cudaMalloc(&out, 64);
  const int iters = 2000;
  for (int parts : {1, 2, 8, 16, 32, 64, 148}) {
    long long r[3];
    for (int v = 0; v < 3; v++) {
      cudaMemset(flags, 0, 2 * kMaxCtas * kFlagStride * 8);
      cudaMemset(bar, 0, 64);
      int it = iters, pp = parts;
      void* args[] = {&flags, &bar, &counts, &it, &pp, &out};
      const void* fn = v == 0 ? (const void*)kbar<0> : (v == 1 ? (const void*)kbar<1> : (const void*)kbar<2>);
      cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(148), dim3(512), args, 0, 0);
      cudaDeviceSynchronize();
      long long h[2];
      cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      r[v] = e == cudaSuccess ? h[0] : -1;
    }
    printf("participants %3d: flag barrier+scan (acquire polls) %lld cycles, (relaxed polls + fence) %lld, atomic barrier + counts read %lld\n",
           parts, r[0], r[1], r[2]);
  }
  return 0;
}
