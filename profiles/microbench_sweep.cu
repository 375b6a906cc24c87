// microbench_sweep.cu -- single-group latency of the pieces of the exchange-slab wavefront's per-candidate chain (sweep, replay,
// own-block staging) on one SM, nothing else running. Build on the GPU box:
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -I include -I isaac_ros_nvblox_b200/csrc -o /tmp/mbs profiles/microbench_sweep.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../isaac_ros_nvblox_b200/csrc/nvb_esdf_wavex.cu"

using namespace nvb;

// the fully unrolled, axis-templated variant of the sweep (three copies of two unrolled passes)
template <int AXIS>
__device__ __forceinline__ bool sweepLineT(unsigned int* R, int a, int b, float max_sq) {
  uint4* A = reinterpret_cast<uint4*>(R);
  const int v0 = AXIS == 0 ? rvox(1, a + 1, b + 1) : (AXIS == 1 ? rvox(a + 1, 1, b + 1) : rvox(a + 1, b + 1, 1));
  constexpr int stride = AXIS == 0 ? kRX : (AXIS == 1 ? kRY : 1);
  constexpr int kNone = 1 << 28, kCap = 1 << 27;
  int T[kVps], pa[kVps], po1[kVps], po2[kVps], pp2[kVps];
  unsigned int tk = 0, dirty = 0;
#pragma unroll
  for (int i = 0; i < kVps; i++) {
    const uint4 q = A[v0 + i * stride];
    const unsigned int fl = R[kFlagBase + v0 + i * stride];
    const float sq = __uint_as_float(q.x);
    const bool o = flagObserved(fl), st = flagSite(fl);
    const int t = min(__float2int_ru(sq), kCap);
    T[i] = (o && !st) ? t : INT_MIN;
    const int qa = (int)(AXIS == 0 ? q.y : (AXIS == 1 ? q.z : q.w));
    const int q1 = (int)(AXIS == 0 ? q.z : q.y);
    const int q2 = (int)(AXIS == 2 ? q.z : q.w);
    pa[i] = st ? 0 : qa, po1[i] = st ? 0 : q1, po2[i] = st ? 0 : q2;
    pp2[i] = po1[i] * po1[i] + po2[i] * po2[i];
    if (o && (st || sq < max_sq)) tk |= 1u << i;
  }
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    int la = 0, lo1 = 0, lo2 = 0, lp2 = kNone;
#pragma unroll
    for (int kk = 0; kk < kVps; kk++) {
      const int k = pass ? (kVps - 1 - kk) : kk;
      const int t = la - k;
      const int pd = t * t + lp2;
      const bool improve = T[k] > pd;
      const bool take = ((tk >> k) & 1u) && !improve;
      const int offer_a = pa[k] + k;
      if (improve) {
        T[k] = pd, pa[k] = t, po1[k] = lo1, po2[k] = lo2, pp2[k] = lp2;
        dirty |= 1u << k;
      }
      la = take ? offer_a : la;
      lo1 = take ? po1[k] : lo1;
      lo2 = take ? po2[k] : lo2;
      lp2 = take ? pp2[k] : lp2;
    }
    tk |= dirty;
  }
#pragma unroll
  for (int i = 0; i < kVps; i++) {
    if ((dirty >> i) & 1u) {
      const unsigned int x = (unsigned)(AXIS == 0 ? pa[i] : po1[i]);
      const unsigned int y = (unsigned)(AXIS == 0 ? po1[i] : (AXIS == 1 ? pa[i] : po2[i]));
      const unsigned int z = (unsigned)(AXIS == 2 ? pa[i] : po2[i]);
      A[v0 + i * stride] = make_uint4(__float_as_uint((float)T[i]), x, y, z);
    }
  }
  return dirty != 0;
}

__global__ void kbench(const unsigned int* init, long long* out, int active_groups, int iters) {
  extern __shared__ __align__(16) unsigned int smem[];
  const int tid = threadIdx.x, group = tid >> 6, lane64 = tid & 63;
  unsigned int* R = smem + group * kXRegionWords;
  if (group >= active_groups) return;
  long long t_roll = 0, t_tmpl = 0, t_line = 0, t_sync = 0;
  for (int it = 0; it < iters; it++) {
    for (int i = lane64; i < kXRegionWords; i += 64) R[i] = init[i];
    groupSync(group);
    long long t0 = clock64();
    sweepBlockX(R, group, lane64, 1600.0f);
    groupSync(group);
    long long t1 = clock64();
    t_roll += t1 - t0;
    for (int i = lane64; i < kXRegionWords; i += 64) R[i] = init[i];
    groupSync(group);
    t0 = clock64();
    const int a = lane64 >> 3, b = lane64 & 7;
    sweepLineT<0>(R, a, b, 1600.0f);
    groupSync(group);
    sweepLineT<1>(R, a, b, 1600.0f);
    groupSync(group);
    sweepLineT<2>(R, a, b, 1600.0f);
    groupSync(group);
    t1 = clock64();
    t_tmpl += t1 - t0;
    t0 = clock64();
    sweepLineT<1>(R, a, b, 1600.0f);
    t1 = clock64();
    t_line += t1 - t0;
    t0 = clock64();
    groupSync(group);
    groupSync(group);
    groupSync(group);
    groupSync(group);
    t1 = clock64();
    t_sync += t1 - t0;
  }
  if (lane64 == 0) {
    out[group * 4 + 0] = t_roll / iters, out[group * 4 + 1] = t_tmpl / iters, out[group * 4 + 2] = t_line / iters, out[group * 4 + 3] = t_sync / iters;
  }
}

int main() {
  std::vector<unsigned int> init(kXRegionWords, 0);
  srand(1);
  for (int v = 0; v < kRegionVox; v++) {
    const bool obs = (rand() % 10) != 0, site = obs && (rand() % 40) == 0;
    float sq = site ? 0.0f : (float)(rand() % 1700);
    init[4 * v] = *reinterpret_cast<unsigned int*>(&sq);
    init[4 * v + 1] = (unsigned)(rand() % 41 - 20), init[4 * v + 2] = (unsigned)(rand() % 41 - 20), init[4 * v + 3] = (unsigned)(rand() % 41 - 20);
    init[kFlagBase + v] = (obs ? 0x100u : 0u) | (site ? 0x10000u : 0u);
  }
  unsigned int* d_init;
  long long* d_out;
  cudaMalloc(&d_init, init.size() * 4);
  cudaMemcpy(d_init, init.data(), init.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&d_out, 8 * 4 * sizeof(long long));
  cudaFuncSetAttribute(kbench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kXSmemBytes);
  for (int groups : {1, 2, 4, 8}) {
    cudaMemset(d_out, 0, 8 * 4 * sizeof(long long));
    kbench<<<1, 512, kXSmemBytes>>>(d_init, d_out, groups, 20);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[32];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("groups active %d (%s): sweep rolled %lld cycles, templated %lld, one templated y-axis pass %lld, 4 group syncs %lld\n", groups,
           cudaGetErrorString(e), h[0], h[1], h[2], h[3]);
  }
  return 0;
}
