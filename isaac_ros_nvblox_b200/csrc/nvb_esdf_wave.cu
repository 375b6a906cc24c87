// nvb_esdf_wave.cu -- the ESDF wavefront (computeEsdf, nvblox/src/integrators/esdf_integrator.cu:1465-1496)
// as ONE cooperative persistent launch.
//
// computeEsdf(list): sweep(list); while (list not empty) { list = updateNeighborBands(list); sweep(list); }
// runs twice per update: for the blocks with sites and for the persistent "cleared" list (:254-257).
// The result depends on the order of the passes (x,y,z in-block sweeps; +x,-x,+y,-y,+z,-z face propagation,
// each seeing the previous ones), so that order is kept: per ring three axis phases and one sweep phase,
// separated by grid barriers.
//
// What a phase costs is the SLOWEST CTA's dependent chain plus a 1.25 us barrier (measured,
// profiles/README.md), ~18 rings x 4 phases per frame. The design therefore minimises round trips and
// imbalance per phase, not bytes:
//   * Ring membership is a per-slot stamp (stamp[r & 1][slot] == r): the face operations test "is my
//     neighbour a member of this ring" with one load that travels with the voxel loads.
//   * Ring r+1's member list is built unique with atomicExch on that stamp. Updated blocks are first recorded
//     in shared memory during the three axis phases and appended once per ring (two dependent L2 atomics per
//     ring instead of per phase); a warp reserves its list range with one atomicAdd. No sort/unique launch.
//   * Members are dealt round-robin from the list (CTA c takes entries c, c+G, ...): every CTA gets
//     ceil(n/G) blocks. (Static slot ownership was tried: max/mean 2.5 and the slowest CTA sets the pace.)
//     A CTA keeps its members and their six neighbour slots (nbr table, built at allocation) in shared
//     memory for the three axis phases: an axis phase is ONE L2 round trip (face voxels + neighbour stamp).
//   * The +dir/-dir passes of an axis are fused per block interface (exact, see axisMembers): 3 phases, not 6;
//     one warp per interface, two face voxels per lane, no block-level synchronisation inside the phase.
//   * A block's three-axis sweep is a ~2 000-instruction dependent chain on two warps; it runs on registers
//     with an integer, branch-free formulation (sweepLineRegs) -- 10 us -> ~2.5 us per block.
//   * 256-thread CTAs with up to 128 registers (half an SM's register file) so that the next frame's
//     raycast / compaction / TSDF kernels co-reside while the wavefront runs on its side stream.
#include "nvb_esdf_common.cuh"
#include "nvb_tma.cuh"

namespace nvb {

namespace {

#ifndef NVB_WAVE_INLINE
#define NVB_WAVE_INLINE 1
#endif
#if NVB_WAVE_INLINE
#define NVB_WAVE_FN __forceinline__
#else
#define NVB_WAVE_FN __noinline__
#endif
#ifndef NVB_WAVE_THREADS
#define NVB_WAVE_THREADS 256
#endif
constexpr int kWT = NVB_WAVE_THREADS;   // threads per CTA. Measured (profiles/wave_variants.sh): 256 and 512 threads are equally fast
                                        // (the phase time is one block's dependent chain), 1024 threads / 64 registers spills and is 50 % slower;
                                        // 256 x 128 registers = half of an SM's register file, so the next frame's kernels co-reside.
constexpr int kWG = kWT / 64;           // groups of 64 threads
constexpr int kWaveMaxMembers = 1024;   // owned candidates scanned per round
constexpr int kPendMax = 1024;         // pending "updated block" records per CTA per ring before a forced flush
constexpr int kNbrCache = 128;          // members whose neighbour slots are cached in smem
constexpr int kSweepBlockWords = kBlockWords;  // (a bank-conflict-free padded image was measured 30 % slower: scalar smem stores)
constexpr size_t kWaveSmemBytes = (size_t)kWG * kSweepBlockWords * sizeof(unsigned int);  // sweep buffers

__device__ __forceinline__ int resolveNeighbor(const EsdfCtx& c, int slot, int dir) {
  int v = __ldcg(c.nbr + 6 * slot + dir);
  if (v < -1) {  // unknown (block created outside the ESDF update path): resolve through the hash once
    const int* bi = c.esdf.block_index + 3 * slot;
    int x = bi[0], y = bi[1], z = bi[2];
    const int d = (dir & 1) ? -1 : 1;
    if ((dir >> 1) == 0) x += d;
    else if ((dir >> 1) == 1) y += d;
    else z += d;
    v = hashFind(c.esdf.hash, x, y, z);
    c.nbr[6 * slot + dir] = v;
  }
  return v;
}

struct WaveShared {
  int members[kWaveMaxMembers];
  int nbr[kNbrCache * 6];
  int scan[kWT / 32];
  int count;
  int changed[kWG];
  int slot[kWG * 2];
  int upd[kWG * 2];
  int npend;
  int pend[kPendMax];  // blocks updated by this CTA's face operations in the current ring (with duplicates)
};

// Members of a ring are dealt round-robin over the CTAs from the ring's global list: CTA c takes entries
// c, c+G, c+2G, ... so every CTA gets ceil(n/G) or floor(n/G) blocks (the static slot-ownership scheme this
// replaces had max/mean of 2.5, and the slowest CTA is what a phase costs). Entries [first, first+cap) of this
// CTA's share are cached in shared memory. Optionally stamps them (initial list of a computeEsdf call).
// `have` leading entries are already in sh.members (fetched speculatively with the ring's count).
__device__ NVB_WAVE_FN int loadMembers(WaveShared& sh, const int* list, int n, int cta, int nctas, int first,
                                       int* stamp_out, int stamp_value, int have = 0) {
  const int tid = threadIdx.x;
  const int mine = (n > cta) ? (n - cta + nctas - 1) / nctas : 0;  // entries of this CTA
  int k = mine - first;
  k = k < 0 ? 0 : (k > kWaveMaxMembers ? kWaveMaxMembers : k);
  for (int j = tid; j < k; j += kWT) {
    const int slot = j < have ? sh.members[j] : __ldcg(list + cta + (first + j) * nctas);
    sh.members[j] = slot;
    if (stamp_out) stamp_out[slot] = stamp_value;
  }
  __syncthreads();
  return k;
}

// Neighbour slots of the first kNbrCache members -> shared memory (one thread per (member, dir)).
__device__ __forceinline__ void prefetchNeighbors(const EsdfCtx& c, WaveShared& sh, int k) {
  const int tid = threadIdx.x;
  const int kc = k < kNbrCache ? k : kNbrCache;
  for (int q = tid; q < kc * 6; q += kWT) sh.nbr[q] = resolveNeighbor(c, sh.members[q / 6], q % 6);
}

// sweepSingleBand (:542-600) on registers, written for a SHORT DEPENDENT CHAIN: one block's sweep is
// 3 axes x 16 sequential steps on two warps, so its latency is (instructions on the chain) x (ALU
// latency), not throughput. Per step the only loop-carried state is the candidate site (l0,l1,l2) and
// `found`; everything that does not depend on it is hoisted:
//   * squared distances are exact integers (or max_sq), so "sq > |d|^2" is evaluated as the integer test
//     ceil(sq) > |d|^2 -- no int->float conversion on the chain; "sq < max_sq" becomes a bit mask;
//   * each voxel's own parent position (parent + voxel) is precomputed;
//   * the four cases of the reference (unobserved / site / first valid voxel / candidate vs own) are
//     folded into branch-free selects.
// The line's 8 voxels are loaded once, walked forward, the register image is reversed and walked again
// (= the backward pass); changed voxels are written back. `axis` and `pass` are run-time values so that
// ONE copy of the 8-step body serves all six passes.
// `sm` is the block image in shared memory, base_w the word offset of the line's first voxel, stride_w the word
// stride along the line; (c0,c1,c2) are the voxel coordinates at position 0.
__device__ __forceinline__ bool sweepLineRegs(unsigned int* sm, int base_w, int stride_w, int c0, int c1, int c2,
                                              int axis, float max_sq) {
  int T[kVps];  // ceil(squared distance): sq > n  <=>  T > n for every integer n
  int p0[kVps], p1[kVps], p2[kVps];
  unsigned int obs = 0, site = 0, valid = 0, dirty = 0;
#pragma unroll
  for (int i = 0; i < kVps; i++) {
    const unsigned int* e = sm + base_w + i * stride_w;
    const float sq = __uint_as_float(e[0]);
    T[i] = __float2int_ru(sq);
    p0[i] = (int)e[1], p1[i] = (int)e[2], p2[i] = (int)e[3];
    const unsigned int fl = e[4];
    if (flagObserved(fl)) obs |= 1u << i;
    if (flagSite(fl)) site |= 1u << i;
    if (sq < max_sq) valid |= 1u << i;
  }
  const int a0 = (axis == 0), a1 = (axis == 1), a2 = (axis == 2);
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    int l0 = 0, l1 = 0, l2 = 0;
    bool found = false;
#pragma unroll
    for (int k = 0; k < kVps; k++) {
      // register slot k holds line position pos = k (pass 0) or 7 - k (pass 1, image reversed)
      const int pos = pass ? (kVps - 1 - k) : k;
      const int x0 = c0 + a0 * pos, x1 = c1 + a1 * pos, x2 = c2 + a2 * pos;  // voxel coordinates
      const bool o = (obs >> k) & 1u, st = (site >> k) & 1u, vl = (valid >> k) & 1u;
      const int own0 = p0[k] + x0, own1 = p1[k] + x1, own2 = p2[k] + x2;    // off the chain
      const int d0 = l0 - x0, d1 = l1 - x1, d2 = l2 - x2;
      const int pd = d0 * d0 + (d1 * d1 + d2 * d2);
      const bool better = found && o && !st && (T[k] > pd);  // candidate site is closer than the voxel's value
      const bool take_site = o && st;
      const bool take_own = o && !st && !better && vl;       // voxel's own parent becomes the running site
      if (better) {
        p0[k] = d0, p1[k] = d1, p2[k] = d2, T[k] = pd;
        dirty |= 1u << k;
        valid |= 1u << k;  // pd < old sq <= max_sq
      }
      l0 = take_site ? x0 : (take_own ? own0 : l0);
      l1 = take_site ? x1 : (take_own ? own1 : l1);
      l2 = take_site ? x2 : (take_own ? own2 : l2);
      found = found || take_site || take_own;
    }
    // reverse the register image (and the bit masks) for the other direction / back to line order
#pragma unroll
    for (int k = 0; k < kVps / 2; k++) {
      const int r = kVps - 1 - k;
      int ti = T[k]; T[k] = T[r]; T[r] = ti;
      ti = p0[k]; p0[k] = p0[r]; p0[r] = ti;
      ti = p1[k]; p1[k] = p1[r]; p1[r] = ti;
      ti = p2[k]; p2[k] = p2[r]; p2[r] = ti;
    }
    obs = __brev(obs) >> 24, site = __brev(site) >> 24, dirty = __brev(dirty) >> 24, valid = __brev(valid) >> 24;
  }
  // after two reversals slot i is line position i again
#pragma unroll
  for (int i = 0; i < kVps; i++) {
    if ((dirty >> i) & 1u) {
      unsigned int* e = sm + base_w + i * stride_w;
      e[0] = __float_as_uint((float)T[i]);
      e[1] = (unsigned)p0[i], e[2] = (unsigned)p1[i], e[3] = (unsigned)p2[i];
    }
  }
  return dirty != 0;
}

// sweepBlockBandKernel (:1390-1431) for the cached members, kWG blocks at a time.
__device__ NVB_WAVE_FN void sweepMembers(const EsdfCtx& c, WaveShared& sh, int k, unsigned int* smem,
                                         bool prefetch_nbr) {
  const int tid = threadIdx.x, group = tid >> 6, lane64 = tid & 63;
  unsigned int* sm = smem + group * kSweepBlockWords;
  const int a = lane64 >> 3, b = lane64 & 7;
  for (int base = 0; base < k; base += kWG) {
    const int item = base + group;
    const int slot = item < k ? sh.members[item] : -1;
    if (lane64 == 0) sh.changed[group] = 0;
    if (slot >= 0) loadBlockGroup(sm, esdfBlockPtr(c.esdf, slot), lane64);
    // neighbour slots for the coming axis phases: issued behind the block loads, not in front of them
    if (prefetch_nbr && base == 0) prefetchNeighbors(c, sh, k);
    __syncthreads();
    bool ch = false;
#pragma unroll 1
    for (int axis = 0; axis < 3; axis++) {
      // x lines: (x, a, b); y lines: (a, y, b); z lines: (a, b, z)
      const int v0 = (axis == 0) ? (a * 8 + b) : ((axis == 1) ? (a * 64 + b) : (a * 64 + b * 8));
      const int stride = (axis == 0) ? 64 : ((axis == 1) ? 8 : 1);
      const int c0 = (axis == 0) ? 0 : a;
      const int c1 = (axis == 0) ? a : ((axis == 1) ? 0 : b);
      const int c2 = (axis == 2) ? 0 : b;
      if (slot >= 0)
        ch |= sweepLineRegs(sm, v0 * kEsdfVoxelWords, stride * kEsdfVoxelWords, c0, c1, c2, axis, c.max_sq);
      __syncthreads();
    }
    if (ch) sh.changed[group] = 1;
    __syncthreads();
    if (slot >= 0 && sh.changed[group]) storeBlockGroup(esdfBlockPtr(c.esdf, slot), sm, lane64);
    __syncthreads();
  }
}

// The +dir and -dir passes of one axis of updateNeighborBands (:1323-1386), fused per block
// interface (see nvb_esdf.cu phaseNeighbors for the ownership rule and why it is exact):
//   group side 0 ("hi"): interface (b, b+d): P = b -> b+d, then Q = b+d -> b if b+d is a member;
//   group side 1 ("lo"): interface (b-d, b) only when b-d is NOT a member: Q = b -> b-d.
// Destination blocks are stamped for ring+1 with a plain store.
__device__ NVB_WAVE_FN void axisMembers(const EsdfCtx& c, WaveShared& sh, int axis, int k, const int* stamp_cur, int ring,
                                        int* stamp_nxt, int* list_nxt, int* count_nxt) {
  // One warp per (member, side) interface, two face voxels per lane: kWT/32 interfaces = kWT/64 members per
  // iteration, so the handful of members a CTA owns are all in flight at once (one L2 round trip per phase).
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kWarps = kWT / 32;
  const int entry_in_cta = warp >> 1, side = warp & 1;
  const int strideA = (axis == 0) ? 64 : ((axis == 1) ? 8 : 1);
  for (int base = 0; base < k; base += kWarps / 2) {
    const int item = base + entry_in_cta;
    int mine = -1, other = -1;
    if (item < k) {
      mine = sh.members[item];
      other = item < kNbrCache ? sh.nbr[item * 6 + axis * 2 + side] : resolveNeighbor(c, mine, axis * 2 + side);
    }
    bool updA = false, updB = false;  // A = low block's hi face, B = high block's lo face
    if (mine >= 0 && other >= 0) {
      // membership of the neighbour and the face voxels are fetched in the same round trip
      const int other_stamp = __ldcg(stamp_cur + other);
      unsigned int* blkA = esdfBlockPtr(c.esdf, side == 0 ? mine : other);
      unsigned int* blkB = esdfBlockPtr(c.esdf, side == 0 ? other : mine);
      VoxelRegs A[2], B[2];
      unsigned int *gHi[2], *gLo[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int f = lane + 32 * h, u = f >> 3, w = f & 7;
        const int faceBase = (axis == 0) ? (u * 8 + w) : ((axis == 1) ? (u * 64 + w) : (u * 64 + w * 8));
        gHi[h] = blkA + (faceBase + (kVps - 1) * strideA) * kEsdfVoxelWords;
        gLo[h] = blkB + faceBase * kEsdfVoxelWords;
        A[h] = loadVoxel(gHi[h]);
        B[h] = loadVoxel(gLo[h]);
      }
#pragma unroll
      for (int h = 0; h < 2; h++) {
        if (side == 0) {
          updB |= updateSingleNeighbor(A[h], B[h], gLo[h], axis, +1, c.max_sq);  // P: mine -> mine + d
          if (other_stamp == ring) updA |= updateSingleNeighbor(B[h], A[h], gHi[h], axis, -1, c.max_sq);  // Q
        } else if (other_stamp != ring) {
          updA |= updateSingleNeighbor(B[h], A[h], gHi[h], axis, -1, c.max_sq);  // Q: mine -> mine - d
        }
      }
    }
    updA = __any_sync(0xffffffffu, updA);
    updB = __any_sync(0xffffffffu, updB);
    if (lane == 0 && mine >= 0 && other >= 0) {
      const int slotA = side == 0 ? mine : other, slotB = side == 0 ? other : mine;
      // Record the updated blocks; they are appended to ring+1 once per ring (flushPending), so the two
      // dependent L2 atomics of the unique append are paid once instead of in each of the three axis phases.
      if (updA) {
        const int q = atomicAdd(&sh.npend, 1);
        if (q < kPendMax) sh.pend[q] = slotA;
        else if (atomicExch(stamp_nxt + slotA, ring + 1) != ring + 1) list_nxt[atomicAdd(count_nxt, 1)] = slotA;
      }
      if (updB) {
        const int q = atomicAdd(&sh.npend, 1);
        if (q < kPendMax) sh.pend[q] = slotB;
        else if (atomicExch(stamp_nxt + slotB, ring + 1) != ring + 1) list_nxt[atomicAdd(count_nxt, 1)] = slotB;
      }
    }
  }
}

// 96 registers x 512 threads = 3/4 of the register file: the wavefront runs on a side stream and must
// leave room for the next frame's raycast / compaction / TSDF CTAs on the same SM.
// Unique append of the recorded blocks to ring+1: the stamp (atomicExch) dedupes across CTAs and doubles as the
// membership flag of ring+1; each warp reserves its range of the list with one atomicAdd.
__device__ NVB_WAVE_FN void flushPending(WaveShared& sh, int* stamp_nxt, int ring, int* list_nxt, int* count_nxt) {
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31;
  const int np = sh.npend < kPendMax ? sh.npend : kPendMax;
  for (int base = 0; base < np; base += kWT) {
    const int q = base + tid;
    int slot = -1;
    bool fresh = false;
    if (q < np) {
      slot = sh.pend[q];
      fresh = atomicExch(stamp_nxt + slot, ring + 1) != ring + 1;
    }
    const unsigned int ballot = __ballot_sync(0xffffffffu, fresh);
    if (ballot) {
      int basepos = 0;
      if (lane == 0) basepos = atomicAdd(count_nxt, __popc(ballot));
      basepos = __shfl_sync(0xffffffffu, basepos, 0);
      if (fresh) list_nxt[basepos + __popc(ballot & ((1u << lane) - 1u))] = slot;
    }
  }
  __syncthreads();
  if (tid == 0) sh.npend = 0;
  __syncthreads();
}

#ifndef NVB_WAVE_MAXREG
#define NVB_WAVE_MAXREG 128
#endif
#ifndef NVB_WAVE_TAIL
#define NVB_WAVE_TAIL 4  // rings with at most this many members (= one sweep round of a CTA) are run by CTA 0 alone; 0 disables. Measured: 0 -> 0.320, 4 -> 0.287, 8 -> 0.300, 16 -> 0.330 ms per frame
#endif
constexpr int kTail = NVB_WAVE_TAIL;
constexpr int kSpec = 8;  // list entries per CTA fetched speculatively together with the ring's member count

// Unique hand-over of the recorded blocks to the next ring when ONE CTA runs the ring (tail mode): duplicates are
// found by comparing the (<= 12 per member) records in shared memory; no atomics, no L2 round trip.
// Next ring's members end up in sh.members[0..n), their stamps and the global list are written with plain stores
// (the list is only read if the ring outgrows the tail mode).
__device__ NVB_WAVE_FN int flushLocal(WaveShared& sh, int* stamp_nxt, int ring, int* list_nxt) {
  const int tid = threadIdx.x;
  if (tid == 0) sh.count = 0;
  __syncthreads();
  const int np = sh.npend;
  for (int q = tid; q < np; q += kWT) {
    const int slot = sh.pend[q];
    bool fresh = true;
    for (int j = 0; j < q; j++) fresh = fresh && (sh.pend[j] != slot);
    if (fresh) {
      const int pos = atomicAdd(&sh.count, 1);
      sh.members[pos] = slot;
      stamp_nxt[slot] = ring + 1;
      list_nxt[pos] = slot;
    }
  }
  __syncthreads();
  const int n = sh.count;
  if (tid == 0) sh.npend = 0;
  __syncthreads();
  return n;
}

__global__ void __maxnreg__(NVB_WAVE_MAXREG) esdfWaveKernel(EsdfCtx c) {
  extern __shared__ __align__(16) unsigned int smem[];
  __shared__ WaveShared sh;
  const int cta = blockIdx.x, nctas = gridDim.x;
  // Empty block list: integrateBlocksTemplate returns before touching anything (:226-228).
  if (*(volatile int*)c.work_count == 0) return;
  if (threadIdx.x == 0) sh.npend = 0;
  __syncthreads();
  unsigned int generation = 0;
  int ring = *(volatile int*)c.ring_id;
  int* stamp[2] = {c.stamp_a, c.stamp_b};
  int* list[2] = {c.ring_a, c.ring_b};
  long long swept = 0, faces = 0, rings = 0;
  // CTA 0 keeps a coarse time split (ns): barriers (incl. waiting for the slowest CTA), axis phases,
  // sweep phases
  long long t_bar = 0, t_axis = 0, t_sweep = 0, n_bar = 0, t0 = globalTimerNs(), t1;
#define NVB_TICK(acc)   \
  t1 = globalTimerNs(); \
  acc += t1 - t0;       \
  t0 = t1;
  // every CTA: duration of its own work in the phase that ends at barrier #n_bar -> max over CTAs
  long long tw0 = globalTimerNs();
#define NVB_PHASE_MAX()                                                                                   \
  if (threadIdx.x == 0 && n_bar < 1000) {                                                                 \
    atomicMax((unsigned long long*)c.phase_max + n_bar, (unsigned long long)(globalTimerNs() - tw0));     \
  }
#define NVB_PHASE_BEGIN() tw0 = globalTimerNs();
#define NVB_BARRIER(acc)                      \
  NVB_TICK(acc)                               \
  NVB_PHASE_MAX()                             \
  gridBarrier(c.barrier, generation, nctas);  \
  NVB_TICK(t_bar)                             \
  n_bar++;                                    \
  NVB_PHASE_BEGIN()
  auto share = [&](int count) { return (count > cta) ? (count - cta + nctas - 1) / nctas : 0; };
  auto roundsOf = [&](int count) { return ((count + nctas - 1) / nctas + kWaveMaxMembers - 1) / kWaveMaxMembers; };
  for (int pass = 0; pass < 2; pass++) {
    // pass 0: blocks with sites; pass 1: the persistent cleared list (:254-257)
    int n = pass ? *(volatile int*)c.cleared_count : *(volatile int*)c.upd_count;
    if (n == 0) continue;
    int ci = ring & 1;
    const int* cur = pass ? c.cleared_list : c.upd_list;  // the first ring's members are read straight from the source list
    bool initial = true;  // members of `cur` still have to be stamped as ring `ring`
    int spec = 0;         // leading entries of this CTA's share already in sh.members (speculative fetch)
    // Loop invariant: the n members of `cur` (ring `ring`) have received their face updates and wait for their sweep.
    while (n > 0) {
      const int ni = ci ^ 1;
      if (kTail > 0 && n <= kTail) {
        // ---- tail mode: the ring fits one CTA. CTA 0 runs whole rings (sweep, three axis phases, hand-over) with
        // block-level synchronisation only, until the wavefront dies out or outgrows the tail; everybody else
        // waits at ONE grid barrier. A ring costs its dependent chain (~6 us) instead of chain + 4 barriers + the
        // list hand-over through L2 (~14 us).
        if (cta == 0) {
          int t = 0, tn = n, tci = ci, tring = ring;
          int k = loadMembers(sh, cur, tn, 0, 1, 0, initial ? stamp[tci] : nullptr, tring);
          while (true) {
            const int tni = tci ^ 1;
            sweepMembers(c, sh, k, smem, true);
            swept += k;
#pragma unroll 1
            for (int axis = 0; axis < 3; axis++) {
              axisMembers(c, sh, axis, k, stamp[tci], tring, stamp[tni], list[tni], c.ring_count + tni);
              __syncthreads();
            }
            faces += 6ll * k;
            rings++, t++;
            k = flushLocal(sh, stamp[tni], tring, list[tni]);
            tring++, tci = tni, tn = k;
            if (tn == 0 || tn > kTail) break;
          }
          if (threadIdx.x == 0) c.tail_state[0] = t, c.tail_state[1] = tn;
        }
        NVB_BARRIER(t_sweep)
        if (threadIdx.x == 0) sh.scan[0] = __ldcg(c.tail_state + 0), sh.scan[1] = __ldcg(c.tail_state + 1);
        __syncthreads();
        const int t = sh.scan[0];
        n = sh.scan[1];
        __syncthreads();
        ring += t;
        ci ^= (t & 1);
        cur = list[ci];
        initial = false, spec = 0;
        continue;
      }
      // ---- grid mode
      const int rounds = roundsOf(n);
      // sweep phase
      for (int r = 0; r < rounds; r++) {
        const int k = loadMembers(sh, cur, n, cta, nctas, r * kWaveMaxMembers, initial ? stamp[ci] : nullptr, ring,
                                  r == 0 ? spec : 0);
        sweepMembers(c, sh, k, smem, rounds == 1);
      }
      if (cta == 0 && threadIdx.x == 0) c.ring_count[ni] = 0;  // append counter of ring+1
      NVB_BARRIER(t_sweep)
      swept += n;
      // axis phases
#pragma unroll 1
      for (int axis = 0; axis < 3; axis++) {
        for (int r = 0; r < rounds; r++) {
          int k = share(n);
          if (rounds > 1) {
            k = loadMembers(sh, cur, n, cta, nctas, r * kWaveMaxMembers, nullptr, 0, 0);
            prefetchNeighbors(c, sh, k);
            __syncthreads();
          }
          axisMembers(c, sh, axis, k, stamp[ci], ring, stamp[ni], list[ni], c.ring_count + ni);
          if (rounds > 1) flushPending(sh, stamp[ni], ring, list[ni], c.ring_count + ni);
        }
        if (axis == 2 && rounds == 1) flushPending(sh, stamp[ni], ring, list[ni], c.ring_count + ni);
        NVB_BARRIER(t_axis)
      }
      faces += 6ll * n;
      rings++;
      // ring+1 = the blocks appended during the three axis phases. Its member count and the first entries of this
      // CTA's share are fetched in the same round trip (entries past the count are ignored).
      {
        const int j = threadIdx.x;
        const long long idx = (long long)cta + (long long)j * nctas;
        int e = 0;
        if (j < kSpec && idx < c.esdf.capacity) e = __ldcg(list[ni] + idx);
        const int n_next = *(volatile int*)(c.ring_count + ni);
        if (j < kSpec) sh.members[j] = e;
        spec = kSpec;
        __syncthreads();
        n = n_next;
      }
      ring++;
      ci = ni;
      cur = list[ni];
      initial = false;
    }
    ring++;
    if (cta == 0 && threadIdx.x == 0) c.ring_count[0] = c.ring_count[1] = 0;
    NVB_BARRIER(t_sweep)
  }
#undef NVB_TICK
  if (cta == 0 && threadIdx.x == 0) {
    *c.ring_id = ring + 1;
    c.stats[4] = *(volatile int*)c.cleared_count;
    c.stats[5] = swept, c.stats[6] = faces, c.stats[7] = rings;
    c.stats[8] = t_bar, c.stats[9] = t_axis, c.stats[10] = t_sweep, c.stats[11] = n_bar;
    long long sum_max = 0;
    for (int q = 0; q < n_bar && q < 1000; q++) sum_max += (long long)c.phase_max[q];
    c.stats[12] = sum_max;  // sum over phases of the slowest CTA's work time
  }
}

// =====================================================================================================
// Gather-emulate-sweep wavefront ("GES"): the same computeEsdf, two grid barriers per ring instead of four.
//
// A ring's six face passes only move information across block boundaries by ONE voxel, so what they do to a
// block B is a function of B and its one-voxel halo (10x10x10 voxels, taken from the 3x3x3 block neighbourhood)
// as they were at the start of the ring, plus which of those 27 blocks are members (sources) of the ring: every
// voxel pair of every pass that touches the region has both voxels inside it. So the CTA that owns a CANDIDATE
// block (= neighbour of a member) gathers the region into shared memory, replays the six passes there in the
// reference's order (halo results are thrown away: their owners compute the same values), and if B changed --
// i.e. B is a member of the next ring -- sweeps it straight away in shared memory. No communication between
// the passes, no unique-append of "updated" blocks (a candidate has exactly one owner). Results are parked in a
// shadow slab until every CTA has finished reading the old state (barrier), then copied over the layer while
// the next ring's candidates and their neighbour rows are fetched (barrier).
// =====================================================================================================
// Region layout (words): "core" = the 10 x 10 z-rows (rx, ry) of the 8 voxels rz = 1..8, 40 words each, in
// the same order as in a block -- so runs of rows that are contiguous in their source block are contiguous here
// and move with ONE TMA bulk copy (30 copies per region); then the two z-halo planes rz = 0 and rz = 9, one
// 8-word cell per voxel, padded so that four of the five voxel words are a 16-byte aligned chunk on both sides.
constexpr int kCoreWords = 100 * 40;
constexpr int kZCell = 8;
constexpr int kZLoBase = kCoreWords;                 // cell: [pad x3][w0][w1 w2 w3 w4]
constexpr int kZHiBase = kCoreWords + 100 * kZCell;  // cell: [w0 w1 w2 w3][w4][pad x3]
constexpr int kRegionWords = kCoreWords + 200 * kZCell;  // 5600
constexpr int kGesMaxCand = 32;                 // candidates of one CTA per chunk
constexpr int kGesDoneMax = 64;                 // changed blocks remembered per 64-thread group and ring
constexpr size_t kGesSmemBytes = (size_t)kWG * kRegionWords * sizeof(unsigned int);

struct GesShared {
  // per-lane constants of the gather and of the pass replay (the same for every candidate): packed descriptors
  unsigned int tz[4][64];  // z-halo voxel copies:  d27 | hi << 5 | zr << 6 | src_off << 13 | valid << 27
  unsigned int tc[4][64];  // core row copies:      d27 | zr << 6 | src_off << 13 | valid << 27
  unsigned int te[6][4][64];  // boundary pairs per pass: src word | dst word << 13 | d27 << 26 | inner << 31; 0 = none
  int cand[kGesMaxCand];
  int rows[kGesMaxCand * 27];
  int done_slots[kWG][kGesDoneMax];
  int done_n[kWG];
  int overflow;
  unsigned int mask[kWG];
  int changed[kWG];
};

__device__ __forceinline__ void groupSync(int group) { asm volatile("bar.sync %0, 64;" ::"r"(group + 1) : "memory"); }
__device__ __forceinline__ void cpAsync16(unsigned int* smem_dst, const void* gsrc, bool valid) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;  // src-size 0: nothing is read, the 16 bytes are zero-filled (block not allocated)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cpAsyncWaitAll() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ int regionWord(int rx, int ry, int rz) {
  const int zr = rx * 10 + ry;
  return rz == 0 ? (kZLoBase + zr * kZCell + 3) : (rz == 9 ? (kZHiBase + zr * kZCell) : (zr * 40 + (rz - 1) * kEsdfVoxelWords));
}
// region coordinate (0..9) -> block offset (-1, 0, +1) and voxel coordinate inside that block
__device__ __forceinline__ int regOff(int r) { return r == 0 ? -1 : (r == 9 ? 1 : 0); }
__device__ __forceinline__ int regLoc(int r) { return r == 0 ? 7 : (r == 9 ? 0 : r - 1); }

// Region of a candidate block <- the 27 blocks of `row` (slots; < 0 = not allocated -> zeros = unobserved voxels).
// All bulk traffic is 16-byte cp.async.cg (L2 -> shared, no L1, no registers). Core: a z-row is 160 contiguous,
// 16-byte aligned bytes on both sides; two neighbouring lanes take alternate chunks of one row, so every request
// of a warp covers whole 32-byte sectors, and the row is decoded once per five copies. z-halo planes: per voxel one
// aligned chunk plus one word. (A TMA bulk-copy version -- 30 copies per region issued by 30 lanes -- was measured
// slower: the per-lane issue of the uniform-datapath copies costs more than the address arithmetic it saves.)
__device__ __forceinline__ void gesInitTables(GesShared& gs, int tid) {
  if (tid < 64) {
    for (int j = 0; j < 4; j++) {
      {
        const int t = tid + 64 * j;
        unsigned int v = 0;
        if (t < 200) {
          const int zr = t >> 1, hi = t & 1;  // hi: rz = 9 <- block +z, voxel z = 0; lo: rz = 0 <- block -z, voxel z = 7
          const int rx = zr / 10, ry = zr % 10;
          const int d = (regOff(rx) + 1) * 9 + (regOff(ry) + 1) * 3 + (hi ? 2 : 0);
          const int off = ((regLoc(rx) * 8 + regLoc(ry)) * 8 + (hi ? 0 : 7)) * 20;
          v = (unsigned)d | ((unsigned)hi << 5) | ((unsigned)zr << 6) | ((unsigned)off << 13) | (1u << 27);
        }
        gs.tz[j][tid] = v;
      }
      {
        const int zr = j * 32 + (tid >> 1);
        unsigned int v = 0;
        if (zr < 100) {
          const int rx = zr / 10, ry = zr % 10;
          const int d = (regOff(rx) + 1) * 9 + (regOff(ry) + 1) * 3 + 1;
          const int off = (regLoc(rx) * 8 + regLoc(ry)) * 160 + (tid & 1) * 16;
          v = (unsigned)d | ((unsigned)zr << 6) | ((unsigned)off << 13) | (1u << 27);
        }
        gs.tc[j][tid] = v;
      }
      for (int pass = 0; pass < 6; pass++) {
        const int t = tid + 64 * j;
        unsigned int v = 0;
        if (t < 200) {
          const int axis = pass >> 1, dir = (pass & 1) ? -1 : 1;
          const int A = axis == 0 ? 9 : (axis == 1 ? 3 : 1), U = axis == 0 ? 3 : 9, W = axis == 2 ? 3 : 1;
          const int p = t / 100, u = (t % 100) / 10, w = t % 10;
          const int sa = dir > 0 ? (p ? 8 : 0) : (p ? 9 : 1);  // source coordinate along the axis
          const int so = dir > 0 ? (p ? 0 : -1) : (p ? 1 : 0);   // block offset of the source along the axis
          const int da = sa + dir;
          const int d = 13 + so * A + regOff(u) * U + regOff(w) * W;
          const int inner = da >= 1 && da <= 8 && u >= 1 && u <= 8 && w >= 1 && w <= 8;
          const int sw = axis == 0 ? regionWord(sa, u, w) : (axis == 1 ? regionWord(u, sa, w) : regionWord(u, w, sa));
          const int dw = axis == 0 ? regionWord(da, u, w) : (axis == 1 ? regionWord(u, da, w) : regionWord(u, w, da));
          v = (unsigned)sw | ((unsigned)dw << 13) | ((unsigned)d << 26) | ((unsigned)inner << 31);
        }
        gs.te[pass][j][tid] = v;
      }
    }
  }
}

__device__ __forceinline__ void gesGather(const EsdfCtx& c, const GesShared& gs, unsigned int* R, const int* row,
                                          int lane64) {
  unsigned int zw[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const unsigned int t = gs.tz[j][lane64];
    zw[j] = 0;
    if (t >> 27) {
      const int slot = row[t & 31u], hi = (t >> 5) & 1u, zr = (t >> 6) & 127u;
      const unsigned char* vox = c.esdf.blocks + (size_t)(slot < 0 ? 0 : slot) * kEsdfBlockBytes + ((t >> 13) & 16383u);
      if (hi) {
        cpAsync16(R + kZHiBase + zr * kZCell, vox, slot >= 0);  // words 0..3
        if (slot >= 0) zw[j] = __ldcg(reinterpret_cast<const unsigned int*>(vox + 16));
      } else {
        cpAsync16(R + kZLoBase + zr * kZCell + 4, vox + 4, slot >= 0);  // words 1..4
        if (slot >= 0) zw[j] = __ldcg(reinterpret_cast<const unsigned int*>(vox));
      }
    }
  }
  const int half = lane64 & 1;
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const unsigned int t = gs.tc[it][lane64];
    if (t >> 27) {
      const int slot = row[t & 31u], zr = (t >> 6) & 127u;
      const unsigned char* src = c.esdf.blocks + (size_t)(slot < 0 ? 0 : slot) * kEsdfBlockBytes + ((t >> 13) & 16383u);
      unsigned int* dst = R + zr * 40 + half * 4;
#pragma unroll
      for (int j = 0; j < 5; j++) cpAsync16(dst + j * 8, src + j * 32, slot >= 0);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const unsigned int t = gs.tz[j][lane64];
    if (t >> 27) {
      const int hi = (t >> 5) & 1u, zr = (t >> 6) & 127u;
      R[hi ? (kZHiBase + zr * kZCell + 4) : (kZLoBase + zr * kZCell + 3)] = zw[j];
    }
  }
  cpAsyncWaitAll();
}

// updateSingleNeighbor (:602-633) on two voxels in shared memory, split into an operand fetch and the update so
// that the fetches of a thread's (up to three) pairs of a plane are in flight together.
struct PairOps {
  unsigned int e0, e1, e2, e3, e4, n0, n4;
  unsigned int* nb;
  bool act;
};
__device__ __forceinline__ PairOps pairLoad(unsigned int* R, unsigned int desc, bool act) {
  PairOps q;
  q.act = act;
  q.nb = R + ((desc >> 13) & 8191u);
  // unconditional fetch (the addresses of an inactive descriptor are valid words of the region): no branch, so the
  // operands of the lane's four pairs are in flight together
  const unsigned int* e = R + (desc & 8191u);
  q.e0 = e[0], q.e1 = e[1], q.e2 = e[2], q.e3 = e[3], q.e4 = e[4];
  q.n0 = q.nb[0], q.n4 = q.nb[4];
  return q;
}
__device__ __forceinline__ bool pairApply(const PairOps& q, int axis, int direction, float max_sq) {
  const bool ok = q.act && flagObserved(q.e4) && flagObserved(q.n4) && !flagSite(q.n4) && !(__uint_as_float(q.e0) >= max_sq);
  const int d0 = (int)q.e1 - (axis == 0 ? direction : 0), d1 = (int)q.e2 - (axis == 1 ? direction : 0),
            d2 = (int)q.e3 - (axis == 2 ? direction : 0);
  const float pdist = (float)(d0 * d0 + (d1 * d1 + d2 * d2));
  if (ok && __uint_as_float(q.n0) > pdist) {
    q.nb[1] = (unsigned)d0, q.nb[2] = (unsigned)d1, q.nb[3] = (unsigned)d2;
    q.nb[0] = __float_as_uint(pdist);
    return true;
  }
  return false;
}

// The six passes of updateLocalNeighborBands (:1323-1386) restricted to the region: +x, -x, +y, -y, +z, -z, each
// seeing the previous ones. A pass along `axis` has two boundary planes (block -1|0 and block 0|+1) of 10 x 10
// voxel pairs; a pair is processed iff its SOURCE block is a member of the ring (bit in `mask`). Inside one pass
// sources and destinations are disjoint planes, so its 200 pairs are independent: each lane owns (up to) four of
// them -- the same for every candidate, so their shared-memory addresses and source-block indices come from a
// table built once per launch (the replay is instruction-bound, not latency-bound) -- fetches the operands of the
// active ones together and then applies them.
__device__ __forceinline__ bool gesEmulate(const GesShared& gs, unsigned int* R, unsigned int mask, int lane64, int group,
                                           float max_sq) {
  bool changed = false;
#pragma unroll 1
  for (int pass = 0; pass < 6; pass++) {
    const int axis = pass >> 1, dir = (pass & 1) ? -1 : 1;
    const int A = axis == 0 ? 9 : (axis == 1 ? 3 : 1);
    // source blocks of this pass: offset along the axis in {-1, 0} (dir +) or {0, +1} (dir -)
    const unsigned int lo = axis == 0 ? 0x000001ffu : (axis == 1 ? 0x001c0e07u : 0x01249249u);  // offset -1 along the axis
    const unsigned int src_blocks = dir > 0 ? (lo | (lo << A)) : ((lo << A) | (lo << (2 * A)));
    if (!(mask & src_blocks)) continue;  // group-uniform
    unsigned int desc[4];
    PairOps q[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      desc[j] = gs.te[pass][j][lane64];
      q[j] = pairLoad(R, desc[j], desc[j] != 0u && ((mask >> ((desc[j] >> 26) & 31u)) & 1u));
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (pairApply(q[j], axis, dir, max_sq)) changed = changed || (desc[j] >> 31);
    groupSync(group);
  }
  return changed;
}

__device__ __forceinline__ void copyShadowToLayer(const EsdfCtx& c, int slot, int lane64) {
  const uint4* src = reinterpret_cast<const uint4*>(c.shadow + (size_t)slot * kEsdfBlockBytes);
  uint4* dst = reinterpret_cast<uint4*>(c.esdf.blocks + (size_t)slot * kEsdfBlockBytes);
  uint4 q[kBlockWords / 4 / 64];
#pragma unroll
  for (int k = 0; k < kBlockWords / 4 / 64; k++) q[k] = __ldcg(src + lane64 + k * 64);
#pragma unroll
  for (int k = 0; k < kBlockWords / 4 / 64; k++) __stcg(dst + lane64 + k * 64, q[k]);
}

__global__ void __maxnreg__(NVB_WAVE_MAXREG) esdfWaveGesKernel(EsdfCtx c) {
  extern __shared__ __align__(16) unsigned int smem[];
  __shared__ WaveShared sh;
  __shared__ GesShared gs;
  const int cta = blockIdx.x, nctas = gridDim.x;
  const int tid = threadIdx.x, group = tid >> 6, lane64 = tid & 63;
  if (*(volatile int*)c.work_count == 0) return;
  if (tid == 0) sh.npend = 0;
  if (tid < kWG) gs.done_n[tid] = 0;
  if (tid == 0) gs.overflow = 0;
  gesInitTables(gs, tid);
  __syncthreads();
  unsigned int generation = 0;
  int ring = *(volatile int*)c.ring_id;
  int* stamp[2] = {c.stamp_a, c.stamp_b};
  int* list[2] = {c.ring_a, c.ring_b};
  int* clist[2] = {c.cand_a, c.cand_b};
  int* ccount = c.ges_counts;      // [parity] candidates of the ring with that parity
  int* mcount = c.ges_counts + 2;  // [parity] members of the ring with that parity
  long long swept = 0, faces = 0, rings = 0, n_bar = 0, t_bar = 0, t_work = 0, t0 = globalTimerNs(), t1;
  unsigned int* R = smem + group * kRegionWords;
  long long tg = 0, te = 0, ts = 0, ncand = 0, nchg = 0, tq;  // CTA 0 / group 0: gather, emulate, sweep+store time (SM clock cycles: %globaltimer costs ~a round trip per read)
#define GES_BARRIER()                                                                                           \
  t1 = globalTimerNs(), t_work += t1 - t0;                                                                      \
  if (tid == 0 && n_bar < 1000) atomicMax((unsigned long long*)c.phase_max + n_bar, (unsigned long long)(t1 - t0)); \
  t0 = t1;                                                                                                      \
  gridBarrier(c.barrier, generation, nctas);                                                                    \
  t1 = globalTimerNs(), t_bar += t1 - t0, t0 = t1;                                                              \
  n_bar++;
  auto roundsOf = [&](int count) { return ((count + nctas - 1) / nctas + kWaveMaxMembers - 1) / kWaveMaxMembers; };
  for (int pass = 0; pass < 2; pass++) {
    const int* src = pass ? c.cleared_list : c.upd_list;
    int n0 = pass ? *(volatile int*)c.cleared_count : *(volatile int*)c.upd_count;
    if (n0 == 0) continue;
    int ci = ring & 1;
    bool initial = true;
    // ---- large rings: four-phase rings (sweep | x | y | z, member lists with unique append), exactly as in
    // esdfWaveKernel. A large ring has several candidates per 64-thread group, which the gather-replay below would
    // process one after the other; here the per-member work is smaller and the four barriers amortise.
    {
      auto share = [&](int count) { return (count > cta) ? (count - cta + nctas - 1) / nctas : 0; };
      int spec = 0;
      while (n0 > c.ges_switch) {
        const int ni = ci ^ 1;
        const int rounds = roundsOf(n0);
        for (int r = 0; r < rounds; r++) {
          const int k = loadMembers(sh, src, n0, cta, nctas, r * kWaveMaxMembers, initial ? stamp[ci] : nullptr, ring,
                                    r == 0 ? spec : 0);
          sweepMembers(c, sh, k, smem, rounds == 1);
        }
        if (cta == 0 && tid == 0) c.ring_count[ni] = 0;  // append counter of ring+1
        GES_BARRIER()
        swept += n0;
#pragma unroll 1
        for (int axis = 0; axis < 3; axis++) {
          for (int r = 0; r < rounds; r++) {
            int k = share(n0);
            if (rounds > 1) {
              k = loadMembers(sh, src, n0, cta, nctas, r * kWaveMaxMembers, nullptr, 0, 0);
              prefetchNeighbors(c, sh, k);
              __syncthreads();
            }
            axisMembers(c, sh, axis, k, stamp[ci], ring, stamp[ni], list[ni], c.ring_count + ni);
            if (rounds > 1) flushPending(sh, stamp[ni], ring, list[ni], c.ring_count + ni);
          }
          if (axis == 2 && rounds == 1) flushPending(sh, stamp[ni], ring, list[ni], c.ring_count + ni);
          GES_BARRIER()
        }
        faces += 6ll * n0;
        rings++;
        {
          const long long idx = (long long)cta + (long long)tid * nctas;
          int e = 0;
          if (tid < kSpec && idx < c.esdf.capacity) e = __ldcg(list[ni] + idx);
          const int n_next = *(volatile int*)(c.ring_count + ni);
          if (tid < kSpec) sh.members[tid] = e;
          spec = kSpec;
          __syncthreads();
          n0 = n_next;
        }
        ring++;
        ci = ni;
        src = list[ni];
        initial = false;
      }
      if (n0 == 0) {  // the wavefront died out in the large-ring regime
        ring++;
        if (cta == 0 && tid == 0) c.ring_count[0] = c.ring_count[1] = 0;
        GES_BARRIER()
        continue;
      }
    }
    // ---- hand-over / initial phase: sweep the current member list in place, stamp it as ring `ring`, register its
    // neighbours as the candidates of that ring
    {
      const int rounds = roundsOf(n0);
      for (int r = 0; r < rounds; r++) {
        const int k = loadMembers(sh, src, n0, cta, nctas, r * kWaveMaxMembers, stamp[ci], ring);
        sweepMembers(c, sh, k, smem, true);
        for (int q = tid; q < k * 6; q += kWT) {
          const int item = q / 6;
          const int nb = item < kNbrCache ? sh.nbr[q] : resolveNeighbor(c, sh.members[item], q % 6);
          if (nb >= 0 && atomicExch(c.cand_stamp + nb, ring) != ring) clist[ci][atomicAdd(ccount + ci, 1)] = nb;
        }
        __syncthreads();
      }
    }
    GES_BARRIER()
    swept += n0;
    int M = n0;
    int K_prev = 0;  // candidates of the previous ring (clist[ni])
    while (true) {
      const int ni = ci ^ 1;
      // ---- phase P: results of the previous ring go from the shadow slab into the layer (nobody reads the layer
      // now); this ring's candidates and their 27-neighbourhood rows come into shared memory
      const int K = *(volatile int*)(ccount + ci);
      const int share = (K > cta) ? (K - cta + nctas - 1) / nctas : 0;
      {
        const int kc = share < kGesMaxCand ? share : kGesMaxCand;
        int myslot = -1;
        if (tid < kc) myslot = __ldcg(clist[ci] + cta + tid * nctas);  // issued before the copies
        if (gs.overflow) {
          // more changed blocks than a group remembers: find them again through the previous ring's candidate
          // list (a candidate that became a member of ring `ring` is a changed block)
          for (int j = group; cta + (long long)j * nctas < K_prev; j += kWG) {
            const int s = __ldcg(clist[ni] + cta + j * nctas);
            if (__ldcg(stamp[ci] + s) == ring) copyShadowToLayer(c, s, lane64);
          }
        } else {
          for (int j = 0; j < gs.done_n[group]; j++) copyShadowToLayer(c, gs.done_slots[group][j], lane64);
        }
        __syncthreads();
        if (tid < kWG) gs.done_n[tid] = 0;
        if (tid == 0) gs.overflow = 0;
        if (tid < kc) gs.cand[tid] = myslot;
        __syncthreads();
        for (int q = tid; q < kc * 27; q += kWT) {
          const int s = gs.cand[q / 27], d = q % 27;
          int v = __ldcg(c.nbr27 + 27 * s + d);
          if (v < -1) {  // never linked (block created outside the ESDF update path): resolve through the hash once
            const int* bi = c.esdf.block_index + 3 * s;
            v = hashFind(c.esdf.hash, bi[0] + d / 9 - 1, bi[1] + (d / 3) % 3 - 1, bi[2] + d % 3 - 1);
            c.nbr27[27 * s + d] = v;
          }
          gs.rows[q] = v;
        }
        if (cta == 0 && tid == 0) ccount[ni] = 0, mcount[ni] = 0;  // filled during the coming phase A
      }
      GES_BARRIER()
      // ---- phase A: every candidate of the ring: gather, replay the six passes, sweep if it changed
      for (int base = 0; base < share; base += kGesMaxCand) {
        const int kc = (share - base) < kGesMaxCand ? (share - base) : kGesMaxCand;
        if (base > 0) {  // further chunks (very large rings only): fetch in place
          __syncthreads();
          if (tid < kc) gs.cand[tid] = __ldcg(clist[ci] + cta + (base + tid) * nctas);
          __syncthreads();
          for (int q = tid; q < kc * 27; q += kWT) {
            const int s = gs.cand[q / 27], d = q % 27;
            int v = __ldcg(c.nbr27 + 27 * s + d);
            if (v < -1) {
              const int* bi = c.esdf.block_index + 3 * s;
              v = hashFind(c.esdf.hash, bi[0] + d / 9 - 1, bi[1] + (d / 3) % 3 - 1, bi[2] + d % 3 - 1);
              c.nbr27[27 * s + d] = v;
            }
            gs.rows[q] = v;
          }
          __syncthreads();
        }
        for (int i = group; i < kc; i += kWG) {
          const int slot = gs.cand[i];
          const int* row = gs.rows + i * 27;
          // membership of the 27 blocks in this ring (sources of the passes); the loads travel with the gather
          int st = ring - 1;
          tq = clock64();
          if (lane64 < 27 && row[lane64] >= 0) st = __ldcg(stamp[ci] + row[lane64]);
          gesGather(c, gs, R, row, lane64);
          const unsigned int m = __ballot_sync(0xffffffffu, lane64 < 27 && st == ring);
          if (lane64 == 0) gs.mask[group] = m, gs.changed[group] = 0;
          groupSync(group);
          tg += clock64() - tq, tq = clock64(), ncand++;
          const bool ch = gesEmulate(gs, R, gs.mask[group], lane64, group, c.max_sq);
          if (ch) gs.changed[group] = 1;
          groupSync(group);
          te += clock64() - tq, tq = clock64();
          if (gs.changed[group]) {
            nchg++;
            // B is a member of ring+1: tell its face neighbours (candidates of ring+1). The exchange is issued
            // now and consumed after the sweep.
            int nb = -1, old = ring + 1;
            if (lane64 < 6) {
              const int face = lane64 == 0 ? 22 : (lane64 == 1 ? 4 : (lane64 == 2 ? 16 : (lane64 == 3 ? 10 : (lane64 == 4 ? 14 : 12))));
              nb = row[face];
              if (nb >= 0) old = atomicExch(c.cand_stamp + nb, ring + 1);
            }
            if (lane64 == 0) {
              stamp[ni][slot] = ring + 1;
              atomicAdd(mcount + ni, 1);
              const int dn = gs.done_n[group];
              if (dn < kGesDoneMax) gs.done_slots[group][dn] = slot, gs.done_n[group] = dn + 1;
              else gs.overflow = 1;
            }
            int pos = -1;
            {
              const int a = lane64 >> 3, b = lane64 & 7;
              sweepLineRegs(R, regionWord(1, a + 1, b + 1), 400, 0, a, b, 0, c.max_sq);
              groupSync(group);
              // the exchange is back by now; the position fetch overlaps the other two axes
              if (nb >= 0 && old != ring + 1) pos = atomicAdd(ccount + ni, 1);
              sweepLineRegs(R, regionWord(a + 1, 1, b + 1), 40, a, 0, b, 1, c.max_sq);
              groupSync(group);
              sweepLineRegs(R, regionWord(a + 1, b + 1, 1), kEsdfVoxelWords, a, b, 0, 2, c.max_sq);
              groupSync(group);
            }
            if (pos >= 0) clist[ni][pos] = nb;
            // inner 8x8x8 -> shadow slab
            uint4* dst = reinterpret_cast<uint4*>(c.shadow + (size_t)slot * kEsdfBlockBytes);
#pragma unroll
            for (int it = 0; it < 2; it++) {
              const int r = it * 32 + (lane64 >> 1);  // block row (lx, ly); two lanes share a row
              const unsigned int* srow = R + (((r >> 3) + 1) * 10 + (r & 7) + 1) * 40 + (lane64 & 1) * 4;
#pragma unroll
              for (int j = 0; j < 5; j++) __stcg(dst + r * 10 + (lane64 & 1) + 2 * j, *reinterpret_cast<const uint4*>(srow + j * 8));
            }
          }
          groupSync(group);
          ts += clock64() - tq;
        }
      }
      GES_BARRIER()
      const int M_next = *(volatile int*)(mcount + ni);
      faces += 6ll * M;
      rings++;
      swept += M_next;
      ring++;
      ci = ni;
      M = M_next;
      K_prev = K;
      if (M == 0) break;
    }
    ring++;
    if (cta == 0 && tid == 0) ccount[0] = ccount[1] = mcount[0] = mcount[1] = 0;
    GES_BARRIER()
  }
#undef GES_BARRIER
  if (cta == 0 && tid == 0) {
    *c.ring_id = ring + 1;
    c.stats[4] = *(volatile int*)c.cleared_count;
    c.stats[5] = swept, c.stats[6] = faces, c.stats[7] = rings;
    c.stats[8] = t_bar, c.stats[9] = t_work, c.stats[10] = 0, c.stats[11] = n_bar;
    long long sum_max = 0;
    for (int q = 0; q < n_bar && q < 1000; q++) sum_max += (long long)c.phase_max[q];
    c.stats[12] = sum_max;
    c.phase_max[3990] = tg, c.phase_max[3991] = te, c.phase_max[3992] = ts, c.phase_max[3993] = ncand, c.phase_max[3994] = nchg;
  }
}

}  // namespace

int esdfPersistentMaxCtas(int num_sms) {
  static int per_sm = -1;
  if (per_sm < 0) {
    cudaFuncSetAttribute(esdfWaveKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWaveSmemBytes);
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, esdfWaveKernel, kWT, kWaveSmemBytes) != cudaSuccess) v = 0;
    per_sm = v;
  }
  return per_sm * num_sms;
}

cudaError_t launchEsdfComputeGes(const EsdfCtx& c, int num_sms, cudaStream_t stream, int* launches) {
  static int per_sm = -1;
  if (per_sm < 0) {
    cudaFuncSetAttribute(esdfWaveGesKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGesSmemBytes);
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, esdfWaveGesKernel, kWT, kGesSmemBytes) != cudaSuccess) v = 0;
    per_sm = v;
  }
  if (per_sm <= 0) return cudaErrorLaunchOutOfResources;
  EsdfCtx cc = c;
  void* args[] = {&cc};
  (*launches)++;
  return cudaLaunchCooperativeKernel((const void*)esdfWaveGesKernel, dim3(num_sms), dim3(kWT), args, kGesSmemBytes, stream);
}

cudaError_t launchEsdfComputePersistent(const EsdfCtx& c, int num_sms, cudaStream_t stream, int* launches) {
  const int max_ctas = esdfPersistentMaxCtas(num_sms);
  if (max_ctas <= 0) return cudaErrorLaunchOutOfResources;
  // One CTA per SM: the wavefront is latency-bound, more CTAs only make the barrier slower.
  int grid = num_sms < max_ctas ? num_sms : max_ctas;
  EsdfCtx cc = c;
  void* args[] = {&cc};
  (*launches)++;
  return cudaLaunchCooperativeKernel((const void*)esdfWaveKernel, dim3(grid), dim3(kWT), args, kWaveSmemBytes, stream);
}

}  // namespace nvb
