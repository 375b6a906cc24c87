"""Multi-GPU plumbing: one process per GPU, one camera stream + one map replica per rank.

The path shards across cameras / time-slices with independent maps: TSDF fusion
is order dependent and the ESDF is a global fixpoint, so there is no voxel-level reduction.
The only exchange is the merge of the ranks' updated-block index lists, so that every rank
knows which blocks the rig touched:

    merged = sort(unique(concat(list_r for r in ranks)))       (bit-identical on every rank)

torch.distributed (NCCL on GPUs, gloo in the CPU tests) moves the lists; the keys are the same
21-bit-per-axis packed indices the device hash uses (csrc/nvb_internal.cuh packIndex).
"""
import numpy as np
import torch
import torch.distributed as dist

INDEX_BIAS = 1 << 20


def pack_indices(xyz):
    """(n,3) int32/int64 tensor -> (n,) int64 keys; order of keys == (x, y, z) lexicographic."""
    x = xyz.to(torch.int64) + INDEX_BIAS
    return (x[:, 0] << 42) | (x[:, 1] << 21) | x[:, 2]


def unpack_indices(keys):
    mask = (1 << 21) - 1
    return torch.stack([((keys >> 42) & mask) - INDEX_BIAS, ((keys >> 21) & mask) - INDEX_BIAS,
                        (keys & mask) - INDEX_BIAS], dim=1).to(torch.int32)


def merge_block_lists(local_xyz, group=None):
    """All ranks pass their (n_r, 3) int32 block list (any n_r, including 0); every rank gets the
    sorted unique union. One all_gather of the counts + one all_gather of the padded lists."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    local_xyz = local_xyz.reshape(-1, 3).to(torch.int32)
    if world == 1:
        keys = torch.unique(pack_indices(local_xyz))
        return unpack_indices(keys)
    dev = local_xyz.device
    n = torch.tensor([local_xyz.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    padded = torch.full((cap,), -1, dtype=torch.int64, device=dev)
    padded[:local_xyz.shape[0]] = pack_indices(local_xyz)
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    keys = torch.cat([g[:c] for g, c in zip(gathered, counts)])
    return unpack_indices(torch.unique(keys))


PAD = -(1 << 31)


def union_on_device(mapper, xyz_dev):
    """Sorted unique union of an (n,3) int32 CUDA tensor of block indices with the library's own kernels
    (mark into the union AABB's bitset + ordered ballot/popc compaction, nvb_blocks_union). Rows whose x is
    INT32_MIN are padding. Order: x fastest, then y, then z (the view calculator's order)."""
    import ctypes as C
    from ._lib import check
    xyz_dev = xyz_dev.reshape(-1, 3).contiguous()
    valid = xyz_dev[:, 0] != PAD
    if xyz_dev.shape[0] == 0 or not bool(valid.any()):
        return torch.zeros((0, 3), dtype=torch.int32, device=xyz_dev.device)
    v = xyz_dev[valid]
    lo = v.amin(dim=0).tolist()
    hi = v.amax(dim=0).tolist()
    out = torch.empty((xyz_dev.shape[0], 3), dtype=torch.int32, device=xyz_dev.device)
    n = C.c_int32(0)
    # order this call after the producer of xyz_dev (torch's current stream)
    torch.cuda.current_stream().synchronize()
    check(mapper._L.nvb_blocks_union(mapper._h, xyz_dev.data_ptr(), xyz_dev.shape[0], (C.c_int32 * 3)(*lo),
                                     (C.c_int32 * 3)(*hi), out.data_ptr(), out.shape[0], C.byref(n)))
    return out[:n.value]


def merge_block_lists_device(mapper, local_xyz, group=None):
    """GPU path of merge_block_lists: NCCL all-gather of the padded lists, then union_on_device.
    Returns the union in view-calculator order (x fastest); as a SET it equals merge_block_lists."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    local_xyz = local_xyz.reshape(-1, 3).to(torch.int32)
    if world == 1:
        return union_on_device(mapper, local_xyz)
    dev = local_xyz.device
    n = torch.tensor([local_xyz.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    cap = max(max(int(c.item()) for c in counts), 1)
    padded = torch.full((cap, 3), PAD, dtype=torch.int32, device=dev)
    padded[:local_xyz.shape[0]] = local_xyz
    gathered = torch.empty((world * cap, 3), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    return union_on_device(mapper, gathered)


def merge_updated_blocks(mapper, stream=None, group=None):
    """Merge of the ranks' TSDF block-index sets (what each camera's map replica touched so far).
    Returns the (m,3) int32 union as a tensor on the mapper's device."""
    idx = mapper.tsdf_layer().get_all_block_indices()  # synchronising read of the slab's index array
    if torch.cuda.is_available():
        local = torch.from_numpy(np.ascontiguousarray(idx)).to(torch.device("cuda", torch.cuda.current_device()))
        return merge_block_lists_device(mapper, local, group=group)
    return merge_block_lists(torch.from_numpy(np.ascontiguousarray(idx)), group=group)


def make_segment(xyz, cap):
    """(n,3) block indices -> the int32 segment [count, x0, y0, z0, ...] of capacity `cap` the device merge exchanges."""
    xyz = np.asarray(xyz, np.int32).reshape(-1, 3)
    seg = np.zeros(1 + 3 * cap, np.int32)
    n = min(len(xyz), cap)
    seg[0] = n
    seg[1:1 + 3 * n] = xyz[:n].reshape(-1)
    return seg


def union_segments_reference(segments, cap):
    """What nvb_blocks_union_segments computes, in numpy: sorted unique union of the gathered segments, x fastest, then y,
    then z. `segments`: (world, 1 + 3 * cap) int32."""
    segments = np.asarray(segments, np.int32).reshape(-1, 1 + 3 * cap)
    rows = [s[1:1 + 3 * int(min(max(s[0], 0), cap))].reshape(-1, 3) for s in segments]
    allxyz = np.concatenate(rows) if rows else np.zeros((0, 3), np.int32)
    if len(allxyz) == 0:
        return allxyz
    u = np.unique(allxyz, axis=0)
    return u[np.lexsort((u[:, 0], u[:, 1], u[:, 2]))]


class BatchMerger:
    """Device-resident merge of the ranks' updated-block lists, one merge per batch of frames.

    Every frame's block list is appended on the device to this rank's segment (no host copy of indices, no count exchange);
    `merge()` enqueues ONE fixed-size all-gather of the segments and the library's union kernels on a side stream, so the next
    batch's frames (on the mapper's streams) overlap it; `result()` is the only call that synchronises. Segments are
    double-buffered: the segment a merge is reading is not the one the next batch appends to.
    """

    def __init__(self, mapper, cap_entries, group=None):
        self.m, self.cap, self.group = mapper, int(cap_entries), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # the all-gather's NCCL kernel must be able to start (and wait for its peers) while a cooperative ESDF wavefront is in
        # flight: leave it SMs (measured, profiles/r2_run8.sh: 0.30 ms per merge with 0 reserved SMs, 0.10 ms with 4)
        if self.world > 1 and hasattr(mapper, "esdf_reserved_sms") and mapper.esdf_reserved_sms() < 4:
            mapper.esdf_reserved_sms(4)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.stride = 1 + 3 * self.cap
        self.local = [torch.zeros(self.stride, dtype=torch.int32, device=dev) for _ in range(2)]
        self.gathered = torch.zeros(self.world * self.stride, dtype=torch.int32, device=dev)
        self.out = torch.zeros((self.world * self.cap, 3), dtype=torch.int32, device=dev)
        self.count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ext = torch.cuda.ExternalStream(mapper.cuda_stream(), device=dev)
        self.comm = torch.cuda.Stream(device=dev)
        self.read_done = [None, None]
        self.cur = 0
        self.t0 = torch.cuda.Event(enable_timing=True)
        self.t1 = torch.cuda.Event(enable_timing=True)
        self.timed = []
        torch.cuda.synchronize()

    def append_last_frame(self):
        """Append the block list of the mapper's last integrated frame to this batch's segment (enqueued, no sync)."""
        from ._lib import check
        check(self.m._L.nvb_mapper_append_frame_blocks(self.m._h, self.local[self.cur].data_ptr(), self.cap))

    def merge(self, timed=False):
        """All-gather the segments of this batch and enqueue the union; start the next batch. No host synchronisation."""
        from ._lib import check
        seg = self.local[self.cur]
        self.comm.wait_stream(self.ext)  # the appends of this batch
        with torch.cuda.stream(self.comm):
            if timed:
                self.t0.record(self.comm)
            if self.world > 1:
                dist.all_gather_into_tensor(self.gathered, seg, group=self.group)
                src = self.gathered
            else:
                src = seg
            ev = torch.cuda.Event()
            ev.record(self.comm)
            self.read_done[self.cur] = ev
            check(self.m._L.nvb_blocks_union_segments(self.m._h, src.data_ptr(), self.world, self.stride, self.cap,
                                                      self.out.data_ptr(), self.out.shape[0], self.count.data_ptr(),
                                                      self.comm.cuda_stream))
            if timed:
                self.t1.record(self.comm)
                self.timed.append((self.t0, self.t1))
                self.t0, self.t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # next batch appends to the other segment, once the merge that last read it is done with it
        self.cur ^= 1
        if self.read_done[self.cur] is not None:
            self.ext.wait_event(self.read_done[self.cur])
        with torch.cuda.stream(self.ext):
            self.local[self.cur][:1].zero_()

    def result(self):
        """The last merge's union as an (n, 3) int32 tensor on the device (synchronises the merge stream)."""
        self.comm.synchronize()
        return self.out[:int(self.count.item())]

    def merge_ms(self):
        """Device time of the timed merges (all-gather + union kernels), in ms each."""
        self.comm.synchronize()
        return [a.elapsed_time(b) for a, b in self.timed]


def shard_frames(num_frames, rank, world):
    """Frame indices of `rank` when one stream of frames is dealt round-robin over `world` map
    replicas (time-slice sharding)."""
    return list(range(rank, num_frames, world))
