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


def shard_frames(num_frames, rank, world):
    """Frame indices of `rank` when one stream of frames is dealt round-robin over `world` map
    replicas (time-slice sharding)."""
    return list(range(rank, num_frames, world))
