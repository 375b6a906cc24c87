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


def merge_updated_blocks(mapper, stream=None, group=None):
    """Merge of the ranks' TSDF block-index sets (what each camera's map replica touched so far).
    Returns the (m,3) int32 union as a tensor on the mapper's device."""
    idx = mapper.tsdf_layer().get_all_block_indices()  # synchronising read of the slab's index array
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    local = torch.from_numpy(np.ascontiguousarray(idx)).to(dev)
    return merge_block_lists(local, group=group)


def shard_frames(num_frames, rank, world):
    """Frame indices of `rank` when one stream of frames is dealt round-robin over `world` map
    replicas (time-slice sharding)."""
    return list(range(rank, num_frames, world))
