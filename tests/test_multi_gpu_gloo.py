"""World-size-2 gloo test of the multi-GPU host logic (block-list merge, frame sharding). CPU only."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from isaac_ros_nvblox_b200 import multi_gpu
    rng = np.random.default_rng(100 + rank)
    n = [700, 0, 333][rank % 3] if rank else 500
    local = rng.integers(-40, 40, size=(n, 3)).astype(np.int32)
    local[: n // 4] = np.array([1, 2, 3], np.int32)  # duplicates inside one list
    merged = multi_gpu.merge_block_lists(torch.from_numpy(local))
    q.put((rank, local, merged.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _run(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res, key=lambda t: t[0])


def test_merge_block_lists_world2():
    res = _run(2)
    allxyz = np.concatenate([r[1] for r in res])
    want = np.unique(allxyz, axis=0)  # lexicographic (x, y, z) sort
    for rank, _, merged in res:
        assert merged.dtype == np.int32
        assert np.array_equal(merged, want), "rank %d union differs from sort(unique(concat))" % rank


def test_merge_block_lists_world3_with_empty_rank():
    res = _run(3)
    want = np.unique(np.concatenate([r[1] for r in res]), axis=0)
    for _, _, merged in res:
        assert np.array_equal(merged, want)


def test_pack_roundtrip_and_order():
    from isaac_ros_nvblox_b200 import multi_gpu
    xyz = torch.tensor([[-(1 << 20), 0, (1 << 20) - 1], [0, 0, 0], [-1, -1, -1], [5, -7, 9]], dtype=torch.int32)
    keys = multi_gpu.pack_indices(xyz)
    assert torch.equal(multi_gpu.unpack_indices(keys), xyz)
    order = torch.argsort(keys)
    want = np.lexsort((xyz[:, 2].numpy(), xyz[:, 1].numpy(), xyz[:, 0].numpy()))
    assert order.tolist() == want.tolist()


def test_shard_frames_partition():
    from isaac_ros_nvblox_b200 import multi_gpu
    for world in (1, 2, 4, 8):
        parts = [multi_gpu.shard_frames(80, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(80))


def _worker_segments(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from isaac_ros_nvblox_b200 import multi_gpu
    cap = 512
    rng = np.random.default_rng(7 + rank)
    n = [300, 0, 512][rank % 3]
    local = rng.integers(-20, 20, size=(n, 3)).astype(np.int32)
    seg = torch.from_numpy(multi_gpu.make_segment(local, cap))
    gathered = torch.zeros(world * (1 + 3 * cap), dtype=torch.int32)
    dist.all_gather_into_tensor(gathered, seg)  # ONE fixed-size collective: no count exchange
    union = multi_gpu.union_segments_reference(gathered.numpy(), cap)
    q.put((rank, local, union))
    dist.barrier()
    dist.destroy_process_group()


def test_fixed_size_segment_exchange_world3():
    """The device merge's protocol (segments [count, xyz...] of fixed capacity, one all-gather, union x fastest) on gloo."""
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_segments, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    u = np.unique(np.concatenate([r[1] for r in res]), axis=0)
    want = u[np.lexsort((u[:, 0], u[:, 1], u[:, 2]))]
    for _, _, union in res:
        assert np.array_equal(union, want)
