"""Layer streamers (isaac_ros_nvblox_b200/streamer.py), restating the reference's own tests
(nvblox/tests/test_layer_streamer.cpp) on the host logic; the serializers and the byte budgets against real layers are in
tests/test_gpu_streamer.py."""
import numpy as np
import pytest

from isaac_ros_nvblox_b200 import streamer as st


class XPriorityStreamer(st.LayerStreamerBase):
    """SimpleLayerStreamer (test_layer_streamer.cpp:24-47): priority = the block's x index."""

    def compute_priorities(self, block_indices):
        return [float(k[0]) for k in block_indices]


class FakeLayer:
    """A voxel layer of 4096-byte blocks that holds the given indices."""
    _dtype = np.dtype([("distance", "<f4"), ("weight", "<f4")])

    def __init__(self, indices):
        self.have = {tuple(int(c) for c in k) for k in indices}

    def is_block_allocated(self, k):
        return tuple(int(c) for c in k) in self.have

    def get_blocks(self, idx):
        found = np.array([tuple(int(c) for c in k) in self.have for k in idx], bool)
        return np.zeros((len(idx), 8, 8, 8), self._dtype), found


def _random_indices(n, rng, unique=False, lo=-1000, hi=1000):
    if not unique:
        return rng.integers(lo, hi + 1, (n, 3))
    s = set()
    while len(s) < n:
        s.add(tuple(int(c) for c in rng.integers(lo, hi + 1, 3)))
    return np.array(sorted(s))


def test_simple_priority():
    """SimplePriorityTest (:66-91): the 95 highest-x blocks come out in non-increasing x, the rest stays tracked."""
    rng = np.random.default_rng(0)
    s = XPriorityStreamer()
    s.mark_indices_candidates(_random_indices(100, rng, unique=True))
    out = s.get_n_blocks(95)
    assert len(out) == 95 and np.all(np.diff(out[:, 0]) <= 0)
    assert all(k[0] <= out[-1, 0] for k in s.index_set()) and s.num_candidates() == 5


def test_request_more_than_available_and_zero():
    """RequestMoreThanAvailable (:93-103), RequestZero (:105-113)."""
    rng = np.random.default_rng(1)
    s = XPriorityStreamer()
    s.mark_indices_candidates(_random_indices(50, rng, unique=True))
    assert len(s.get_n_blocks(0)) == 0 and s.num_candidates() == 50
    assert len(s.get_n_blocks(100)) == 50 and s.num_candidates() == 0
    s.mark_indices_candidates([[1, 2, 3], [1, 2, 3]])  # a set
    assert s.num_candidates() == 1
    s.clear()
    assert s.num_candidates() == 0


def test_oldest_blocks_first():
    """LayerStreamerOldestBlocks (:177-208): two halves, re-marked, come out in the same halves again."""
    rng = np.random.default_rng(2)
    s = st.LayerStreamerOldestBlocks()
    s.mark_indices_candidates(_random_indices(100, rng, unique=True))
    assert s.num_candidates() == 100
    first, second = s.get_n_blocks(50), s.get_n_blocks(50)
    assert len(first) == 50 and len(second) == 50 and s.num_candidates() == 0
    as_set = lambda a: {tuple(k) for k in a.tolist()}
    assert not (as_set(first) & as_set(second))
    s.mark_indices_candidates(first), s.mark_indices_candidates(second)
    first2, second2 = s.get_n_blocks(50), s.get_n_blocks(50)
    assert as_set(first2) == as_set(first) and as_set(second2) == as_set(second)
    pm = s.last_published_map()
    assert {pm[k] for k in as_set(first)} == {2} and {pm[k] for k in as_set(second)} == {3}
    # a block never streamed beats every streamed one (priority float(int64 max) vs -index, layer_streamer_impl.h:240-249)
    s.mark_indices_candidates(first), s.mark_indices_candidates([[5000, 0, 0]])
    assert s.get_n_blocks(1).tolist() == [[5000, 0, 0]]


def test_exclusion_functors():
    """getExcludeAboveHeightFunctor / getExcludeOutsideRadiusFunctor (layer_streamer_impl.h:286-312): a block is excluded if
    its LOW z face is above the height / its centre is outside the radius; excluded blocks leave the tracking set."""
    s = st.LayerStreamerOldestBlocks()
    idx = np.array([[x, 0, z] for x in range(-5, 6) for z in range(-2, 6)])
    s.mark_indices_candidates(idx)
    p = st.BlockExclusionParams(exclusion_height_m=1.0, block_size_m=0.4)
    out = s.get_n_blocks(1000, p)
    assert len(out) > 0 and out[:, 2].max() == 2  # 0.4 * 3 = 1.2 > 1.0 is excluded, 0.4 * 2 = 0.8 is not
    assert s.num_candidates() == 0  # the excluded ones are gone too
    s.mark_indices_candidates(idx)
    p = st.BlockExclusionParams(exclusion_center_m=(0.0, 0.0, 0.0), exclusion_radius_m=1.0, block_size_m=0.4)
    out = s.get_n_blocks(1000, p)
    c = 0.4 * (out + 0.5)
    assert len(out) > 0 and np.all((c * c).sum(axis=1) <= 1.0 + 1e-6)
    far = 0.4 * (np.array([[3, 0, 0]]) + 0.5)
    assert (far * far).sum() > 1.0 and [3, 0, 0] not in out.tolist()
    # params without a block size install nothing (:266-283)
    s.mark_indices_candidates(idx)
    assert len(s.get_n_blocks(1000, st.BlockExclusionParams(exclusion_height_m=1.0))) == len(idx)


def test_byte_budget_semantics():
    """getNBytesOfBlocks (layer_streamer_impl.h:69-92): blocks are streamed while the running sum (which includes the block
    under test) stays BELOW the budget; a block that is not in the layer is dropped from the tracking set."""
    have = np.array([[i, 0, 0] for i in range(10)])
    layer = FakeLayer(have)
    s = st.LayerStreamerOldestBlocks()
    s.mark_indices_candidates(have)
    s.mark_indices_candidates([[100, 0, 0]])  # not in the layer
    out = s.get_n_bytes_of_blocks(4096 * 4, layer)
    assert len(out) == 3  # 3 * 4096 < 16384; the fourth makes the sum equal to the budget -> not streamed
    assert s.num_candidates() == 8 and (100, 0, 0) in s.index_set()  # blocks behind the limit are not even looked at
    out2 = s.get_n_bytes_of_blocks(1 << 40, layer)
    assert len(out2) == 7 and s.num_candidates() == 0
    assert st.size_in_bytes(layer, [0, 0, 0]) == 4096 and st.size_in_bytes(layer, [100, 0, 0]) is None


def test_bandwidth_estimate():
    """estimateBandwidthAndSerialize (layer_streamer_impl.h:314-352): budget = limit [Mbit/s] x the measured call period."""
    have = np.array([[i, 0, 0] for i in range(200)])
    layer = FakeLayer(have)
    s = st.LayerStreamerOldestBlocks()
    got = s.estimate_bandwidth_and_serialize(layer, have, bandwidth_limit_mbps=-1.0, now_s=0.0)
    assert len(got["block_indices"]) == 200  # unlimited
    # 10 Hz calls at 8 Mbit/s -> 100 000 bytes per call -> 24 blocks of 4096 B (the 25th reaches 102 400 >= 100 000)
    n = []
    for i in range(1, 6):
        n.append(len(s.estimate_bandwidth_and_serialize(layer, have, bandwidth_limit_mbps=8.0, now_s=0.1 * i)["block_indices"]))
    assert n[-1] == 24 and s.num_candidates() > 0


def test_layer_cake_streamer():
    """LayerCakeStreamer (nvblox/tests/test_layer_cake_streamer.cpp): a streamer per layer kind; unknown kinds do nothing."""
    have = np.array([[i, 0, 0] for i in range(20)])
    layer = FakeLayer(have)
    cake = st.LayerCakeStreamer("tsdf", "mesh")
    cake.add("tsdf")  # already there
    assert cake.get("tsdf") is not None and cake.get("esdf") is None
    got = cake.estimate_bandwidth_and_serialize("tsdf", layer, have)
    assert len(got["block_indices"]) == 20 and cake.get("tsdf").num_candidates() == 0
    assert cake.estimate_bandwidth_and_serialize("esdf", layer, have) is None
    assert len(cake.serialize_all_blocks("tsdf", layer, have[:3])["block_indices"]) == 3
