"""Layer serializers and layer streamers (second half).

Mirrors nvblox/include/nvblox/serialization/: `LayerSerializerGpu` (layer_serializer_gpu.h:52-77), `MeshSerializerGpu`
(mesh_serializer_gpu.h:96-134), `LayerStreamerBase` / `LayerStreamerOldestBlocks` (layer_streamer.h:52-262,
internal/impl/layer_streamer_impl.h). The serializers pack the requested blocks back to back with ONE gather kernel and one
device-to-host copy per array (nvb_layer_get_blocks / nvb_mesh_get_blocks of libnvblox_b200.so); the streamers are the host-side
bookkeeping that decides WHICH blocks go out under a block or byte budget: candidates not streamed yet, oldest-published first,
optional exclusion above a height / outside a radius.

Where the reference leaves an order undefined -- it walks an unordered_set and sorts equal priorities with std::sort -- this
implementation is deterministic: equal priorities go out in (x, y, z) order.
"""
import time

import numpy as np

INT64_MAX_AS_FLOAT = float(np.float32(np.iinfo(np.int64).max))


class BlockExclusionParams:
    """layer_streamer.h:31-47: all four optional; a functor is only installed when its inputs are set and > 0."""

    def __init__(self, exclusion_center_m=None, exclusion_height_m=None, exclusion_radius_m=None, block_size_m=None):
        self.exclusion_center_m = exclusion_center_m
        self.exclusion_height_m = exclusion_height_m
        self.exclusion_radius_m = exclusion_radius_m
        self.block_size_m = block_size_m


# ---------------------------------------------------------------------------------------------------------------------------
# Serializers
# ---------------------------------------------------------------------------------------------------------------------------
def serialize_voxel_layer(layer, block_indices):
    """LayerSerializerGpu::serialize (layer_serializer_gpu_impl.h:25-44): {"block_indices", "voxels", "block_offsets"}.
    voxels: (total_blocks_found * 512,) structured voxel array; block_offsets: (n + 1,) int32 in voxels, a block that is not in
    the layer contributes nothing (size 0), like a null block pointer in the reference."""
    idx = np.ascontiguousarray(block_indices, dtype=np.int32).reshape(-1, 3)
    n = idx.shape[0]
    offsets = np.zeros(n + 1, np.int32)
    if n == 0:
        return {"block_indices": idx, "voxels": np.zeros(0, layer._dtype), "block_offsets": offsets}
    blocks, found = layer.get_blocks(idx)
    offsets[1:] = np.cumsum(np.where(found, 512, 0))
    return {"block_indices": idx, "voxels": blocks[found].reshape(-1), "block_offsets": offsets}


def serialize_mesh_layer(mesh_layer, block_indices):
    """MeshSerializerGpu::serialize (src/serialization/mesh_serializer_gpu.cu:27-68): {"block_indices", "vertices" (v, 3) f32,
    "vertex_appearances" (c, 4) u8 RGBA, "triangle_indices" (t,) i32 relative to the block's first vertex,
    "vertex_block_offsets", "triangle_index_block_offsets" (n + 1,) i32}. Like in the reference the appearances share the
    vertices' offsets (a coloured block has one colour per vertex)."""
    idx = np.ascontiguousarray(block_indices, dtype=np.int32).reshape(-1, 3)
    n = idx.shape[0]
    v_off, t_off = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.int32)
    out = {"block_indices": idx, "vertices": np.zeros((0, 3), np.float32), "vertex_appearances": np.zeros((0, 4), np.uint8),
           "triangle_indices": np.zeros(0, np.int32), "vertex_block_offsets": v_off, "triangle_index_block_offsets": t_off}
    if n == 0:
        return out
    blocks = mesh_layer.get_blocks(idx)
    v_off[1:] = np.cumsum([0 if b is None else len(b["vertices"]) for b in blocks])
    t_off[1:] = np.cumsum([0 if b is None else len(b["triangles"]) for b in blocks])
    live = [b for b in blocks if b is not None]
    if live:
        out["vertices"] = np.concatenate([b["vertices"] for b in live]).reshape(-1, 3)
        out["vertex_appearances"] = np.concatenate([b["colors"] for b in live]).reshape(-1, 4)
        out["triangle_indices"] = np.concatenate([b["triangles"] for b in live])
    return out


def size_in_bytes(layer, block_index):
    """sizeInBytes(block) of the streamers' byte budget: sizeof(VoxelBlock) for voxel layers (blox.h), the four vectors of a
    MeshBlock (mesh_block.h:86-93: 12 B per vertex and per normal, 3 B per Color, 4 B per index). None if the block is absent."""
    if hasattr(layer, "block_sizes"):  # mesh layer
        sz = layer.block_sizes(np.asarray(block_index, np.int32).reshape(1, 3))[0]
        if sz[0] < 0:
            return None
        return int(sz[0]) * 12 * 2 + int(sz[2]) * 3 + int(sz[1]) * 4
    if not layer.is_block_allocated(block_index):
        return None
    return int(layer._dtype.itemsize) * 512


class _SizeCache:
    """Sizes of many blocks with one device query (the reference asks the host-side layer per block)."""

    def __init__(self, layer, indices):
        self.sizes = {}
        idx = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1, 3)
        if idx.shape[0] == 0:
            return
        if hasattr(layer, "block_sizes"):
            sz = layer.block_sizes(idx)
            for k, s in zip(idx, sz):
                self.sizes[tuple(int(c) for c in k)] = None if s[0] < 0 else int(s[0]) * 24 + int(s[2]) * 3 + int(s[1]) * 4
        else:
            _, found = layer.get_blocks(idx)
            per = int(layer._dtype.itemsize) * 512
            for k, f in zip(idx, found):
                self.sizes[tuple(int(c) for c in k)] = per if f else None


# ---------------------------------------------------------------------------------------------------------------------------
# Streamers
# ---------------------------------------------------------------------------------------------------------------------------
class LayerStreamerBase:
    """LayerStreamerBase (layer_streamer.h:65-172). Subclasses define compute_priorities()."""

    def __init__(self):
        self._index_set = set()
        self._exclude_block_functors = []

    # -- bookkeeping
    def mark_indices_candidates(self, block_indices):
        for k in np.asarray(block_indices, dtype=np.int64).reshape(-1, 3):
            self._index_set.add((int(k[0]), int(k[1]), int(k[2])))

    def num_candidates(self):
        return len(self._index_set)

    def clear(self):
        self._index_set.clear()

    def index_set(self):
        return set(self._index_set)

    def set_exclusion_functors(self, functors):
        self._exclude_block_functors = list(functors)

    def compute_priorities(self, block_indices):
        raise NotImplementedError

    # -- queries
    def get_n_blocks(self, num_blocks):
        """getNBlocks (layer_streamer_impl.h:51-67)."""
        state = {"n": 0}

        def status(_idx):
            ok = state["n"] < num_blocks
            if ok:
                state["n"] += 1
            return ok, False, not ok

        return self._get_highest_priority_blocks(status)

    def get_n_bytes_of_blocks(self, num_bytes, layer):
        """getNBytesOfBlocks (:69-92): the running sum INCLUDES the block that crosses the limit, which is not streamed."""
        state = {"bytes": 0}
        cache = _SizeCache(layer, list(self._index_set))

        def status(idx):
            size = cache.sizes.get(idx)
            if size is None:
                return False, True, False  # not in the layer: stop tracking it
            state["bytes"] += size
            ok = state["bytes"] < num_bytes
            return ok, False, not ok

        return self._get_highest_priority_blocks(status)

    def _exclude_blocks(self, indices):
        if not self._exclude_block_functors:
            return indices
        return [k for k in indices if not any(f(k) for f in self._exclude_block_functors)]

    def _get_highest_priority_blocks(self, get_stream_status):
        """getHighestPriorityBlocks (:104-165): excluded blocks are dropped from the tracking set for good."""
        index_vec = sorted(self._index_set)
        self._index_set.clear()
        index_vec = self._exclude_blocks(index_vec)
        priorities = self.compute_priorities(index_vec)
        order = sorted(range(len(index_vec)), key=lambda i: (-priorities[i], index_vec[i]))
        out = []
        rest_from = -1
        for pos, i in enumerate(order):
            k = index_vec[i]
            stream, invalid, limit = get_stream_status(k)
            if stream:
                out.append(k)
            elif not invalid:
                self._index_set.add(k)
            if limit:
                rest_from = pos + 1
                break
        if rest_from > 0:
            for i in order[rest_from:]:
                self._index_set.add(index_vec[i])
        return np.asarray(out, dtype=np.int32).reshape(-1, 3)

    # -- serialization
    @staticmethod
    def serialize_all_blocks(layer, block_indices):
        """serializeAllBlocks (:186-192)."""
        return serialize_mesh_layer(layer, block_indices) if hasattr(layer, "block_sizes") else serialize_voxel_layer(layer, block_indices)


class LayerStreamerOldestBlocks(LayerStreamerBase):
    """LayerStreamerOldestBlocks (layer_streamer.h:177-262): blocks never streamed first, then the longest-ago streamed."""

    def __init__(self):
        super().__init__()
        self._publishing_index = 0
        self._last_published = {}
        self._tick_times = []

    def last_published_map(self):
        return dict(self._last_published)

    def compute_priorities(self, block_indices):
        """computePriority (:240-249): float(int64 max) for a block never streamed, else float(-last publishing index)."""
        return [INT64_MAX_AS_FLOAT if k not in self._last_published else float(np.float32(-1 * self._last_published[k]))
                for k in block_indices]

    def _update_last_publish_index(self, block_indices):
        for k in np.asarray(block_indices, dtype=np.int64).reshape(-1, 3):
            self._last_published[(int(k[0]), int(k[1]), int(k[2]))] = self._publishing_index
        self._publishing_index += 1

    def _setup_exclusion_functors(self, p):
        """setupExclusionFunctors (:262-284) + the two functor factories (:286-312), in binary32 like the reference."""
        functors = []
        if p.exclusion_height_m is not None and p.block_size_m is not None and p.exclusion_height_m > 0.0:
            h, bs = np.float32(p.exclusion_height_m), np.float32(p.block_size_m)
            functors.append(lambda k: bool(np.float32(k[2]) * bs > h))
        if p.block_size_m is not None and p.exclusion_center_m is not None and p.exclusion_radius_m is not None and p.exclusion_radius_m > 0.0:
            bs = np.float32(p.block_size_m)
            c = np.asarray(p.exclusion_center_m, np.float32)
            r2 = np.float32(p.exclusion_radius_m) * np.float32(p.exclusion_radius_m)

            def outside(k):
                d = bs * (np.asarray(k, np.float32) + np.float32(0.5)) - c  # getCenterPositionFromBlockIndex
                return bool(d[0] * d[0] + (d[1] * d[1] + d[2] * d[2]) > r2)  # Eigen's 3-vector squaredNorm order

            functors.append(outside)
        self.set_exclusion_functors(functors)

    def get_n_blocks(self, num_blocks, block_exclusion_params=None):
        self._setup_exclusion_functors(block_exclusion_params or BlockExclusionParams())
        out = super().get_n_blocks(num_blocks)
        self._update_last_publish_index(out)
        return out

    def get_n_bytes_of_blocks(self, num_bytes, layer, block_exclusion_params=None):
        self._setup_exclusion_functors(block_exclusion_params or BlockExclusionParams())
        out = super().get_n_bytes_of_blocks(num_bytes, layer)
        self._update_last_publish_index(out)
        return out

    def get_n_bytes_of_serialized_blocks(self, num_bytes, layer, block_exclusion_params=None):
        return self.serialize_all_blocks(layer, self.get_n_bytes_of_blocks(num_bytes, layer, block_exclusion_params))

    def estimate_bandwidth_and_serialize(self, layer, blocks_to_serialize, block_exclusion_params=None, bandwidth_limit_mbps=-1.0,
                                         now_s=None):
        """estimateBandwidthAndSerialize (:314-352): the byte budget of this call = bandwidth limit x the measured period of the
        calls (mean rate clamped to 1..100 Hz; timing::Rates keeps a window of the last ticks), unlimited if the limit is < 0."""
        self.mark_indices_candidates(blocks_to_serialize)
        t = time.monotonic() if now_s is None else float(now_s)
        self._tick_times = (self._tick_times + [t])[-100:]
        rate = 0.0
        if len(self._tick_times) >= 2 and self._tick_times[-1] > self._tick_times[0]:
            rate = (len(self._tick_times) - 1) / (self._tick_times[-1] - self._tick_times[0])
        rate = max(1.0, min(100.0, rate))
        if bandwidth_limit_mbps < 0:
            num_bytes = np.iinfo(np.uint64).max
        else:
            num_bytes = int(np.float32(bandwidth_limit_mbps) * np.float32(1.0 / rate) * np.float32(1e6 / 8.0))
        return self.get_n_bytes_of_serialized_blocks(num_bytes, layer, block_exclusion_params)


class LayerCakeStreamer:
    """LayerCakeStreamer (layer_cake_streamer.h:25-98): one LayerStreamerOldestBlocks per layer kind ("tsdf", "esdf", "mesh", ...)
    behind one object; a request for a kind that was not added does nothing and returns None."""

    def __init__(self, *kinds):
        self._streamers = {}
        for k in kinds:
            self.add(k)

    def add(self, kind):
        self._streamers.setdefault(kind, LayerStreamerOldestBlocks())  # one of each kind at most

    def get(self, kind):
        return self._streamers.get(kind)

    def estimate_bandwidth_and_serialize(self, kind, layer, blocks_to_serialize, block_exclusion_params=None, bandwidth_limit_mbps=-1.0,
                                         now_s=None):
        s = self._streamers.get(kind)
        return None if s is None else s.estimate_bandwidth_and_serialize(layer, blocks_to_serialize, block_exclusion_params,
                                                                         bandwidth_limit_mbps, now_s)

    def serialize_all_blocks(self, kind, layer, block_indices):
        s = self._streamers.get(kind)
        return None if s is None else s.serialize_all_blocks(layer, block_indices)
