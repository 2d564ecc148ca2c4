"""Ray sharding and the per-frame histogram reduce of the multi-GPU path.

Rays are independent (the reference walks the lanes one by one with no cross-ray state except counters and the
1-based rayCount, vulkan_ray_tracing.cc:1512,1665), so a frame is split into contiguous ray-id blocks, one per
rank, aligned to the reference's warp size so that a warp's lanes never straddle two ranks.  The BVH and the
treelet tables are replicated.  The only exchange is the reduce of the functional counters (SUM over the first
N_SUM fields, MAX over the last N_MAX) and of the per-treelet visit histogram (SUM)."""
from . import _abi

WARP = 32


def shard_range(total, world, rank, align=WARP):
    """[first, first+count) of rank `rank`: contiguous, multiples of `align` except the tail of the last rank."""
    units = (total + align - 1) // align
    lo = units * rank // world * align
    hi = units * (rank + 1) // world * align
    return min(lo, total), max(0, min(hi, total) - min(lo, total))


def pack_reduce_header(counters, prev_counters, world, rank):
    """The header the library's reduce sends (csrc/reduce.cu, k_snapshot), restated for the CPU tests: uint64 words
    [0, N_SUM) deltas of the SUM counters since the previous reduce | [N_SUM, N_SUM + N_MAX * world) the MAX counters in this
    rank's own pair, zeros elsewhere -- so that ONE sum-reduce delivers every rank's maxima -- | records added by this rank | 0."""
    import numpy as np
    h = np.zeros(_abi.N_SUM + _abi.N_MAX * world + 2, dtype=np.int64)
    h[:_abi.N_SUM] = np.asarray(counters[:_abi.N_SUM], dtype=np.int64) - np.asarray(prev_counters[:_abi.N_SUM], dtype=np.int64)
    h[_abi.N_SUM + _abi.N_MAX * rank:_abi.N_SUM + _abi.N_MAX * (rank + 1)] = counters[_abi.N_SUM:_abi.N_SUM + _abi.N_MAX]
    h[-2] = h[:9].sum()
    return h


def fold_reduce_header(total, header, world):
    """k_fold of csrc/reduce.cu: global totals += reduced deltas; MAX over the per-rank pairs.  Returns (totals, overflow)."""
    import numpy as np
    out = np.array(total, dtype=np.int64)
    out[:_abi.N_SUM] += header[:_abi.N_SUM]
    pairs = np.asarray(header[_abi.N_SUM:_abi.N_SUM + _abi.N_MAX * world]).reshape(world, _abi.N_MAX)
    out[_abi.N_SUM:] = np.maximum(out[_abi.N_SUM:], pairs.max(axis=0))
    return out, bool(header[-2] >= (1 << 32) or header[-1] != 0)


def reduce_counters(dist, csum, cmax, hist, group=None):
    """In-place all-reduce of the three buffers (torch int64 tensors on the backend's device)."""
    assert csum.numel() == _abi.N_SUM and cmax.numel() == _abi.N_MAX
    dist.all_reduce(csum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(cmax, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return csum, cmax, hist
