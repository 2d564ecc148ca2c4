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


def reduce_counters(dist, csum, cmax, hist, group=None):
    """In-place all-reduce of the three buffers (torch int64 tensors on the backend's device)."""
    assert csum.numel() == _abi.N_SUM and cmax.numel() == _abi.N_MAX
    dist.all_reduce(csum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(cmax, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return csum, cmax, hist
