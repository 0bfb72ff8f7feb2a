"""One-process-per-GPU plumbing (torch.distributed over NCCL / gloo).

The ensemble shards trivially: rank r integrates the trajectory block [lo, hi) of the global index
range (Philox counters are global indices, so the sample set does not depend on the world size) and
the per-time arrays are summed with ONE all-reduce per arena (f64 sums, u64 counts).  This mirrors
the reference's `@distributed (+)` over Julia workers (src/TWA.jl:115-128,163,369).
"""
from __future__ import annotations


def shard_bounds(N: int, rank: int, world: int):
    """contiguous block [lo, hi) of rank `rank`, balanced: N // world trajectories each, the first N % world ranks
    one more.  (src/TWA.jl:119-128 hands out ceil(N/workers) per work item and leaves the last ones short or empty;
    the balanced split covers the same index range with the same samples and never leaves a GPU idle.)"""
    N, rank, world = int(N), int(rank), int(world)
    base, extra = divmod(N, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def chunk_plan(N: int, world: int, maxNBatch=None):
    """(chunk, nchunks) for maxNBatch-bounded launches (src/TWA.jl:119).  Both depend on N and world only, never on the
    rank: every rank runs the same number of launches -- and therefore the same number of collectives -- even when its own
    block is shorter or empty (it then launches an empty chunk)."""
    per_max = -(-int(N) // int(world))
    chunk = per_max if (maxNBatch is None or maxNBatch == float("inf")) else int(max(1, min(per_max, maxNBatch)))
    return chunk, -(-per_max // chunk)


class _CudaArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def torch_allreduce(device_index: int, group=None):
    """Returns the callable monte_carlo_integrate hands the two device arenas to: wraps the raw
    pointers as torch tensors (no copy) and sums them over the ranks on the current stream."""
    import torch
    import torch.distributed as dist

    def allreduce(p_f64, n_f64, p_u64, n_u64):
        dev = torch.device("cuda", device_index)
        if n_f64:
            t = torch.as_tensor(_CudaArray(p_f64, n_f64, "<f8"), device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        if n_u64:
            t = torch.as_tensor(_CudaArray(p_u64, n_u64, "<i8"), device=dev)   # two's complement: same bits as u64 sum
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        # dtwa_plan_fetch reads the arenas on the library's stream: make the collectives visible first
        torch.cuda.current_stream(dev).synchronize()
    return allreduce


def reduce_host_arrays(arrays, group=None):
    """all-reduce (sum) a list of numpy arrays in place through torch.distributed -- the CPU-side twin
    used by the gloo tests and by the reference arm."""
    import numpy as np
    import torch
    import torch.distributed as dist
    for a in arrays:
        if a.dtype == np.uint64:
            t = torch.from_numpy(a.view(np.int64))
        else:
            t = torch.from_numpy(a)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return arrays
