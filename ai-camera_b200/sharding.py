"""Multi-GPU sharding of the hot path: independent video streams are partitioned across ranks
(one process per GPU); there is NO collective on the data path.  The only exchange is a final
gather of a few per-rank statistics (NCCL on GPUs, gloo in the CPU tests)."""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def stream_partition(n_streams: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [first, first + count) of the global streams owned by `rank`
    (stream s -> GPU floor(s / ceil(S / world))).  Blocks differ by at most one."""
    if not (0 <= rank < world_size) or n_streams < 0:
        raise ValueError("bad partition arguments")
    base, extra = divmod(n_streams, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def owner_of_stream(stream: int, n_streams: int, world_size: int) -> int:
    for r in range(world_size):
        first, count = stream_partition(n_streams, world_size, r)
        if first <= stream < first + count:
            return r
    raise ValueError("stream out of range")


def gather_stats(values: Sequence[float], device=None) -> List[List[float]]:
    """All ranks contribute a short vector of float64 statistics; every rank gets all of them
    (row r = rank r).  Without an initialised process group returns [values]."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [t.cpu().tolist()]
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.cpu().tolist() for o in out]


def job_throughput(units_per_rank: Sequence[float], seconds_per_rank: Sequence[float]) -> float:
    """Whole-job units/s: all units divided by the slowest rank's time."""
    return float(sum(units_per_rank)) / max(seconds_per_rank)
