"""Launch plumbing of the multi-GPU runs (one process per GPU): rank / world size from the torchrun environment, the
read range of a rank under weak scaling, and the max / sum reductions of bench.py's timing contract (`backend` is "nccl"
on GPUs and "gloo" in the CPU tests).  The data path of a multi-GPU encode - ONE global container from N ranks - is
uq_b200/multigpu.py; `bench.py --multi shards` (independent container shards per rank, no data-path collective) uses only
what is here."""
import os


def world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(rank, world_size, reads_per_rank):
    """(first read index, number of reads) of a rank under weak scaling."""
    return rank * reads_per_rank, reads_per_rank


def split_evenly(n_total, rank, world_size):
    """(first, count) of a rank when a fixed total is divided (strong scaling, remainder to the low ranks)."""
    base, rem = divmod(n_total, world_size)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


class Reducer:
    """max / sum over ranks of python floats through torch.distributed (no-op for a single rank)."""

    def __init__(self, dist=None, device="cpu"):
        self.dist, self.device = dist, device

    def _reduce(self, x, op):
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, None if self.dist is None else self.dist.ReduceOp.MAX)

    def sum(self, x):
        return self._reduce(x, None if self.dist is None else self.dist.ReduceOp.SUM)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()


def job_throughput(reducer, reads_this_rank, seconds_this_rank):
    """whole-job reads/s = reads of all ranks / slowest rank's time (the bench contract)."""
    return reducer.sum(reads_this_rank) / reducer.max(seconds_this_rank)
