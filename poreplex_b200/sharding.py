"""Multi-GPU plumbing: reads shard embarrassingly, one collective at the end.

The reference's only parallelism is a process pool over independent 128-read batches
(pipeline.py:96,204-205) and its only aggregation is FinalSummaryTracker.feed_results
(io.py:274-278).  Here: contiguous block partition of read indices over ranks, no
data-path exchange, and ONE all-reduce(sum) of the int64[4,5,11] count tensor
(NCCL on GPUs, gloo in the CPU tests)."""
import torch.distributed as dist

__all__ = ['shard_range', 'reduce_counts']


def shard_range(n_reads, rank, world):
    """Reads [lo, hi) owned by `rank`: GPU g gets [g*N/G, (g+1)*N/G)."""
    return (n_reads * rank) // world, (n_reads * (rank + 1)) // world


def reduce_counts(counts):
    """In-place all-reduce(sum) of the per-(label, barcode slot, status) counts."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts
