"""Multi-GPU sharding of independent system instances and assembly of the gathered draw stream.

System instances are fully independent (own RNG, own id spaces; ParticleSystem.cpp:241-251,:265-267),
so instance i lives wholly on one rank and no collective is needed during emit/update/compact/sort.
Once per rendered frame the per-rank runs of compacted 32-byte records are all-gathered (rank-major)
and expanded 4x on every rank. This module holds the host-side bookkeeping
only; it is exercised on CPU with gloo (tests/test_multi_rank.py) and on GPUs by bench.py.
"""
import numpy as np


def shard_instances(num_instances, world, rank):
    """Contiguous block partition: instance ids owned by `rank` (seed[1] = 2 + instance id)."""
    base, rem = divmod(num_instances, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def run_layout(counts_per_rank, capacity):
    """Offsets of every rank's run inside the rank-major padded all-gather buffer.

    counts_per_rank[r] = records rank r contributed (<= capacity). Returns (offsets, total) where
    offsets[r] is the record index of rank r's first record in the gathered buffer."""
    counts = np.asarray(counts_per_rank, dtype=np.int64)
    if (counts > capacity).any():
        raise ValueError("a rank's run exceeds the padded capacity")
    return (np.arange(len(counts), dtype=np.int64) * capacity), int(counts.sum())


def compact_gathered(gathered, counts_per_rank, capacity):
    """Drop the padding of a rank-major gathered buffer: [world*capacity, ...] -> [sum(counts), ...]."""
    offsets, _ = run_layout(counts_per_rank, capacity)
    return np.concatenate([gathered[o:o + c] for o, c in zip(offsets, counts_per_rank)], axis=0)


def expand4(records):
    """4x replication into the vertex layout the renderer consumes (ParticleSystem.cpp:607-665)."""
    return np.repeat(records, 4, axis=0)
