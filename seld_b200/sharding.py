"""Clip sharding for the multi-GPU form of the path: clips are independent through
extraction, normalisation and masking, so clip ``i`` goes to rank ``i mod world`` and the only cross-rank step is
the all-reduce of the per-bin statistics (pipeline.allreduce_statistics)."""


def shard_indices(n_clips: int, rank: int, world: int):
    """Indices of the clips owned by ``rank`` (round-robin: sizes differ by at most one)."""
    if not 0 <= rank < world:
        raise ValueError('rank out of range')
    return list(range(rank, n_clips, world))
