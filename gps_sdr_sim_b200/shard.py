"""Time-sharding of epochs across the GPUs of one box.

Once the host has produced the rows, every epoch is independent: the only quantity the
reference carried from one epoch to the next, the carrier phase, is in the row.  So the
path shards by time with no exchange step and no collective: rank r of W generates a
contiguous range of epochs and writes it at byte offset first_epoch*epoch_bytes.
"""
from __future__ import annotations


def epoch_range(rank: int, world_size: int, n_epochs: int) -> tuple[int, int]:
    """(first_epoch, count) of `rank`: contiguous, balanced to within one epoch, covers all."""
    if not (0 <= rank < world_size) or n_epochs < 0:
        raise ValueError("bad rank / world_size / n_epochs")
    base, extra = divmod(n_epochs, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def batches(first: int, count: int, max_batch: int):
    """Split [first, first+count) into consecutive (first, count) batches of <= max_batch epochs."""
    if max_batch < 1:
        raise ValueError("max_batch must be >= 1")
    done = 0
    while done < count:
        n = min(max_batch, count - done)
        yield first + done, n
        done += n
