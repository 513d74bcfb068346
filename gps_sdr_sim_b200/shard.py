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


def link_aware_shares(total: int, rates) -> list[int]:
    """Cut `total` epochs among ranks in proportion to their measured host-link rates (the GPUs of a box do not
    share the host link evenly: profiles/r02_pcie_8gpu.md).  Every rank gets at least one epoch; the shares sum to
    `total` exactly (the rounding remainder goes to the fastest link).  Used by bench.py's e2e step."""
    rates = [float(r) for r in rates]
    if total < len(rates) or not rates or min(rates) <= 0.0:
        raise ValueError("need at least one epoch per rank and positive rates")
    s = sum(rates)
    shares = [max(1, int(total * r / s)) for r in rates]
    fastest = max(range(len(rates)), key=lambda i: rates[i])
    shares[fastest] += total - sum(shares)
    if shares[fastest] < 1:      # the floor of 1 epoch for very slow links overdrew the budget
        raise ValueError("rates too uneven for this many epochs")
    return shares


def repeats_of(share: int, table_epochs: int) -> list[int]:
    """Epoch counts of the passes a rank makes over a table of `table_epochs` epochs to generate `share` epochs
    (whole passes, then the remainder from the start of the scenario)."""
    reps, rest = divmod(share, table_epochs)
    return [table_epochs] * reps + ([rest] if rest else [])
