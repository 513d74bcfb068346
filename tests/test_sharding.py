"""Multi-GPU host logic on CPU: time-sharding of epochs, world_size 2 over gloo.

The path shards by epoch with no data-path collective; what has to be right is
the partition and that a rank's slice of rows generates exactly its byte range of the whole
output.  The sample generation here is done by the oracle - this test is about the sharding.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib
from conftest import load_golden
from gps_sdr_sim_b200.shard import batches, epoch_range, link_aware_shares, repeats_of


def test_epoch_range_partitions_exactly():
    for n in (0, 1, 7, 299, 2999, 863999):
        for w in (1, 2, 3, 4, 8):
            got = [epoch_range(r, w, n) for r in range(w)]
            assert got[0][0] == 0 and sum(c for _, c in got) == n
            for (f0, c0), (f1, _) in zip(got, got[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in got) - min(c for _, c in got) <= 1
    with pytest.raises(ValueError):
        epoch_range(2, 2, 10)


def test_batches_cover_range_in_order():
    assert list(batches(5, 0, 4)) == []
    assert list(batches(5, 10, 4)) == [(5, 4), (9, 4), (13, 2)]
    assert list(batches(0, 8, 8)) == [(0, 8)]


def test_link_aware_shares_sum_and_follow_the_rates():
    # the 8-GPU box of profiles/r02_pcie_8gpu.md: four links at 11.6 GB/s, four at 18.5
    rates = [11.6] * 4 + [18.5] * 4
    sh = link_aware_shares(8 * 2999, rates)
    assert sum(sh) == 8 * 2999 and min(sh) >= 1
    assert sh[4] > 2999 > sh[0]                                  # the fast links take more than one table's worth
    assert abs(sh[0] / sh[5] - 11.6 / 18.5) < 0.01
    assert link_aware_shares(10, [1.0, 1.0]) == [5, 5]
    assert sum(link_aware_shares(7, [1.0, 1e-6, 1.0])) == 7
    with pytest.raises(ValueError):
        link_aware_shares(1, [1.0, 1.0])
    assert repeats_of(3686, 2999) == [2999, 687] and repeats_of(2999, 2999) == [2999] and repeats_of(5, 2999) == [5]
    assert sum(repeats_of(3 * 2999 + 1, 2999)) == 3 * 2999 + 1


def _worker(rank, world, port, name, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        table, _, _ = load_golden(name)
        first, count = epoch_range(rank, world, table.n_epochs)
        mine = oracle_lib.generate(table.slice(first, count), nthreads=1)
        # the only "communication" of the path: every rank reports how many epochs it produced
        n = torch.tensor([count], dtype=torch.int64)
        dist.all_reduce(n)
        assert int(n) == table.n_epochs
        np.save(os.path.join(tmp, f"part{rank}.npy"), mine)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_rank_bytes(tmp_path):
    name = "circle_int_b8"
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, name, str(tmp_path)), nprocs=2, join=True)
    table, _, _ = load_golden(name)
    whole = oracle_lib.generate(table)
    parts = np.concatenate([np.load(tmp_path / f"part{r}.npy") for r in range(2)])
    assert np.array_equal(parts, whole)
