"""world_size-2 gloo tests (CPU) of the sharding / gather / statistics plumbing used at N > 1."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mat_mul_b200 import dist as tgd


def test_shard_ranges_partition_the_index_space():
    for n in (0, 1, 7, 8, 1000, 1 << 20):
        for ws in (1, 2, 3, 8):
            ranges = [tgd.shard_range(n, r, ws) for r in range(ws)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = tgd.shard_sizes(n, ws)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, ws: int, port: int, n_total: int):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        lo, hi = tgd.shard_range(n_total)
        # fixed-size demo records: tape (R, n, TP) sharded on dim 1, slab (n, GP) on dim 0; record i is filled with i
        R, TP, GP = 3, 32, 768
        idx = torch.arange(lo, hi)
        tape = (idx.view(1, -1, 1) % 251).to(torch.uint8).expand(R, -1, TP).contiguous()
        slab = (idx.view(-1, 1) % 100).to(torch.int8).expand(-1, GP).contiguous()
        full_tape = tgd.gather_shards(tape, n_total, dim=1)
        full_slab = tgd.gather_shards(slab, n_total, dim=0)
        want = torch.arange(n_total)
        assert full_tape.shape == (R, n_total, TP) and torch.equal(full_tape[1, :, 5], (want % 251).to(torch.uint8))
        assert full_slab.shape == (n_total, GP) and torch.equal(full_slab[:, 7], (want % 100).to(torch.int8))
        if n_total % ws == 0:  # equal shards: gathered in place, the shard being the rank's slice of the output
            out_slab = torch.zeros((n_total, GP), dtype=torch.int8)
            out_slab[lo:hi] = slab
            assert tgd.gather_shards(out_slab[lo:hi], n_total, dim=0, out=out_slab) is out_slab and torch.equal(out_slab, full_slab)
            out_tape = torch.zeros((R, n_total, TP), dtype=torch.uint8)
            out_tape[:, lo:hi] = tape
            assert tgd.gather_shards(out_tape[:, lo:hi], n_total, dim=1, out=out_tape) is out_tape and torch.equal(out_tape, full_tape)
        # statistics: game i is solved iff i % 3 == 0, nnz = i + 5, steps = i % 7, range flag iff i % 10 == 0
        flags = ((idx % 3 == 0).to(torch.uint8) * 1) | ((idx % 10 == 0).to(torch.uint8) * 4)
        st = tgd.reduce_episode_stats(flags, (idx + 5).to(torch.int32), (idx % 7).to(torch.int32))
        assert st.games == n_total and st.solved == int((want % 3 == 0).sum()) and st.steps == int((want % 7).sum())
        assert st.min_nnz == 5 and st.out_of_range == int((want % 10 == 0).sum())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [10, 11])
def test_gather_and_stats_world_size_2(n_total):
    mp.spawn(_worker, args=(2, _free_port(), n_total), nprocs=2, join=True)
