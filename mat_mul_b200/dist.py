"""Multi-GPU plumbing: one process per GPU, games sharded by index, no
collective on the data path.  NCCL (torch.distributed) is used only to
all-gather generated demonstrations and to reduce episode statistics
(SURVEY.md 8e).  Every function also works on the gloo backend with CPU
tensors, which is how the host-side logic is tested without GPUs.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_total: int, rank: int | None = None, world_size: int | None = None) -> tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of the game indices 0..n_total-1 owned by `rank`."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, extra = divmod(n_total, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_total: int, world_size: int) -> list[int]:
    return [shard_range(n_total, r, world_size)[1] - shard_range(n_total, r, world_size)[0] for r in range(world_size)]


def gather_shards(shard: torch.Tensor, n_total: int, dim: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
    """All-gather per-rank shards (split along `dim` by shard_range) into the full tensor on every rank.

    Equal shard sizes (n_total divisible by the world size -- the bench and the datasets arrange that): ONE
    all_gather_into_tensor per leading index, straight into the caller's (or a freshly allocated) full tensor --
    no staging copy, no torch.cat.  `shard` may already be the rank's slice of `out` (NCCL's in-place all-gather).
    dim > 0 (a step-major tape (R, N, TP) sharded over N) issues one collective per leading index, all in flight
    at once.  Unequal shards take the padded path (one collective, then the pad rows are dropped)."""
    rank, ws = world()
    if ws == 1:
        if out is not None and out.data_ptr() != shard.data_ptr():
            out.copy_(shard)
            return out
        return shard
    sizes = shard_sizes(n_total, ws)
    assert shard.shape[dim] == sizes[rank], (shard.shape, sizes, rank)
    full_shape = list(shard.shape)
    full_shape[dim] = n_total
    lead1 = all(shard.shape[d] == 1 for d in range(dim))  # the shards are contiguous pieces of the full tensor
    if len(set(sizes)) == 1 and (lead1 or dim == 1):
        if out is None:
            out = shard.new_empty(full_shape)
        assert list(out.shape) == full_shape and out.is_contiguous()
        if lead1:
            dist.all_gather_into_tensor(out.view(-1), shard.contiguous().view(-1))
        else:  # dim == 1: row r of the full tensor is the concatenation of row r of every shard
            works = [dist.all_gather_into_tensor(out[r].view(-1), shard[r].contiguous().view(-1), async_op=True)
                     for r in range(shard.shape[0])]
            for w in works:
                w.wait()
        return out
    x = shard.movedim(dim, 0).contiguous()
    pad = max(sizes) - x.shape[0]
    if pad:
        x = torch.cat((x, x.new_zeros((pad, *x.shape[1:]))))
    buf = x.new_empty((ws * max(sizes), *x.shape[1:]))
    dist.all_gather_into_tensor(buf, x)
    parts = [buf[r * max(sizes) : r * max(sizes) + sizes[r]] for r in range(ws)]
    res = torch.cat(parts).movedim(0, dim).contiguous()
    if out is not None:
        out.copy_(res)
        return out
    return res


@dataclass
class EpisodeStats:
    games: int
    solved: int          # games whose head reached the zero tensor
    steps: int           # actions applied over all games
    min_nnz: int         # best (smallest) non-zero count of a final head
    out_of_range: int    # games flagged TG_FLAG_RANGE

    @property
    def mean_steps(self) -> float:
        return self.steps / max(self.games, 1)


def reduce_episode_stats(flags: torch.Tensor, nnz: torch.Tensor, steps: torch.Tensor | None = None) -> EpisodeStats:
    """Whole-job episode statistics from per-rank per-game outputs: two tiny all_reduces (SUM, MIN)."""
    from ._lib import FLAG_RANGE, FLAG_TERMINAL

    n = flags.numel()
    sums = torch.stack([
        torch.tensor(n, device=flags.device, dtype=torch.int64),
        ((flags & FLAG_TERMINAL) != 0).sum().to(torch.int64),
        (steps.sum().to(torch.int64) if steps is not None else torch.tensor(0, device=flags.device, dtype=torch.int64)),
        ((flags & FLAG_RANGE) != 0).sum().to(torch.int64),
    ])
    mn = (nnz.min().to(torch.int64) if n else torch.tensor(2 ** 31 - 1, device=flags.device, dtype=torch.int64)).reshape(1)
    if world()[1] > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    g, s, st, rg = (int(v) for v in sums.tolist())
    return EpisodeStats(g, s, st, int(mn.item()), rg)


def make_synthetic_demos_sharded(n_total: int, max_actions: int, S: int, values, probs, shift: int, seed: int = 0,
                                 device=None, gather: bool = False):
    """Each rank generates demos [lo, hi) of the global index space (Philox keyed by the global demo index, so the
    union is byte-identical to a 1-GPU run).  gather=True all-gathers tape and slab to every rank."""
    from . import env

    rank, ws = world()
    lo, hi = shard_range(n_total, rank, ws)
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    tape, slab, flags = env.make_synthetic_demos(hi - lo, max_actions, S, values, probs, shift, seed=seed, first_demo=lo,
                                                 device=device)
    if gather and ws > 1:
        tape = gather_shards(tape, n_total, dim=1)
        slab = gather_shards(slab, n_total, dim=0)
        flags = gather_shards(flags, n_total, dim=0)
    return tape, slab, flags


def make_synthetic_demos_gathered(n_total: int, max_actions: int, S: int, values, probs, shift: int, seed: int = 0, device=None):
    """Every rank ends up with ALL n_total demos: each generates its contiguous slice [lo, hi) straight into its
    rows of the full slab / columns of the full step-major tape, then the slices are all-gathered in place (the
    shard handed to NCCL is a view of the output).  Requires n_total divisible by the world size."""
    from . import env

    rank, ws = world()
    if n_total % ws:
        raise ValueError("make_synthetic_demos_gathered needs n_total divisible by the world size")
    lo, hi = shard_range(n_total, rank, ws)
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    lay = env.layout(S)
    tape = torch.empty((max_actions, n_total, lay.token_pitch), dtype=torch.uint8, device=device)
    slab = torch.empty((n_total, lay.game_pitch), dtype=torch.int8, device=device)
    flags = torch.empty(n_total, dtype=torch.uint8, device=device)
    _, _, f = env.make_synthetic_demos(hi - lo, max_actions, S, values, probs, shift, seed=seed, first_demo=lo, device=device,
                                       tape=tape[:, lo:hi], slab=slab[lo:hi])
    flags[lo:hi] = f
    if ws > 1:
        gather_shards(slab[lo:hi], n_total, dim=0, out=slab)
        gather_shards(flags[lo:hi], n_total, dim=0, out=flags)
        gather_shards(tape[:, lo:hi], n_total, dim=1, out=tape)
    return tape, slab, flags
