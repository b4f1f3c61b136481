"""BASELINE.json configs 3, 4 and 5 as parity cases (small sizes; configs[1] is the bench workload, configs[0] is
covered by the Strassen / same-seed golden tests)."""
import numpy as np
import pytest
import torch

from mat_mul_b200 import dist as tgd
from oracle import tg_oracle as orc
from tests.helpers import dense_to_slab, slab_to_dense, tape3_to_tokens, tokens_to_tape3

pytestmark = pytest.mark.gpu

V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
V3, P3 = (-1, 0, 1), (0.15, 0.7, 0.15)


@pytest.fixture(scope="module")
def env():
    from mat_mul_b200 import env as e

    assert torch.cuda.is_available()
    return e


def test_config3_variable_rank_demos_with_change_of_basis(env):
    """4x4 matmul size (16x16x16): synthetic demos of rank <= R (rank drawn per demo, the tail of the action list padded
    with null actions) followed by the change-of-basis augmentation, one (A, B, C) triple per demo."""
    S, R, N, shift = 16, 49, 96, 2
    rng = np.random.default_rng(3)
    tok, _, _ = orc.demos_philox(33, 0, N, V5, P5, R, S, shift)
    rank = rng.integers(1, R + 1, N)
    for n in range(N):
        tok[n, rank[n]:] = shift  # null actions: every coefficient zero
    fac = tok.reshape(N, R, 3, S) - shift
    tgt = np.stack([sum(orc.uvw_to_tensor(*fac[n, r]) for r in range(rank[n])) for n in range(N)])
    tape = torch.from_numpy(tokens_to_tape3(tok)).cuda()
    slab, flags = env.accumulate_demos(tape, S, shift)
    slab_tc, flags_tc = env.accumulate_demos_tc(tape, shift)
    assert np.array_equal(slab_to_dense(slab.cpu().numpy(), S), tgt) and not (flags.cpu().numpy() & 4).any()
    assert torch.equal(slab, slab_tc) and torch.equal(flags, flags_tc)
    mats = env.sample_unimodular(N, S, seed=5, p_nonzero=0.03)
    m = mats.cpu().numpy().astype(np.int64)
    out, tape2, f = env.change_of_basis(slab, mats, S, tape=tape, shift=shift, shift_out=100)
    want = np.einsum("nia,njb,nkc,nabc->nijk", m[:, 0], m[:, 1], m[:, 2], tgt)
    ok = np.abs(want.reshape(N, -1)).max(1) <= 127
    assert ok.sum() > N // 2
    assert np.array_equal(slab_to_dense(out.cpu().numpy()[ok], S), want[ok])
    assert np.array_equal((f.cpu().numpy() & 4) != 0, (np.abs(want.reshape(N, -1) + 0.5) > 64).any(1))
    # the transformed factors are a rank <= R decomposition of the transformed tensor; null actions stay null
    fac2 = tape3_to_tokens(tape2.cpu().numpy(), S).reshape(N, R, 3, S) - 100
    for n in (0, N // 3, N - 1):
        assert np.array_equal(sum(orc.uvw_to_tensor(*fac2[n, r]) for r in range(R)), want[n])
        assert not fac2[n, rank[n]:].any()
    # replaying the transformed demo through the transition reaches the zero tensor after rank[n] real steps
    tape3 = torch.from_numpy(tokens_to_tape3(np.ascontiguousarray((fac2 + 4).reshape(N, R, 3 * S)))).cuda()
    small = torch.from_numpy(np.abs(fac2).reshape(N, -1).max(1) <= 4).cuda() & torch.from_numpy(ok).cuda()
    assert int(small.sum()) > N // 4
    res, fl, nnz, steps = env.rollout(out[small].contiguous(), tape3[:, small].contiguous(), S, 4)
    assert not res.any() and (fl & 1).all()
    assert np.array_equal(steps.cpu().numpy() <= rank[small.cpu().numpy()], np.ones(int(small.sum()), bool))


@pytest.mark.parametrize("world", [2, 8])
def test_config4_mixed_sizes_sharded_by_game_index(env, world):
    """Mixed 2x2 / 3x3 / 4x4 batch sharded over `world` ranks by game index: every rank generates and steps ITS slice
    (emulated here one rank after the other on one GPU); the concatenation of the shards -- what the NCCL all-gather of
    fixed-size records returns, tests/test_dist_gloo.py -- is byte-identical to the one-rank result."""
    for S, R, vals, probs, shift, n_total in [(4, 7, V3, P3, 1, 203), (9, 23, V5, P5, 2, 101), (16, 12, V5, P5, 2, 37)]:
        full_tape, full_slab, full_flags = env.make_synthetic_demos(n_total, R, S, vals, probs, shift, seed=11)
        one_step, f1, n1 = env.step_batch(full_slab, full_tape[R - 1], S, shift)
        tapes, slabs, steps_, nnzs = [], [], [], []
        for r in range(world):
            lo, hi = tgd.shard_range(n_total, r, world)
            t, s, _ = env.make_synthetic_demos(hi - lo, R, S, vals, probs, shift, seed=11, first_demo=lo)
            o, f, n = env.step_batch(s, t[R - 1], S, shift)
            tapes.append(t), slabs.append(s), steps_.append(o), nnzs.append(n)
        assert torch.equal(torch.cat(tapes, dim=1), full_tape) and torch.equal(torch.cat(slabs), full_slab)
        assert torch.equal(torch.cat(steps_), one_step) and torch.equal(torch.cat(nnzs), n1)


@pytest.mark.parametrize("B", [1, 1 << 6, 1 << 10, (1 << 12) + 5])
def test_config5_rollout_sweep_fused_equals_per_step_launches(env, B):
    """Rollout-heavy sweep (9x9x9, K = 64 steps, batch size swept): the action tape is the demo replayed in reverse and
    padded with null actions up to K; the fused K-step kernel and K single-step launches agree, every game is solved
    after exactly R applied actions."""
    S, R, K, shift = 9, 23, 64, 2
    tape, slab, _ = env.make_synthetic_demos(B, R, S, V5, P5, shift, seed=B)
    lay = env.layout(S)
    pad = torch.zeros((K - R, B, lay.token_pitch), dtype=torch.uint8, device="cuda")
    pad[:, :, : 3 * S] = shift
    tapeK = torch.cat([tape.flip(0), pad]).contiguous()
    fused, ff, fn, steps = env.rollout(slab, tapeK, S, shift)
    cur = slab.clone()
    for k in range(K):
        cur, f, n = env.step_batch(cur, tapeK[k], S, shift)
    assert torch.equal(fused, cur) and not fused.any()
    assert (ff & 1).all() and not fn.any() and (f & 1).all() and (f & 2).all()  # the last single steps are null actions
    assert int(steps.max()) <= R and int(steps.min()) >= 1
    # the un-frozen replay (tg_replay) applies all K actions and ends in the same place
    rep, rf, rn = env.replay(slab, tapeK, S, shift)
    assert torch.equal(rep, fused)


def test_full_size_16_tensor_core_paths_by_invariants(env):
    """BASELINE.json configs[2] at scale, through size-independent properties (no oracle at this size):
    (1) 2^16 generated 16x16x16 demos of rank 49 (targets summed on the tensor cores) replay to the zero tensor;
    (2) accumulate(tape) rebuilds the generated targets bit for bit;
    (3) change of basis by random signed permutation matrices, then by their transposes, is the identity, keeps the
        multiset of entries of every game, and the sum of squares is a checksum of checksums."""
    S, R, shift, N = 16, 49, 2, 1 << 16
    tape, slab, flags = env.make_synthetic_demos(N, R, S, V5, P5, shift, seed=2024)
    assert not (flags & 8).any()
    inr = (flags & 4) == 0  # demos whose target stays in the int8 guaranteed zone
    assert int(inr.sum()) > N * 0.99
    out, f, nnz, steps = env.rollout(slab, tape.flip(0).contiguous(), S, shift)
    assert not out[inr].any() and bool((f[inr] & 1).all()) and not nnz[inr].any()
    slab2, flags2 = env.accumulate_demos(tape, S, shift)
    assert torch.equal(slab2, slab) and torch.equal(flags2 & 4, flags & 4)
    # (3)
    g = torch.Generator(device="cuda").manual_seed(7)
    perm = torch.rand((N, 3, S), device="cuda", generator=g).argsort(dim=-1)  # one permutation per game and mode
    sign = (torch.randint(0, 2, (N, 3, S), device="cuda", generator=g) * 2 - 1).to(torch.int8)
    mats = torch.zeros((N, 3, S, S), dtype=torch.int8, device="cuda")
    mats.scatter_(3, perm.unsqueeze(-1), sign.unsqueeze(-1))
    t1, f1 = env.change_of_basis(slab, mats, S)
    back, f2 = env.change_of_basis(t1, mats.transpose(2, 3).contiguous(), S)
    assert torch.equal(back, slab)
    assert torch.equal(f1 & 4, flags & 4) and torch.equal(f2 & 4, flags & 4) and not ((f1 | f2) & 0x80).any()
    sq = lambda x: (x.to(torch.int32) ** 2).sum(dim=1)
    assert torch.equal(sq(t1), sq(slab))
    assert torch.equal(t1.to(torch.int16).abs().sort(dim=1).values[:: 257], slab.to(torch.int16).abs().sort(dim=1).values[:: 257])


@pytest.mark.parametrize("S,R,N", [(9, 23, 1 << 18), (4, 7, 1 << 20)])
def test_full_size_change_of_basis_round_trip(env, S, R, N):
    """Change of basis at bench scale (no oracle at this size): signed permutation matrices per game, then their transposes,
    give the original tensors back; the sum of squares of every game is unchanged; range flags follow the entries."""
    vals, probs, shift = (V3, P3, 1) if S == 4 else (V5, P5, 2)
    _, slab, flags = env.make_synthetic_demos(N, R, S, vals, probs, shift, seed=77)
    g = torch.Generator(device="cuda").manual_seed(S)
    perm = torch.rand((N, 3, S), device="cuda", generator=g).argsort(dim=-1)
    sign = (torch.randint(0, 2, (N, 3, S), device="cuda", generator=g) * 2 - 1).to(torch.int8)
    mats = torch.zeros((N, 3, S, S), dtype=torch.int8, device="cuda")
    mats.scatter_(3, perm.unsqueeze(-1), sign.unsqueeze(-1))
    t1, f1 = env.change_of_basis(slab, mats, S)
    back, f2 = env.change_of_basis(t1, mats.transpose(2, 3).contiguous(), S)
    assert torch.equal(back, slab)
    assert torch.equal(f1 & 4, flags & 4) and torch.equal(f2 & 4, flags & 4) and not ((f1 | f2) & 0x80).any()
    sq = lambda x: (x.to(torch.int32) ** 2).sum(dim=1)
    assert torch.equal(sq(t1), sq(slab))
