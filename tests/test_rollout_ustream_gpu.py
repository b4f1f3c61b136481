"""GPU parity of K2 (tg_rollout) and of same-seed demo generation (tg_demo_from_ustream)."""
import numpy as np
import pytest
import torch

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


@pytest.mark.parametrize("S,R,values,probs,shift,N,extra", [
    (4, 7, V3, P3, 1, 1000, 3), (9, 23, V5, P5, 2, 700, 5), (16, 12, V5, P5, 2, 37, 2), (9, 4, V5, P5, 2, 13, 0),
])
def test_rollout_matches_oracle(env, S, R, values, probs, shift, N, extra):
    rng = np.random.default_rng(S + R)
    tok, tgt, _ = orc.demos_seeded(2, values, probs, R, S, shift, N)
    # replay in reverse; a third of the games get a wrong action in the middle, junk actions follow the end
    tape_tok = np.concatenate([tok[:, ::-1], rng.integers(0, 2 * shift + 1, (N, extra, 3 * S))], axis=1)
    wrong = rng.random(N) < 0.33
    tape_tok[wrong, R // 2] = rng.integers(0, 2 * shift + 1, (int(wrong.sum()), 3 * S))
    tgt[5] = 0  # a game that starts solved
    want, wflags, wnnz, wsteps = orc.rollout_batch(tgt, tape_tok, shift)
    ok = np.abs(want.reshape(N, -1)).max(1) <= 63
    slab = torch.from_numpy(dense_to_slab(tgt)).cuda()
    tape = torch.from_numpy(tokens_to_tape3(tape_tok)).cuda()
    out, flags, nnz, steps = env.rollout(slab, tape, S, shift)
    torch.cuda.synchronize()
    f = flags.cpu().numpy()
    assert np.array_equal(slab_to_dense(out.cpu().numpy(), S)[ok], want[ok])
    assert np.array_equal(f[ok] & 1, wflags[ok] & 1) and np.array_equal(nnz.cpu().numpy()[ok], wnnz[ok])
    assert np.array_equal(steps.cpu().numpy()[ok], wsteps[ok])
    assert wsteps[5] == 0 and (wsteps[~wrong] <= R).all() and (f[~ok] & 4).all()
    # in place, and equal to K single steps for games that never terminate early
    out2, _, _, _ = env.rollout(slab.clone(), tape, S, shift, out=None)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("S", [4, 9, 16])
def test_rollout_and_replay_ragged_batches(env, S):
    """Batches that do not fill a warp / a CTA tile of the row-owner kernels (three or two games per warp, 24 or 16 per CTA)
    and step counts around the ring depth: tg_rollout against the oracle, tg_replay against K single tg_step launches."""
    shift, R = 2, 6
    rng = np.random.default_rng(300 + S)
    for B in (1, 2, 3, 4, 23, 24, 25, 49):
        for K in (1, 3, 4, 5, 9):
            tok, tgt, _ = orc.demos_seeded(7 + B + K, V5, P5, R, S, shift, B)
            tape_tok = np.concatenate([tok[:, ::-1], rng.integers(0, 2 * shift + 1, (B, 4, 3 * S))], axis=1)[:, :K]
            want, wflags, wnnz, wsteps = orc.rollout_batch(tgt, tape_tok, shift)
            ok = np.abs(want.reshape(B, -1)).max(1) <= 63
            slab = torch.from_numpy(dense_to_slab(tgt)).cuda()
            tape = torch.from_numpy(tokens_to_tape3(tape_tok)).cuda()
            out, flags, nnz, steps = env.rollout(slab, tape, S, shift)
            assert np.array_equal(slab_to_dense(out.cpu().numpy(), S)[ok], want[ok]), (B, K)
            assert np.array_equal(steps.cpu().numpy()[ok], wsteps[ok]) and np.array_equal(nnz.cpu().numpy()[ok], wnnz[ok]), (B, K)
            assert np.array_equal(flags.cpu().numpy()[ok] & 1, wflags[ok] & 1), (B, K)
            rep, rflags, rnnz = env.replay(slab, tape, S, shift)
            cur = slab
            for t in range(K):
                cur, sf, sn = env.step_batch(cur, tape[t], S, shift)
            assert torch.equal(rep, cur) and torch.equal(rnnz, sn), (B, K)
            assert torch.equal(rflags & 1, sf & 1), (B, K)


def test_rollout_zero_steps_and_empty(env):
    S, shift = 9, 2
    lay = env.layout(S)
    slab = torch.randint(-2, 3, (50, lay.game_pitch), dtype=torch.int8, device="cuda")
    env.slab_view(slab, S)  # arbitrary padding is allowed on input; it is masked out
    tape = torch.zeros((0, 50, lay.token_pitch), dtype=torch.uint8, device="cuda")
    out, flags, nnz, steps = env.rollout(slab, tape, S, shift)
    assert not steps.any()
    assert torch.equal(env.slab_view(out, S), env.slab_view(slab, S))
    assert torch.equal(nnz.long(), (env.slab_view(slab, S) != 0).flatten(1).sum(1))


@pytest.mark.parametrize("name", ["S4", "S9", "S16", "S4u"])
def test_same_seed_demos_match_reference(env, golden, name):
    """utils.py:203-233 after torch.manual_seed(s): identical tokens, targets and stream position."""
    g = golden["demos"]
    R, S, shift, n, _ = (int(x) for x in g[f"{name}_cfg"])
    for seed in (0, 1, 7):
        tape, slab, flags, consumed = env.demos_from_seed(n, R, S, g[f"{name}_values"], g[f"{name}_probs"], shift, seed=seed)
        assert np.array_equal(tape3_to_tokens(tape.cpu().numpy(), S), g[f"{name}_seed{seed}_tokens"])
        assert np.array_equal(slab_to_dense(slab.cpu().numpy(), S), g[f"{name}_seed{seed}_targets"])
        tail = env.torch_cpu_stream(2, seed=seed, skip=consumed)
        assert np.array_equal(tail, g[f"{name}_seed{seed}_tail"])


def test_same_seed_demos_continue_the_global_generator(env, golden):
    g = golden["demos"]
    torch.manual_seed(5)
    tape, slab, _, _ = env.demos_from_seed(5, 7, 4)  # SyntheticDemoDataset defaults (datasets.py:30-32)
    assert np.array_equal(tape3_to_tokens(tape.cpu().numpy(), 4), g["dataset_seed5_tokens"])
    assert np.array_equal(slab_to_dense(slab.cpu().numpy(), 4), g["dataset_seed5_targets"])
    after = torch.rand(2, dtype=torch.float64).numpy()  # the global generator stands where the reference left it
    _, _, used = orc.demos_seeded(5, (-1, 0, 1), (0.15, 0.7, 0.15), 7, 4, 1, 5)
    assert np.array_equal(after, orc.mt_doubles(5, used + 2)[used:])


def test_ustream_large_vs_oracle(env):
    S, R, N, shift = 9, 23, 3000, 2
    tok, tgt, used = orc.demos_seeded(11, V5, P5, R, S, shift, N)
    tape, slab, flags, consumed = env.demos_from_seed(N, R, S, V5, P5, shift, seed=11)
    assert consumed == used
    assert np.array_equal(tape3_to_tokens(tape.cpu().numpy(), S), tok)
    assert np.array_equal(slab_to_dense(slab.cpu().numpy(), S), tgt)


def test_ustream_short_stream_reports_progress(env):
    u = orc.mt_doubles(3, 12 * 40)  # far too short for 100 demos
    tape, slab, flags, done, consumed = env.demos_from_ustream(u, 100, 7, 4, V3, P3, 1)
    tok, tgt, used = orc.demos_from_ustream(u, V3, P3, 7, 4, 1, 100)
    assert done == len(tok) and consumed == used and done < 100
    assert np.array_equal(tape3_to_tokens(tape.cpu().numpy(), 4)[:done], tok)
    assert np.array_equal(slab_to_dense(slab.cpu().numpy()[:done], 4), tgt)
