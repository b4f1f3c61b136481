"""Larger randomized differential run of the GPU kernels against the oracle (not part of the test suite)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from mat_mul_b200 import env
from oracle import tg_oracle as orc
from tests.helpers import dense_to_slab, slab_to_dense, tape3_to_tokens, tokens_to_tape3, tokens_to_tape
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
V3, P3 = (-1, 0, 1), (0.15, 0.7, 0.15)
t0 = time.time()
for S, R, vals, probs, shift, N in [(4, 7, V3, P3, 1, 200000), (9, 23, V5, P5, 2, 60000), (16, 49, V5, P5, 2, 4000), (9, 40, V5, (0.2,) * 5, 2, 5000),
                                    (4, 7, V5, (0.3, 0.05, 0.3, 0.05, 0.3), 2, 50000)]:
    for seed in (1, 0xABCDEF0123):
        tape, slab, flags = env.make_synthetic_demos(N, R, S, vals, probs, shift, seed=seed, first_demo=77)
        tok, tgt, ex = orc.demos_philox(seed, 77, N, vals, probs, R, S, shift)
        assert ex == 0 and np.array_equal(tape3_to_tokens(tape.cpu().numpy(), S), tok), (S, R, seed, "tokens")
        ok = np.abs(tgt.reshape(N, -1)).max(1) <= 127
        assert np.array_equal(slab_to_dense(slab.cpu().numpy()[ok], S), tgt[ok]), (S, R, seed, "targets")
        f = flags.cpu().numpy()
        assert np.array_equal((f & 4) != 0, (np.abs(tgt.reshape(N, -1) + 0.5) > 64).any(1)), (S, R, seed, "flags")
        # replay through the step kernel, the fused rollout and K8
        okt = torch.from_numpy(ok).cuda()
        cur = slab[okt].contiguous()
        tp = tape[:, okt].contiguous()
        res, fl, nnz, steps = env.rollout(cur, tp.flip(0).contiguous(), S, shift)
        assert not res.any() and bool((fl & 1).all()), (S, R, seed, "rollout")
        k = min(R, 6)
        kids, kf, kn, kk = env.expand_children(cur, tp[:k].permute(1, 0, 2).contiguous(), S, shift)
        for c in range(k):
            w, wf, wn = env.step_batch(cur, tp[c], S, shift)
            assert torch.equal(kids[:, c], w) and torch.equal(kf[:, c], wf) and torch.equal(kn[:, c], wn), (S, R, seed, "expand", c)
            assert torch.equal(kk[:, c], env.state_keys(w, S)), (S, R, seed, "keys", c)
    print(f"S={S} R={R} N={N} ok  ({time.time() - t0:.1f} s)", flush=True)
for S, N, p in [(4, 20000, 0.3), (9, 4000, 0.08), (9, 3000, 0.3), (16, 600, 0.03), (16, 300, 0.15)]:
    rng = np.random.default_rng(S)
    T = rng.integers(-25, 26, (N, S, S, S)) * (rng.random((N, S, S, S)) < 0.4)
    mats = env.sample_unimodular(N, S, seed=9, p_nonzero=p)
    m = mats.cpu().numpy().astype(np.int64)
    assert np.array_equal(m, orc.sample_unimodular(9, 0, N, S, p))
    out, flags = env.change_of_basis(torch.from_numpy(dense_to_slab(T)).cuda(), mats, S)
    want = np.einsum("nia,njb,nkc,nabc->nijk", m[:, 0], m[:, 1], m[:, 2], T)
    ok = np.abs(want.reshape(N, -1)).max(1) <= 127
    assert np.array_equal(slab_to_dense(out.cpu().numpy()[ok], S), want[ok]), (S, "basis")
    assert np.array_equal((flags.cpu().numpy() & 4) != 0, (np.abs(want.reshape(N, -1) + 0.5) > 64).any(1)), (S, "basis flags")
    print(f"basis S={S} N={N} p={p}: ok, {ok.mean():.2f} in int8  ({time.time() - t0:.1f} s)", flush=True)
print("soak ok")
