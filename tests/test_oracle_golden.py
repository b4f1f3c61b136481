"""Pins the CPU oracle (oracle/tg_oracle.c) to outputs of the live reference
(tests/golden/*.npz, produced by oracle/gen_golden.py from /root/reference)."""
import numpy as np
import pytest

from oracle import tg_oracle as orc


def test_mt19937_doubles_match_torch_rand(golden):
    g = golden["rng"]
    for seed, want in zip(g["seeds"], g["rand"]):
        got = orc.mt_doubles(int(seed), len(want))
        assert np.array_equal(got, want)  # bit-exact float64


def test_categorical_matches_factor_sample(golden):
    # utils.py:197-200 through torch.multinomial, for 6 prob vectors x 5 seeds x 4 lengths
    g = golden["rng"]
    for ip, (probs, n) in enumerate(zip(g["prob_sets"], g["prob_lens"])):
        probs = probs[:n].astype(np.float32)
        cdf = orc.categorical_cdf(probs)
        values = np.arange(n) - n // 2
        for seed in g["seeds"]:
            for length in (4, 9, 16, 64):
                u = orc.mt_doubles(int(seed), length)
                got = values[[orc.categorical_pick(cdf, x) for x in u]]
                assert np.array_equal(got, g[f"p{ip}_s{seed}_n{length}"]), (ip, seed, length)


@pytest.mark.parametrize("name", ["S4", "S9", "S16", "S4u"])
def test_synthetic_demos_same_seed(golden, name):
    # utils.py:203-233: identical tokens, targets AND stream position for the same seed
    g = golden["demos"]
    R, S, shift, n, _ = g[f"{name}_cfg"]
    for seed in (0, 1, 7):
        tok, tgt, used = orc.demos_seeded(seed, g[f"{name}_values"], g[f"{name}_probs"], int(R), int(S), int(shift), int(n))
        assert np.array_equal(tok, g[f"{name}_seed{seed}_tokens"])
        assert np.array_equal(tgt, g[f"{name}_seed{seed}_targets"])
        tail = orc.mt_doubles(seed, used + 2)[used:]
        assert np.array_equal(tail, g[f"{name}_seed{seed}_tail"])


def test_dataset_class_uses_same_loop(golden):
    # datasets.py:124-142 with the class defaults values=(-1,0,1), probs=(.15,.7,.15), shift=1
    g = golden["demos"]
    tok, tgt, _ = orc.demos_seeded(5, (-1, 0, 1), (0.15, 0.7, 0.15), 7, 4, 1, 5)
    assert np.array_equal(tok, g["dataset_seed5_tokens"])
    assert np.array_equal(tgt, g["dataset_seed5_targets"])


@pytest.mark.parametrize("S", [4, 9, 16])
def test_transition_matches_get_child_states(golden, S):
    g = golden["steps"]
    states, actions = g[f"S{S}_states"], g[f"S{S}_actions"]
    n, k = actions.shape[:2]
    for b in range(n):
        head = np.repeat(states[b, 0][None], k, axis=0)
        out, flags, nnz = orc.step_batch(head, actions[b], shift=1)  # action_to_tensor: shift fixed at 1 (Q1)
        assert np.array_equal(out, g[f"S{S}_child_heads"][b])
        assert np.array_equal((flags & orc.FLAG_NULL) == 0, g[f"S{S}_not_null"][b])
        assert np.array_equal((flags & orc.FLAG_TERMINAL) != 0, g[f"S{S}_terminal"][b])
        assert np.array_equal(nnz, (out != 0).reshape(k, -1).sum(1))
    assert g[f"S{S}_terminal"][0, 1] and not g[f"S{S}_not_null"][:, 0].any()


@pytest.mark.parametrize("S", [4, 9, 16])
def test_take_action_arithmetic(golden, S):
    # training.py:253-267: tokens - 2, head update, history shift, nnz, min over n_samples
    g = golden["steps"]
    sb, aa = g[f"S{S}_ta_states"], g[f"S{S}_ta_tokens"][:, 0]
    out, flags, nnz = orc.step_batch(sb[:, 0], aa, shift=2)
    new = g[f"S{S}_ta_new_states"]
    assert np.array_equal(out, new[:, 0])
    assert np.array_equal(new[:, 1:], sb[:, :-1])
    grouped = nnz.reshape(-1, 4)
    assert np.array_equal(grouped.min(1), g[f"S{S}_ta_best_values"])
    assert np.array_equal(grouped.argmin(1), g[f"S{S}_ta_best_indices"])
    assert flags[3] & orc.FLAG_TERMINAL


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_demo_getitem(golden, tag):
    g = golden["getitem"]
    R, dim_t, S, shift, n_demos = g[f"{tag}_cfg"]
    assert g[f"{tag}_len"] == n_demos * R
    for idx in range(n_demos * R):
        d, a = divmod(idx, R)
        state, scalar, action, reward = orc.demo_getitem(g[f"{tag}_tokens"][d], g[f"{tag}_targets"][d], int(dim_t), a, replay_shift=1)
        assert np.array_equal(state, g[f"{tag}_states"][idx]), (tag, idx)
        assert scalar == g[f"{tag}_scalars"][idx, 0] and reward == g[f"{tag}_rewards"][idx, 0]
        assert np.array_equal(action, g[f"{tag}_actions"][idx])


def test_getitem_replay_reaches_zero_only_with_right_shift(golden):
    # with shift=2 demos the reference's replay (shift 1) is NOT the true residual
    g = golden["getitem"]
    tok, tgt = g["c_tokens"][0], g["c_targets"][0]
    assert not orc.take_actions(tok, tgt, shift=2).any()
    assert orc.take_actions(tok, tgt, shift=1).any()


def test_strassen(golden):
    g = golden["strassen"]
    uu, vv, ww = orc.strassen_factors()
    assert np.array_equal(uu, g["uu"]) and np.array_equal(vv, g["vv"]) and np.array_equal(ww, g["ww"])
    tensor, actions = orc.uvw_to_demo(uu, vv, ww, shift=1)
    assert np.array_equal(tensor, g["tensor"]) and np.array_equal(actions, g["action_list"])
    assert np.array_equal(tensor, orc.build_matmul_tensor(2))  # notebook: strassen == 2x2 matmul tensor
    states, acts, rewards, bits = orc.strassen_dataset()
    assert g["n"] == 448
    assert np.array_equal(states, g["states"][:, 0]) and np.array_equal(acts, g["actions"])
    assert np.array_equal(rewards, g["rewards"][:, 0]) and np.array_equal(bits, g["bits"])
    assert list(rewards[-10:]) == [-3, -2, -2, -2, -2, -1, -2, -2, -1, -1]  # notebooks/strassen_example.ipynb:344
    assert list(acts[0]) == [3, 2, 2, 3, 3, 2, 2, 3, 3, 2, 2, 3] and rewards[0] == -7
    assert np.array_equal(states[447], orc.action_to_tensor(acts[447], shift=2))


def test_build_matmul_tensor(golden):
    g = golden["matmul"]
    for n in (2, 3, 4):
        t = orc.build_matmul_tensor(n)
        assert np.array_equal(t, g[f"n{n}"][0]) and not g[f"n{n}"][1].any()
        assert t.sum() == n ** 3
        rng = np.random.default_rng(n)
        A, B = rng.integers(-3, 4, (n, n)), rng.integers(-3, 4, (n, n))
        assert np.array_equal(np.einsum("abc,a,b->c", t, A.ravel(), B.ravel()), (A @ B).ravel())


@pytest.mark.parametrize("S", [4, 9, 16])
def test_slice_rank_matches_get_rank(golden, S):
    g = golden["ranks"]
    assert np.array_equal(orc.slice_rank_batch(g[f"S{S}_T"]), g[f"S{S}_rank"])


def test_slice_rank_matmul(golden):
    got = [orc.slice_rank_batch(orc.build_matmul_tensor(n)[None])[0] for n in (2, 3, 4)]
    assert np.array_equal(got, golden["ranks"]["mm_rank"])


def test_rollout_freezes_at_terminal():
    tok, tgt, _ = orc.demos_seeded(3, (-2, -1, 0, 1, 2), (0.05, 0.1, 0.7, 0.1, 0.05), 5, 4, 2, 16)
    tape = np.concatenate([tok[:, ::-1], np.full((16, 3, 12), 4, np.int32)], axis=1)  # junk after the end
    out, flags, nnz, steps = orc.rollout_batch(tgt, tape, shift=2)
    assert not out.any() and (flags & orc.FLAG_TERMINAL).all() and (steps == 5).all() and not nnz.any()


def test_philox_known_answers():
    # Random123 v1.14 kat_vectors, philox4x32 10 rounds
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        assert list(orc.philox4x32_10(ctr, key)) == want


def test_philox_demos_invariants():
    vals, probs = (-2, -1, 0, 1, 2), (0.05, 0.1, 0.7, 0.1, 0.05)
    tok, tgt, ex = orc.demos_philox(0x5EED, 0, 64, vals, probs, 23, 9, 2)
    assert ex == 0 and tok.min() >= 0 and tok.max() <= 4
    assert not orc.take_actions(tok[5], tgt[5], shift=2).any()  # replaying a demo reaches zero
    tok2, tgt2, _ = orc.demos_philox(0x5EED, 40, 8, vals, probs, 23, 9, 2)  # partition independence
    assert np.array_equal(tok2, tok[40:48]) and np.array_equal(tgt2, tgt[40:48])
    # no accepted term has an all-zero factor (utils.py:229)
    f = tok.reshape(64, 23, 3, 9) - 2
    assert (f != 0).any(-1).all()
    # degenerate distribution: bounded retries instead of the reference's endless loop (Q11)
    _, _, ex = orc.demos_philox(1, 0, 2, (-1, 0, 1), (0.0, 1.0, 0.0), 3, 4, 1, max_tries=4)
    assert ex == 6


def test_change_of_basis_invariants():
    rng = np.random.default_rng(0)
    S, R = 4, 5
    f = rng.integers(-2, 3, (R, 3, S))
    T = sum(orc.uvw_to_tensor(*f[r]) for r in range(R))
    A, B, Cm = (np.triu(rng.integers(-1, 2, (S, S)), 1) + np.eye(S, dtype=np.int64) for _ in range(3))
    Tp = orc.change_of_basis(T, A, B, Cm)
    assert np.array_equal(Tp, np.einsum("ia,jb,kc,abc->ijk", A, B, Cm, T))
    fp = orc.change_of_basis_factors(f, A, B, Cm)
    assert np.array_equal(Tp, sum(orc.uvw_to_tensor(*fp[r]) for r in range(R)))  # sum (Au)(Bv)(Cw) = T'
    inv = [np.rint(np.linalg.inv(M)).astype(np.int64) for M in (A, B, Cm)]
    assert np.array_equal(orc.change_of_basis(Tp, *inv), T)  # unimodular => integer inverse


def test_state_key():
    rng = np.random.default_rng(1)
    T = rng.integers(-2, 3, (50, 4, 4, 4)) * (rng.random((50, 4, 4, 4)) < 0.2)
    T[7] = T[3]
    T[9] = 0
    k = orc.state_key_batch(T)
    assert k[7] == k[3] and k[9] == 0
    assert len(set(k.tolist())) == len({t.tobytes() for t in T.astype(np.int32)})


def test_misc_reference_helpers(golden):
    g = golden["misc"]
    assert g["scalars_b"].shape == (5, 1) and (g["scalars_b"] == 3).all() and g["scalars_s"].tolist() == [3.0]
    assert str(g["str_key"]) == "1_-2_0_3"


def test_alias_sampler_contract_v2():
    """Throughput-mode contract v2 (group alias tables): the library's host-side table builder and the oracle's
    restatement agree bit for bit; the tables encode the reference's distribution (i.i.d. entries, factor conditioned on
    being non-zero, utils.py:222-232); the sampled demos follow it."""
    import ctypes as C
    import itertools

    from mat_mul_b200 import _lib

    L = _lib.lib()
    cases = [((-1, 0, 1), (0.15, 0.7, 0.15)), ((-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)), ((-2, -1, 0, 1, 2), (0.2,) * 5),
             ((-1, 0, 1), (0.1, 0.8, 0.1)), ((1, 2), (0.3, 0.7)), ((-2, -1, 0, 1, 2), (1.0, 2.0, 3.0, 2.0, 1.0))]
    for values, probs in cases:
        for S in (4, 9, 16):
            assert orc.alias_applies(values, probs, S)
            want = orc.alias_tables(values, probs, S)
            got = np.zeros((8, 128), dtype=np.uint16)
            v = np.array(values, dtype=np.int8)
            p = np.array(probs, dtype=np.float64)
            assert L.tg_demo_alias_tables(v.ctypes.data, p.ctypes.data, len(v), S, got.ctypes.data) == 0
            assert np.array_equal(got, want), (values, probs, S)
            # the distribution a table encodes: P(outcome) = (sum over buckets of thr / 512 for own + (1 - thr / 512) for alias) / 128
            n, pn = len(values), np.array(probs, dtype=np.float64) / np.sum(probs)
            def table_dist(t):
                d = np.zeros(128)
                for o in range(128):
                    thr, al = int(t[o]) & 511, int(t[o]) >> 9
                    q = 1.0 if (thr == 511 and al == o) else thr / 512.0
                    d[o] += q / 128
                    d[al] += (1 - q) / 128
                return d
            p3 = np.array([pn[o % n] * pn[(o // n) % n] * pn[o // (n * n)] for o in range(n ** 3)])
            assert np.abs(table_dist(want[0])[: n ** 3] - p3).max() < 2.0 ** -13  # 9-bit thresholds per bucket: outcome probabilities to ~2^-13 absolute
            assert np.abs(table_dist(want[1])[:n] - pn).max() < 2.0 ** -13
            if 0 in values:  # the last group's tilted table never yields the zero outcome: a factor cannot end all zero
                ng, z = (S + 2) // 3, values.index(0)
                last_zero = z if S - 3 * (ng - 1) == 1 else z * (1 + n + n * n)
                assert table_dist(want[2 + ng - 1])[last_zero] == 0.0
    assert not orc.alias_applies((-1, 0, 1), (0.0, 1.0, 0.0), 4) and not orc.alias_applies(tuple(range(-3, 3)), (1,) * 6, 9)
    # sampled factors: never all zero, entry frequencies = the conditional distribution of the reference's accepted factors
    values, probs, S, R, N = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05), 4, 7, 6000
    tok, tgt, ex = orc.demos_philox(1, 0, N, values, probs, R, S, 2)
    fac = tok.reshape(N * R * 3, S) - 2
    assert ex == 0 and fac.any(axis=1).all()
    pn = np.array(probs)
    exact = {}
    for combo in itertools.product(range(5), repeat=S):
        if any(values[c] != 0 for c in combo):
            exact[combo] = np.prod(pn[list(combo)])
    zsum = sum(exact.values())
    p_entry0_zero = sum(v for c, v in exact.items() if values[c[0]] == 0) / zsum  # P(first entry = 0 | factor non-zero)
    assert abs((fac[:, 0] == 0).mean() - p_entry0_zero) < 0.006 and abs((fac[:, 3] == 0).mean() - p_entry0_zero) < 0.006
    assert abs((fac == 2).mean() - (fac == -2).mean()) < 0.004
    assert np.array_equal(tgt[5], sum(orc.uvw_to_tensor(*(tok[5, r].reshape(3, S) - 2)) for r in range(R)))
