"""The int16 residual format and the demo-major store (round 2): change of basis at SURVEY 8(d)'s configuration (one
unimodular triple per game, off-diagonal density 0.3, int8 in / int16 out) with EVERY game compared with the int64 einsum;
targets beyond the int8 zone accumulated as int16; the TMA batcher (tg_demo_sample_dm) on int8 and int16 targets."""
import numpy as np
import pytest
import torch

from oracle import tg_oracle as orc
from tests.helpers import dense_to_slab, geo, slab_to_dense, tokens_to_tape3

pytestmark = pytest.mark.gpu

V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
U5 = (0.2, 0.2, 0.2, 0.2, 0.2)


@pytest.fixture(scope="module")
def env():
    from mat_mul_b200 import env as e

    return e


def slab16_to_dense(slab16: np.ndarray, S: int) -> np.ndarray:
    """int16 (B, GP) -> int32 (B,S,S,S); asserts padding elements are zero."""
    rp, gp, _ = geo(S)
    B = slab16.shape[0]
    assert slab16.shape[1] == gp and slab16.dtype == np.int16
    rows = slab16[:, : S * rp].reshape(B, S, rp)
    assert not rows[:, :, S * S:].any() and not slab16[:, S * rp:].any(), "slab16 padding must stay zero"
    return rows[:, :, : S * S].reshape(B, S, S, S).astype(np.int32)


@pytest.mark.parametrize("S,R,N", [(4, 7, 300), (9, 23, 301), (16, 49, 131)])
def test_change_of_basis_config3_every_game_int16(env, S, R, N):
    shift = 2
    tok, tgt, _ = orc.demos_philox(5, 0, N, V5, P5, R, S, shift)
    mats = env.sample_unimodular(N, S, seed=9, first=3, p_nonzero=0.3)
    m = mats.cpu().numpy().astype(np.int64)
    assert np.array_equal(m, orc.sample_unimodular(9, 3, N, S, 0.3))
    want = np.einsum("nia,njb,nkc,nabc->nijk", m[:, 0], m[:, 1], m[:, 2], tgt.astype(np.int64))
    assert np.abs(want).max() > 127 or S == 4  # the configuration really leaves int8
    slab = torch.from_numpy(dense_to_slab(tgt)).cuda()
    out16, flags, stats = env.change_of_basis(slab, mats, S, out_dtype=torch.int16, return_path_stats=True)
    assert out16.dtype == torch.int16
    f = flags.cpu().numpy()
    assert not (f & 0x80).any() and not (f & 4).any()  # nothing left marked, everything fits int16
    assert np.array_equal(slab16_to_dense(out16.cpu().numpy(), S), want)  # EVERY game
    assert stats["games"] == N and stats["fast_path"] + stats["byte_plane_path"] + stats["exact_int32_redo"] == N
    assert stats["fast_path"] >= 0.9 * N, stats  # the tensor-core / packed fast path holds this distribution
    # the int8 entry point on the same games: low byte always, range flag exactly for the games that left [-64, 63]
    out8, f8 = env.change_of_basis(slab, mats, S)
    f8 = f8.cpu().numpy()
    assert not (f8 & 0x80).any()
    assert np.array_equal((f8 & 4) != 0, ((want < -64) | (want > 63)).reshape(N, -1).any(1))
    assert np.array_equal(slab_to_dense(out8.cpu().numpy(), S).astype(np.int8), want.astype(np.int8))
    # one shared triple
    o1, f1 = env.change_of_basis(slab, mats[7], S, out_dtype=torch.int16)
    want1 = np.einsum("ia,jb,kc,nabc->nijk", m[7, 0], m[7, 1], m[7, 2], tgt.astype(np.int64))
    assert np.array_equal(slab16_to_dense(o1.cpu().numpy(), S), want1) and not (f1.cpu().numpy() & 0x84).any()


@pytest.mark.parametrize("S", [4, 9, 16])
def test_change_of_basis_int16_extreme_inputs(env, S):
    # int8 extremes and large matrices: whatever path a game takes (f16 tensor cores with checked operands, byte planes,
    # exact int32), the int16 result is the int64 einsum, and the flag says exactly which games do not fit int16
    N = 150
    rng = np.random.default_rng(100 + S)
    T = rng.integers(-128, 128, (N, S, S, S)) * (rng.random((N, S, S, S)) < 0.5)
    T[0], T[1], T[2] = -128, 127, 0
    m = rng.integers(-3, 4, (N, 3, S, S)) * (rng.random((N, 3, S, S)) < 0.35)
    m[::4] = orc.sample_unimodular(1, 0, N, S, 0.3)[::4]
    m[3] = np.eye(S, dtype=np.int64)
    m[5, 0] = 127  # far beyond every fast path
    m[6, 2] = -128
    want = np.einsum("nia,njb,nkc,nabc->nijk", m[:, 0], m[:, 1], m[:, 2], T.astype(np.int64))
    slab = torch.from_numpy(dense_to_slab(T)).cuda()
    out16, flags, stats = env.change_of_basis(slab, torch.from_numpy(m.astype(np.int8)).cuda(), S, out_dtype=torch.int16,
                                              return_path_stats=True)
    f = flags.cpu().numpy()
    fits = (np.abs(want.reshape(N, -1) + 0.5) < 32768).all(1)
    assert not (f & 0x80).any()
    assert np.array_equal((f & 4) == 0, fits)
    # 4x4x4 runs one thread per game in exact int32 from the start: nothing is left to redo there
    assert fits.sum() > N // 2 and (stats["exact_int32_redo"] > 0 or S == 4)
    assert np.array_equal(slab16_to_dense(out16.cpu().numpy(), S)[fits], want[fits])


@pytest.mark.parametrize("S,R", [(4, 7), (9, 23), (16, 49), (16, 128)])
def test_accumulate_int16_targets_beyond_int8(env, S, R):
    # uniform coefficient probabilities: targets reach |T| ~ 86 (9x9x9) / 174 (16x16x16 rank 128), SURVEY 7.3
    N, shift = 90, 2
    tok, tgt, _ = orc.demos_philox(11, 0, N, V5, U5, R, S, shift)
    assert S == 4 or np.abs(tgt).max() > (63 if S == 16 else 40)  # well beyond what the int8 demos of the bench reach (~25)
    tape = torch.from_numpy(tokens_to_tape3(tok)).cuda()
    slab16, flags = env.accumulate_demos16(tape, S, shift)
    assert np.array_equal(slab16_to_dense(slab16.cpu().numpy(), S), tgt) and not flags.any()
    # conversions at the boundary
    f32 = env.expand_states16(slab16, S)
    assert np.array_equal(f32.cpu().numpy(), tgt.astype(np.float32))
    assert torch.equal(env.pack_states16(f32, S), slab16)
    with pytest.raises(env.TensorGameError):
        env.pack_states16(f32 + 0.5, S)


@pytest.mark.parametrize("S,R,dim_t,probs", [(4, 7, 1, P5), (4, 7, 2, U5), (9, 23, 2, P5), (9, 23, 3, U5), (9, 5, 4, P5), (16, 12, 3, P5),
                                             (16, 49, 2, U5)])
def test_demo_store_samples_match_oracle(env, S, R, dim_t, probs):
    # the TMA batcher on the demo-major store: int8 targets, and int16 targets for the wide (uniform-probability) demos;
    # every sample against the oracle's __getitem__ restatement, with the states buffer at every 4-byte alignment and an
    # odd batch size (tail CTA), called through the C ABI
    from mat_mul_b200 import _lib
    from mat_mul_b200.env import _p, _stream, check

    N, shift = 14, 2
    tok, tgt, _ = orc.demos_seeded(9, V5, probs, R, S, shift, N)
    tape = torch.from_numpy(tokens_to_tape3(tok)).cuda()
    wide = probs is U5
    if wide:
        targets, _ = env.accumulate_demos16(tape, S, shift)
    else:
        targets = torch.from_numpy(dense_to_slab(tgt)).cuda()
    store = env.DemoStore.from_tape(tape, targets, S, shift)
    assert torch.equal(store.tape(), tape) and store.records.is_contiguous()
    idx_np = np.random.default_rng(1).permutation(N * R)[:min(N * R, 99)]
    idx = torch.from_numpy(idx_np).cuda()
    for rs in (shift, 1):  # the demos' own shift (true residual) and the reference's fixed replay shift
        items = [orc.demo_getitem(tok[i // R], tgt[i // R], dim_t, i % R, replay_shift=rs) for i in idx_np]
        st, sc, ac, rw = store.samples(idx, dim_t, replay_shift=rs)
        assert np.array_equal(st.cpu().numpy(), np.stack([it[0] for it in items]).astype(np.float32))
        assert np.array_equal(sc.cpu().numpy().ravel(), [it[1] for it in items])
        assert np.array_equal(ac.cpu().numpy(), np.stack([it[2] for it in items]))
        assert np.array_equal(rw.cpu().numpy().ravel(), [it[3] for it in items])
    want = np.stack([orc.demo_getitem(tok[i // R], tgt[i // R], dim_t, i % R, replay_shift=shift)[0] for i in idx_np]).astype(np.float32)
    nb, per = len(idx_np), dim_t * S ** 3
    for off in range(4):
        buf = torch.full((nb * per + 8,), -77.0, dtype=torch.float32, device="cuda")
        states = buf[off:off + nb * per]
        scalars = torch.empty((nb, 1), dtype=torch.float32, device="cuda")
        actions = torch.empty((nb, 3 * S), dtype=torch.int64, device="cuda")
        rewards = torch.empty((nb, 1), dtype=torch.float32, device="cuda")
        check(_lib.lib().tg_demo_sample_dm(_p(store.records), _p(store.targets), int(wide), store.target_bound, N, R, S, dim_t, shift,
                                           _p(idx), nb, _p(states), _p(scalars), _p(actions), _p(rewards), _stream()), "tg_demo_sample_dm")
        got = buf.cpu().numpy()
        assert np.array_equal(got[off:off + nb * per].reshape(want.shape), want), off
        assert (got[:off] == -77.0).all() and (got[off + nb * per:] == -77.0).all(), off  # nothing outside the batch
    # out-of-range indices give all-zero states
    bad = torch.tensor([0, N * R + 5, -3, 1], dtype=torch.int64, device="cuda")
    st, _, _, _ = store.samples(bad, dim_t, replay_shift=shift)
    assert not st[1].any() and not st[2].any() and st[0].any()


def test_demo_store_large_bound_takes_int32_accumulators(env):
    # |target| + R * cmax^3 beyond int16: the 32-bit accumulator variant (replay_shift 0 makes cmax = 8)
    S, R, N, shift = 9, 80, 6, 2
    tok, tgt, _ = orc.demos_seeded(3, V5, U5, R, S, shift, N)
    tape = torch.from_numpy(tokens_to_tape3(tok)).cuda()
    targets, _ = env.accumulate_demos16(tape, S, shift)
    store = env.DemoStore.from_tape(tape, targets, S, shift)
    idx_np = np.arange(0, N * R, 7)
    st, _, _, _ = store.samples(torch.from_numpy(idx_np).cuda(), 2, replay_shift=0)
    want = np.stack([orc.demo_getitem(tok[i // R], tgt[i // R], 2, i % R, replay_shift=0)[0] for i in idx_np]).astype(np.float32)
    assert store.target_bound + R * 8 ** 3 > 32767 and np.array_equal(st.cpu().numpy(), want)


def test_synthetic_demo_dataset_with_wide_targets(env, tmp_path, monkeypatch):
    # the reference's dataset with uniform coefficient probabilities: targets leave the int8 zone, the mirror keeps them as
    # int16 instead of raising (reference: utils.py:218-232 accumulates without limit), items equal the reference's rule
    from mat_mul_b200 import datasets as ds

    monkeypatch.chdir(tmp_path)
    S, R, n = 9, 48, 40
    tok, tgt, _ = orc.demos_seeded(4, V5, U5, R, S, 2, n)  # == the reference loop after torch.manual_seed(4)
    assert np.abs(tgt).max() > 63  # beyond the int8 slab's zone
    torch.manual_seed(4)
    d = ds.SyntheticDemoDataset(R, n, 2, S, "cpu", values=V5, probs=U5, shift=2, save_dir=tmp_path / "demos")
    assert d._slab.dtype == torch.int16 and int(d._slab.abs().max()) == np.abs(tgt).max()
    for i in (0, 5, R - 1, R, 17 * R + 3, n * R - 1):
        st, sc, ac, rw = d[i]
        w = orc.demo_getitem(tok[i // R], tgt[i // R], 2, i % R, replay_shift=1)
        assert np.array_equal(st.numpy(), w[0].astype(np.float32)) and sc.item() == w[1] and rw.item() == w[3]
        assert np.array_equal(ac.numpy(), w[2])
    with pytest.raises(IndexError):
        d.get_batch([n * R])
