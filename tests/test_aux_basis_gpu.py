"""GPU parity of K4 demo_sample, K5 change of basis, K6 slice rank, K7 state key."""
import numpy as np
import pytest
import torch

from oracle import tg_oracle as orc
from tests.helpers import dense_to_slab, slab_to_dense, tape3_to_tokens, tokens_to_tape3

pytestmark = pytest.mark.gpu

V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)


@pytest.fixture(scope="module")
def env():
    from mat_mul_b200 import env as e

    assert torch.cuda.is_available()
    return e


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_demo_sample_matches_reference_getitem(env, golden, tag):
    # datasets.py:77-122 for EVERY index of three reference demos, incl. the shift-1 replay quirk (tag c)
    g = golden["getitem"]
    R, dim_t, S, shift, n_demos = (int(x) for x in g[f"{tag}_cfg"])
    tape = torch.from_numpy(tokens_to_tape3(g[f"{tag}_tokens"])).cuda()
    slab = torch.from_numpy(dense_to_slab(g[f"{tag}_targets"])).cuda()
    idx = torch.arange(n_demos * R, device="cuda")
    idx = idx[torch.randperm(len(idx), device="cuda")]
    states, scalars, actions, rewards = env.demo_samples(tape, slab, idx, S, dim_t, replay_shift=1)
    order = idx.cpu().numpy()
    assert np.array_equal(states.cpu().numpy(), g[f"{tag}_states"][order].astype(np.float32))
    assert np.array_equal(scalars.cpu().numpy(), g[f"{tag}_scalars"][order])
    assert np.array_equal(actions.cpu().numpy(), g[f"{tag}_actions"][order])
    assert np.array_equal(rewards.cpu().numpy(), g[f"{tag}_rewards"][order])
    assert states.dtype == torch.float32 and actions.dtype == torch.int64 and scalars.shape == (len(order), 1)


@pytest.mark.parametrize("S,R,dim_t", [(4, 7, 3), (9, 23, 2), (16, 12, 4)])
def test_demo_sample_random_vs_oracle(env, S, R, dim_t):
    N, shift = 40, 2
    tok, tgt, _ = orc.demos_seeded(4, V5, P5, R, S, shift, N)
    tape = torch.from_numpy(tokens_to_tape3(tok)).cuda()
    slab = torch.from_numpy(dense_to_slab(tgt)).cuda()
    rng = np.random.default_rng(0)
    idx = rng.integers(0, N * R, 300)
    for rs in (1, 2):
        st, sc, ac, rw = env.demo_samples(tape, slab, torch.from_numpy(idx).cuda(), S, dim_t, replay_shift=rs)
        for b in (0, 7, 150, 299):
            d, a = divmod(int(idx[b]), R)
            ws, wsc, wa, wr = orc.demo_getitem(tok[d], tgt[d], dim_t, a, replay_shift=rs)
            assert np.array_equal(st[b].cpu().numpy(), ws.astype(np.float32))
            assert sc[b].item() == wsc and rw[b].item() == wr and np.array_equal(ac[b].cpu().numpy(), wa)
    # with the demos' own shift the head is the true residual: all actions up to a, i.e. zero before action 0 is undone
    st, _, _, _ = env.demo_samples(tape, slab, torch.arange(0, N * R, R, device="cuda"), S, 1, replay_shift=shift)
    first = np.stack([orc.action_to_tensor(tok[d, 0], shift) for d in range(N)])
    assert np.array_equal(st[:, 0].cpu().numpy(), first.astype(np.float32))


@pytest.mark.parametrize("S,R,dim_t", [(4, 7, 1), (9, 23, 3), (9, 23, 2), (16, 12, 3)])
def test_demo_sample_every_store_alignment(env, S, R, dim_t):
    # the output rows leave as 16-byte / 8-byte / 4+8+4-byte stores depending on the address of each run of four floats:
    # every sample of the batch against the oracle, with the states buffer starting at each 4-byte offset of a 16-byte
    # line (called through the C ABI with a caller-owned, offset buffer)
    from mat_mul_b200 import _lib
    from mat_mul_b200.env import _p, _stream, check, layout

    N, shift = 12, 2
    tok, tgt, _ = orc.demos_seeded(9, V5, P5, R, S, shift, N)
    tape = torch.from_numpy(tokens_to_tape3(tok)).cuda()
    slab = torch.from_numpy(dense_to_slab(tgt)).cuda()
    idx_np = np.random.default_rng(1).permutation(N * R)[:97]
    idx = torch.from_numpy(idx_np).cuda()
    nb, per = len(idx_np), dim_t * S**3
    want = np.stack([orc.demo_getitem(tok[i // R], tgt[i // R], dim_t, i % R, replay_shift=shift)[0] for i in idx_np]).astype(np.float32)
    for off in range(4):
        buf = torch.full((nb * per + 8,), -77.0, dtype=torch.float32, device="cuda")
        states = buf[off:off + nb * per]
        scalars = torch.empty((nb, 1), dtype=torch.float32, device="cuda")
        actions = torch.empty((nb, 3 * S), dtype=torch.int64, device="cuda")
        rewards = torch.empty((nb, 1), dtype=torch.float32, device="cuda")
        check(_lib.lib().tg_demo_sample(_p(tape), N * layout(S).token_pitch, _p(slab), N, R, S, dim_t, shift, _p(idx), nb,
                                        _p(states), _p(scalars), _p(actions), _p(rewards), _stream()), "tg_demo_sample")
        got = buf.cpu().numpy()
        assert np.array_equal(got[off:off + nb * per].reshape(want.shape), want), off
        assert (got[:off] == -77.0).all() and (got[off + nb * per:] == -77.0).all(), off  # nothing outside the batch


@pytest.mark.parametrize("S", [4, 9, 16])
def test_slice_rank_matches_reference_get_rank(env, golden, S):
    g = golden["ranks"]
    slab = torch.from_numpy(dense_to_slab(g[f"S{S}_T"])).cuda()
    assert np.array_equal(env.slice_rank(slab, S).cpu().numpy(), g[f"S{S}_rank"])


def test_slice_rank_matmul_and_large(env, golden):
    for n, want in zip((2, 3, 4), golden["ranks"]["mm_rank"]):
        slab = torch.from_numpy(dense_to_slab(orc.build_matmul_tensor(n)[None])).cuda()
        assert env.slice_rank(slab, n * n).item() == want
    rng = np.random.default_rng(3)
    T = rng.integers(-60, 60, (3000, 9, 9, 9)) * (rng.random((3000, 9, 9, 9)) < 0.25)
    T[::7] = 0
    got = env.slice_rank(torch.from_numpy(dense_to_slab(T)).cuda(), 9).cpu().numpy()
    assert np.array_equal(got, orc.slice_rank_batch(T))


@pytest.mark.parametrize("S", [4, 9, 16])
def test_state_keys(env, S):
    rng = np.random.default_rng(S)
    B = 2000
    T = rng.integers(-3, 4, (B, S, S, S)) * (rng.random((B, S, S, S)) < 0.1)
    T[10] = T[3]
    T[11] = 0
    keys = env.state_keys(torch.from_numpy(dense_to_slab(T)).cuda(), S).cpu().numpy().view(np.uint64)
    assert np.array_equal(keys, orc.state_key_batch(T))
    assert keys[10] == keys[3] and keys[11] == 0
    assert len(set(keys.tolist())) == len({t.tobytes() for t in T.astype(np.int32)})


@pytest.mark.parametrize("S,R,p", [(4, 7, 0.3), (9, 23, 0.08), (16, 12, 0.03)])
def test_change_of_basis_vs_oracle_and_invariants(env, S, R, p):
    N, shift = 60, 2
    tok, tgt, _ = orc.demos_seeded(8, V5, P5, R, S, shift, N)
    mats = env.sample_unimodular(N, S, seed=21, first=5, p_nonzero=p)
    m = mats.cpu().numpy().astype(np.int64)
    assert np.array_equal(m, orc.sample_unimodular(21, 5, N, S, p))  # sampler contract
    assert all(round(abs(np.linalg.det(x))) == 1 for x in m.reshape(-1, S, S).astype(np.float64))  # unimodular
    slab = torch.from_numpy(dense_to_slab(tgt)).cuda()
    tape = torch.from_numpy(tokens_to_tape3(tok)).cuda()
    out, tape2, flags = env.change_of_basis(slab, mats, S, tape=tape, shift=shift, shift_out=100)
    f = flags.cpu().numpy()
    want = np.stack([orc.change_of_basis(tgt[n], m[n, 0], m[n, 1], m[n, 2]) for n in range(N)])
    assert np.array_equal((f & 4) != 0, (np.abs(want.reshape(N, -1) + 0.5) > 64).any(1))
    ok = np.abs(want.reshape(N, -1)).max(1) <= 127
    assert ok.sum() > N // 2
    assert np.array_equal(slab_to_dense(out.cpu().numpy()[ok], S), want[ok])
    # factors follow: u' = A u, v' = B v, w' = C w and sum u'(x)v'(x)w' = T'
    fac = tok.reshape(N, R, 3, S) - shift
    fac2 = tape3_to_tokens(tape2.cpu().numpy(), S).reshape(N, R, 3, S) - 100
    for n in (0, N // 2, N - 1):
        assert np.array_equal(fac2[n], orc.change_of_basis_factors(fac[n], m[n, 0], m[n, 1], m[n, 2]))
        assert np.array_equal(sum(orc.uvw_to_tensor(*fac2[n, r]) for r in range(R)), want[n])
    assert not (f & 16).any()
    # narrow output alphabet: entries that leave [-shift, shift] are flagged
    _, _, f2 = env.change_of_basis(slab, mats, S, tape=tape, shift=shift, shift_out=shift)
    assert np.array_equal((f2.cpu().numpy() & 16) != 0, (np.abs(fac2) > shift).reshape(N, -1).any(1))


def test_change_of_basis_shared_triple_and_identity(env):
    S, N = 9, 33
    rng = np.random.default_rng(2)
    T = rng.integers(-2, 3, (N, S, S, S)) * (rng.random((N, S, S, S)) < 0.2)
    slab = torch.from_numpy(dense_to_slab(T)).cuda()
    eye = torch.eye(S, dtype=torch.int8, device="cuda").expand(3, S, S).contiguous()
    out, flags = env.change_of_basis(slab, eye, S)
    assert torch.equal(out, slab) and not flags.any()
    P = np.eye(S, dtype=np.int64)[rng.permutation(S)]  # a permutation of the first mode, shared by all games
    mats = torch.from_numpy(np.stack([P, np.eye(S, dtype=np.int64), np.eye(S, dtype=np.int64)]).astype(np.int8)).cuda()
    out, _ = env.change_of_basis(slab, mats, S)
    assert np.array_equal(slab_to_dense(out.cpu().numpy(), S), np.einsum("ia,nabc->nibc", P, T))
    inv = torch.from_numpy(np.stack([P.T, np.eye(S, dtype=np.int64), np.eye(S, dtype=np.int64)]).astype(np.int8)).cuda()
    back, _ = env.change_of_basis(out, inv, S)
    assert torch.equal(back, slab)


@pytest.mark.parametrize("off", [0, 1, 4, 8])
def test_change_of_basis_4_matrices_at_any_alignment_and_shared(env, off):
    # 4x4x4 runs one thread per game with 16-byte matrix loads; matrices that are not 16-byte aligned (a view into a larger
    # buffer) and the shared triple take other routes -- all of them equal the int64 einsum, int8 and int16 out
    S, N = 4, 777
    rng = np.random.default_rng(40 + off)
    T = rng.integers(-128, 128, (N, S, S, S)) * (rng.random((N, S, S, S)) < 0.6)
    m = rng.integers(-2, 3, (N, 3, S, S))
    want = np.einsum("nia,njb,nkc,nabc->nijk", m[:, 0], m[:, 1], m[:, 2], T.astype(np.int64))
    slab = torch.from_numpy(dense_to_slab(T)).cuda()
    buf = torch.zeros(N * 3 * S * S + 64, dtype=torch.int8, device="cuda")
    base = (-buf.data_ptr()) % 16 + off  # byte offset `off` from a 16-byte boundary
    mats = buf[base:base + N * 3 * S * S].view(N, 3, S, S)
    mats.copy_(torch.from_numpy(m.astype(np.int8)))
    assert mats.data_ptr() % 16 == off
    out16, f16 = env.change_of_basis(slab, mats, S, out_dtype=torch.int16)
    fits16 = (np.abs(want.reshape(N, -1) + 0.5) < 32768).all(1)
    assert np.array_equal((f16.cpu().numpy() & 4) == 0, fits16)
    from tests.test_int16_gpu import slab16_to_dense
    assert np.array_equal(slab16_to_dense(out16.cpu().numpy(), S)[fits16], want[fits16])
    out8, f8 = env.change_of_basis(slab, mats, S)
    zone = (np.abs(want.reshape(N, -1) + 0.5) < 64).all(1)
    assert np.array_equal((f8.cpu().numpy() & 4) == 0, zone)
    assert np.array_equal(slab_to_dense(out8.cpu().numpy(), S)[zone], want[zone])
    shared, fs = env.change_of_basis(slab, mats[5:6].contiguous(), S, out_dtype=torch.int16)
    want_s = np.einsum("ia,jb,kc,nabc->nijk", m[5, 0], m[5, 1], m[5, 2], T.astype(np.int64))
    ok = (np.abs(want_s.reshape(N, -1) + 0.5) < 32768).all(1)
    assert np.array_equal(slab16_to_dense(shared.cpu().numpy(), S)[ok], want_s[ok])


@pytest.mark.parametrize("S,amp,dens", [(4, 3, 1.0), (9, 3, 0.6), (9, 1, 0.2), (16, 2, 0.5), (16, 7, 1.0)])
def test_change_of_basis_large_matrices_take_the_exact_path(env, S, amp, dens):
    # dense / large matrices overflow the packed lanes of the fast kernel: those games are redone by the exact
    # int32 kernel, and the result is the same int64 einsum either way
    N = 97
    rng = np.random.default_rng(S * 10 + amp)
    T = rng.integers(-20, 21, (N, S, S, S)) * (rng.random((N, S, S, S)) < 0.3)
    m = rng.integers(-amp, amp + 1, (N, 3, S, S)) * (rng.random((N, 3, S, S)) < dens)
    m[::5] = np.eye(S, dtype=np.int64)  # some games stay on the fast path
    slab = torch.from_numpy(dense_to_slab(T)).cuda()
    out, flags = env.change_of_basis(slab, torch.from_numpy(m.astype(np.int8)).cuda(), S)
    want = np.einsum("nia,njb,nkc,nabc->nijk", m[:, 0], m[:, 1], m[:, 2], T)
    f = flags.cpu().numpy()
    assert not (f & 0x80).any()
    assert np.array_equal((f & 4) != 0, (np.abs(want.reshape(N, -1) + 0.5) > 64).any(1))
    ok = np.abs(want.reshape(N, -1)).max(1) <= 127
    assert ok.sum() >= N // 5
    assert np.array_equal(slab_to_dense(out.cpu().numpy()[ok], S), want[ok])
    assert np.array_equal(slab_to_dense(out.cpu().numpy(), S).astype(np.int8), want.astype(np.int8))  # low byte always


def test_change_of_basis_tensor_core_path_16(env):
    # S = 16 runs on mma.sync (tg_basis_mma.cu): modulo-2^16 byte planes, exact range flag under its norm guard, the
    # exact int32 kernel for the games the guard rejects.  Mix of identity / permutation / sparse unimodular /
    # denser matrices and extreme int8 entries; the result is the int64 einsum either way.
    S, N = 16, 403
    rng = np.random.default_rng(16)
    T = rng.integers(-25, 26, (N, S, S, S)) * (rng.random((N, S, S, S)) < 0.4)
    T[5] = rng.integers(-128, 128, (S, S, S))
    T[6] = -128
    T[7] = 127
    T[8] = 0
    m = np.zeros((N, 3, S, S), dtype=np.int64)
    uni = orc.sample_unimodular(3, 0, N, S, 0.03)  # small norms: the 12-bit K = 32 path
    uni2 = orc.sample_unimodular(4, 0, N, S, 0.12)  # larger norms: the 16-bit path (or, beyond its guard, the exact kernel)
    for n in range(N):
        kind = n % 6
        if kind == 0:
            m[n] = np.eye(S, dtype=np.int64)
        elif kind == 1:
            m[n] = np.stack([np.eye(S, dtype=np.int64)[rng.permutation(S)] * rng.choice([-1, 1], (S, 1)) for _ in range(3)])
        elif kind == 2:
            m[n] = uni[n]
        elif kind == 3:
            m[n] = uni2[n]
        elif kind == 4:
            m[n] = np.eye(S, dtype=np.int64) + rng.integers(-1, 2, (3, S, S)) * (rng.random((3, S, S)) < 0.04)
        else:
            m[n] = rng.integers(-3, 4, (3, S, S)) * (rng.random((3, S, S)) < 0.3)
    m[9, 2] = -128  # ||C||inf far beyond the guard
    m[10, 0, 3] = 127
    m[12, 1, 5, 5] = 8  # an entry of B too large for the 12-bit path's [M | 16 M] fragment
    m[18, 0, 2, 2] = -8
    slab = torch.from_numpy(dense_to_slab(T)).cuda()
    out, flags = env.change_of_basis(slab, torch.from_numpy(m.astype(np.int8)).cuda(), S)
    want = np.einsum("nia,njb,nkc,nabc->nijk", m[:, 0], m[:, 1], m[:, 2], T)
    f = flags.cpu().numpy()
    assert not (f & 0x80).any()
    assert np.array_equal((f & 4) != 0, ((want < -64) | (want > 63)).reshape(N, -1).any(1))
    assert np.array_equal(f & ~np.uint8(4 | 32 | 64), np.zeros(N, dtype=np.uint8))  # besides RANGE only the informational path bits
    assert (f & 64).any() and (f[::6] & (32 | 64) == 0).all()  # some games need the exact kernel; identities stay on the f16 path
    assert np.array_equal(slab_to_dense(out.cpu().numpy(), S).astype(np.int8), want.astype(np.int8))
    # one shared triple for every game
    out1, f1 = env.change_of_basis(slab, torch.from_numpy(m[2].astype(np.int8)).cuda(), S)
    want1 = np.einsum("ia,jb,kc,nabc->nijk", m[2, 0], m[2, 1], m[2, 2], T)
    assert np.array_equal(slab_to_dense(out1.cpu().numpy(), S).astype(np.int8), want1.astype(np.int8))
    assert np.array_equal((f1.cpu().numpy() & 4) != 0, ((want1 < -64) | (want1 > 63)).reshape(N, -1).any(1))
