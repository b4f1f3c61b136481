"""GPU parity of K1 (tg_step) against the CPU oracle, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import tg_oracle as orc
from tests.helpers import dense_to_slab, geo, slab_to_dense, tokens_to_tape

pytestmark = pytest.mark.gpu

CFG = {4: ((-1, 0, 1), (0.15, 0.7, 0.15), 7, 1), 9: ((-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05), 23, 2),
       16: ((-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05), 12, 2)}


@pytest.fixture(scope="module")
def env():
    from mat_mul_b200 import env as e

    assert torch.cuda.is_available()
    return e


def gpu_step(env, T, tok, S, shift):
    dev = torch.device("cuda:0")
    slab = torch.from_numpy(dense_to_slab(T)).to(dev)
    tape = torch.from_numpy(tokens_to_tape(tok)).to(dev)
    out, flags, nnz = env.step_batch(slab, tape, S, shift)
    torch.cuda.synchronize()
    return slab_to_dense(out.cpu().numpy(), S), flags.cpu().numpy(), nnz.cpu().numpy()


@pytest.mark.parametrize("S", [4, 9, 16])
def test_golden_transitions(env, golden, S):
    g = golden["steps"]
    states, actions = g[f"S{S}_states"], g[f"S{S}_actions"]
    n, k = actions.shape[:2]
    heads = np.repeat(states[:, 0][:, None], k, axis=1).reshape(n * k, S, S, S)
    out, flags, nnz = gpu_step(env, heads, actions.reshape(n * k, -1), S, 1)
    assert np.array_equal(out.reshape(n, k, S, S, S), g[f"S{S}_child_heads"])
    assert np.array_equal((flags & 2) == 0, g[f"S{S}_not_null"].reshape(-1))
    assert np.array_equal((flags & 1) != 0, g[f"S{S}_terminal"].reshape(-1))
    # training.py:253-267
    sb, aa = g[f"S{S}_ta_states"], g[f"S{S}_ta_tokens"][:, 0]
    out, flags, nnz = gpu_step(env, sb[:, 0], aa, S, 2)
    assert np.array_equal(out, g[f"S{S}_ta_new_states"][:, 0])
    assert np.array_equal(nnz.reshape(-1, 4).min(1), g[f"S{S}_ta_best_values"])
    assert np.array_equal(nnz.reshape(-1, 4).argmin(1), g[f"S{S}_ta_best_indices"])


@pytest.mark.parametrize("S", [4, 9, 16])
@pytest.mark.parametrize("B", [1, 3, 257, 5000])
def test_random_vs_oracle(env, S, B):
    rng = np.random.default_rng(S * 1000 + B)
    shift = 2
    T = rng.integers(-60, 60, (B, S, S, S)) * (rng.random((B, S, S, S)) < 0.3)
    tok = rng.integers(0, 5, (B, 3 * S))
    tok[rng.random((B, 3 * S)) < 0.5] = 2
    tok[0, :S] = 2  # null action
    want, wflags, wnnz = orc.step_batch(T, tok, shift)
    out, flags, nnz = gpu_step(env, T, tok, S, shift)
    assert np.array_equal(out, want)
    assert np.array_equal(flags & 3, wflags)
    assert np.array_equal(nnz, wnnz)
    want_range = (np.abs(want.reshape(B, -1) + 0.5) > 64).any(1)  # outside [-64, 63]
    assert np.array_equal((flags & 4) != 0, want_range)


@pytest.mark.parametrize("S", [4, 9, 16])
@pytest.mark.parametrize("shift", [1, 2, 3, 4])
def test_shifts_and_dense_coefficients(env, S, shift):
    rng = np.random.default_rng(shift)
    B = 300
    T = rng.integers(-63, 64, (B, S, S, S))
    tok = rng.integers(0, 2 * shift + 1, (B, 3 * S))
    want, wflags, wnnz = orc.step_batch(T, tok, shift)
    keep = np.abs(want.reshape(B, -1)).max(1) <= 127  # int8 contract: |T| <= 63 and |uvw| <= 64 never overflow
    assert keep.all()
    out, flags, nnz = gpu_step(env, T, tok, S, shift)
    assert np.array_equal(out, want) and np.array_equal(flags & 3, wflags) and np.array_equal(nnz, wnnz)


@pytest.mark.parametrize("S", [4, 9, 16])
def test_demo_replay_reaches_zero(env, S):
    values, probs, R, shift = CFG[S]
    B = 777
    tok, tgt, _ = orc.demos_seeded(1, values, probs, R, S, shift, B)
    dev = torch.device("cuda:0")
    slab = torch.from_numpy(dense_to_slab(tgt)).to(dev)
    cur = tgt
    for r in reversed(range(R)):
        tape = torch.from_numpy(tokens_to_tape(tok[:, r])).to(dev)
        slab, flags, nnz = env.step_batch(slab, tape, S, shift, out=slab)  # in place
        cur, wflags, wnnz = orc.step_batch(cur, tok[:, r], shift)
        assert np.array_equal(flags.cpu().numpy() & 3, wflags) and np.array_equal(nnz.cpu().numpy(), wnnz)
    assert not slab.any() and (flags.cpu().numpy() & 1).all() and not cur.any()


def test_empty_and_bad_args(env):
    dev = torch.device("cuda:0")
    out, flags, nnz = env.step_batch(torch.zeros((0, 768), dtype=torch.int8, device=dev),
                                     torch.zeros((0, 32), dtype=torch.uint8, device=dev), 9, 2)
    assert out.shape == (0, 768) and flags.numel() == 0
    with pytest.raises(env.TensorGameError):
        env.step_batch(torch.zeros((2, 100), dtype=torch.int8, device=dev), torch.zeros((2, 32), dtype=torch.uint8, device=dev), 9, 2)


@pytest.mark.parametrize("S", [4, 9, 16])
def test_boundary_conversions(env, S):
    rng = np.random.default_rng(5)
    B = 1000
    dev = torch.device("cuda:0")
    T = rng.integers(-128, 128, (B, S, S, S))
    state = torch.zeros((B, 2, S, S, S), dtype=torch.float32, device=dev)
    state[:, 0] = torch.from_numpy(T).float().to(dev)
    slab = env.pack_states(state[:, 0], S)  # strided batch, dense games
    assert np.array_equal(slab.cpu().numpy(), dense_to_slab(T))
    assert torch.equal(env.slab_view(slab, S).float(), state[:, 0])
    back = torch.full((B, 2, S, S, S), 7.0, device=dev)
    env.expand_states(slab, S, out=back[:, 0])
    assert torch.equal(back[:, 0], state[:, 0]) and (back[:, 1] == 7).all()
    acts = torch.from_numpy(rng.integers(0, 5, (B, 3 * S))).to(dev)
    tape = env.pack_actions(acts, S)
    assert np.array_equal(tape.cpu().numpy(), tokens_to_tape(acts.cpu().numpy()))
    assert torch.equal(env.unpack_actions(tape, S), acts)
    with pytest.raises(env.TensorGameError):
        env.pack_states(state[:, 0] + 0.5, S)


@pytest.mark.parametrize("S", [4, 9, 16])
def test_boundary_conversions_every_alignment(env, S):
    # the float32 side is read / written with 16-, 8- or 4-byte accesses chosen per run of four floats: games at every
    # 4-byte alignment (odd batch stride, offset base pointer), nothing outside the games touched
    rng = np.random.default_rng(6)
    B, S3 = 257, S**3
    dev = torch.device("cuda:0")
    T = rng.integers(-128, 128, (B, S, S, S))
    for off in range(4):
        for stride in (S3 + 1, S3 + 2):
            buf = torch.full((off + B * stride + 8,), -77.0, dtype=torch.float32, device=dev)
            games = buf[off:off + B * stride].view(B, stride)[:, :S3].view(B, S, S, S)
            games.copy_(torch.from_numpy(T).float().to(dev))
            slab = env.pack_states(games, S)
            assert np.array_equal(slab.cpu().numpy(), dense_to_slab(T)), (off, stride)
            out = torch.full_like(buf, -77.0)
            oview = out[off:off + B * stride].view(B, stride)[:, :S3].view(B, S, S, S)
            env.expand_states(slab, S, out=oview)
            assert torch.equal(out, buf), (off, stride)  # games identical, the gaps and both ends untouched


@pytest.mark.parametrize("S", [4, 9])
def test_host_path_matches_device_path(env, S):
    rng = np.random.default_rng(6)
    B, shift = 70000, 2
    rp, gp, tp = geo(S)
    T = rng.integers(-3, 4, (B, S, S, S)) * (rng.random((B, S, S, S)) < 0.3)
    tok = rng.integers(0, 5, (B, 3 * S))
    slab = torch.from_numpy(dense_to_slab(T)).pin_memory()
    tape = torch.from_numpy(tokens_to_tape(tok)).pin_memory()
    out = torch.empty_like(slab).pin_memory()
    flags = torch.empty(B, dtype=torch.uint8).pin_memory()
    nnz = torch.empty(B, dtype=torch.int32).pin_memory()
    hs = env.HostStepper(S, 0, chunk=1 << 14)
    hs.step(slab, tape, out, flags, nnz, shift)
    hs.close()
    want, wflags, wnnz = orc.step_batch(T, tok, shift)
    assert np.array_equal(slab_to_dense(out.numpy(), S), want)
    assert np.array_equal(flags.numpy() & 3, wflags) and np.array_equal(nnz.numpy(), wnnz)


@pytest.mark.parametrize("S,B,k,shift", [(4, 1, 1, 1), (4, 777, 8, 1), (9, 13, 3, 2), (9, 500, 8, 2), (16, 5, 2, 2), (16, 67, 8, 2)])
def test_expand_children_equals_k_single_steps(env, S, B, k, shift):
    """K8: batched leaf expansion = tg_step on the k-fold replicated parents + tg_state_key of the children."""
    rng = np.random.default_rng(S * 100 + k)
    T = rng.integers(-3, 4, (B, S, S, S)) * (rng.random((B, S, S, S)) < 0.15)
    tok = rng.integers(0, 2 * shift + 1, (B, k, 3 * S))
    tok[rng.random((B, k, 3 * S)) < 0.5] = shift
    tok[0, 0] = shift                      # a null action
    if B > 2:                              # a child that solves the game: parent = the rank-1 tensor of its action
        f = tok[2, k - 1].reshape(3, S) - shift
        f[:, 0] = 1
        tok[2, k - 1] = (f + shift).reshape(-1)
        T[2] = orc.uvw_to_tensor(f[0], f[1], f[2])
    slab = torch.from_numpy(dense_to_slab(T)).cuda()
    tape = torch.from_numpy(np.stack([tokens_to_tape(tok[:, c]) for c in range(k)], axis=1)).cuda().contiguous()
    children, flags, nnz, keys = env.expand_children(slab, tape, S, shift)
    rep = slab.unsqueeze(1).expand(B, k, slab.shape[1]).reshape(B * k, -1).contiguous()
    want, wf, wn = env.step_batch(rep, tape.reshape(B * k, -1).contiguous(), S, shift)
    assert torch.equal(children.reshape(B * k, -1), want) and torch.equal(flags.reshape(-1), wf)
    assert torch.equal(nnz.reshape(-1), wn) and torch.equal(keys.reshape(-1), env.state_keys(want, S))
    # and against the oracle
    od, of, on = orc.step_batch(np.repeat(T, k, axis=0), tok.reshape(B * k, -1), shift)
    assert np.array_equal(slab_to_dense(children.reshape(B * k, -1).cpu().numpy(), S), od)
    assert np.array_equal(flags.reshape(-1).cpu().numpy() & 3, of)
    assert (flags[0, 0] & 2) and (B <= 2 or (flags[2, k - 1] & 1))
    c2, f2, n2, k2 = env.expand_children(slab, tape, S, shift, with_keys=False)
    assert k2 is None and torch.equal(c2, children) and torch.equal(f2, flags)


def test_new_entry_points_empty_and_bad_args(env):
    """N = 0 is a no-op; wrong sizes / misaligned pointers / unsupported shapes come back as TG_E_ARG, never a crash."""
    from mat_mul_b200 import _lib
    from mat_mul_b200.env import _p, _stream

    L = _lib.lib()
    S = 9
    lay = env.layout(S)
    slab = torch.zeros((4, lay.game_pitch), dtype=torch.int8, device="cuda")
    tape = torch.full((4, 2, lay.token_pitch), 2, dtype=torch.uint8, device="cuda")
    tape[:, :, 3 * S:] = 0
    kids = torch.empty((4, 2, lay.game_pitch), dtype=torch.int8, device="cuda")
    fl = torch.empty((4, 2), dtype=torch.uint8, device="cuda")
    nz = torch.empty((4, 2), dtype=torch.int32, device="cuda")
    assert L.tg_expand_children(_p(slab), _p(tape), 2, _p(kids), _p(fl), _p(nz), None, 0, S, 2, _stream()) == 0       # B = 0
    assert L.tg_expand_children(_p(slab), _p(tape), 0, _p(kids), _p(fl), _p(nz), None, 4, S, 2, _stream()) != 0       # k = 0
    assert L.tg_expand_children(_p(slab), _p(tape), 2, _p(kids), _p(fl), _p(nz), None, 4, 5, 2, _stream()) != 0       # S = 5
    assert L.tg_expand_children(slab.data_ptr() + 1, _p(tape), 2, _p(kids), _p(fl), _p(nz), None, 4, S, 2, _stream()) != 0
    assert L.tg_expand_children(_p(slab), _p(tape), 100000, _p(kids), _p(fl), _p(nz), None, 4, S, 2, _stream()) != 0  # smem
    c, f, n, k = env.expand_children(slab, tape, S, 2)  # null actions on the zero state: terminal and null
    torch.cuda.synchronize()
    assert not c.any() and bool(((f & 3) == 3).all()) and not n.any() and not k.any()
    # tensor-core accumulate: only S = 16 and R <= 64
    t16 = torch.zeros((65, 3, 48), dtype=torch.uint8, device="cuda")
    s16 = torch.empty((3, 4096), dtype=torch.int8, device="cuda")
    f16 = torch.empty(3, dtype=torch.uint8, device="cuda")
    assert L.tg_demo_accumulate_tc(_p(t16), 3 * 48, 3, 65, 16, 2, _p(s16), _p(f16), _stream()) != 0
    assert L.tg_demo_accumulate_tc(_p(t16), 3 * 48, 3, 64, 9, 2, _p(s16), _p(f16), _stream()) != 0
    assert L.tg_demo_accumulate_tc(_p(t16), 3 * 48, 0, 64, 16, 2, _p(s16), _p(f16), _stream()) == 0
    torch.cuda.synchronize()
