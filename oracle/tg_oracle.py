"""ctypes/numpy front end of the CPU oracle (oracle/tg_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from mat_mul_b200/.
Every function restates reference behaviour; the C side cites file:line.
Residuals are int32 numpy arrays (..., S, S, S); tokens int32 (..., 3S).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libtg_oracle.so"

FLAG_TERMINAL = 1
FLAG_NULL = 2


def build(force: bool = False) -> Path:
    """Compile the oracle with the recipe in oracle/Makefile."""
    src = _HERE / "tg_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s", "-B"], check=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.orc_mt_f64.restype = C.c_double
        _lib.orc_demos_from_ustream.restype = C.c_int64
        _lib.orc_state_key.restype = C.c_uint64
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


# ---------------------------------------------------------------- RNG
def mt_doubles(seed: int, n: int) -> np.ndarray:
    """First n doubles of torch.manual_seed(seed); torch.rand(n, dtype=float64) on CPU."""
    out = np.empty(n, dtype=np.float64)
    lib().orc_mt_fill_f64(C.c_uint32(seed), C.c_int64(n), _p(out))
    return out


def categorical_cdf(probs) -> np.ndarray:
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    cdf = np.empty_like(probs)
    lib().orc_categorical_cdf(_p(probs), C.c_int(len(probs)), _p(cdf))
    return cdf


def categorical_pick(cdf: np.ndarray, u: float) -> int:
    cdf = np.ascontiguousarray(cdf, dtype=np.float32)
    return int(lib().orc_categorical_pick(_p(cdf), C.c_int(len(cdf)), C.c_double(u)))


# ---------------------------------------------------------------- env
def uvw_to_tensor(u, v, w) -> np.ndarray:
    u, v, w = _i32(u), _i32(v), _i32(w)
    S = u.shape[-1]
    out = np.empty((S, S, S), dtype=np.int32)
    lib().orc_uvw_to_tensor(_p(u), _p(v), _p(w), C.c_int(S), _p(out))
    return out


def action_to_tensor(action, shift: int = 1) -> np.ndarray:
    action = _i32(action)
    S = action.shape[-1] // 3
    out = np.empty((S, S, S), dtype=np.int32)
    lib().orc_action_to_tensor(_p(action), C.c_int(S), C.c_int(shift), _p(out))
    return out


def step_batch(T, actions, shift: int):
    """T (B,S,S,S) int32, actions (B,3S) -> (T_out, flags uint32 (B,), nnz int32 (B,))."""
    T, actions = _i32(T), _i32(actions)
    B, S = T.shape[0], T.shape[-1]
    out = np.empty_like(T)
    flags = np.empty(B, dtype=np.uint32)
    nnz = np.empty(B, dtype=np.int32)
    lib().orc_step_batch(_p(T), _p(actions), C.c_int64(B), C.c_int(S), C.c_int(shift), _p(out), _p(flags), _p(nnz))
    return out, flags, nnz


def step_batch_f32(T, actions, shift: int, out=None, flags=None, nnz=None):
    """Reference data types: float32 residuals, int64 tokens (bench CPU baseline leg)."""
    T = np.ascontiguousarray(T, dtype=np.float32)
    actions = np.ascontiguousarray(actions, dtype=np.int64)
    B, S = T.shape[0], T.shape[-1]
    out = np.empty_like(T) if out is None else out
    flags = np.empty(B, dtype=np.uint8) if flags is None else flags
    nnz = np.empty(B, dtype=np.int32) if nnz is None else nnz
    lib().orc_step_batch_f32(_p(T), _p(actions), C.c_int64(B), C.c_int(S), C.c_int(shift), _p(out), _p(flags), _p(nnz))
    return out, flags, nnz


def take_actions(actions, T, shift: int = 1) -> np.ndarray:
    actions = _i32(actions)
    T = _i32(T).copy()
    S = T.shape[-1]
    lib().orc_take_actions(_p(actions), C.c_int(actions.shape[0]), C.c_int(S), C.c_int(shift), _p(T))
    return T


def rollout_batch(T, tape, shift: int):
    """T (B,S,S,S), tape (B,K,3S) -> (T_out, flags, nnz, steps)."""
    T, tape = _i32(T), _i32(tape)
    B, S, K = T.shape[0], T.shape[-1], tape.shape[1]
    out = np.empty_like(T)
    flags = np.empty(B, dtype=np.uint32)
    nnz = np.empty(B, dtype=np.int32)
    steps = np.empty(B, dtype=np.int32)
    lib().orc_rollout_batch(_p(T), _p(tape), C.c_int64(B), C.c_int(K), C.c_int(S), C.c_int(shift), _p(out), _p(flags), _p(nnz), _p(steps))
    return out, flags, nnz, steps


# ---------------------------------------------------------------- demos
def demos_from_ustream(ustream, values, probs, R: int, S: int, shift: int, n_demos: int):
    """Reference-stream demo generation -> (tokens (n,R,3S), targets (n,S,S,S), consumed)."""
    ustream = np.ascontiguousarray(ustream, dtype=np.float64)
    values = _i32(values)
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    tokens = np.zeros((n_demos, R, 3 * S), dtype=np.int32)
    targets = np.zeros((n_demos, S, S, S), dtype=np.int32)
    consumed = C.c_int64(0)
    n = lib().orc_demos_from_ustream(
        _p(ustream), C.c_int64(len(ustream)), _p(values), _p(probs), C.c_int(len(values)), C.c_int(R), C.c_int(S),
        C.c_int(shift), C.c_int64(n_demos), _p(tokens), _p(targets), C.byref(consumed))
    return tokens[:n], targets[:n], int(consumed.value)


def demos_seeded(seed: int, values, probs, R: int, S: int, shift: int, n_demos: int, slack: float = 4.0):
    """Same as running the reference loop after torch.manual_seed(seed)."""
    probs32 = np.asarray(probs, dtype=np.float32)
    p0 = float(probs32[np.asarray(values) == 0].sum() / probs32.sum()) if 0 in list(values) else 0.0
    acc = max((1.0 - p0 ** S) ** 3, 1e-3)
    n_u = int(n_demos * R * 3 * S / acc * slack) + 4096
    while True:
        tok, tgt, used = demos_from_ustream(mt_doubles(seed, n_u), values, probs, R, S, shift, n_demos)
        if len(tok) == n_demos:
            return tok, tgt, used
        n_u *= 2


def demo_getitem(tokens, target, dim_t: int, idx_action: int, replay_shift: int = 1):
    tokens, target = _i32(tokens), _i32(target)
    R, S = tokens.shape[0], target.shape[-1]
    state = np.empty((dim_t, S, S, S), dtype=np.int32)
    scalar, reward = C.c_float(0), C.c_float(0)
    action = np.empty(3 * S, dtype=np.int32)
    lib().orc_demo_getitem(_p(tokens), _p(target), C.c_int(R), C.c_int(S), C.c_int(dim_t), C.c_int(idx_action),
                           C.c_int(replay_shift), _p(state), C.byref(scalar), _p(action), C.byref(reward))
    return state, float(scalar.value), action, float(reward.value)


def philox4x32_10(ctr, key) -> np.ndarray:
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out


def philox_thresholds(probs) -> np.ndarray:
    probs = np.ascontiguousarray(probs, dtype=np.float64)
    thr = np.empty(len(probs), dtype=np.uint32)
    lib().orc_philox_thresholds(_p(probs), C.c_int(len(probs)), _p(thr))
    return thr


def alias_applies(values, probs, S: int) -> bool:
    """Contract v2 (group alias sampler) covers alphabets of at most five values with P(0) <= 0.999."""
    v, p = _i32(values), np.ascontiguousarray(probs, dtype=np.float64)
    return bool(lib().orc_alias_applies(_p(v), _p(p), C.c_int(len(v)), C.c_int(S)))


def alias_tables(values, probs, S: int) -> np.ndarray:
    """The alias tables of contract v2: uint16 (8, 128) -- [0] plain group of three, [1] plain single entry, [2 + g] tilted
    table of group g."""
    v, p = _i32(values), np.ascontiguousarray(probs, dtype=np.float64)
    out = np.zeros((8, 128), dtype=np.uint16)
    lib().orc_alias_tables(_p(v), _p(p), C.c_int(len(v)), C.c_int(S), _p(out))
    return out


def demos_philox(seed: int, d0: int, n: int, values, probs, R: int, S: int, shift: int, max_tries: int = 64):
    """Throughput-mode contracts (ours) -> (tokens (n,R,3S), targets (n,S,S,S), exhausted): v2 (group alias tables, no
    rejection loop) where it applies, else v1 (thresholded 15-bit draws, bounded rejection loop)."""
    values = _i32(values)
    if alias_applies(values, probs, S):
        p64 = np.ascontiguousarray(probs, dtype=np.float64)
        tokens = np.empty((n, R, 3 * S), dtype=np.int32)
        targets = np.empty((n, S, S, S), dtype=np.int32)
        lib().orc_demos_philox_v2_batch(C.c_uint64(seed), C.c_uint64(d0), C.c_int64(n), _p(values), _p(p64), C.c_int(len(values)),
                                        C.c_int(R), C.c_int(S), C.c_int(shift), _p(tokens), _p(targets))
        return tokens, targets, 0
    thr = philox_thresholds(probs)
    tokens = np.empty((n, R, 3 * S), dtype=np.int32)
    targets = np.empty((n, S, S, S), dtype=np.int32)
    ex = C.c_int32(0)
    lib().orc_demos_philox_batch(C.c_uint64(seed), C.c_uint64(d0), C.c_int64(n), _p(values), _p(thr),
                                 C.c_int(len(values)), C.c_int(R), C.c_int(S), C.c_int(shift), C.c_int(max_tries),
                                 _p(tokens), _p(targets), C.byref(ex))
    return tokens, targets, int(ex.value)


# ---------------------------------------------------------------- Strassen / matmul tensor
def strassen_factors():
    uu, vv, ww = (np.empty((7, 4), dtype=np.int32) for _ in range(3))
    lib().orc_strassen_factors(_p(uu), _p(vv), _p(ww))
    return uu, vv, ww


def uvw_to_demo(uu, vv, ww, shift: int = 1):
    uu, vv, ww = _i32(uu), _i32(vv), _i32(ww)
    n, S = uu.shape
    tensor = np.empty((S, S, S), dtype=np.int32)
    actions = np.empty((n, 3 * S), dtype=np.int32)
    lib().orc_uvw_to_demo(_p(uu), _p(vv), _p(ww), C.c_int(n), C.c_int(S), C.c_int(shift), _p(tensor), _p(actions))
    return tensor, actions


def strassen_dataset():
    states = np.empty((448, 4, 4, 4), dtype=np.int32)
    actions = np.empty((448, 12), dtype=np.int32)
    rewards = np.empty(448, dtype=np.float32)
    bits = np.empty(448, dtype=np.int32)
    n = lib().orc_strassen_dataset(_p(states), _p(actions), _p(rewards), _p(bits))
    assert n == 448
    return states, actions, rewards, bits


def build_matmul_tensor(n: int) -> np.ndarray:
    S = n * n
    out = np.empty((S, S, S), dtype=np.int32)
    lib().orc_build_matmul_tensor(C.c_int(n), _p(out))
    return out


# ---------------------------------------------------------------- rank / basis / key
def slice_rank_batch(T) -> np.ndarray:
    T = _i32(T)
    B, S = T.shape[0], T.shape[-1]
    ranks = np.empty(B, dtype=np.int32)
    lib().orc_slice_rank_batch(_p(T), C.c_int64(B), C.c_int(S), _p(ranks))
    return ranks


def change_of_basis(T, A, Bm, Cm) -> np.ndarray:
    T, A, Bm, Cm = _i32(T), _i32(A), _i32(Bm), _i32(Cm)
    S = T.shape[-1]
    out = np.empty((S, S, S), dtype=np.int64)
    lib().orc_change_of_basis(_p(T), _p(A), _p(Bm), _p(Cm), C.c_int(S), _p(out))
    return out


def change_of_basis_factors(factors, A, Bm, Cm) -> np.ndarray:
    """factors (R,3,S) coefficient values -> transformed (R,3,S) int64."""
    factors, A, Bm, Cm = _i32(factors), _i32(A), _i32(Bm), _i32(Cm)
    R, S = factors.shape[0], factors.shape[-1]
    out = np.empty((R, 3, S), dtype=np.int64)
    lib().orc_change_of_basis_factors(_p(factors), C.c_int(R), _p(A), _p(Bm), _p(Cm), C.c_int(S), _p(out))
    return out


def sample_unimodular(seed: int, first: int, n: int, S: int, p_nonzero: float = 0.3) -> np.ndarray:
    """Our unimodular sampler contract -> int32 (n, 3, S, S)."""
    mats = np.empty((n, 3, S, S), dtype=np.int32)
    thr = int(p_nonzero * 128.0)
    for i in range(n):
        lib().orc_sample_unimodular(C.c_uint64(seed), C.c_uint64(first + i), C.c_int(S), C.c_uint32(thr), _p(mats[i]))
    return mats


def state_key_batch(T) -> np.ndarray:
    T = _i32(T)
    B, S = T.shape[0], T.shape[-1]
    keys = np.empty(B, dtype=np.uint64)
    lib().orc_state_key_batch(_p(T), C.c_int64(B), C.c_int(S), _p(keys))
    return keys


def num_threads() -> int:
    return int(lib().orc_num_threads())


def use_all_threads() -> int:
    """OpenMP threads = every core the process may run on (torchrun sets OMP_NUM_THREADS=1 for its workers)."""
    return int(lib().orc_use_all_threads())
