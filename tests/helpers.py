"""numpy converters between the oracle's dense arrays and the device formats
(an independent restatement of the layout in include/tensorgame.h)."""
import numpy as np


def geo(S: int):
    rp = (S * S + 3) & ~3
    gp = (S * rp + 15) & ~15
    tp = (3 * S + 15) & ~15
    return rp, gp, tp


def dense_to_slab(T: np.ndarray) -> np.ndarray:
    """(B,S,S,S) ints -> int8 (B,GP)."""
    B, S = T.shape[0], T.shape[-1]
    rp, gp, _ = geo(S)
    assert T.min(initial=0) >= -128 and T.max(initial=0) <= 127
    slab = np.zeros((B, gp), dtype=np.int8)
    rows = slab[:, : S * rp].reshape(B, S, rp)
    rows[:, :, : S * S] = T.reshape(B, S, S * S).astype(np.int8)
    return slab


def slab_to_dense(slab: np.ndarray, S: int) -> np.ndarray:
    """int8 (B,GP) -> int32 (B,S,S,S); asserts padding bytes are zero."""
    B = slab.shape[0]
    rp, gp, _ = geo(S)
    assert slab.shape[1] == gp
    rows = slab[:, : S * rp].reshape(B, S, rp)
    assert not rows[:, :, S * S :].any() and not slab[:, S * rp :].any(), "slab padding must stay zero"
    return rows[:, :, : S * S].reshape(B, S, S, S).astype(np.int32)


def tokens_to_tape(tok: np.ndarray) -> np.ndarray:
    """(B,3S) ints -> uint8 (B,TP)."""
    B, n = tok.shape
    _, _, tp = geo(n // 3)
    tape = np.zeros((B, tp), dtype=np.uint8)
    tape[:, :n] = tok.astype(np.uint8)
    return tape


def tokens_to_tape3(tok: np.ndarray) -> np.ndarray:
    """(N,R,3S) ints -> step-major uint8 (R,N,TP)."""
    N, R, n = tok.shape
    _, _, tp = geo(n // 3)
    tape = np.zeros((R, N, tp), dtype=np.uint8)
    tape[:, :, :n] = np.transpose(tok, (1, 0, 2)).astype(np.uint8)
    return tape


def tape3_to_tokens(tape: np.ndarray, S: int) -> np.ndarray:
    """step-major uint8 (R,N,TP) -> int32 (N,R,3S); asserts padding is zero."""
    assert not tape[:, :, 3 * S :].any(), "tape padding must stay zero"
    return np.transpose(tape[:, :, : 3 * S], (1, 0, 2)).astype(np.int32)
