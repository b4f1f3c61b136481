"""Batched TensorGame environment on device formats (additions alongside the
reference-named API; the mirrors in utils.py / datasets.py / act.py call these).

All tensors are CUDA tensors owned by PyTorch; kernels run on the current
torch stream through the C ABI (include/tensorgame.h).  Nothing here computes
on the CPU: without the CUDA library or a CUDA tensor these functions raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import FLAG_EXHAUSTED, FLAG_NULL, FLAG_RANGE, FLAG_TERMINAL, TensorGameError, check  # noqa: F401


@dataclass(frozen=True)
class Layout:
    S: int
    row_pitch: int    # RP: bytes per i-row
    game_pitch: int   # GP: bytes per game in a slab
    token_pitch: int  # TP: bytes per game in a tape

    @property
    def algorithmic_step_bytes(self) -> int:
        """2*S^3 + 3S + 1 + 4 bytes per env step."""
        return 2 * self.S ** 3 + 3 * self.S + 5


def layout(S: int) -> Layout:
    rp, gp, tp = C.c_int(), C.c_int(), C.c_int()
    check(_lib.lib().tg_layout(S, C.byref(rp), C.byref(gp), C.byref(tp)), f"tg_layout(S={S})")
    return Layout(S, rp.value, gp.value, tp.value)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t: torch.Tensor, name: str, dtype: torch.dtype) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TensorGameError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise TensorGameError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise TensorGameError(f"{name} must be contiguous")


def _p(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


# ---------------------------------------------------------------- formats
def new_slab(B: int, S: int, device) -> torch.Tensor:
    """Zeroed residual slab, int8 (B, GP)."""
    return torch.zeros((B, layout(S).game_pitch), dtype=torch.int8, device=device)


def slab_view(slab: torch.Tensor, S: int) -> torch.Tensor:
    """Strided (B, S, S, S) int8 view of a slab (no copy)."""
    lay = layout(S)
    return slab.as_strided((slab.shape[0], S, S, S), (lay.game_pitch, lay.row_pitch, S, 1))


def pack_states(heads: torch.Tensor, S: int | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """float32 heads (B, S, S, S) (any batch stride, dense game) -> slab.  Raises if a value
    is not an integer in [-128, 127]."""
    if heads.dtype != torch.float32 or not heads.is_cuda:
        raise TensorGameError("heads must be a CUDA float32 tensor")
    S = S or heads.shape[-1]
    B = heads.shape[0]
    if heads.shape[1:] != (S, S, S) or heads[0].numel() and not heads[0].is_contiguous():
        raise TensorGameError("heads must be (B, S, S, S) with dense games")
    out = new_slab(B, S, heads.device) if out is None else out
    flag = torch.zeros(1, dtype=torch.int32, device=heads.device)
    stride = heads.stride(0) if B > 1 else S ** 3
    check(_lib.lib().tg_pack_f32(_p(heads), stride, _p(out), B, S, _p(flag), _stream()), "tg_pack_f32")
    if int(flag.item()):
        raise TensorGameError("pack_states: residual entries must be integers in [-128, 127]")
    return out


def expand_states(slab: torch.Tensor, S: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """slab -> float32 (B, S, S, S) (the dtype model.py consumes)."""
    _need_cuda(slab, "slab", torch.int8)
    B = slab.shape[0]
    if out is None:
        out = torch.empty((B, S, S, S), dtype=torch.float32, device=slab.device)
    stride = out.stride(0) if B > 1 else S ** 3
    check(_lib.lib().tg_expand_f32(_p(slab), _p(out), stride, B, S, _stream()), "tg_expand_f32")
    return out


def pack_actions(actions: torch.Tensor, S: int | None = None) -> torch.Tensor:
    """int64 tokens (B, 3S) -> tape uint8 (B, TP)."""
    _need_cuda(actions, "actions", torch.int64)
    S = S or actions.shape[-1] // 3
    B = actions.shape[0]
    tape = torch.empty((B, layout(S).token_pitch), dtype=torch.uint8, device=actions.device)
    flag = torch.zeros(1, dtype=torch.int32, device=actions.device)
    check(_lib.lib().tg_pack_actions_i64(_p(actions), _p(tape), B, S, _p(flag), _stream()), "tg_pack_actions_i64")
    if int(flag.item()):
        raise TensorGameError("pack_actions: tokens must be in [0, 255]")
    return tape


def unpack_actions(tape: torch.Tensor, S: int) -> torch.Tensor:
    _need_cuda(tape, "tape", torch.uint8)
    B = tape.shape[0]
    out = torch.empty((B, 3 * S), dtype=torch.int64, device=tape.device)
    check(_lib.lib().tg_unpack_actions_i64(_p(tape), _p(out), B, S, _stream()), "tg_unpack_actions_i64")
    return out


# ---------------------------------------------------------------- K1
def step_batch(slab: torch.Tensor, tape: torch.Tensor, S: int, shift: int, out: torch.Tensor | None = None,
               flags: torch.Tensor | None = None, nnz: torch.Tensor | None = None):
    """One transition for every game: out = slab - u(x)v(x)w.  Returns (out, flags, nnz).

    flags uint8 (B,): FLAG_TERMINAL | FLAG_NULL | FLAG_RANGE; nnz int32 (B,).
    Reference: act.py:266-275, training.py:253-267, utils.py:181-194.
    """
    _need_cuda(slab, "slab", torch.int8)
    _need_cuda(tape, "tape", torch.uint8)
    B = slab.shape[0]
    lay = layout(S)
    if slab.shape != (B, lay.game_pitch) or tape.shape != (B, lay.token_pitch):
        raise TensorGameError(f"expected slab (B,{lay.game_pitch}) and tape (B,{lay.token_pitch})")
    out = torch.empty_like(slab) if out is None else out
    flags = torch.empty(B, dtype=torch.uint8, device=slab.device) if flags is None else flags
    nnz = torch.empty(B, dtype=torch.int32, device=slab.device) if nnz is None else nnz
    check(_lib.lib().tg_step(_p(slab), _p(tape), _p(out), _p(flags), _p(nnz), B, S, shift, _stream()), "tg_step")
    return out, flags, nnz


# ---------------------------------------------------------------- K3
def _cat_arrays(values, probs):
    import numpy as np

    v = np.ascontiguousarray(values, dtype=np.int8)
    p = np.ascontiguousarray(probs, dtype=np.float64)
    if v.ndim != 1 or v.shape != p.shape or not 1 <= len(v) <= 8:
        raise TensorGameError("values/probs must be 1-D of equal length <= 8")
    return v, p


def make_synthetic_demos(n_demos: int, max_actions: int, S: int, values=(-1, 0, 1), probs=(0.15, 0.7, 0.15),
                         shift: int = 1, seed: int = 0, first_demo: int = 0, device="cuda", max_tries: int = 64,
                         tape: torch.Tensor | None = None, slab: torch.Tensor | None = None):
    """Throughput-mode demo generation (device Philox stream, tg_demo_gen_philox).

    Returns (tape uint8 (R, N, TP) step-major, slab int8 (N, GP), flags uint8 (N,)).
    Same distribution and rejection rule as utils.py:203-233; NOT the torch RNG stream --
    use demos_from_seed for same-seed parity with the reference."""
    lay = layout(S)
    dev = torch.device(device)
    if dev.type != "cuda":
        raise TensorGameError("make_synthetic_demos needs a CUDA device (there is no CPU path)")
    v, p = _cat_arrays(values, probs)
    if tape is None:
        tape = torch.empty((max_actions, n_demos, lay.token_pitch), dtype=torch.uint8, device=dev)
    if slab is None:
        slab = torch.empty((n_demos, lay.game_pitch), dtype=torch.int8, device=dev)
    flags = torch.empty(n_demos, dtype=torch.uint8, device=dev)
    stride = tape.stride(0) if max_actions > 1 else n_demos * lay.token_pitch
    with torch.cuda.device(dev):
        check(_lib.lib().tg_demo_gen_philox(seed, first_demo, n_demos, max_actions, S, shift, v.ctypes.data, p.ctypes.data,
                                            len(v), max_tries, _p(tape), stride, _p(slab), _p(flags), _stream()),
              "tg_demo_gen_philox")
    return tape, slab, flags


def accumulate_demos(tape: torch.Tensor, S: int, shift: int, slab: torch.Tensor | None = None):
    """slab[n] = sum_r rank1(tape[r, n]) (tg_demo_accumulate).  Returns (slab, flags)."""
    _need_cuda(tape, "tape", torch.uint8)
    lay = layout(S)
    R, N = tape.shape[0], tape.shape[1]
    if tape.shape[2] != lay.token_pitch:
        raise TensorGameError(f"tape must be (R, N, {lay.token_pitch})")
    if slab is None:
        slab = torch.empty((N, lay.game_pitch), dtype=torch.int8, device=tape.device)
    flags = torch.empty(N, dtype=torch.uint8, device=tape.device)
    check(_lib.lib().tg_demo_accumulate(_p(tape), N * lay.token_pitch, N, R, S, shift, _p(slab), _p(flags), _stream()),
          "tg_demo_accumulate")
    return slab, flags


class HostStepper:
    """End-to-end step for callers whose data lives in host memory (pinned
    numpy/torch CPU buffers): chunked H2D -> tg_step -> D2H over three streams
    inside the C library (tg_step_host)."""

    def __init__(self, S: int, device: int = 0, chunk: int = 1 << 16):
        self.S, self.lay = S, layout(S)
        self._ctx = C.c_void_p()
        check(_lib.lib().tg_host_ctx_create(C.byref(self._ctx), device, S, chunk), "tg_host_ctx_create")

    def step(self, slab: torch.Tensor, tape: torch.Tensor, out: torch.Tensor, flags: torch.Tensor, nnz: torch.Tensor,
             shift: int) -> None:
        for t in (slab, tape, out, flags, nnz):
            if t.is_cuda or not t.is_contiguous():
                raise TensorGameError("HostStepper takes contiguous CPU tensors")
        B = slab.shape[0]
        check(_lib.lib().tg_step_host(self._ctx, _p(slab), _p(tape), _p(out), _p(flags), _p(nnz), B, shift), "tg_step_host")

    def close(self) -> None:
        if self._ctx:
            _lib.lib().tg_host_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
