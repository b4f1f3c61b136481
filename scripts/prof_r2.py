"""One launch set of a round-2 kernel at bench size (for ncu): python scripts/prof_r2.py sample|basis9|basis16|basis4|demo9|demo4|demo16|acc9|acc4|step|rollout|rank|expand"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mat_mul_b200 import env
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
which = sys.argv[1]
if which in ("sample", "sample4"):
    S, R, N = (9, 23, 1 << 18) if which == "sample" else (4, 7, 1 << 20)
    vals, probs, shift = (V5, P5, 2) if S == 9 else ((-1, 0, 1), (0.15, 0.7, 0.15), 1)
    tape, slab, _ = env.make_synthetic_demos(N, R, S, vals, probs, shift, seed=1)
    store = env.DemoStore.from_tape(tape, slab, S, shift)
    idx = torch.randint(0, N * R, (1 << (16 if S == 9 else 18),), device="cuda")
    for _ in range(3):
        store.samples(idx, 2, replay_shift=shift)
elif which.startswith("basis"):
    S = int(which[5:])
    R, N = {4: (7, 1 << 20), 9: (23, 1 << 18), 16: (49, 1 << 17)}[S]
    _, slab, _ = env.make_synthetic_demos(N, R, S, V5, P5, 2, seed=1)
    mats = env.sample_unimodular(N, S, seed=3, p_nonzero=0.3)
    out = torch.empty(slab.shape, dtype=torch.int16, device="cuda")
    for _ in range(3):
        env.change_of_basis(slab, mats, S, out=out)
elif which.startswith("acc"):
    S = int(which[3:])
    R, N = {4: (7, 1 << 22), 9: (23, 1 << 20), 16: (49, 1 << 17)}[S]
    vals, probs, shift = ((-1, 0, 1), (0.15, 0.7, 0.15), 1) if S == 4 else (V5, P5, 2)
    tape, slab, _ = env.make_synthetic_demos(N, R, S, vals, probs, shift, seed=1)
    out = torch.empty_like(slab)
    for _ in range(3):
        env.accumulate_demos(tape, S, shift, slab=out)
elif which.startswith("demo"):
    S = int(which[4:])
    R, N = {4: (7, 1 << 22), 9: (23, 1 << 20), 16: (49, 1 << 17)}[S]
    vals, probs, shift = ((-1, 0, 1), (0.15, 0.7, 0.15), 1) if S == 4 else (V5, P5, 2)
    lay = env.layout(S)
    tape = torch.empty((R, N, lay.token_pitch), dtype=torch.uint8, device="cuda")
    slab = torch.empty((N, lay.game_pitch), dtype=torch.int8, device="cuda")
    for _ in range(3):
        env.make_synthetic_demos(N, R, S, vals, probs, shift, seed=1, tape=tape, slab=slab)
elif which in ("step", "rollout", "rank", "expand"):
    S, R, N = 9, 23, 1 << 20
    tape, slab, _ = env.make_synthetic_demos(N, R, S, V5, P5, 2, seed=1)
    out = torch.empty_like(slab)
    if which == "step":
        fl, nz = torch.empty(N, dtype=torch.uint8, device="cuda"), torch.empty(N, dtype=torch.int32, device="cuda")
        for r in range(3):
            env.step_batch(slab, tape[R - 1 - r], S, 2, out=out, flags=fl, nnz=nz)
    elif which == "rollout":
        rev = tape.flip(0).contiguous()
        for _ in range(3):
            env.rollout(slab, rev, S, 2, out=out)
    elif which == "rank":
        for _ in range(3):
            env.slice_rank(slab[: 1 << 17], S)
    else:
        tb = tape[:8, : 1 << 17].permute(1, 0, 2).contiguous()
        for _ in range(3):
            env.expand_children(slab[: 1 << 17], tb, S, 2)
torch.cuda.synchronize()
print("ok")
