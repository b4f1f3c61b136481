"""Launch each non-headline kernel a few times on bench-sized inputs (for ncu captures; see profiles/README.md)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from mat_mul_b200 import env

S, R, shift, B = 9, 23, 2, 1 << 18
V, P = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
for _ in range(2):
    tape, slab, flags = env.make_synthetic_demos(B, R, S, V, P, shift, seed=1)
    env.accumulate_demos(tape, S, shift)
    rev = tape.flip(0).contiguous()
    env.rollout(slab, rev, S, shift)
    mats = env.sample_unimodular(B, S, seed=3, p_nonzero=0.08)
    env.change_of_basis(slab, mats, S)
    idx = torch.randint(0, B * R, (1 << 16,), device="cuda")
    env.demo_samples(tape, slab, idx, S, 2, replay_shift=shift)
torch.cuda.synchronize()
print("ok")
