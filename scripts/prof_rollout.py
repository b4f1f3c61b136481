"""One launch of the fused rollout at bench size (for ncu)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mat_mul_b200 import env
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
S, R, N = 9, 23, 1 << 18
tape, slab, _ = env.make_synthetic_demos(N, R, S, V5, P5, 2, seed=1)
rev = tape.flip(0).contiguous()
out = torch.empty_like(slab)
for _ in range(3):
    env.rollout(slab, rev, S, 2, out=out)
torch.cuda.synchronize()
print("ok")
