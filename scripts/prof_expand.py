"""A few launches of the slab -> float32 boundary conversion at bench size (for ncu)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mat_mul_b200 import env
S, N = 9, 1 << 18
slab = (torch.randint(-3, 4, (N, env.layout(S).game_pitch), device="cuda", dtype=torch.int8))
out = env.expand_states(slab, S)
for _ in range(3):
    env.expand_states(slab, S, out=out)
torch.cuda.synchronize()
print("ok")
