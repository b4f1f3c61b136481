"""One launch each of the K3 kernels at bench size (for ncu): python scripts/prof_demo.py [S]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from mat_mul_b200 import env

V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 9
R, N = {4: (7, 1 << 20), 9: (23, 1 << 18), 16: (49, 1 << 15)}[S]
vals, probs, shift = ((-1, 0, 1), (0.15, 0.7, 0.15), 1) if S == 4 else (V5, P5, 2)
lay = env.layout(S)
tape = torch.empty((R, N, lay.token_pitch), dtype=torch.uint8, device="cuda")
slab = torch.empty((N, lay.game_pitch), dtype=torch.int8, device="cuda")
for _ in range(2):
    env.make_synthetic_demos(N, R, S, vals, probs, shift, seed=1, tape=tape, slab=slab)
    env.accumulate_demos(tape, S, shift, slab=slab)
torch.cuda.synchronize()
print("ok")
