"""One launch of the change-of-basis kernels at bench size (for ncu): python scripts/prof_basis.py [S]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from mat_mul_b200 import env

V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 9
R, N, p = {4: (7, 1 << 20, 0.3), 9: (23, 1 << 18, 0.08), 16: (49, 1 << 15, 0.03)}[S]
vals, probs, shift = ((-1, 0, 1), (0.15, 0.7, 0.15), 1) if S == 4 else (V5, P5, 2)
tape, slab, _ = env.make_synthetic_demos(N, R, S, vals, probs, shift, seed=1)
mats = env.sample_unimodular(N, S, seed=3, p_nonzero=p)
for _ in range(3):
    out, flags = env.change_of_basis(slab, mats, S)
torch.cuda.synchronize()
print("ok")
