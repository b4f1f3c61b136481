import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mat_mul_b200 import env
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
S, R, N = 16, 49, 1 << 16
tape, slab, _ = env.make_synthetic_demos(N, R, S, V5, P5, 2, seed=3)
mats = env.sample_unimodular(N, S, seed=5, p_nonzero=0.03)
for _ in range(3):
    env.change_of_basis(slab, mats, S, tape=tape, shift=2, shift_out=100)
torch.cuda.synchronize()
print("ok")
