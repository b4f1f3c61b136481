"""One launch of the K4 sample batcher at bench size (for ncu)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mat_mul_b200 import env
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
S, R, N = 9, 23, 1 << 18
tape, slab, _ = env.make_synthetic_demos(N, R, S, V5, P5, 2, seed=1)
idx = torch.randint(0, N * R, (1 << 16,), device="cuda")
for _ in range(3):
    env.demo_samples(tape, slab, idx, S, 2, replay_shift=2)
torch.cuda.synchronize()
print("ok")
