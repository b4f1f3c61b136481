"""CUDA-event timing of K4 demo_sample (training-sample batcher) at the bench sizes.  python scripts/time_sample.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mat_mul_b200 import env

torch.manual_seed(0)
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
ONLY = int(os.environ.get('TS_ONLY', '0'))
for S, R, N, vals, probs, shift in [(4, 7, 1 << 20, (-1, 0, 1), (0.15, 0.7, 0.15), 1), (9, 23, 1 << 18, V5, P5, 2), (16, 49, 1 << 15, V5, P5, 2)]:
    if ONLY and S != ONLY:
        continue
    tape, slab, _ = env.make_synthetic_demos(N, R, S, vals, probs, shift, seed=1)
    for T in (2, 4):
        idx = torch.randint(0, N * R, (1 << 16,), device="cuda")
        for _ in range(3):
            out = env.demo_samples(tape, slab, idx, S, T, replay_shift=shift)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = env.demo_samples(tape, slab, idx, S, T, replay_shift=shift)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nbytes = idx.numel() * T * S**3 * 4
        print(f"S={S} R={R} demo_sample T={T}: {ms:.3f} ms {idx.numel() / ms / 1e6:.4f} G samples/s  write {nbytes / ms / 1e6:.0f} GB/s = {nbytes / ms / 1e6 / 6549.1:.3f} of HBM peak "
              f"checksum {float(out[0].sum()):.0f}")
