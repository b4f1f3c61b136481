"""CUDA-event timings of the 4x4x4 kernels on batches large enough that the launch overhead of the Python wrapper does not
hide them: leaf expansion (2^20 parents x 8), rollout, step, sample batcher."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mat_mul_b200 import env
PEAK = 6549.1

def t_ms(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

S, R, N, shift = 4, 7, 1 << 22, 1
tape, slab, _ = env.make_synthetic_demos(N, R, S, (-1, 0, 1), (0.15, 0.7, 0.15), shift, seed=1)
nbp = 1 << 20
tb = tape[:, :nbp].permute(1, 0, 2).contiguous()  # (parents, k = 7, TP)
k = tb.shape[1]
moved = 64 * (1 + 1 / k) + 16 + 13
for wk in (False, True):
    ms = t_ms(lambda: env.expand_children(slab[:nbp], tb, S, shift, with_keys=wk))
    print(f"S=4 expand_children k={k} keys={wk}: {ms:.3f} ms {nbp * k / ms / 1e6:.2f} G children/s hbm_frac={nbp * k * moved / ms / 1e6 / PEAK:.3f}")
out = torch.empty_like(slab)
rev = tape.flip(0).contiguous()
ms = t_ms(lambda: env.rollout(slab, rev, S, shift, out=out))
print(f"S=4 rollout K={R}: {ms:.3f} ms {N * R / ms / 1e6:.1f} G game-steps/s hbm_frac={N * (128 + R * 12 + 8) / ms / 1e6 / PEAK:.3f}")
fl, nz = torch.empty(N, dtype=torch.uint8, device="cuda"), torch.empty(N, dtype=torch.int32, device="cuda")
ms = t_ms(lambda: env.step_batch(slab, tape[R - 1], S, shift, out=out, flags=fl, nnz=nz))
print(f"S=4 step: {ms:.3f} ms {N / ms / 1e6:.1f} G steps/s hbm_frac={N * 145 / ms / 1e6 / PEAK:.3f}")
