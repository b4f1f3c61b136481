"""CUDA-event timing of 16x16x16 demo generation and accumulation: f16 tensor-core kernel (default) vs the packed-IMAD
kernel (TG_DEMO_MMA=0) vs the tcgen05 kernel.  Run under gpurun:  python scripts/time_demo16.py [log2 N] [R]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mat_mul_b200 import env

S, shift = 16, 2
N = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 17)
R = int(sys.argv[2]) if len(sys.argv) > 2 else 49
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
lay = env.layout(S)
tape = torch.empty((R, N, lay.token_pitch), dtype=torch.uint8, device="cuda")
slab = torch.empty((N, lay.game_pitch), dtype=torch.int8, device="cuda")


def t_ms(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


tag = f"TG_DEMO_MMA={os.environ.get('TG_DEMO_MMA', '1')} S=16 R={R} N={N}"
ms = t_ms(lambda: env.make_synthetic_demos(N, R, S, V5, P5, shift, seed=1, tape=tape, slab=slab))
algo = S**3 + R * 3 * S
print(f"{tag} demo_gen: {ms:.3f} ms {N / ms / 1e6:.4f} G demos/s hbm_frac={N * algo / ms / 1e6 / 6549.1:.3f}")
chk = int(slab.to(torch.int64).sum()), int(tape.to(torch.int64).sum())
out = torch.empty_like(slab)
ms = t_ms(lambda: env.accumulate_demos(tape, S, shift, slab=out))
print(f"{tag} accumulate: {ms:.3f} ms {N / ms / 1e6:.4f} G demos/s moved {N * (algo + 0) / ms / 1e6:.0f} GB/s = {N * algo / ms / 1e6 / 6549.1:.3f} of HBM peak")
assert torch.equal(out, slab)
if R <= 64:
    ms = t_ms(lambda: env.accumulate_demos_tc(tape, shift, slab=out))
    print(f"{tag} accumulate (tcgen05 kernel): {ms:.3f} ms {N / ms / 1e6:.4f} G demos/s")
print("checksum", chk)
