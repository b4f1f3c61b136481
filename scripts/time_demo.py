"""CUDA-event timings of demo generation / accumulation (K3) at bench sizes: python scripts/time_demo.py [sizes, e.g. 9 or 4,9,16]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mat_mul_b200 import env
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
PEAK = 6549.1

def t_ms(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4, 9, 16]
CFG = {4: (7, 1 << 22, (-1, 0, 1), (0.15, 0.7, 0.15), 1), 9: (23, 1 << 20, V5, P5, 2), 16: (49, 1 << 17, V5, P5, 2)}
for S in sizes:
    R, N, vals, probs, shift = CFG[S]
    lay = env.layout(S)
    tape = torch.empty((R, N, lay.token_pitch), dtype=torch.uint8, device="cuda")
    slab = torch.empty((N, lay.game_pitch), dtype=torch.int8, device="cuda")
    ms = t_ms(lambda: env.make_synthetic_demos(N, R, S, vals, probs, shift, seed=1, tape=tape, slab=slab))
    print(f"S={S} demo_gen R={R} N={N}: {ms:.4f} ms {N / ms / 1e6:.3f} G demos/s hbm_frac={(N * (S**3 + R * 3 * S)) / ms / 1e6 / PEAK:.3f}")
    out = torch.empty_like(slab)
    ms = t_ms(lambda: env.accumulate_demos(tape, S, shift, slab=out))
    assert torch.equal(out, slab)
    print(f"S={S} demo_accumulate: {ms:.4f} ms {N / ms / 1e6:.3f} G demos/s")
