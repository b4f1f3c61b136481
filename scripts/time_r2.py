"""CUDA-event timings of the round-2 kernels at bench sizes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mat_mul_b200 import env
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
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

for S, R, N, vals, probs, shift in [(4, 7, 1 << 22, (-1, 0, 1), (0.15, 0.7, 0.15), 1), (9, 23, 1 << 20, V5, P5, 2)]:
    lay = env.layout(S)
    tape = torch.empty((R, N, lay.token_pitch), dtype=torch.uint8, device="cuda")
    slab = torch.empty((N, lay.game_pitch), dtype=torch.int8, device="cuda")
    ms = t_ms(lambda: env.make_synthetic_demos(N, R, S, vals, probs, shift, seed=1, tape=tape, slab=slab))
    print(f"S={S} demo_gen R={R} N={N}: {ms:.3f} ms {N / ms / 1e6:.3f} G demos/s hbm_frac={(N * (S**3 + R * 3 * S)) / ms / 1e6 / PEAK:.3f}")
    store = env.DemoStore.from_tape(tape, slab, S, shift)
    for nb in (1 << 16, 1 << 18):
        idx = torch.randint(0, N * R, (nb,), device="cuda")
        for T in (2, 4):
            ms = t_ms(lambda: store.samples(idx, T, replay_shift=shift))
            print(f"S={S} demo_sample nb={nb} T={T}: {ms:.3f} ms {nb / ms / 1e6:.4f} G samples/s hbm_frac={nb * T * S**3 * 4 / ms / 1e6 / PEAK:.3f}")
    nbg = min(N, 1 << 18 if S == 9 else 1 << 20)
    mats = env.sample_unimodular(nbg, S, seed=3, p_nonzero=0.3)
    for dt in (torch.int16, torch.int8):
        out = torch.empty((nbg, lay.game_pitch), dtype=dt, device="cuda")
        ms = t_ms(lambda: env.change_of_basis(slab[:nbg], mats, S, out=out))
        by = S**3 + 3 * S * S + (2 if dt == torch.int16 else 1) * S**3
        print(f"S={S} change_of_basis p=0.3 out={dt}: {ms:.3f} ms {nbg / ms / 1e6:.4f} G games/s hbm_frac={nbg * by / ms / 1e6 / PEAK:.3f}")
