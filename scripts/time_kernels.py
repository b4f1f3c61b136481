"""CUDA-event timings of the non-headline kernels at bench sizes (per GPU)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from mat_mul_b200 import env

V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)


def t_ms(fn, n=5):
    for _ in range(3):  # the first launches of a kernel carry module-load / allocator latency on the host side
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for S, R, N, vals, probs, shift, p in [(4, 7, 1 << 22, (-1, 0, 1), (0.15, 0.7, 0.15), 1, 0.3), (9, 23, 1 << 20, V5, P5, 2, 0.08),
                                       (16, 49, 1 << 16, V5, P5, 2, 0.03)]:
    lay = env.layout(S)
    tape = torch.empty((R, N, lay.token_pitch), dtype=torch.uint8, device="cuda")
    slab = torch.empty((N, lay.game_pitch), dtype=torch.int8, device="cuda")
    out = torch.empty_like(slab)
    ms = t_ms(lambda: env.make_synthetic_demos(N, R, S, vals, probs, shift, seed=1, tape=tape, slab=slab))
    print(f"S={S} demo_gen R={R} N={N}: {ms:.3f} ms {N / ms / 1e6:.3f} G demos/s  hbm_frac={(N * (S**3 + R * 3 * S)) / ms / 1e6 / 6549.1:.3f}")
    ms = t_ms(lambda: env.accumulate_demos(tape, S, shift, slab=out))
    print(f"S={S} accumulate: {ms:.3f} ms {N / ms / 1e6:.3f} G demos/s")
    rev = tape.flip(0).contiguous()
    ms = t_ms(lambda: env.rollout(slab, rev, S, shift, out=out))
    print(f"S={S} rollout K={R}: {ms:.3f} ms {N * R / ms / 1e6:.3f} G game-steps/s")
    nb = min(N, 1 << 18)
    mats = env.sample_unimodular(nb, S, seed=3, p_nonzero=p)
    ms = t_ms(lambda: env.change_of_basis(slab[:nb], mats, S))
    print(f"S={S} change_of_basis: {ms:.3f} ms {nb / ms / 1e6:.4f} G games/s hbm_frac={(nb * (2 * S**3 + 3 * S * S)) / ms / 1e6 / 6549.1:.3f} "
          f"(int8 out; {(nb * (3 * S**3 + 3 * S * S)) / ms / 1e6 / 6549.1:.3f} on SURVEY 8(d)'s int16-out bytes)")
    idx = torch.randint(0, N * R, (1 << 16,), device="cuda")
    ms = t_ms(lambda: env.demo_samples(tape, slab, idx, S, 2, replay_shift=shift))
    print(f"S={S} demo_sample T=2: {ms:.3f} ms {idx.numel() / ms / 1e6:.4f} G samples/s")
    nbp = min(N, 1 << 17)
    tape_bk = tape[:8, :nbp].permute(1, 0, 2).contiguous()  # (B, k=8, TP): the first 8 actions of each demo as candidates
    ms = t_ms(lambda: env.expand_children(slab[:nbp], tape_bk, S, shift))
    lay_b = lay.game_pitch * (1 + 1 / 8) + lay.token_pitch + 13
    print(f"S={S} expand_children k=8: {ms:.3f} ms {nbp * 8 / ms / 1e6:.3f} G children/s hbm_frac={nbp * 8 * lay_b / ms / 1e6 / 6549.1:.3f} (moved bytes)")
    ms = t_ms(lambda: env.expand_children(slab[:nbp], tape_bk, S, shift, with_keys=False))
    print(f"S={S} expand_children k=8 (no keys): {ms:.3f} ms {nbp * 8 / ms / 1e6:.3f} G children/s hbm_frac={nbp * 8 * (lay_b - 8) / ms / 1e6 / 6549.1:.3f}")
    del tape, slab, out, rev
