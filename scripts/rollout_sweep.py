"""BASELINE.json config 5: rollout-heavy sweep, 9x9x9, K = 64 steps, throughput vs batch size (one GPU).

Every game replays its own rank-23 demo in reverse and is then fed null actions up to K = 64 (the padding SURVEY 8(d)
describes): fused K-step kernel (tg_rollout, games freeze when solved) against K single-step launches (tg_step).
Also BASELINE config 3 at scale: 2^18 demos of 16x16x16 with rank <= 49 followed by the change of basis."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from mat_mul_b200 import env

V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
PEAK = 6549.1


def t_ms(fn, n=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


S, R, K, shift = 9, 23, 64, 2
lay = env.layout(S)
print(f"rollout sweep {S}x{S}x{S}, K={K} (R={R} real actions + null padding), one B200")
print(f"{'games':>10} {'fused ms':>10} {'fused Gsteps/s':>15} {'hbm_frac':>9} {'per-step ms':>12} {'per-step Gsteps/s':>18} {'solved':>8}")
LBS = tuple(int(x) for x in sys.argv[1].split(",")) if len(sys.argv) > 1 else (14, 16, 18, 20, 22, 24)
for lb in LBS:
    B = 1 << lb
    tape, slab, _ = env.make_synthetic_demos(B, R, S, V5, P5, shift, seed=lb)
    tapeK = torch.empty((K, B, lay.token_pitch), dtype=torch.uint8, device="cuda")
    tapeK[:R] = tape.flip(0)
    tapeK[R:] = 0
    tapeK[R:, :, : 3 * S] = shift
    del tape
    out = torch.empty_like(slab)
    ms_f = t_ms(lambda: env.rollout(slab, tapeK, S, shift, out=out))
    _, fl, _, steps = env.rollout(slab, tapeK, S, shift, out=out)
    solved = int((fl & 1).sum())
    a, b = slab.clone(), torch.empty_like(slab)
    flags = torch.empty(B, dtype=torch.uint8, device="cuda")
    nnz = torch.empty(B, dtype=torch.int32, device="cuda")

    def per_step():
        src, dst = a, b
        for k in range(K):
            env.step_batch(src, tapeK[k], S, shift, out=dst, flags=flags, nnz=nnz)
            src, dst = dst, src

    ms_s = t_ms(per_step, n=1 if lb >= 22 else 2, warm=1)
    algo = 2 * S**3 + K * 3 * S + 8
    print(f"{B:>10} {ms_f:>10.3f} {B * K / ms_f / 1e6:>15.2f} {B * algo / ms_f / 1e6 / PEAK:>9.3f} {ms_s:>12.3f} {B * K / ms_s / 1e6:>18.2f} {solved:>8}")
    del tapeK, slab, out, a, b

S, R, N, shift = 16, 49, 1 << 18, 2
lay = env.layout(S)
tape, slab, _ = env.make_synthetic_demos(N, R, S, V5, P5, shift, seed=3)
rank = torch.randint(1, R + 1, (N,), device="cuda")
mask = torch.arange(R, device="cuda").view(R, 1) >= rank.view(1, N)          # (R, N): term r of demo n is padding
null = torch.zeros(lay.token_pitch, dtype=torch.uint8, device="cuda")
null[: 3 * S] = shift
tape[mask] = null
ms_acc = t_ms(lambda: env.accumulate_demos(tape, S, shift, slab=slab))
mats = env.sample_unimodular(N, S, seed=5, p_nonzero=0.03)
ms_cb = t_ms(lambda: env.change_of_basis(slab, mats, S, tape=tape, shift=shift, shift_out=100))
print(f"config 3: {N} demos 16x16x16, rank uniform in [1,{R}]: accumulate {ms_acc:.3f} ms ({N / ms_acc / 1e6:.3f} G demos/s), "
      f"change of basis tensor+factors {ms_cb:.3f} ms ({N / ms_cb / 1e6:.3f} G demos/s)")
