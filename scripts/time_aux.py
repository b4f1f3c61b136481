"""CUDA-event timings of the boundary / auxiliary kernels (conversions, unimodular sampler, slice rank, keys)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mat_mul_b200 import env
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
def t_ms(fn, n=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for S, R, N, p in [(4, 7, 1 << 20, 0.3), (9, 23, 1 << 18, 0.08), (16, 49, 1 << 16, 0.03)]:
    lay = env.layout(S)
    vals, probs, shift = ((-1, 0, 1), (0.15, 0.7, 0.15), 1) if S == 4 else (V5, P5, 2)
    tape, slab, _ = env.make_synthetic_demos(N, R, S, vals, probs, shift, seed=1)
    ms = t_ms(lambda: env.sample_unimodular(N, S, seed=3, p_nonzero=p))
    print(f"S={S} sample_unimodular: {ms:.3f} ms {N / ms / 1e6:.3f} G games/s  write {N * 3 * S * S / ms / 1e6:.0f} GB/s")
    f32 = env.expand_states(slab, S)
    ms = t_ms(lambda: env.expand_states(slab, S, out=f32))
    print(f"S={S} expand_states (slab -> f32): {ms:.3f} ms  {N * (lay.game_pitch + 4 * S**3) / ms / 1e6:.0f} GB/s")
    ms = t_ms(lambda: env.pack_states(f32, S))
    print(f"S={S} pack_states (f32 -> slab): {ms:.3f} ms  {N * (lay.game_pitch + 4 * S**3) / ms / 1e6:.0f} GB/s")
    acts = env.unpack_actions(tape[0], S)
    ms = t_ms(lambda: env.pack_actions(acts, S))
    print(f"S={S} pack_actions (i64 -> tape): {ms:.3f} ms  {N * (lay.token_pitch + 24 * S) / ms / 1e6:.0f} GB/s")
    ms = t_ms(lambda: env.slice_rank(slab, S))
    print(f"S={S} slice_rank: {ms:.3f} ms {N / ms / 1e6:.4f} G games/s")
    ms = t_ms(lambda: env.state_keys(slab, S))
    print(f"S={S} state_keys: {ms:.3f} ms {N / ms / 1e6:.3f} G games/s  read {N * lay.game_pitch / ms / 1e6:.0f} GB/s")
