"""CUDA-event timing of the S = 16 change of basis (K5): tensor-core kernel (default) vs the packed-IMAD kernel
(TG_BASIS_VARIANT=1), same inputs.  Run under gpurun:  python scripts/time_basis16.py [log2 N] [p_nonzero]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mat_mul_b200 import _lib, env
from mat_mul_b200.env import _p, _stream

S, R, shift = 16, 49, 2
N = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 18)
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.03
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
tape, slab, _ = env.make_synthetic_demos(N, R, S, V5, P5, shift, seed=1)
del tape
mats = env.sample_unimodular(N, S, seed=3, p_nonzero=p)
out = torch.empty_like(slab)
flags = torch.zeros(N, dtype=torch.uint8, device="cuda")
L = _lib.lib()


def call():
    _lib.check(L.tg_change_of_basis(_p(slab), _p(mats), 1, _p(out), _p(flags), N, S, _stream()), "tg_change_of_basis")


for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
tot = 0.0
for _ in range(10):
    e0.record()
    call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best, tot = min(best, ms), tot + ms
ms = tot / 10
moved = N * (2 * S**3 + 3 * S * S)
print(f"variant={os.environ.get('TG_BASIS_VARIANT', '0')} S=16 N={N} p={p}: mean {ms:.3f} ms (best {best:.3f}) "
      f"{N / ms / 1e6:.4f} G games/s  moved {moved / ms / 1e6:.0f} GB/s = {moved / ms / 1e6 / 6549.1:.3f} of HBM peak; "
      f"{N * (3 * S**3 + 3 * S * S) / ms / 1e6 / 6549.1:.3f} on SURVEY 8(d) bytes; range-flagged {int((flags & 4).ne(0).sum())}")
print("checksum", int(out.to(torch.int64).sum()), int(flags.to(torch.int64).sum()))
