"""tg_demo_accumulate_tc (tcgen05) against tg_demo_accumulate (packed IMAD) on the same tapes."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from mat_mul_b200 import _lib, env
from mat_mul_b200.env import _p, _stream, check

V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
S = 16
lay = env.layout(S)
for R, N in [(1, 5), (12, 300), (32, 64), (33, 700), (49, 1000), (64, 333)]:
    tape, slab, fl = env.make_synthetic_demos(N, R, S, V5, P5, 2, seed=R)
    out = torch.full_like(slab, 7)
    flags = torch.full((N,), 99, dtype=torch.uint8, device="cuda")
    check(_lib.lib().tg_demo_accumulate_tc(_p(tape), N * lay.token_pitch, N, R, S, 2, _p(out), _p(flags), _stream()), "tc")
    torch.cuda.synchronize()
    same = torch.equal(out, slab)
    print(R, N, "slab equal:", same, "flags equal:", torch.equal(flags, fl), flush=True)
    if not same:
        bad = (out != slab).nonzero()
        print("  first diffs", bad[:8].tolist(), out[bad[0, 0], bad[0, 1]].item(), slab[bad[0, 0], bad[0, 1]].item())
        a = out[bad[0, 0]].view(16, 256)[:2, :16].cpu(); b = slab[bad[0, 0]].view(16, 256)[:2, :16].cpu()
        print(a.tolist()); print(b.tolist())

def t_ms(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

R, N = 49, 1 << 16
tape, slab, fl = env.make_synthetic_demos(N, R, S, V5, P5, 2, seed=1)
out = torch.empty_like(slab); flags = torch.empty(N, dtype=torch.uint8, device="cuda")
L = _lib.lib()
ms_tc = t_ms(lambda: L.tg_demo_accumulate_tc(_p(tape), N * lay.token_pitch, N, R, S, 2, _p(out), _p(flags), _stream()))
ms_old = t_ms(lambda: L.tg_demo_accumulate(_p(tape), N * lay.token_pitch, N, R, S, 2, _p(out), _p(flags), _stream()))
print(f"S=16 R=49 N=65536 accumulate: tcgen05 {ms_tc:.3f} ms ({N / ms_tc / 1e6:.3f} G demos/s)   packed IMAD {ms_old:.3f} ms")
import ctypes, numpy as np
if hasattr(L, "tg_debug_tc"):
    buf = (ctypes.c_ulonglong * 16)()
    L.tg_debug_tc(buf, 1)
    L.tg_demo_accumulate_tc(_p(tape), N * lay.token_pitch, N, R, S, 2, _p(out), _p(flags), _stream())
    L.tg_debug_tc(buf, 0)
    names = ["mma:wait ready", "mma:wait tmfree", "mma:issue", "prod:wait opfree", "prod:expand", "prod:fence", "cons:wait full", "cons:ld"]
    nwarps = [1, 1, 1, 8, 8, 8, 8, 8]
    for n, k, b in zip(names, nwarps, buf):
        print(f"  {n:18s} {b / k / N:9.1f} cycles per demo per warp (x grid/SM concurrency)")
