import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mat_mul_b200 import _lib, env
from mat_mul_b200.env import _p, _stream
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
S, R, N = 16, 49, 1 << 15
lay = env.layout(S)
tape, slab, fl = env.make_synthetic_demos(N, R, S, V5, P5, 2, seed=1)
out = torch.empty_like(slab); flags = torch.empty(N, dtype=torch.uint8, device="cuda")
for _ in range(3):
    _lib.lib().tg_demo_accumulate_tc(_p(tape), N * lay.token_pitch, N, R, S, 2, _p(out), _p(flags), _stream())
torch.cuda.synchronize()
print("ok")
