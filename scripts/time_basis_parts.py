import sys; sys.path.insert(0,'/root/repo')
import torch
from mat_mul_b200 import env, _lib
from mat_mul_b200.env import _p, _stream
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
def t_ms(fn, n=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for S,R,N,p in [(16,49,1<<18,0.03),(9,23,1<<20,0.08)]:
    lay=env.layout(S)
    tape, slab, _ = env.make_synthetic_demos(N, R, S, V5, P5, 2, seed=3)
    mats = env.sample_unimodular(N, S, seed=5, p_nonzero=p)
    print(S, 'tensor only', t_ms(lambda: env.change_of_basis(slab, mats, S)))
    tape_out = torch.empty_like(tape); flags = torch.zeros(N, dtype=torch.uint8, device='cuda')
    L=_lib.lib()
    f = lambda: L.tg_change_of_basis_factors(_p(tape), N*lay.token_pitch, 2, _p(mats), 1, _p(tape_out), N*lay.token_pitch, 100, _p(flags), N, R, S, _stream())
    ms = t_ms(f)
    print(S, 'factors only', ms, 'GB/s', (2*R*N*lay.token_pitch + N*3*S*S)/ms/1e6)
    print(S, 'copy of the tape (torch)', t_ms(lambda: tape_out.copy_(tape)))
