import sys; sys.path.insert(0,'/root/repo')
import torch
from mat_mul_b200 import env
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
for S,R,N in [(4,7,1<<22),(9,23,1<<20),(16,49,1<<16)]:
    vals,probs,shift = ((-1,0,1),(0.15,0.7,0.15),1) if S==4 else (V5,P5,2)
    tape,slab,_=env.make_synthetic_demos(N,R,S,vals,probs,shift,seed=1)
    for _ in range(3): env.state_keys(slab,S)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): env.state_keys(slab,S)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    nz=(slab!=0).float().mean().item()
    print(f"S={S} state_keys: {ms:.3f} ms {N/ms/1e6:.3f} G games/s  nonzero fraction {nz:.3f}")
