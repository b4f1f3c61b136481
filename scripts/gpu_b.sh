timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
for v in 0 1 2 3; do for c in 8 32; do python bench.py --steps 20 --no-e2e --no-cpu --size 4 --games 8388608 --variant $v --ctas-per-sm $c | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('S4 v$v ctas', $c, round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"; done; done
for v in 0 1 2 3; do for c in 8 32; do python bench.py --steps 20 --no-e2e --no-cpu --size 16 --games 262144 --variant $v --ctas-per-sm $c | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('S16 v$v ctas', $c, round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"; done; done
python bench.py --steps 20 --no-e2e --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('S9 default', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
python - <<'PY'
import torch, time
from mat_mul_b200 import env
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
for S,R,N,vals,probs,shift in [(4,7,1<<22,(-1,0,1),(0.15,0.7,0.15),1),(9,23,1<<20,V5,P5,2),(16,49,1<<16,V5,P5,2)]:
    lay=env.layout(S)
    tape=torch.empty((R,N,lay.token_pitch),dtype=torch.uint8,device='cuda'); slab=torch.empty((N,lay.game_pitch),dtype=torch.int8,device='cuda')
    for _ in range(2): env.make_synthetic_demos(N,R,S,vals,probs,shift,seed=1,tape=tape,slab=slab)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): env.make_synthetic_demos(N,R,S,vals,probs,shift,seed=1,tape=tape,slab=slab)
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/5
    print(f"demo S={S} R={R} N={N}: {ms:.3f} ms  {N/ms*1e3/1e9:.3f} G demos/s  write {(N*(S**3+R*3*S))/ms/1e6:.1f} GB/s algorithmic")
    e0.record()
    for _ in range(5): env.accumulate_demos(tape,S,shift,slab=slab)
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/5
    print(f"accumulate S={S} R={R}: {ms:.3f} ms  {N/ms*1e3/1e9:.3f} G demos/s")
PY
