import sys; sys.path.insert(0,'/root/repo')
import torch, numpy as np
from mat_mul_b200 import env
from tests.helpers import slab_to_dense
V5, P5 = (-2, -1, 0, 1, 2), (0.05, 0.10, 0.70, 0.10, 0.05)
for S,R,p,LMAX in [(9,23,0.08,511),(16,49,0.03,32767),(4,7,0.3,32767)]:
    N=4096
    vals,probs,shift = ((-1,0,1),(0.15,0.7,0.15),1) if S==4 else (V5,P5,2)
    tape,slab,fl = env.make_synthetic_demos(N,R,S,vals,probs,shift,seed=1)
    mats = env.sample_unimodular(N,S,seed=3,p_nonzero=p)
    T = slab_to_dense(slab.cpu().numpy(),S).astype(np.int64); m = mats.cpu().numpy().astype(np.int64)
    Y = np.einsum('nkc,nabc->nabk', m[:,2], T)
    mx = np.abs(Y).reshape(N,-1).max(1); nA = np.abs(m[:,0]).sum(2).max(1); nB = np.abs(m[:,1]).sum(2).max(1)
    ok = (mx*nA<=LMAX)&(mx*nA*nB<=LMAX)
    print(S, 'fast fraction', ok.mean(), 'mx', mx.mean(), mx.max(), 'nA', nA.mean(), nA.max(), 'prod mean', (mx*nA*nB).mean())
    want = np.einsum("nia,njb,nkc,nabc->nijk", m[:, 0], m[:, 1], m[:, 2], T)
    print('   final max', np.abs(want).reshape(N,-1).max(1).mean(), ' in int8 frac', (np.abs(want).reshape(N,-1).max(1)<=127).mean())
