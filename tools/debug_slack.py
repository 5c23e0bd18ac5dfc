import sys, numpy as np
sys.path.insert(0, '.')
from tests.helpers import D, to_arrays, from_arrays, bits_equal
from oracle import oracle as O
from examples_b200.distributed import MgpuGroup
from examples_b200 import Field2D
W=float(np.float32(np.float32(5376)*np.float32(D))); H=float(np.float32(np.float32(480)*np.float32(D)))
print('W',W,'H',H)
rng=np.random.default_rng(3)
n=int(0.1*W*H)
b=np.zeros(n,dtype=O.BIRD); b['id']=np.arange(n)
b['x']=(rng.random(n)*W).astype(np.float32); b['y']=(rng.random(n)*H).astype(np.float32)
b['x']=np.minimum(b['x'],np.nextafter(np.float32(W),np.float32(0))); b['y']=np.minimum(b['y'],np.nextafter(np.float32(H),np.float32(0)))
for nstrips in (8,):
    g=MgpuGroup(W,H,D,capacity_per_rank=int(n/nstrips*1.3)+50000,devices=[0]*nstrips)
    g.scatter_objects(*to_arrays(b)); g.commit()
    f=Field2D(W,H,D,True,capacity=n); f.set_objects(*to_arrays(b)); f.lazy_update()
    for t0 in range(0,60,10):
        g.run(t0,10); f.run(t0,10)
        owned=g.num_owned()
        fi,fl,fd=f.snapshot_by_id()
        atW=(fl[:,0]==np.float32(W)).sum(); atH=(fl[:,1]==np.float32(H)).sum()
        gi,gl,gd=g.snapshot_by_id()
        same = len(gi)==n and bits_equal(from_arrays(gi,gl,gd), from_arrays(fi,fl,fd))
        print('strips',nstrips,'t',t0+10,'owned sum',sum(owned),'of',n,'x==W now',atW,'y==H now',atH,'same',same)
        if len(gi)!=n:
            missing=np.setdiff1d(np.arange(n),gi)
            print('  missing',missing[:10],'their single-field state',[(fl[m].tolist(),fd[m].tolist()) for m in missing[:5]])
            break
