import sys, numpy as np
sys.path.insert(0, '.')
from tests.helpers import D, uniform_birds, to_arrays
from examples_b200.distributed import MgpuGroup
from examples_b200 import Field2D
W=500.0; n=20000
birds=uniform_birds(n,W,W)
g=MgpuGroup(W,W,D,capacity_per_rank=n,devices=[0]*4)
g.set_objects(*to_arrays(birds)); g.commit()
f=Field2D(W,W,D,True,capacity=n); f.set_objects(*to_arrays(birds)); f.lazy_update()
for q in g.fields:
    ids,loc,ld=q.snapshot_owned(); ai,al,ad=q.snapshot()
    print('after commit rank',q.rank,'owned',len(ids),'zero ids owned',(ids==0).sum(),'all',len(ai),'zero ids all',(ai==0).sum())
g.run(0,1); f.run(0,1)
fi,fl,fd=f.snapshot_by_id()
pos2id={ (float(x),float(y)):int(i) for i,(x,y) in zip(fi,fl)}
for q in g.fields:
    ids,loc,ld=q.snapshot_owned()
    z=np.nonzero(ids==0)[0]
    cols=np.floor(loc[z,0]/np.float32(D)).astype(int)
    true=[pos2id.get((float(x),float(y)),-1) for x,y in loc[z]]
    print('rank',q.rank,q.strip(),'zero-id entries',len(z),'cols',sorted(set(cols.tolist())),'true ids',true[:8])
    old=birds[[t for t in true if t>0]]
    print('   came from cols', sorted(set(np.floor(old['x']/np.float32(D)).astype(int).tolist())))
