import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import grmp_b200 as G, oracle as O
L=int(sys.argv[1]) if len(sys.argv)>1 else 4
g=G.uniform_refine(G.grid_unitcube("Tetrahedron3D"),L); g=G.perturb_interior_nodes(g)
s=G.FESpace(G.H1P2(1,3),g)
A=O.OracleMatrix(s.ndofs,s.ndofs)
t=time.time(); O.blf_assemble(A,g,s,s,O.OP_GRAD,O.OP_GRAD,apt=O.APT_SYMMETRIC,factor=1.0); A.flush(); print('oracle',time.time()-t)
cp,rv,nz=A.csc(); cp=cp-1; rv=rv-1
nn=g.nnodes; ndofs=s.ndofs
# all entries (row=e edge dof, col=v vertex dof) and (row=w vertex, col=v vertex, w!=v) are mirrors; destination slot = position in column v
cols=np.repeat(np.arange(ndofs),np.diff(cp))
vmask=(cols<nn)&(rv!=cols)          # off-diagonal entries of vertex columns
slots=np.nonzero(vmask)[0]
rows=rv[slots]
print('mirrored entries',slots.size,'of nnz',rv.size)
# producing edge of each mirrored entry: edge rows -> that edge; vertex rows (w,v) -> edge (v,w): need edge index of the pair
# edge dof index e = nn + edge id; find edge (v,w) through the edge columns: column e has rows v_P, v_Q as its two smallest vertex rows? use celldofs
dofs=s.celldofs.astype(np.int64)-1
cn=g.cellnodes.astype(np.int64)-1
pairs=[(0,1),(0,2),(0,3),(1,2),(1,3),(2,3)]
import collections
ekey={}
for k,(a,b) in enumerate(pairs):
    lo=np.minimum(cn[:,a],cn[:,b]); hi=np.maximum(cn[:,a],cn[:,b])
    key=lo*nn+hi
    ed=dofs[:,4+k]
    for kk,e in zip(key[::1],ed[::1]):
        ekey[kk]=e
colv=cols[slots]
prod=np.where(rows>=nn, rows, -1)
vv=np.nonzero(rows<nn)[0]
lo=np.minimum(rows[vv],colv[vv]); hi=np.maximum(rows[vv],colv[vv])
prod[vv]=np.array([ekey[k] for k in lo*nn+hi])
edge_id=prod-nn
for TILE in (224, 448, 896, 1792):
    tile=edge_id//TILE
    # sectors: (tile, slot//4) unique count ; full sectors: count==4
    key=tile*(rv.size//4+1)+slots//4
    u,c=np.unique(key,return_counts=True)
    # runs: consecutive slots within same tile & column
    order=np.lexsort((slots,tile)); ts=tile[order]; ss=slots[order]
    newrun=np.ones(ss.size,bool); newrun[1:]=(ts[1:]!=ts[:-1])|(ss[1:]!=ss[:-1]+1)
    print('tile',TILE,'sectors',u.size,'values/sector %.2f'%(slots.size/u.size),'full %.2f'%((c==4).mean()),'runs',newrun.sum(),'avg run %.1f'%(ss.size/newrun.sum()))
TILE=224
tile=edge_id//TILE
ntiles=tile.max()+1
line=slots//16          # 128-byte lines
order=np.argsort(line,kind='stable')
ls=line[order]; ts=tile[order]
start=np.r_[0,np.nonzero(np.diff(ls))[0]+1]
tmin=np.minimum.reduceat(ts,start); tmax=np.maximum.reduceat(ts,start)
span=tmax-tmin
print('ntiles',ntiles,'lines',start.size)
for q in (50,75,90,95,99): print('span percentile',q,np.percentile(span,q))
# bytes of partially written lines alive at a time: for window w tiles: fraction of lines with span <= w
for w in (1,4,16,64,296,1000): print('lines complete within',w,'tiles: %.3f'%((span<=w).mean()))
