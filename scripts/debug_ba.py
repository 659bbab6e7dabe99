import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ba as oba, pnp as opnp, synth
from tests import helpers as H
from polychase_b200 import capi
F=np.float32
ctx=capi.Context(max_width=1920,max_height=1088)
w,h,NF=320,240,12
clip=synth.Clip(w,h,NF,seed=4)
go=capi.default_gftt(max_corners=250)
kps,flows={},{}
ctx.analyze_begin(w,h,0,NF,go)
for k in range(NF):
    ctx.analyze_push(k,clip.rgb(k))
    if ctx.analyze_pending()>=3:
        r=ctx.analyze_pop(); kps[r["frame_id"]]=r["keypoints"]
        for (a,b,rows,idx,tgt,err) in r["pairs"]: flows[(a,b)]=(idx,tgt,err)
while ctx.analyze_pending():
    r=ctx.analyze_pop(); kps[r["frame_id"]]=r["keypoints"]
    for (a,b,rows,idx,tgt,err) in r["pairs"]: flows[(a,b)]=(idx,tgt,err)
ctx.analyze_end()
verts,tris=H.bumpy_mesh(clip,quads=8,amp=0.03)
scene=dict(clip=clip,kps=kps,flows=flows,verts=verts,tris=tris,NF=NF)
rng=np.random.default_rng(3)
model=np.eye(4,dtype=F)
traj=[H.oracle_cam(clip,k) for k in range(NF)]
for k in range(1,NF-1): traj[k]=H.perturb(traj[k],rng,rot_deg=0.05,trans=0.004)
edges=[oba.Edge(a,b,flows[(a,b)][0],flows[(a,b)][1]) for (a,b) in sorted(flows) if len(flows[(a,b)][0])]
prob=oba.RefineProblem([kps[k] for k in range(NF)],edges,verts,tris,None,model,False,False,traj[0].intrinsics.bounds())
ctx.mesh_set(verts,tris)
ctx.ba_load([kps[k] for k in range(NF)],[(e.src,e.tgt,e.src_kps_indices,e.tgt_kps) for e in edges],model)
lo=opnp.Loss(2,1.0); bo=capi.default_bundle(loss_type=2)
atraj=[H.to_abi(c) for c in traj]
prob.total_cost(traj,lo); ctx.ba_cost(atraj,bo)
oc=np.concatenate(prob.cache); gc=ctx.ba_read_cache(len(oc))
bad=np.nonzero(oc!=gc)[0]
print('cache mismatches',len(bad),'of',len(oc), 'referenced', prob.referenced.sum())
for g in bad[:10]:
    f=prob.kp_frame[g]; print(' kp',g,'frame',f,'xy',prob.all_kps[g],'oracle prim',oc[g],'gpu prim',gc[g], 'tris', tris[oc[g]] if oc[g]!=0xFFFFFFFF else None, tris[gc[g]] if gc[g]!=0xFFFFFFFF else None)
for loss in (2,):
    lo=opnp.Loss(loss,1.0); bo=capi.default_bundle(loss_type=loss)
    atraj=[H.to_abi(c) for c in traj]
    print('cost',prob.total_cost(traj,lo),ctx.ba_cost(atraj,bo))
    A,g=prob.normal_equations(traj,lo)
    band,jtr=ctx.ba_normal_equations(atraj,bo)
    Ag=capi.band_to_dense(band); At=np.tril(A)
    d=np.abs(Ag-At); i,j=np.unravel_index(np.argmax(d),d.shape)
    print('loss',loss,'max diff',d.max(),'at',i,j,'vals',Ag[i,j],At[i,j],'scale',np.abs(A).max())
    # per 6x6 block relative diffs
    p=6
    worst=[]
    for bi in range(NF):
        for bj in range(bi+1):
            blkA=At[bi*p:(bi+1)*p,bj*p:(bj+1)*p]; blkG=Ag[bi*p:(bi+1)*p,bj*p:(bj+1)*p]
            if np.abs(blkA).max()>0:
                worst.append((np.abs(blkA-blkG).max()/np.abs(blkA).max(),bi,bj))
    worst.sort(reverse=True); print(worst[:6])
    print('jtr rel', np.abs(jtr-g).max()/np.abs(g).max())
    bi,bj=worst[0][1],worst[0][2]
    np.set_printoptions(precision=4,suppress=True,linewidth=200)
    print(At[bi*p:(bi+1)*p,bj*p:(bj+1)*p]); print(Ag[bi*p:(bi+1)*p,bj*p:(bj+1)*p])
