"""Phase-level timings on the GPU: each SpMV alone, the inner CG iteration."""
import sys, time, json
sys.path[:0]=['.','ip-nonlinear-solver_b200']
import numpy as np
from ipsolver_b200 import synthetic, device
from ipsolver_b200.projections import projections
n=int(float(sys.argv[1])) if len(sys.argv)>1 else 10_000_000
pattern=sys.argv[2] if len(sys.argv)>2 else 'window'
H,A,c,b=synthetic.make_qp(n,pattern,1e4,0); m=n//2
ctx=device.Context.default()
Z,LS,Y=projections(A); proj=Z.projector
dH=device.DeviceCSR(ctx,H,False)
dc=device.DeviceVector(ctx,c); db=device.DeviceVector(ctx,b); dn=device.DeviceVector(ctx,length=n); dm=device.DeviceVector(ctx,length=m)
def t(f,reps=20):
    f(); f(); ctx.timer_start()
    for _ in range(reps): f()
    return ctx.timer_stop()/reps
lib=ctx.lib
B=lambda nnz,r,cc: 12*nnz+4*(r+1)+8*cc+8*r
res={}
ms=t(lambda: lib.ipb_csr_matvec_dev(ctx.h,proj.A.h,dc.h,dm.h)); res['A.x']=(ms,B(A.nnz,m,n)/ms/1e6)
ms=t(lambda: lib.ipb_csr_rmatvec_dev(ctx.h,proj.A.h,db.h,dn.h)); res['At.y']=(ms,B(A.nnz,n,m)/ms/1e6)
ms=t(lambda: lib.ipb_csr_matvec_dev(ctx.h,dH.h,dc.h,dn.h)); res['H.p']=(ms,B(H.nnz,n,n)/ms/1e6)
it=50
ms=device.bench_inner_iterations(proj,dc,it); ms=device.bench_inner_iterations(proj,dc,it)
Bin=B(A.nnz,m,n)+B(A.nnz,n,m)+56*m
res['inner_iter']=(ms/(it+1),Bin/(ms/(it+1))/1e6)
for k,(ms,gbs) in res.items(): print('%-12s %8.1f us  %7.1f GB/s  (%.1f%% of 6545.6)'%(k,ms*1e3,gbs,100*gbs/6545.6))
print('grid',ctx.grid())
# latency floor: tiny problem, same grid logic
for nn in (2000, 200000):
    H2,A2,c2,b2=synthetic.make_qp(nn,'window',1e4,0)
    Z2,_,_=projections(A2); p2=Z2.projector; d2=device.DeviceVector(ctx,c2)
    ms=device.bench_inner_iterations(p2,d2,200); ms=device.bench_inner_iterations(p2,d2,200)
    print('inner iteration at n=%d: %.2f us'%(nn,1e3*ms/201))
