import sys; sys.path[:0]=['.','ip-nonlinear-solver_b200','tests']
import numpy as np, cases, scipy.sparse as sps, oracle
from ipsolver_b200.projections import projections
from ipsolver_b200.qp_subproblem import projected_cg
H,A,c,b,kw,_=cases.PCG_CASES['tr_barely_feasible']
Hs,As=sps.csc_matrix(H),sps.csc_matrix(A)
Z,_,Y=projections(As)
Zo,_,Yo=oracle.build_projectors(As)
x0=Y.dot(-np.array(b,float)); x0o=Yo.dot(-np.array(b,float))
print('x0',x0,'oracle',x0o,'norms %.17g %.17g radius %.17g'%(np.linalg.norm(x0),np.linalg.norm(x0o),kw['trust_radius']))
try:
    print(projected_cg(Hs,np.array(c),Z,Y,np.array(b),**kw))
except Exception as e: print('EXC',repr(e))
print(oracle.eqp_projected_cg(Hs,np.array(c),Zo,Yo,np.array(b),**kw))
