import sys; sys.path[:0]=['.','ip-nonlinear-solver_b200','tests']
import numpy as np, cases, scipy.sparse as sps
from ipsolver_b200.projections import projections
A=sps.csc_matrix(np.array(cases.A38,float))
for tol,mr in ((1e-12,3),(1e-18,100)):
    Z,LS,Y=projections(A,orth_tol=tol,max_refin=mr)
    for p in cases.P38_N:
        x=Z.matvec(np.array(p,float)); i=Z.projector.last_info
        print(tol,mr,'res',np.abs(A.dot(x)).max(),'inner',i.inner_iters,'solves',i.n_solves,'refin',i.n_refinements,'orth',i.orth)
