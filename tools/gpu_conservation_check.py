import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from functools import partial
from scipy.linalg import expm
import acropolis_b200.flags as flags; flags.verbose=False
from acropolis_b200.models import DecayModel
from acropolis_b200.engine import Engine
fixed = dict(temp0=10., n0a=1e-10, bree=0., braa=1.)
grid = [dict(mphi=float(m), tau=float(t)) for m in np.logspace(0, 3, 200) for t in np.logspace(3, 10, 200)]
par = grid[7::40]
ms=[DecayModel(**p, **fixed) for p in par]
todo=[m for m in ms if not m.is_trivial()]
eng=Engine.get(todo[0]._sII)
specs=DecayModel.point_specs(todo)
r=eng.run(specs, want=("Yf","mpdi","mdcy"))
A = np.array([1., 1., 2., 3., 3., 4., 6., 7., 7.])
Y0=todo[0]._sII.bbn_abundances()
dev=np.abs((np.einsum('pik,i->pk', r["Yf"], A))/(A@Y0)[None,:]-1).max(axis=1)
w=np.argsort(-dev)[:5]
for k in w:
    m=todo[k]
    M=r["mpdi"][k]+r["mdcy"][k]
    print("mphi %.3g tau %.3g dev %.2e  |mpdi|1 %.2e |mdcy|1 %.2e  colsum(A.mpdi) %.2e colsum(A.mdcy) %.2e" % (m._sMphi, m._sTau, dev[k], np.abs(r["mpdi"][k]).sum(0).max(), np.abs(r["mdcy"][k]).sum(0).max(), np.abs(A@r["mpdi"][k]).max(), np.abs(A@r["mdcy"][k]).max()))
    Ys = m.run_disintegration()
    # scipy on the same matrices
    pred, postd = m._pred_matrix(), m._postd_matrix()
    T = postd.dot(expm(M).dot(pred))
    Yscipy = T.dot(Y0)
    print("    conservation: ours %.2e  scipy-expm on our matrices %.2e" % (np.abs(A@Ys/(A@Y0)-1).max(), np.abs(A@Yscipy/(A@Y0)-1).max()))
