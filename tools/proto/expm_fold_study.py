"""CPU study of the accuracy of expm(mpdi + mdcy) inside the abundance transfer: a NumPy
transcription of csrc expm9 (Pade-13 scaling and squaring), SciPy's expm and a 60-digit mpmath evaluation on the matrices
of DecayModel(54.2, 7.49e8, ...) (||mdcy||_1 = 1.45e9), and the fold of the free-neutron decay that apply_transfer() uses.
Needs the built oracle (make -C oracle).  usage: python tools/proto/expm_fold_study.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np
from scipy.linalg import expm
import acropolis_b200.flags as flags; flags.verbose=False
from acropolis_b200.models import DecayModel
from pyoracle import Oracle
b=[64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800., 129060195264000.,10559470521600., 670442572800., 33522128640., 1323241920., 40840800., 960960., 16380., 182., 1.]
def expm9(Ain, extra=0):
    nrm=np.abs(Ain).sum(0).max()
    s=0
    if nrm>5.371920351148152: s=int(max(0,np.ceil(np.log2(nrm/5.371920351148152))))
    s+=extra
    A=Ain*2.0**-s
    I=np.eye(9)
    A2=A@A; A4=A2@A2; A6=A4@A2
    U=A@(A6@(b[13]*A6+b[11]*A4+b[9]*A2)+b[7]*A6+b[5]*A4+b[3]*A2+b[1]*I)
    V=A6@(b[12]*A6+b[10]*A4+b[8]*A2)+b[6]*A6+b[4]*A4+b[2]*A2+b[0]*I
    R=np.linalg.solve(V-U,V+U)
    for _ in range(s): R=R@R
    return R
m=DecayModel(54.2, 7.49e8, 10., 1e-10, 0., 1.)
sp=m.point_spec()
o=Oracle(eps=1e-3)
r=o.run_point(sp)
M=r["mpdi"]+r["mdcy"]
A=np.array([1.,1.,2.,3.,3.,4.,6.,7.,7.])
Y0=m._sII.bbn_abundances()
pred,postd=m._pred_matrix(),m._postd_matrix()
print("norm1", np.abs(M).sum(0).max(), " A@M (column conservation of the generator):", np.abs(A@M).max(), np.abs(A@r["mdcy"]).max())
for name,E in (("ours",expm9(M)),("ours s+2",expm9(M,2)),("scipy",expm(M))):
    Y=postd@E@pred@Y0
    print(name, "conservation", np.abs(A@Y/(A@Y0)-1).max(), " A@E - A:", np.abs(A@E-A).max())
np.set_printoptions(linewidth=200, precision=3)
print("A@M per column:", A@M)
for name,E in (("ours",expm9(M)),("scipy",expm(M))):
    print(name, "A@E-A per column:", A@E-A)
# how many squarings does scipy use? emulate: check ours with fewer squarings
for ex in (-2,-4,-6,-8):
    E=expm9(M,ex); print("ours s%+d"%ex, "A@E-A cols 0..5:", (A@E-A)[:6])
E=expm9(M)
print("M[:,1] (p column):", M[:,1])
print("ours E[:,1]-e_p:", E[:,1]-np.eye(9)[:,1])
print("M[:,0] (n column):", M[:,0])
print("ours E[:,0]:", E[:,0], " scipy:", expm(M)[:,0])
def pade(Ain):
    nrm=np.abs(Ain).sum(0).max(); s=int(max(0,np.ceil(np.log2(nrm/5.371920351148152))))
    A=Ain*2.0**-s; I=np.eye(9)
    A2=A@A; A4=A2@A2; A6=A4@A2
    U=A@(A6@(b[13]*A6+b[11]*A4+b[9]*A2)+b[7]*A6+b[5]*A4+b[3]*A2+b[1]*I)
    V=A6@(b[12]*A6+b[10]*A4+b[8]*A2)+b[6]*A6+b[4]*A4+b[2]*A2+b[0]*I
    return U,V,s
U,V,s=pade(M)
print("s",s,"(V-U)[:,1]",(V-U)[:,1],"(V+U)[:,1]",(V+U)[:,1])
R=np.linalg.solve(V-U,V+U)
print("R[:,1]-e_p", R[:,1]-np.eye(9)[:,1])
print("R[1,:]", R[1,:])
R2=R.copy()
for k in range(s):
    R2=R2@R2
    if k<3 or k>s-3: print(k, R2[1,1]-1)
print("---- fold study")
import mpmath as mp
mp.mp.dps = 60
def mp_expm(Mat):
    return np.array(mp.expm(mp.matrix(Mat.tolist())).tolist(), dtype=object)
def fold(Min):
    B=Min.copy()
    B[:,0]=0.0                    # neutrons have no reactions of their own: the column is pure decay
    B[1,:]+=B[0,:]; B[0,:]=0.0    # whatever produces a neutron produces a proton
    return B
def transfer(E):
    return postd@E@pred@Y0
Etrue=mp_expm(M)
Ytrue=np.array((mp.matrix(postd.tolist())*mp.matrix(Etrue.tolist())*mp.matrix(pred.tolist())*mp.matrix(Y0.tolist())).tolist(),dtype=float)
for name,E in (("ours",expm9(M)),("scipy",expm(M)),("fold+ours",expm9(fold(M)))):
    Y=transfer(E)
    rel=np.abs(Y-Ytrue)/np.maximum(np.abs(Ytrue),1e-300)
    rel[Ytrue==0]=0
    print(name,"max rel err of Yf vs 60-digit:", rel.max(), " conservation:", np.abs(A@Y/(A@Y0)-1).max(), "norm used:", np.abs(fold(M)).sum(0).max() if 'fold' in name else np.abs(M).sum(0).max())
