"""Error of the n-node Gauss-Laguerre rule per band of v_a = x_a/T for the three thermal integrals (long spans): the table
behind the orders in thermal_integral (acro_fastquad.cuh)."""
import numpy as np, sys
from numpy.polynomial.laguerre import laggauss
from numpy.polynomial.legendre import leggauss
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from wien_short_span import core, limits, me
rng=np.random.default_rng(2)
Eg=np.geomspace(1.5,300,280)
bands=[(7.389,9.0),(9.0,12.18),(12.18,16),(16,20.09),(20.09,30),(30,54.6),(54.6,100),(100,168)]
orders=[2,3,4,5,6,7,8,10,12]
LG={n:laggauss(n) for n in orders}
xg,wg=leggauss(400)
def exact(kind,E,Ep,T,xa,xb):
    # composite: [0,60] in s with 400 GL nodes
    S=min((xb-xa)/T,80.0); s=0.5*S*(xg+1); X=xa+s*T
    return 0.5*S*np.sum(wg*core(kind,E,Ep,X,xa)*np.exp(-s)/(1-np.exp(-xa/T-s)))
worst={}
for kind in range(3):
    for (lo,hi) in bands:
        cnt=0
        while cnt<400:
            i,j=sorted(rng.integers(0,280,2)); E,Ep=Eg[i],Eg[j]
            if kind<2 and i==j: continue
            if kind==0 and Ep<E+me: continue
            xa,xb0=limits(kind,E,Ep)
            if not (xb0>xa>0): continue
            va=rng.uniform(lo,hi); T=xa/va
            xb=min(xb0,200*T)
            if (xb-xa)/T<32: continue
            ex=exact(kind,E,Ep,T,xa,xb)
            for n in orders:
                s,w=LG[n]; X=xa+s*T; ok=X<xb
                v=np.sum((w*core(kind,E,Ep,np.where(ok,X,xa*1.0001),xa)/(1-np.exp(-va-s)))[ok])
                key=(kind,lo,n); worst[key]=max(worst.get(key,0),abs(v/ex-1))
            cnt+=1
for kind in range(3):
    print("kind",kind)
    for (lo,hi) in bands:
        print("  va %6.2f-%6.2f: "%(lo,hi)+"  ".join("n%d %.1e"%(n,worst[(kind,lo,n)]) for n in orders))
