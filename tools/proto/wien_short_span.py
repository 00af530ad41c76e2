"""Prototype behind thermal_integral's rule for Wien-tail ranges that end at the thermal cut-off (acro_fastquad.cuh): error of
the difference of two Gauss-Laguerre sums and of short Gauss-Legendre rules against a 300-node reference, per span S."""
import numpy as np
from numpy.polynomial.laguerre import laggauss
from numpy.polynomial.legendre import leggauss
me=0.51099895; me2=me*me
def core(kind,E,Ep,x,xa):
    if kind==0:
        r=E/(Ep-E); c9=r*r/(2+2*r); q=xa/x; omq=1-q
        return (2*q*np.log(q)+(1+2*q)*omq+c9*omq)*x
    if kind==1:
        dE=Ep-E; c10=0.25*me2/Ep; i2=0.5/Ep
        A=dE+x; B=E-x; q=c10*A/(x*B); omq=1-q
        return (2*q*np.log(q)+(1+2*q)*omq+A*A*omq*i2/B)*x
    s=Ep+x; Eep=s-E; ee=E*Eep; s2=s*s
    sud=4*s2*np.log(4*x*ee/s/me2)/ee
    sud+= (me2/(x*s)-1)*(s2*s2)/(ee*ee)
    sud+= 2*(2*x*s-me2)*s2/ee/me2
    return sud-8/me2*x*s
def limits(kind,E,Ep):
    dE=Ep-E; ub=Ep-me2/(4*Ep)
    if kind==0: return 0.25*me2*E/(Ep*Ep-Ep*E), min(E,ub)
    if kind==1:
        pf=0.25*me2/Ep-E; qf=0.25*me2*dE/Ep; sq=np.sqrt((0.5*pf)**2-qf); return -0.5*pf-sq, min(-0.5*pf+sq,ub)
    rt=np.sqrt(E*E-me2); den=4*Ep*dE+me2; z1=Ep*(me2-2*dE*(rt-E))/den; z2=Ep*(me2+2*dE*(rt+E))/den
    return max(me2/Ep,z1), min(z2,Ep)
def exact(kind,E,Ep,T,xa,S):
    x,w=leggauss(300); s=0.5*S*(x+1); X=xa+s*T
    return 0.5*S*np.sum(w*core(kind,E,Ep,X,xa)*np.exp(-s)/(1-np.exp(-xa/T-s)))
def lag(kind,E,Ep,T,xa,x0,n):
    s,w=laggauss(n); X=x0+s*T
    return np.sum(w*core(kind,E,Ep,X,xa)/(1-np.exp(-x0/T-s)))
def glq(kind,E,Ep,T,xa,S,n):
    x,w=leggauss(n); s=0.5*S*(x+1); X=xa+s*T
    return 0.5*S*np.sum(w*core(kind,E,Ep,X,xa)*np.exp(-s)/(1-np.exp(-xa/T-s)))
def main():
    rng=np.random.default_rng(1)
    Eg=np.geomspace(1.5,100,150)
    worst={}
    for kind in range(3):
        cnt=0
        for trial in range(200000):
            i,j=sorted(rng.integers(0,150,2)); E,Ep=Eg[i],Eg[j]
            if kind<2 and i==j: continue
            if kind==0 and Ep<E+me: continue
            xa,xb0=limits(kind,E,Ep)
            if not (xb0>xa>0): continue
            va=rng.uniform(168,200); T=xa/va; S=200-va
            if xb0<200*T: continue
            xb=200*T
            ex=exact(kind,E,Ep,T,xa,S)
            for n in (4,6):
                ok = xb+laggauss(n)[0][-1]*T < xb0
                if ok:
                    v=lag(kind,E,Ep,T,xa,xa,n)-np.exp(-S)*lag(kind,E,Ep,T,xa,xb,n)
                    key=(kind,'lagdiff',n,int(min(S,31)//2)*2)
                    worst[key]=max(worst.get(key,0),abs(v/ex-1))
                else:
                    worst[(kind,'lag-invalid',n)]=worst.get((kind,'lag-invalid',n),0)+1
            for n in (6,8,12,16):
                v=glq(kind,E,Ep,T,xa,S,n); key=(kind,'gl',n,int(min(S,31)//2)*2)
                worst[key]=max(worst.get(key,0),abs(v/ex-1))
            cnt+=1
            if cnt>=1500: break
        print(kind,cnt)
    for k in sorted(worst, key=str): print(k, "%.2e"%worst[k])


if __name__ == "__main__":
    main()
