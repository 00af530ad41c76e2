"""Numpy prototype of the T-shared kernel quadrature ("v2"): one global panel grid in y = log(eps_bar) per model point,
F/G evaluated once per node for all temperatures, Bose factors tabulated per (T, node); partial end panels through
Legendre moments.  Compares against a converged per-(pair,T) Gauss-Legendre rule on the exact limits."""
import sys
import numpy as np
from numpy.polynomial.legendre import leggauss, legvander

me = 0.511; me2 = me * me; pi = np.pi


def F(Eph, Ee, x):
    G = 4 * x * Ee / me2
    q = Eph / (G * (Ee - Eph))
    ok = (x <= Eph) & (Eph * (me2 + 4 * x * Ee) <= 4 * x * Ee * Ee)
    q = np.where(ok, q, 0.5)
    return np.where(ok, 2 * q * np.log(q) + (1 + 2 * q) * (1 - q) + (G * q)**2 * (1 - q) / (2 + 2 * G * q), 0.0)


def Gf(Ee, Eph, x):
    s = Eph + x; Eep = s - Ee
    arg = np.maximum(1 - me2 / (Eph * x), 0)
    dEs = (Eph - x) * np.sqrt(arg)
    lm, lp = 0.5 * (s - dEs), 0.5 * (s + dEs)
    ok = (me < lm) & (lm <= Ee) & (Ee <= lp)
    ee = Ee * Eep
    with np.errstate(all="ignore"):
        sud = 4 * s * s * np.log(4 * x * ee / (me2 * s)) / ee + (me2 / (x * s) - 1) * s**4 / ee**2 + 2 * (2 * x * s - me2) * s * s / (me2 * ee) - 8 * x * s / me2
    return np.where(ok, sud, 0.0)


def limits(kind, E, Ep, T):
    xmax = 200 * T
    if kind == 9:
        if Ep < E + me: return None
        a = 0.25 * me2 * E / (Ep * Ep - Ep * E); b = min(E, Ep - me2 / (4 * Ep), xmax)
    elif kind == 10:
        if E == Ep: return None
        pf = 0.25 * me2 / Ep - E; qf = 0.25 * me2 * (Ep - E) / Ep; sq = np.sqrt((pf / 2)**2 - qf)
        a = -pf / 2 - sq; b = min(-pf / 2 + sq, Ep - me2 / (4 * Ep), xmax)
    else:
        if Ep < me2 / (50 * T): return None
        dE = Ep - E; rt = np.sqrt(E * E - me2); den = 4 * Ep * dE + me2
        z1 = Ep * (me2 - 2 * dE * (rt - E)) / den; z2 = Ep * (me2 + 2 * dE * (rt + E)) / den
        a = max(me2 / Ep, z1); b = min(z2, Ep, xmax)
    return (a, b) if b > a else None


def fy(kind, E, Ep, y):
    x = np.exp(y)
    if kind == 9: return F(E, Ep, x)
    if kind == 10: return F(Ep - E + x, Ep, x)
    return Gf(E, Ep, x)


def bose(kind, y, T):
    x = np.exp(y)
    p = 2 if kind != 11 else 1
    return x**p / np.expm1(x / T) / pi**2


def exact(kind, E, Ep, T, n=200):
    lim = limits(kind, E, Ep, T)
    if lim is None: return 0.0
    a, b = np.log(lim[0]), np.log(lim[1])
    x, w = leggauss(n)
    y = 0.5 * (a + b) + 0.5 * (b - a) * x
    return 0.5 * (b - a) * np.sum(w * fy(kind, E, Ep, y) * bose(kind, y, T))


class Grid:
    """Panels: coarse (width hc, nc nodes) below y_split, fine (hf, nf) above, up to y_hi."""
    def __init__(self, Ts, y_min, y_max, hf=1.0, nf=16, hc=2.0, nc=16, umin=-25.0):
        Tmin, Tmax = Ts.min(), Ts.max()
        y_hi = min(y_max, np.log(200 * Tmax))
        y_split = np.log(Tmin) - 1.0
        y_lo = max(y_min, np.log(Tmin) + umin)
        edges = [y_hi]
        while edges[-1] > max(y_split, y_lo):
            edges.append(edges[-1] - hf)
        nfine = len(edges) - 1
        while edges[-1] > y_lo:
            edges.append(edges[-1] - hc)
        self.edges = np.array(edges[::-1]); self.npan = len(self.edges) - 1
        self.order = np.array([nc] * (self.npan - nfine) + [nf] * nfine)
        self.Ts = Ts

    def integrate(self, kind, E, Ep):
        """Integral for every T with T-independent limits (the 200T cut applied through the Bose table)."""
        out = np.zeros(len(self.Ts))
        lim = limits(kind, E, Ep, 1e9)          # T-independent part of the limits
        if lim is None: return out
        a, b = np.log(lim[0]), np.log(lim[1])
        a = max(a, self.edges[0]); b = min(b, self.edges[-1])
        nev = 0
        for p in range(self.npan):
            p0, p1 = self.edges[p], self.edges[p + 1]
            lo, hi = max(a, p0), min(b, p1)
            if hi <= lo: continue
            n = self.order[p]
            x, w = leggauss(n)
            yk = 0.5 * (p0 + p1) + 0.5 * (p1 - p0) * x          # global nodes of this panel
            Bk = np.array([bose(kind, yk, T) * (yk <= np.log(200 * T)) for T in self.Ts])   # [NT, n]  (table)
            if lo == p0 and hi == p1:                           # full panel: nodal rule
                f = fy(kind, E, Ep, yk); nev += n
                out += 0.5 * (p1 - p0) * (Bk * (w * f)[None, :]).sum(1)
            else:                                               # partial panel: Legendre moments of f on [lo,hi]
                ys = 0.5 * (lo + hi) + 0.5 * (hi - lo) * x
                f = fy(kind, E, Ep, ys); nev += n
                t = (ys - 0.5 * (p0 + p1)) / (0.5 * (p1 - p0))  # position inside the panel, in [-1,1]
                V = legvander(t, n - 1)                         # [n, n] P_j(t_k)
                mu = 0.5 * (hi - lo) * (V * (w * f)[:, None]).sum(0)        # mu_j = int f P_j
                # modal coefficients of B on the panel: beta_j = (2j+1)/2 sum_k w_k B_k P_j(x_k)
                Vg = legvander(x, n - 1)
                beta = (Bk * w[None, :]) @ Vg * (2 * np.arange(n) + 1) / 2       # [NT, n]
                out += beta @ mu
        self.nev = nev
        return out


def cut_mask(kind, E, Ep, T):
    return limits(kind, E, Ep, T) is not None


if __name__ == "__main__":
    z = np.load("/root/repo/tests/golden/%s.npz" % (sys.argv[1] if len(sys.argv) > 1 else "cfg1_decay"))
    Eg, Ts = z["E_grid"], z["T"]
    NE = len(Eg)
    pairs = [(i, j) for i in range(0, NE, max(1, NE // 9)) for j in sorted(set(list(range(i, min(i + 4, NE))) + list(range(i, NE, max(1, NE // 7))) + [NE - 1]))]
    # exact values once
    EX = {}
    for kind in (9, 10, 11):
        for (i, j) in pairs:
            for it, T in enumerate(Ts):
                lim = limits(kind, Eg[i], Eg[j], T)
                if lim is not None:
                    EX[(kind, i, j, it)] = (exact(kind, Eg[i], Eg[j], T), np.log(lim[0]) - np.log(T))
    print("combos", len(EX))
    for usw in (0.5, 1.0, 1.5, 2.0, 2.5):
        print("u_switch %.1f: fraction of (pair,T) in the steep regime: %.3f" % (usw, np.mean([v[1] > usw for v in EX.values()])))
    for hf, nf, hc, nc in ((1.0, 16, 2.0, 16), (0.5, 8, 2.0, 16), (1.0, 12, 2.0, 12), (0.5, 8, 1.0, 8), (1.5, 16, 3.0, 16)):
        g = Grid(Ts, np.log(1e-9), np.log(Eg[-1]), hf, nf, hc, nc)
        worst = {}; nevs = []
        for kind in (9, 10, 11):
            for (i, j) in pairs:
                got = g.integrate(kind, Eg[i], Eg[j])
                if got.any(): nevs.append(g.nev)
                for it in range(len(Ts)):
                    k = (kind, i, j, it)
                    if k in EX and EX[k][0] > 1e-250:
                        e = abs(got[it] / EX[k][0] - 1)
                        for usw in (0.5, 1.0, 1.5, 2.0, 2.5):
                            if EX[k][1] <= usw:
                                worst[(kind, usw)] = max(worst.get((kind, usw), 0), e)
        print("hf=%.2f nf=%d hc=%.1f nc=%d panels=%d nodes=%d avg f-evals=%.0f" % (hf, nf, hc, nc, g.npan, g.order.sum(), np.mean(nevs)))
        for usw in (0.5, 1.0, 1.5, 2.0, 2.5):
            print("     u_a<=%.1f: a9 %.1e  a10 %.1e  a11 %.1e" % (usw, worst.get((9, usw), 0), worst.get((10, usw), 0), worst.get((11, usw), 0)))
