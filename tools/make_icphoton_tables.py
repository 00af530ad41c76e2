#!/usr/bin/env python3
"""Generate acropolis_b200/csrc/acro_icphoton_tab.h: polynomial tables of the four one-variable functions in which the
e -> gamma inverse-Compton kernel (reference cascade.py:406-435) has a closed form.

With q = x_a/x the integrand factor F(E, E', x) (cascade.py:29-43) is 2 q ln q + (1+2q)(1-q) + c9 (1-q), c9 = r^2/(2+2r),
r = E/(E'-E): it depends on x only through q.  In v = x/T therefore

    int_{x_a}^{x_b} F x / (e^{x/T} - 1) dx  =  T^2 [ Phi0(v_a, v_b) + c9 Phi1(v_a, v_b) ]
    Phi1 = [B(v_a)] - [B(v_b) + (v_b - v_a) L(v_b)]
    Phi0 = [A(v_a)] - [B(v_b) + v_b L(v_b) + v_a (1 + 2 ln v_a) L(v_b) - 2 v_a^2 M(v_b) - 2 v_a N(v_b)]

    L(v) = -ln(1 - e^-v)                          B(v) = Li2(e^-v)
    M(v) = int_v^inf dt / (t (e^t - 1))           N(v) = int_v^inf ln t / (e^t - 1) dt
    A(a) = int_a^inf [2 q ln q + (1+2q)(1-q)]_{q=a/v} v / (e^v - 1) dv
         = B(a) + a L(a) (2 + 2 ln a) - 2 a^2 M(a) - 2 a N(a)      (tabulated directly: the pieces cancel for large a)

Tables (all functions times e^v, in u = ln v, monomial coefficients in the local variable of each interval):
    A, B on u in [-40, -8) in steps of 2 and [-8, 5.5] in steps of 1/2   (v up to 244 > Ephb_T_max = 200)
    M, N on u in [2, 5.5] in steps of 1/2                                 (v >= e^2: only the upper limit needs them)
Computed with mpmath at 30 digits (Chebyshev interpolation of degree DEG, converted to monomials), checked below against
direct 30-digit quadrature of the original integral in emulated double-precision evaluation.

    python tools/make_icphoton_tables.py            # writes the header, prints the check
"""
import os
import sys
import time

import mpmath as mp
import numpy as np

mp.mp.dps = 30
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "acropolis_b200", "csrc", "acro_icphoton_tab.h")
DEG = 11
U_MIN, U_SPLIT, U_MAX, W_LO, W_HI = -40.0, -8.0, 5.5, 2.0, 0.5
U_TAIL = 2.0
PTS = [0, mp.mpf(1) / 4, 1, 4, 12, 30, 70, 160]


def h(q):
    return 2 * q * mp.log(q) + (1 + 2 * q) * (1 - q)


def bose(v, s):
    return 1 / (-mp.expm1(-(v + s)))


def A_scaled(a):
    return mp.quad(lambda s: h(a / (a + s)) * (a + s) * mp.exp(-s) * bose(a, s), PTS)


def B_scaled(a):
    return mp.polylog(2, mp.exp(-a)) * mp.exp(a)


def M_scaled(v):
    return mp.quad(lambda s: mp.exp(-s) * bose(v, s) / (v + s), PTS)


def N_scaled(v):
    return mp.quad(lambda s: mp.log(v + s) * mp.exp(-s) * bose(v, s), PTS)


def cheb_to_monomial(c):
    """Chebyshev coefficients -> monomial coefficients (mp arithmetic)."""
    n = len(c)
    T = [[mp.mpf(0)] * n for _ in range(n)]
    T[0][0] = mp.mpf(1)
    if n > 1:
        T[1][1] = mp.mpf(1)
    for k in range(2, n):
        for j in range(n):
            T[k][j] = (2 * T[k - 1][j - 1] if j else 0) - T[k - 2][j]
    return [sum(c[k] * T[k][j] for k in range(n)) for j in range(n)]


def fit(fun, lo, hi):
    n = DEG + 1
    xs = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / n) for k in range(n)]
    fs = [fun(mp.exp((lo + hi) / 2 + (hi - lo) / 2 * x)) for x in xs]
    c = []
    for j in range(n):
        s = sum(fs[k] * mp.cos(mp.pi * j * (k + mp.mpf(1) / 2) / n) for k in range(n)) * 2 / n
        c.append(s if j else s / 2)
    return [float(m) for m in cheb_to_monomial(c)], float(max(abs(c[-1]), abs(c[-2])) / abs(c[0]))


def intervals_main():
    iv = []
    u = U_MIN
    while u < U_SPLIT - 1e-9:
        iv.append((u, u + W_LO)); u += W_LO
    while u < U_MAX - 1e-9:
        iv.append((u, u + W_HI)); u += W_HI
    return iv


def intervals_tail():
    iv = []
    u = U_TAIL
    while u < U_MAX - 1e-9:
        iv.append((u, u + W_HI)); u += W_HI
    return iv


def build():
    tabs, tails = {}, {}
    for name, fun, iv in (("A", A_scaled, intervals_main()), ("B", B_scaled, intervals_main()),
                          ("M", M_scaled, intervals_tail()), ("N", N_scaled, intervals_tail())):
        rows, worst = [], 0.0
        for lo, hi in iv:
            m, t = fit(fun, mp.mpf(lo), mp.mpf(hi))
            rows.append(m); worst = max(worst, t)
        tabs[name] = np.array(rows); tails[name] = worst
    return tabs, tails


# ---- double-precision emulation of the device evaluation (same operation order as acro_icphoton.cuh) ---------------

def _poly(tab, k, x):
    p = tab[k, DEG]
    for j in range(DEG - 1, -1, -1):
        p = p * x + tab[k, j]
    return p


def _locate_main(u):
    if u >= U_SPLIT:
        k = int((U_SPLIT - U_MIN) / W_LO) + min(int((u - U_SPLIT) / W_HI), int(round((U_MAX - U_SPLIT) / W_HI)) - 1)
        lo = U_SPLIT + (k - int((U_SPLIT - U_MIN) / W_LO)) * W_HI
        return k, (u - lo) * (2.0 / W_HI) - 1.0
    uu = max(u, U_MIN)
    k = min(int((uu - U_MIN) / W_LO), int((U_SPLIT - U_MIN) / W_LO) - 1)
    return k, (uu - (U_MIN + k * W_LO)) * (2.0 / W_LO) - 1.0


def _locate_tail(u):
    k = min(int((u - U_TAIL) / W_HI), int(round((U_MAX - U_TAIL) / W_HI)) - 1)
    return k, (u - (U_TAIL + k * W_HI)) * (2.0 / W_HI) - 1.0


def phi_double(tabs, va, vb, tail_from=36.0):
    """(Phi0, Phi1) in float64 from the tables."""
    ua = float(np.log(va))
    k, x = _locate_main(ua)
    ea = float(np.exp(-va))
    p0 = _poly(tabs["A"], k, x) * ea
    p1 = _poly(tabs["B"], k, x) * ea
    if vb - va < tail_from:
        ub = float(np.log(vb))
        eb = float(np.exp(-vb))
        kb, xb = _locate_main(ub)
        kt, xt = _locate_tail(ub)
        Bb = _poly(tabs["B"], kb, xb) * eb
        Mb = _poly(tabs["M"], kt, xt) * eb
        Nb = _poly(tabs["N"], kt, xt) * eb
        Lb = -float(np.log1p(-eb))
        p0 -= Bb + vb * Lb + va * (1.0 + 2.0 * ua) * Lb - 2.0 * va * va * Mb - 2.0 * va * Nb
        p1 -= Bb + (vb - va) * Lb
    return p0, p1


def phi_exact(va, vb):
    """30-digit quadrature of the original integrals, in s = v - v_a with the factor e^{-v_a} taken out of the integrand."""
    va, vb = mp.mpf(va), mp.mpf(vb)
    S = vb - va
    cut = sorted(set([mp.mpf(0), S] + [S * mp.mpf(f) for f in ("1e-3", "1e-2", "0.1", "0.3")] + [p for p in PTS if p < S]))
    f0 = mp.quad(lambda s: h(va / (va + s)) * (va + s) * mp.exp(-s) * bose(va, s), cut)
    f1 = mp.quad(lambda s: s * mp.exp(-s) * bose(va, s), cut)
    return f0 * mp.exp(-va), f1 * mp.exp(-va)


def check(tabs, n=500, seed=0):
    rng = np.random.default_rng(seed)
    worst = {}
    for i in range(n):
        va = float(np.exp(rng.uniform(np.log(1e-9), np.log(199.0))))
        mode = i % 5
        if mode == 0:
            vb = 200.0
        elif mode == 1:
            vb = min(200.0, va + float(np.exp(rng.uniform(np.log(2.0), np.log(36.0)))))
        elif mode == 2:
            vb = min(200.0, max(8.0, va + 2.0) + float(rng.uniform(0, 30)))
        elif mode == 3:
            vb = min(200.0, va * float(np.exp(rng.uniform(0.1, 6))) + 2.0)
        else:                                                   # short ranges at large v (the thermal cut-off): cancellation
            va = float(rng.uniform(8.0, 199.7))
            vb = va + float(np.exp(rng.uniform(np.log(0.25), np.log(2.0))))
        if vb < 8.0 and vb - va < 36.0:
            vb = 8.0 + float(rng.uniform(0, 5))
        if vb - va < 0.25:
            continue
        e0, e1 = phi_exact(va, vb)
        g0, g1 = phi_double(tabs, va, vb)
        band = "va<1" if va < 1 else "va<7.4" if va < 7.39 else "va<50" if va < 50 else "va>=50"
        key = (band, "short" if vb - va < 2 else "tail" if vb - va < 36 else "no tail")
        w = worst.setdefault(key, [0.0, 0.0, 0])
        w[0] = max(w[0], abs(float(g0 / e0 - 1))); w[1] = max(w[1], abs(float(g1 / e1 - 1))); w[2] += 1
    return worst


def load_header(path=OUT):
    """The four tables back from the generated header (tests/test_host.py checks them against scipy quadrature)."""
    import re
    txt = open(path).read()
    tabs = {}
    for name in "ABMN":
        m = re.search(r"kIcp%s\[(\d+)\]\[(\d+)\] = \{(.*?)\};" % name, txt, re.S)
        vals = [float(v) for v in re.findall(r"[-+]?\d\.\d+e[-+]\d+", m.group(3))]
        tabs[name] = np.array(vals).reshape(int(m.group(1)), int(m.group(2)))
    return tabs


def write_header(tabs):
    def arr(name, a):
        rows = ",\n".join("    {" + ", ".join("%.17e" % v for v in r) + "}" for r in a)
        return "ACRO_ICP_TABLE double %s[%d][%d] = {\n%s};\n" % (name, a.shape[0], a.shape[1], rows)
    with open(OUT, "w") as f:
        f.write("// GENERATED by tools/make_icphoton_tables.py -- do not edit.  Polynomial tables (monomial coefficients in the local\n"
                "// variable x in [-1, 1] of each interval of u = ln v) of A(v) e^v, B(v) e^v = Li2(e^-v) e^v, M(v) e^v, N(v) e^v: see the\n"
                "// generator for the definitions and acro_icphoton.cuh for the use.  The includer defines ACRO_ICP_TABLE (storage class).\n"
                "#pragma once\n\n")
        f.write("#define ACRO_ICP_DEG %d\n#define ACRO_ICP_U_MIN %.1f\n#define ACRO_ICP_U_SPLIT %.1f\n#define ACRO_ICP_U_MAX %.1f\n"
                "#define ACRO_ICP_W_LO %.1f\n#define ACRO_ICP_W_HI %.1f\n#define ACRO_ICP_U_TAIL %.1f\n#define ACRO_ICP_N_MAIN %d\n#define ACRO_ICP_N_TAIL %d\n\n"
                % (DEG, U_MIN, U_SPLIT, U_MAX, W_LO, W_HI, U_TAIL, tabs["A"].shape[0], tabs["M"].shape[0]))
        for name in "ABMN":
            f.write(arr("kIcp" + name, tabs[name]) + "\n")


if __name__ == "__main__":
    t0 = time.time()
    tabs, tails = build()
    print("tables built in %.0f s; largest trailing Chebyshev coefficient relative to c0:" % (time.time() - t0),
          {k: "%.1e" % v for k, v in tails.items()})
    if "--no-write" not in sys.argv:
        write_header(tabs)
        print("wrote", OUT)
    np.savez("/tmp/icphoton_tabs.npz", **tabs)
    for k, w in sorted(check(tabs).items()):
        print("%-8s %-8s n=%3d  worst rel. error Phi0 %.1e  Phi1 %.1e" % (k[0], k[1], w[2], w[0], w[1]))
