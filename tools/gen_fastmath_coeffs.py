"""Derives the polynomial coefficients of branson_b200/csrc/fastmath.cuh with a weighted Remez exchange in 80-digit
arithmetic (mpmath) and prints them as C hexfloat literals together with the achieved approximation error.

    python tools/gen_fastmath_coeffs.py            # prints the tables pasted into fastmath.cuh

Approximations (all minimise the RELATIVE error of the final function value):
  exp : e^r          = 1 + r + r^2 * E(r),       |r| <= 0.3470  (ln2/2 plus slack), E degree 10
  log : log(m)       = s + s^3 * L(s^2),         s = 2(m-1)/(m+1), m in [sqrt(1/2), sqrt(2)],  L degree 7
  sin : sin(r)       = r + r^3 * S(r^2),         |r| <= 0.7860 (pi/4 plus slack), S degree 5
  cos : cos(r)       = 1 - r^2/2 + r^4 * C(r^2), same range, C degree 5
"""
import struct
import sys

from mpmath import mp, mpf, matrix, lu_solve, exp, log, sin, cos, atanh, sqrt, pi, findroot, diff

mp.dps = 80


def remez(g, w, a, b, deg, iters=12, grid=4000):
    """minimax of w(x) * (g(x) - p(x)) on [a, b]; returns (coeffs low->high, max weighted error)"""
    n = deg + 2
    xs = [(a + b) / 2 - (b - a) / 2 * mp.cos(mp.pi * i / (n - 1)) for i in range(n)]
    # the weights vanish at isolated points (r = 0, z = 0): the weighted error is 0 there, never an extremum
    xs = [x if w(x) > mpf(10) ** -30 else x + (b - a) / (7 * n) for x in xs]
    coef = None
    for _ in range(iters):
        A = matrix(n, n)
        rhs = matrix(n, 1)
        for i, x in enumerate(xs):
            for j in range(deg + 1):
                A[i, j] = x ** j
            A[i, deg + 1] = (-1) ** i / w(x)
            rhs[i] = g(x)
        sol = lu_solve(A, rhs)
        coef = [sol[j] for j in range(deg + 1)]

        def err(x):
            return w(x) * (g(x) - sum(c * x ** j for j, c in enumerate(coef)))
        # locate the extrema of err on a grid, refine by parabolic steps
        pts = [a + (b - a) * i / grid for i in range(grid + 1)]
        ev = [err(x) for x in pts]
        ext = []
        for i in range(grid + 1):
            l = ev[i - 1] if i > 0 else None
            r = ev[i + 1] if i < grid else None
            v = ev[i]
            if w(pts[i]) <= mpf(10) ** -30:
                continue
            if (l is None or abs(v) >= abs(l)) and (r is None or abs(v) >= abs(r)):
                x = pts[i]
                if 0 < i < grid:
                    h = (b - a) / grid
                    for _k in range(3):
                        f0, f1, f2 = err(x - h), err(x), err(x + h)
                        den = f0 - 2 * f1 + f2
                        if den == 0:
                            break
                        x = x - h * (f2 - f0) / (2 * den)
                        x = min(max(x, a), b)
                        h /= 8
                ext.append((x, err(x)))
        # keep an alternating subsequence of the n largest
        alt = []
        for x, v in ext:
            if alt and (v > 0) == (alt[-1][1] > 0):
                if abs(v) > abs(alt[-1][1]):
                    alt[-1] = (x, v)
            else:
                alt.append((x, v))
        while len(alt) > n:
            if abs(alt[0][1]) < abs(alt[-1][1]):
                alt.pop(0)
            else:
                alt.pop()
        if len(alt) < n:
            break
        xs = [x for x, _ in alt]
    emax = max(abs(v) for v in ev)
    return coef, emax


def to_double(x):
    return float(x)


def hexf(x):
    return float(x).hex()


def check(name, f_true, f_apx, w_rel, a, b, n=20001):
    worst = mpf(0)
    for i in range(n):
        x = a + (b - a) * i / (n - 1)
        t = f_true(x)
        if t == 0:
            continue
        e = abs((f_apx(x) - t) / t)
        worst = max(worst, e)
    print(f"// {name}: max relative approximation error with double-rounded coefficients = 2^{float(mp.log(worst, 2)):.2f}")


def emit(name, coef):
    print(f"// {name}")
    for j, c in enumerate(coef):
        print(f"  {hexf(c)},  // [{j}] {float(c)!r}")


def main():
    # ---- exp
    L = mpf("0.3470")
    g = lambda r: (exp(r) - 1 - r) / (r * r) if r != 0 else mpf(1) / 2
    w = lambda r: (r * r) / exp(r) if r != 0 else mpf(10) ** -40
    ce, e = remez(g, w, -L, L, 10)
    ced = [mpf(to_double(c)) for c in ce]
    emit("EXP: E(r), e^r = 1 + r + r^2 E(r)", ced)
    check("exp", exp, lambda r: 1 + r + r * r * sum(c * r ** j for j, c in enumerate(ced)), None, -L, L)
    # ---- log
    zmax = (2 * (sqrt(2) - 1) / (sqrt(2) + 1)) ** 2 * mpf("1.0001")

    def gl(z):
        if z == 0:
            return mpf(1) / 12
        s = sqrt(z)
        return (2 * atanh(s / 2) - s) / (s * z)
    wl = lambda z: z if z != 0 else mpf(10) ** -40
    cl, e = remez(gl, wl, mpf(0), zmax, 7)
    cld = [mpf(to_double(c)) for c in cl]
    emit("LOG: L(z), log(m) = s + s^3 L(s^2), s = 2(m-1)/(m+1)", cld)
    smax = sqrt(zmax)
    check("log", lambda s: 2 * atanh(s / 2), lambda s: s + s ** 3 * sum(c * (s * s) ** j for j, c in enumerate(cld)),
          None, -smax, smax)
    # ---- sin / cos
    R = mpf("0.7860")
    gs = lambda z: (sin(sqrt(z)) - sqrt(z)) / (sqrt(z) * z) if z != 0 else -mpf(1) / 6
    ws = lambda z: z if z != 0 else mpf(10) ** -40
    cs, e = remez(gs, ws, mpf(0), R * R, 5)
    csd = [mpf(to_double(c)) for c in cs]
    emit("SIN: S(z), sin(r) = r + r^3 S(r^2)", csd)
    check("sin", sin, lambda r: r + r ** 3 * sum(c * (r * r) ** j for j, c in enumerate(csd)), None, -R, R)
    gc = lambda z: (cos(sqrt(z)) - 1 + z / 2) / (z * z) if z != 0 else mpf(1) / 24
    wc = lambda z: z * z / cos(sqrt(z)) if z != 0 else mpf(10) ** -40
    cc, e = remez(gc, wc, mpf(0), R * R, 5)
    ccd = [mpf(to_double(c)) for c in cc]
    emit("COS: C(z), cos(r) = 1 - r^2/2 + r^4 C(r^2)", ccd)
    check("cos", cos, lambda r: 1 - r * r / 2 + r ** 4 * sum(c * (r * r) ** j for j, c in enumerate(ccd)), None, -R, R)
    # ---- constants
    ln2 = log(mpf(2))
    hi = mpf(float(ln2))
    # ln2 split for exp: hi with 32 significant bits so that k * hi is exact for |k| < 2^20
    def trunc_bits(x, bits):
        m, e = mp.frexp(x)
        return mp.ldexp(mp.floor(mp.ldexp(m, bits)), e - bits)
    ln2_hi = trunc_bits(ln2, 32)
    ln2_lo = ln2 - ln2_hi
    print("// constants")
    print(f"  LN2_HI32 = {hexf(ln2_hi)}, LN2_LO = {hexf(ln2_lo)}, LOG2E = {hexf(1 / ln2)}")
    print(f"  LN2_D = {hexf(hi)}, LN2_D_LO = {hexf(ln2 - hi)}")
    p2 = pi / 2
    p2_hi = trunc_bits(p2, 33)
    p2_mid = trunc_bits(p2 - p2_hi, 45)  # q * mid stays exact for |q| < 2^8
    p2_lo = mpf(float(p2 - p2_hi - p2_mid))
    print(f"  PIO2_HI33 = {hexf(p2_hi)}, PIO2_MID45 = {hexf(p2_mid)}, PIO2_LO = {hexf(p2_lo)}, TWO_OVER_PI = {hexf(2 / pi)}")


if __name__ == "__main__":
    main()
