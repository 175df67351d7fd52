"""Generate the polynomial coefficients of the device asin/acos/atan2 used by the geodetic
transform (turtle_b200/csrc/tb_math.cuh) and measure their accuracy against mpmath.

  atan(t) = t + t^3 P(t^2)          t in [0, 1]
  asin(s) = s + s^3 Q(s^2)          s in [0, 0.55]

Both are near-minimax fits (Chebyshev interpolation in 60-digit arithmetic, then rounded
to double) of the RELATIVE-to-leading-term remainder. Run: python tools/fit_libm.py
"""
import sys
import numpy as np
import mpmath as mp

mp.mp.dps = 60


def fit(g, lo, hi, n):
    """Chebyshev interpolant of degree n-1 of g on [lo, hi] -> monomial coefficients."""
    poly, err = mp.chebyfit(g, [lo, hi], n, error=True)
    return [mp.mpf(c) for c in poly[::-1]], err  # ascending powers


def horner_double(coefs, u):
    acc = np.full_like(u, float(coefs[-1]))
    for c in coefs[-2::-1]:
        acc = acc * u + float(c)  # numpy: separate mul/add roundings (pessimistic vs FMA)
    return acc


def ulp_err(approx, exact_fn, x):
    out = []
    for a, xi in zip(approx, x):
        e = exact_fn(mp.mpf(float(xi)))
        u = mp.mpf(2) ** (mp.floor(mp.log(abs(e), 2)) - 52) if e != 0 else mp.mpf(1)
        out.append(float(abs(mp.mpf(float(a)) - e) / u))
    return np.array(out)


def main():
    rng = np.random.default_rng(0)
    # ---- atan ----
    def g_atan(u):
        u = mp.mpf(u)
        if u == 0:
            return mp.mpf(-1) / 3
        t = mp.sqrt(u)
        return (mp.atan(t) - t) / (t * u)
    for n in (19, 20, 21, 22):
        P, err = fit(g_atan, 0, 1, n)
        t = np.concatenate([rng.uniform(0, 1, 4000), [1.0, 0.5, 1e-3, 0.999999]])
        u = t * t
        approx = t + (t * u) * horner_double(P, u)
        e = ulp_err(approx, mp.atan, t)
        print("atan n=%d fit err %.2e  max ulp %.3f  mean %.3f" % (n, float(err), e.max(), e.mean()))
    # ---- asin ----
    def g_asin(u):
        u = mp.mpf(u)
        if u == 0:
            return mp.mpf(1) / 6
        s = mp.sqrt(u)
        return (mp.asin(s) - s) / (s * u)
    for n in (12, 13, 14, 15):
        Q, err = fit(g_asin, 0, mp.mpf("0.3025"), n)
        s = np.concatenate([rng.uniform(0, 0.55, 4000), [0.55, 0.5, 1e-3]])
        u = s * s
        approx = s + (s * u) * horner_double(Q, u)
        e = ulp_err(approx, mp.asin, s)
        print("asin n=%d fit err %.2e  max ulp %.3f  mean %.3f" % (n, float(err), e.max(), e.mean()))
    if "--emit" in sys.argv:
        P, _ = fit(g_atan, 0, 1, int(sys.argv[sys.argv.index("--emit") + 1]))
        Q, _ = fit(g_asin, 0, mp.mpf("0.3025"), int(sys.argv[sys.argv.index("--emit") + 2]))
        print("ATAN", ", ".join(float(c).hex() for c in P))
        print("ASIN", ", ".join(float(c).hex() for c in Q))


if __name__ == "__main__":
    main()


def check_combined():
    """Accuracy of the composed functions as the kernel evaluates them."""
    rng = np.random.default_rng(1)
    def g_atan(u):
        u = mp.mpf(u)
        if u == 0:
            return mp.mpf(-1) / 3
        t = mp.sqrt(u)
        return (mp.atan(t) - t) / (t * u)
    def g_asin(u):
        u = mp.mpf(u)
        if u == 0:
            return mp.mpf(1) / 6
        s = mp.sqrt(u)
        return (mp.asin(s) - s) / (s * u)
    P, _ = fit(g_atan, 0, 1, 21)
    Q, _ = fit(g_asin, 0, mp.mpf("0.3025"), 14)
    pio4 = mp.pi / 4
    PIO4_HI = float(pio4); PIO4_LO = float(pio4 - mp.mpf(PIO4_HI))
    pio2 = mp.pi / 2
    PIO2_HI = float(pio2); PIO2_LO = float(pio2 - mp.mpf(PIO2_HI))
    PI_HI = float(mp.pi); PI_LO = float(mp.pi - mp.mpf(PI_HI))
    R2 = float(mp.sqrt(mp.mpf(1) / 2))
    # asin on (0.55, 0.84) through w = (s - c) / sqrt 2
    s = rng.uniform(0.55, 0.8367, 6000)
    c = np.sqrt(1. - s * s)
    w = (s - c) * R2
    u = w * w
    a = (PIO4_HI + (w + (w * u) * horner_double(Q, u))) + PIO4_LO
    e = ulp_err(a, mp.asin, s)
    print("asin via pi/4: max ulp %.3f mean %.3f" % (e.max(), e.mean()))
    cc = rng.uniform(0., 0.548, 6000)
    u = cc * cc
    a = (PIO2_HI - (cc + (cc * u) * horner_double(Q, u))) + PIO2_LO
    e = ulp_err(a, mp.acos, cc)
    print("acos: max ulp %.3f mean %.3f" % (e.max(), e.mean()))
    # atan2 over all quadrants
    x = rng.normal(size=6000) * 10 ** rng.uniform(-3, 7, 6000)
    y = rng.normal(size=6000) * 10 ** rng.uniform(-3, 7, 6000)
    ax, ay = np.abs(x), np.abs(y)
    t = np.minimum(ax, ay) / np.maximum(ax, ay)
    u = t * t
    r = t + (t * u) * horner_double(P, u)
    r = np.where(ay > ax, (PIO2_HI - r) + PIO2_LO, r)
    r = np.where(x < 0, (PI_HI - r) + PI_LO, r)
    r = np.copysign(r, y)
    out = []
    for ri, xi, yi in zip(r, x, y):
        ex = mp.atan2(mp.mpf(float(yi)), mp.mpf(float(xi)))
        ulp = mp.mpf(2) ** (mp.floor(mp.log(abs(ex), 2)) - 52)
        out.append(float(abs(mp.mpf(float(ri)) - ex) / ulp))
    out = np.array(out)
    print("atan2: max ulp %.3f mean %.3f" % (out.max(), out.mean()))
    print("constants:", PIO4_HI.hex(), PIO4_LO.hex(), PIO2_HI.hex(), PIO2_LO.hex(), PI_HI.hex(),
          PI_LO.hex(), R2.hex())


if __name__ == "__main__" and "--combined" in sys.argv:
    check_combined()
