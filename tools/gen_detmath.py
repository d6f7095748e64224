"""Generate the polynomial coefficients used by include/pcr_detmath.h.

The deterministic elementary functions are *specification*: both the CPU oracle and the CUDA
kernels evaluate the same polynomials with the same IEEE-754 basic operations (no FMA
contraction), so their results are bit-identical.  Coefficients are near-minimax (Chebyshev
interpolation in 80-digit arithmetic) and printed as C hex-float literals.

Run:  python tools/gen_detmath.py   (prints the tables; paste into pcr_detmath.h)
"""
import mpmath as mp

mp.mp.dps = 80


def cheb_fit(f, a, b, deg):
    """Monomial coefficients (in x) of the degree-`deg` Chebyshev interpolant of f on [a,b]."""
    n = deg + 1
    nodes = [mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    xs = [(a + b) / 2 + (b - a) / 2 * t for t in nodes]
    ys = [f(x) for x in xs]
    # solve Vandermonde in high precision (n <= 16, fine at 80 digits)
    V = mp.matrix(n, n)
    for i, x in enumerate(xs):
        for j in range(n):
            V[i, j] = x ** j
    c = mp.lu_solve(V, mp.matrix(ys))
    return [c[i] for i in range(n)]


def hexf(x):
    return float(x).hex()


def show(name, coeffs):
    print(f"/* {name} */")
    for i, c in enumerate(coeffs):
        print(f"    {hexf(c)}, /* {mp.nstr(c, 20)} */")


if __name__ == "__main__":
    q = mp.pi / 4
    zmax = (q * mp.mpf("1.0001")) ** 2

    def fs(z):  # (sin(r) - r)/r^3 with z=r^2
        if z == 0:
            return -mp.mpf(1) / 6
        r = mp.sqrt(z)
        return (mp.sin(r) - r) / (r ** 3)

    def fc(z):  # (cos(r) - 1 + z/2)/z^2
        if z == 0:
            return mp.mpf(1) / 24
        r = mp.sqrt(z)
        return (mp.cos(r) - 1 + z / 2) / (z * z)

    a = mp.sqrt(2) - 1  # tan(pi/8)
    amax = (a * mp.mpf("1.0001")) ** 2

    def fa(z):  # (atan(t)/t - 1)/z with z=t^2
        if z == 0:
            return -mp.mpf(1) / 3
        t = mp.sqrt(z)
        return (mp.atan(t) / t - 1) / z

    show("SIN: sin(r) = r + r^3 * S(z)", cheb_fit(fs, mp.mpf(0), zmax, 6))
    show("COS: cos(r) = 1 - z/2 + z^2 * C(z)", cheb_fit(fc, mp.mpf(0), zmax, 6))
    show("ATAN: atan(t) = t + t*z*A(z), |t| <= tan(pi/8)", cheb_fit(fa, mp.mpf(0), amax, 13))
    print("PIO2_HI/LO:")
    pio2 = mp.pi / 2
    hi = mp.mpf(int(pio2 * 2 ** 32)) / 2 ** 32  # 33 significant bits
    print(hexf(hi), hexf(pio2 - hi), hexf(pio2 - hi - mp.mpf(float(pio2 - hi))))
    print("PI, PI/2, PI/4, 2/PI:", hexf(mp.pi), hexf(mp.pi / 2), hexf(mp.pi / 4), hexf(2 / mp.pi))
    print("PI_LO (pi - double(pi)):", hexf(mp.pi - mp.mpf(float(mp.pi))))
