import numpy as np
from scipy.special import erfc, erf
from scipy.optimize import least_squares
ZMAX = 4.0
z = np.linspace(1e-6, ZMAX, 20001)
T = -np.log2(erfc(z))            # target exponent
for deg in (5, 6, 7):
    # initial linear LSQ on T(z)/z
    V = np.vander(z, deg + 1, increasing=True)[:, :deg]   # Q degree deg-1 => P = z*Q degree deg
    q0 = np.linalg.lstsq(V * z[:, None], T, rcond=None)[0]
    def resid(q):
        P = z * (V @ q)
        return (np.exp2(-P) - erfc(z)) * z     # weight by z ~ |x| (gelu error = 0.5|x| * err)
    sol = least_squares(resid, q0, xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=20000)
    q = sol.x
    # fp32 emulation of the kernel formula on x grid
    x = np.linspace(-9, 9, 400001).astype(np.float32)
    c = np.float32(1 / np.sqrt(2))
    zz = np.minimum(np.abs(x) * c, np.float32(ZMAX)).astype(np.float32)
    acc = np.float32(q[-1]) * np.ones_like(zz)
    for k in range(deg - 2, -1, -1):
        acc = (acc * zz + np.float32(q[k])).astype(np.float32)
    u = (-(zz * acc)).astype(np.float32)
    e = np.exp2(u.astype(np.float64)).astype(np.float32)
    h = (np.float32(0.5) * x).astype(np.float32)
    a = np.abs(h)
    out = ((h + a) - a * e).astype(np.float32)
    ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
    err = np.abs(out - ref)
    print(deg, "max abs gelu err", err.max(), "at x", x[err.argmax()], "max rel-to-fp16ulp", (err / np.maximum(np.abs(ref), 6e-5) / 4.9e-4).max())
    print("  coeffs", ", ".join(f"{v:.9e}f" for v in q))
