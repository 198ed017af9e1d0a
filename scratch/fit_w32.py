"""Single-precision seed polynomial for w(delta)/delta (see gfp_cubic_w in gf_physics.cuh):
Chebyshev interpolation in t = 2 delta - 1, monomial coefficients rounded to fp32, error measured
with fp32 Horner evaluation."""
import mpmath as mp, numpy as np
mp.mp.dps = 40
def wtrue(d):
    d = mp.mpf(d)
    if d == 0: return mp.mpf(1) / 9
    w = 1 - mp.cos(mp.acos(1 - d) / 3)
    for _ in range(3):
        f = ((4 * w - 12) * w + 9) * w - d; fp = (12 * w - 24) * w + 9; w -= f / fp
    return w / d
for deg in (4, 5, 6, 7, 8):
    n = deg + 1
    nodes = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / n) for k in range(n)]
    fv = [wtrue((t + 1) / 2) for t in nodes]
    c = [2 / mp.mpf(n) * sum(fv[k] * mp.cos(mp.pi * j * (k + mp.mpf(1) / 2) / n) for k in range(n)) for j in range(n)]
    c[0] /= 2
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for j in range(2, n):
        a = [mp.mpf(0)] + [2 * x for x in T[-1]]
        b = T[-2] + [mp.mpf(0)] * (len(a) - len(T[-2]))
        T.append([x - y for x, y in zip(a, b)])
    mono = [mp.mpf(0)] * n
    for j in range(n):
        for k, x in enumerate(T[j]): mono[k] += c[j] * x
    m32 = np.array([float(x) for x in mono], dtype=np.float32)
    ds = np.concatenate([np.linspace(0, 1, 4001), 10.0 ** np.linspace(-9, 0, 901)])
    t = (2 * ds - 1).astype(np.float32)
    p = np.zeros_like(t)
    for coef in m32[::-1]: p = (p * t + coef).astype(np.float32)
    ref = np.array([float(wtrue(d)) for d in ds])
    print(deg, 'max rel err fp32 eval', np.abs(p.astype(np.float64) / ref - 1).max())
    if deg in (6, 7):
        print(', '.join('%.9ef' % x for x in m32))
