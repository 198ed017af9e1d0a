import mpmath as mp, numpy as np
mp.mp.dps=50
def wtrue(d):
    d=mp.mpf(d)
    if d==0: return mp.mpf(1)/9  # P(0)
    # z = cos(acos(1-d)/3); w=1-z
    r=1-d
    z=mp.cos(mp.acos(r)/3)
    w=1-z
    # refine by newton on 4w^3-12w^2+9w=d
    for _ in range(3):
        f=((4*w-12)*w+9)*w-d; fp=(12*w-24)*w+9; w-=f/fp
    return w/d
for deg in [16,18,20,22,24]:
    n=deg+1
    nodes=[mp.cos(mp.pi*(k+mp.mpf(1)/2)/n) for k in range(n)]
    fv=[wtrue((t+1)/2) for t in nodes]
    # chebyshev coeffs
    c=[2/mp.mpf(n)*sum(fv[k]*mp.cos(mp.pi*j*(k+mp.mpf(1)/2)/n) for k in range(n)) for j in range(n)]
    c[0]/=2
    # convert to monomial in t
    import numpy.polynomial.chebyshev as C
    # do conversion in mp
    T=[[mp.mpf(1)],[mp.mpf(0),mp.mpf(1)]]
    for j in range(2,n):
        a=[mp.mpf(0)]+[2*x for x in T[-1]]
        b=T[-2]+[mp.mpf(0)]*(len(a)-len(T[-2]))
        T.append([x-y for x,y in zip(a,b)])
    mono=[mp.mpf(0)]*n
    for j in range(n):
        for k,x in enumerate(T[j]): mono[k]+=c[j]*x
    mono64=np.array([float(x) for x in mono])
    # test in fp64
    ds=np.concatenate([np.linspace(0,1,20001),10.0**np.linspace(-12,0,2001)])
    t=2*ds-1
    p=np.zeros_like(t)
    for coef in mono64[::-1]: p=p*t+coef
    ref=np.array([float(wtrue(d)) for d in ds[::50]])
    err=np.abs(p[::50]/ref-1).max()
    print(deg, err, float(abs(c[-1])))
    if deg==22:
        print(','.join('%.17e'%x for x in mono64))
