"""Fit and check of the polynomial the loss kernel uses for log1p(e), e = exp(-|x|) in (0, 1]:
    log1p(e) = e * (1 + e * Q(e)),  Q of degree 7 fitted at Chebyshev nodes of [0, 1].
Prints the fp32 coefficients (ODK_Q0..ODK_Q7 in csrc/odk_loss.cu), the maximum relative error of the fp32
Horner evaluation against float64 log1p, and the relative error of the SUM of softplus over N(-4.6, 1.5)
and N(-7, 1) logits (what the 1e-5 loss tolerance is about).  CPU only (numpy)."""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P


def g(e):
    e = np.asarray(e, dtype=np.float64)
    out = np.empty_like(e)
    small = e < 1e-3
    es = e[small]
    out[small] = -0.5 + es / 3 - es ** 2 / 4 + es ** 3 / 5
    eb = e[~small]
    out[~small] = (np.log1p(eb) / eb - 1) / eb
    return out


def fit(deg=7):
    x = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000)
    e = (x + 1) / 2
    p = C.cheb2poly(C.chebfit(x, g(e), deg))
    pe, pw = np.zeros(1), np.ones(1)
    for ck in p:                      # substitute x = 2e - 1
        pe = P.polyadd(pe, ck * pw)
        pw = P.polymul(pw, np.array([-1.0, 2.0]))
    return pe.astype(np.float32)


def log1p_poly(e32, q):
    acc = np.full_like(e32, q[-1])
    for ck in q[-2::-1]:
        acc = (acc * e32 + ck).astype(np.float32)
    acc = (acc * e32 + np.float32(1)).astype(np.float32)
    return (acc * e32).astype(np.float32)


if __name__ == '__main__':
    q = fit()
    for k, v in enumerate(q):
        print(f'ODK_Q{k} = {float(v)!r}')
    e = np.linspace(0, 1, 400001)[1:].astype(np.float32)
    rel = (log1p_poly(e, q) - np.log1p(e.astype(np.float64))) / np.log1p(e.astype(np.float64))
    print(f'max |relative error| on (0,1]: {np.abs(rel).max():.3e}   mean signed: {rel.mean():.3e}')
    rs = np.random.RandomState(0)
    for mu, sd in ((-4.6, 1.5), (-7.0, 1.0), (0.0, 3.0)):
        x = rs.normal(mu, sd, 4_000_000).astype(np.float32)
        ee = np.exp(-np.abs(x).astype(np.float64)).astype(np.float32)
        sp = np.maximum(x, 0).astype(np.float64) + log1p_poly(ee, q).astype(np.float64)
        ref = np.maximum(x.astype(np.float64), 0) + np.log1p(np.exp(-np.abs(x.astype(np.float64))))
        print(f'logits N({mu},{sd}): relative error of sum(softplus) = {(sp.sum() - ref.sum()) / ref.sum():.3e}')
