#!/usr/bin/env python3
"""Derives the FMA-pipe tanh used by the lattice kernel for a fraction of its activations.

    tanh(|h|) ~= hc * q(hc^2),  hc = min(|h|, A),  q = polynomial of degree n in u = hc^2

so that  silu(2h) = h + |h| tanh(|h|) = fma(|h|, hc*q(u), h)  costs n+4 FMA/ALU-pipe instructions and
no MUFU slot.  q minimises the maximum ABSOLUTE error of tanh on [0, A] (Lawson-weighted least
squares in the Chebyshev basis, converted to the monomial basis, then checked with the fp32 Horner
evaluation the kernel performs).  Prints a C initialiser.
"""
import sys

import numpy as np
from numpy.polynomial import chebyshev as C
from numpy.polynomial import polynomial as P


def fit(A: float, n: int, iters: int = 400):
    U = A * A
    k = np.arange(6000)
    x = np.cos(np.pi * (k + 0.5) / 6000)
    u = (x + 1) / 2 * U
    h = np.sqrt(u)
    f = np.where(h > 1e-9, np.tanh(h) / np.maximum(h, 1e-9), 1.0)
    V = C.chebvander(x, n)
    lw = np.ones_like(x)
    best = None
    for _ in range(iters):
        w = h * lw + 1e-12
        cf = np.linalg.lstsq(V * w[:, None], f * w, rcond=None)[0]
        err = np.abs((V @ cf - f) * h)
        m = err.max()
        if best is None or m < best[0]:
            best = (m, cf.copy())
        lw = lw * (err / m + 1e-4)
        lw /= lw.max()
    cheb = best[1]
    # monomial basis in u: x = 2u/U - 1
    mono_x = C.cheb2poly(cheb)
    mono_u = np.zeros(1)
    lin = np.array([-1.0, 2.0 / U])
    pw = np.ones(1)
    for c in mono_x:
        mono_u = P.polyadd(mono_u, c * pw)
        pw = P.polymul(pw, lin)
    return mono_u


def horner_f32(coef, hh):
    hh = hh.astype(np.float32)
    u = hh * hh
    p = np.full_like(u, np.float32(coef[-1]))
    for c in coef[-2::-1]:
        p = (p * u + np.float32(c)).astype(np.float32)  # fp32 rounding of each step (FMA rounds once: this is pessimistic)
    return (hh * p).astype(np.float32)


if __name__ == "__main__":
    A = float(sys.argv[1]) if len(sys.argv) > 1 else 4.2
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 9
    coef = fit(A, n)
    hh = np.linspace(0, A, 400001)
    e32 = np.abs(horner_f32(coef, hh).astype(np.float64) - np.tanh(hh)).max()
    e64 = np.abs(hh * P.polyval(hh * hh, coef) - np.tanh(hh)).max()
    print(f"// tanh(h) ~= h*q(h^2) on |h| <= {A}: degree {n} in h^2, max abs error {e64:.3e} (exact), {e32:.3e} (fp32 Horner); 1-tanh(A) = {1 - np.tanh(A):.2e}")
    print("{" + ", ".join(f"{c:.9e}f" for c in coef) + "}")
