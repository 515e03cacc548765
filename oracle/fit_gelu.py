"""Test infrastructure (not product code): the erf-GELU polynomial of the GEMM epilogues, its fit and its error sweep.

The GELU / GELU' epilogues of csrc/gemm_tc_kernel.cuh (``phi2``) evaluate the normal CDF of the reference's exact
``nn.GELU()`` (common.py:13,21: ``0.5 x (1 + erf(x / sqrt 2))``) without MUFU as

    Phi(x) = 0.5 + xc * Q(xc^2),   xc = clamp(x, -3 sqrt 2, +3 sqrt 2),   Q = degree-8 polynomial.

``fit()`` re-derives a table of the same quality (near-minimax: Lawson-reweighted least squares of Phi(x) - 0.5 =
x Q(x^2) on Chebyshev nodes, float64);
``COEFFS`` are the constants compiled into the kernel (highest degree first, Horner order); ``sweep()`` evaluates the
float32 Horner form exactly as the kernel does and returns the worst |Phi| and |gelu| errors over [-8, 8].

    python oracle/fit_gelu.py          # prints the fitted coefficients next to the compiled ones and both errors
"""
from __future__ import annotations

import math

import numpy as np

Z = 4.242640687  # 3 sqrt 2
# highest degree first: q = c0; q = q * s + c1; ...  (csrc/gemm_tc_kernel.cuh: phi2)
COEFFS = (5.6236895431e-11, -5.3744284878e-09, 2.2710010238e-07, -5.6547267380e-06, 9.3721011908e-05,
          -1.1104664642e-03, 9.8226745766e-03, -6.6355885986e-02, 3.9890877892e-01)


def _phi(x: np.ndarray) -> np.ndarray:
    return 0.5 * (1.0 + np.vectorize(math.erf)(x / math.sqrt(2.0)))


def fit(degree: int = 8, nodes: int = 4001, lawson_iters: int = 60) -> np.ndarray:
    """Near-minimax fit of Phi(x) - 0.5 = x * Q(x^2) on (0, Z]: weighted least squares on Chebyshev nodes with Lawson's
    reweighting (weights grow where the error is largest), which is how the compiled table was obtained."""
    k = np.arange(nodes)
    x = 0.5 * Z * (1.0 + np.cos(np.pi * (k + 0.5) / nodes))
    x = x[x > 1e-6]
    y = _phi(x) - 0.5
    smax = Z * Z
    V = np.vander(x * x / smax, degree + 1) * x[:, None]  # columns: x * (s / smax)^p, highest degree first
    w = np.ones_like(x)
    best, best_err = None, np.inf
    for _ in range(lawson_iters):
        sw = np.sqrt(w)
        c, *_ = np.linalg.lstsq(V * sw[:, None], y * sw, rcond=None)
        err = np.abs(V @ c - y)
        if err.max() < best_err:
            best, best_err = c, err.max()
        w = w * (err / err.max() + 1e-3)
        w /= w.sum()
    return best / smax ** np.arange(degree, -1, -1)


def phi_poly_f32(x: np.ndarray, coeffs=COEFFS) -> np.ndarray:
    """float32 Horner evaluation, operation for operation what phi2() does (fma rounding differences aside)."""
    x = x.astype(np.float32)
    xc = np.clip(x, np.float32(-Z), np.float32(Z))
    s = xc * xc
    q = np.full_like(xc, np.float32(coeffs[0]))
    for c in coeffs[1:]:
        q = q * s + np.float32(c)
    return xc * q + np.float32(0.5)


def sweep(lo: float = -8.0, hi: float = 8.0, n: int = 400001, coeffs=COEFFS):
    x = np.linspace(lo, hi, n)
    ref = _phi(x)
    got = phi_poly_f32(x, coeffs).astype(np.float64)
    phi_err = float(np.max(np.abs(got - ref)))
    gelu_err = float(np.max(np.abs(x * got - x * ref)))
    return phi_err, gelu_err


if __name__ == "__main__":
    c = fit()
    print("fitted (float64, Lawson-reweighted least squares) vs compiled:")
    for a, b in zip(c, COEFFS):
        print(f"  {a: .10e}   {b: .10e}")
    pe, ge = sweep()
    print(f"compiled coefficients, float32 Horner: max |Phi err| {pe:.3e}, max |gelu err| {ge:.3e} on [-8, 8]")
    pe2, ge2 = sweep(coeffs=tuple(c))
    print(f"re-fitted coefficients:               max |Phi err| {pe2:.3e}, max |gelu err| {ge2:.3e}")
