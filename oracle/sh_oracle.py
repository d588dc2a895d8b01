"""TEST INFRASTRUCTURE — numpy restatement of the spherical-harmonic side of the reference's self-gravity /
shell-pressure term. NOT product code: only tests/, __graft_entry__.smoke() and bench.py may import this.

PARITY UNPINNED for this term: it is commented out at reference HEAD (src/spatialOperators.cpp:387-462,
src/sphericalHarmonics.cpp:1-176, src/mesh.cpp:2154-2260) and depends on SHTOOLS 4.0 (Fortran, un-vendored; PlmBar,
SHExpandLSQ — Makefile:51), so neither the reference nor its dependency can be run here. What is restated, from the
commented code and SHTOOLS' published definitions:
  * PlmBar(lmax, z, csphase=-1)  (src/legendre.f95): 4-pi normalised associated Legendre functions
        Pbar_lm = sqrt((2 - delta_m0)(2l+1)(l-m)!/(l+m)!) P_lm,  P_lm with the Condon-Shortley phase (-1)^m;
  * SHExpandLSQ (src/extractSHCoeffGG.f95): unweighted least squares over the points for all degrees 0..lmax;
  * sh_matrix(i, .) = factor_l Pbar_lm cos/sin(m lon_i), l >= 2 (mesh.cpp:2228-2238) and
    soln += g * sh_matrix * coeffs (spatialOperators.cpp:446).
tests/test_sh_oracle.py pins plm_bar against scipy.special.lpmv (an independent implementation of P_lm).
Row order: degree-major; per degree m = 0, then (cos, sin) for m = 1..l."""
from __future__ import annotations

import math

import numpy as np


def plm_bar(l_max: int, z: np.ndarray) -> np.ndarray:
    """[l_max+1][l_max+1][n] (entries m > l are 0): unnormalised upward recurrences, then the 4-pi factor."""
    z = np.asarray(z, dtype=np.float64)
    u = np.sqrt((1.0 - z) * (1.0 + z))
    P = np.zeros((l_max + 1, l_max + 1) + z.shape)
    for m in range(l_max + 1):
        dfact = 1.0
        for k in range(1, 2 * m, 2):
            dfact *= k                                     # (2m-1)!!
        P[m, m] = (-1.0) ** m * dfact * u ** m
        if m + 1 <= l_max:
            P[m + 1, m] = z * (2 * m + 1) * P[m, m]
        for l in range(m + 2, l_max + 1):
            P[l, m] = ((2 * l - 1) * z * P[l - 1, m] - (l + m - 1) * P[l - 2, m]) / (l - m)
    for l in range(l_max + 1):
        for m in range(l + 1):
            P[l, m] *= math.sqrt((2 - (m == 0)) * (2 * l + 1) * math.factorial(l - m) / math.factorial(l + m))
    return P


def rows(l_max: int) -> int:
    return (l_max + 1) ** 2


def row_degree(l_max: int) -> np.ndarray:
    return np.concatenate([np.full(2 * l + 1, l) for l in range(l_max + 1)])


def basis(pos_sph: np.ndarray, l_max: int) -> np.ndarray:
    """Y [rows][n] at (lat, lon) in radians."""
    lat, lon = np.asarray(pos_sph)[:, 0], np.asarray(pos_sph)[:, 1]
    P = plm_bar(l_max, np.cos(0.5 * np.pi - lat))          # mesh.cpp:2175
    out = []
    for l in range(l_max + 1):
        out.append(P[l, 0])
        for m in range(1, l + 1):
            out.append(P[l, m] * np.cos(m * lon))
            out.append(P[l, m] * np.sin(m * lon))
    return np.array(out)


def lsq_coefficients(Y: np.ndarray, eta: np.ndarray) -> np.ndarray:
    """SHExpandLSQ: argmin |Y^T c - eta|."""
    return np.linalg.lstsq(Y.T, eta, rcond=None)[0]


def apply_operator(Y: np.ndarray, factor) -> np.ndarray:
    """T [rows][rows] with T b = factor_l * coefficients for b = Y eta; rows of degree < 2 are zero."""
    l_max = int(round(math.sqrt(Y.shape[0]))) - 1
    f = np.asarray(factor, dtype=np.float64)[row_degree(l_max)].copy()
    f[:4] = 0.0
    G = (Y.astype(np.longdouble) @ Y.T.astype(np.longdouble)).astype(np.float64)
    return f[:, None] * np.linalg.inv(G)


def self_gravity_potential(Y: np.ndarray, factor, g: float, eta: np.ndarray) -> np.ndarray:
    """g * sum_{l>=2} factor_l sum_m c_lm Y_lm, c from the least-squares fit of eta (independent of apply_operator)."""
    l_max = int(round(math.sqrt(Y.shape[0]))) - 1
    c = lsq_coefficients(Y, eta)
    f = np.asarray(factor, dtype=np.float64)[row_degree(l_max)].copy()
    f[:4] = 0.0
    return g * (Y.T @ (f * c))
