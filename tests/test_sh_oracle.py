"""Spherical-harmonic self-gravity term, CPU side: the numpy oracle (oracle/sh_oracle.py) against an independent
Legendre implementation (scipy), the product's host-side basis / normal inverse (C ABI, no GPU) against the oracle, and the
domain's own size-independent properties (orthonormality, band-limited round trip).

The term is commented out at reference HEAD and needs SHTOOLS: parity for it is UNPINNED (oracle/sh_oracle.py header)."""
import math

import numpy as np
import pytest

from oracle import sh_oracle as so


def random_points(n, seed=1):
    rng = np.random.default_rng(seed)
    return np.stack([np.arcsin(rng.uniform(-1, 1, n)), rng.uniform(0, 2 * np.pi, n)], 1)


def test_oracle_legendre_matches_scipy():
    sp = pytest.importorskip("scipy.special")
    pos = random_points(400)
    z = np.cos(0.5 * np.pi - pos[:, 0])
    L = 16
    P = so.plm_bar(L, z)
    for l in range(L + 1):
        for m in range(l + 1):
            norm = math.sqrt((2 - (m == 0)) * (2 * l + 1) * math.factorial(l - m) / math.factorial(l + m))
            ref = sp.lpmv(m, l, z) * norm               # lpmv carries the Condon-Shortley phase, as csphase = -1 does
            assert np.abs(P[l, m] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), (l, m)


def test_known_values():
    # Pbar_20 = sqrt(5) (3 z^2 - 1)/2 ; Pbar_22 cos(2 lon) = sqrt(15)/2 cos^2(lat) cos(2 lon) ... with (-1)^2 = +1;
    # Pbar_21 = -sqrt(15) z sqrt(1-z^2)  (phase (-1)^1)
    pos = np.array([[0.3, 1.1], [-1.0, 4.0]])
    Y = so.basis(pos, 2)
    z, c = np.sin(pos[:, 0]), np.cos(pos[:, 0])
    assert np.allclose(Y[4], math.sqrt(5) * (3 * z * z - 1) / 2, rtol=1e-14)
    assert np.allclose(Y[5], -math.sqrt(15) * z * c * np.cos(pos[:, 1]), rtol=1e-14)
    assert np.allclose(Y[7], math.sqrt(15) / 2 * c * c * np.cos(2 * pos[:, 1]), rtol=1e-14)
    assert np.allclose(Y[8], math.sqrt(15) / 2 * c * c * np.sin(2 * pos[:, 1]), rtol=1e-14)


@pytest.mark.parametrize("l_max", [2, 8, 20])
def test_product_basis_matches_oracle(odis, l_max):
    pos = random_points(1000, seed=l_max)
    Y = odis.sh_basis(pos, l_max)
    assert Y.shape == (so.rows(l_max), 1000)
    ref = so.basis(pos, l_max)
    assert np.abs(Y - ref).max() <= 1e-12 * np.abs(ref).max()


def test_basis_is_orthonormal_on_the_geodesic_grid(odis):
    """4-pi normalisation: the area-weighted mean of Y_a Y_b over the sphere is delta_ab."""
    pos, fr, cen = odis.generate_grid(6)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, 1.0)
    w = mesh.tables["control_volume_surf_area_map"] / mesh.tables["control_volume_surf_area_map"].sum()
    Y = odis.sh_basis(pos, 6)
    gram = (Y * w) @ Y.T
    assert np.abs(gram - np.eye(Y.shape[0])).max() < 2e-3


@pytest.mark.parametrize("l_max", [2, 5, 12])
def test_normal_inverse_and_band_limited_round_trip(odis, l_max):
    pos, fr, cen = odis.generate_grid(5)
    Y = odis.sh_basis(pos, l_max)
    Gi = odis.sh_normal_inverse(pos, l_max)
    R = so.rows(l_max)
    assert np.abs(Gi @ (Y @ Y.T) - np.eye(R)).max() <= 1e-11
    rng = np.random.default_rng(3)
    c = rng.uniform(-1, 1, R)
    eta = Y.T @ c
    assert np.abs(Gi @ (Y @ eta) - c).max() <= 1e-12                    # analysis of a band-limited field is exact
    assert np.abs(so.lsq_coefficients(so.basis(pos, l_max), eta) - c).max() <= 1e-12
    # the operator the oracle time loop uses = factor * coefficients, degrees 0 and 1 dropped
    factor = 0.5 / (1.0 + np.arange(l_max + 1))
    T = so.apply_operator(so.basis(pos, l_max), factor)
    f = factor[so.row_degree(l_max)].copy(); f[:4] = 0
    assert np.abs(T @ (Y @ eta) - f * c).max() <= 1e-12
    u = so.self_gravity_potential(so.basis(pos, l_max), factor, 0.113, eta)
    assert np.abs(u - 0.113 * (Y.T @ (f * c))).max() <= 1e-12 * np.abs(u).max() + 1e-15


def test_argument_errors(odis):
    pos = random_points(10)
    with pytest.raises(odis.OdisError):
        odis.sh_basis(pos, 40)
    with pytest.raises(odis.OdisError):
        odis.sh_normal_inverse(pos[:3], 2)          # 3 points cannot fix 9 coefficients


def test_c_oracle_term_matches_numpy_restatement(odis):
    """oracle/lte_oracle.c: self_gravity() (what the GPU parity tests compare with) against oracle/sh_oracle.py computed
    independently (numpy lstsq instead of the normal-matrix operator): potential 'NONE', so forcing_potential is the term alone."""
    from oracle.lte_oracle import LteOracle
    pos, fr, cen = odis.generate_grid(4)
    r = 1.0e6
    mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
    prm = dict(g=1.3, h=1.0e3, alpha=1e-6, dt=5.0, radius=r, omega=2e-5, love_reduct=1.0, ecc=0.0, obl=0.0, shell_thickness=0.0,
               potential=16, friction=0, surface=0, init_load=0)
    l_max = 5
    factor = np.linspace(0.9, 0.1, l_max + 1)
    Y = so.basis(pos, l_max)
    o = LteOracle(mesh.tables, prm)
    o.set_self_gravity(Y, so.apply_operator(Y, factor))
    eta = np.random.default_rng(9).uniform(-1, 1, mesh.n_cells)
    o.set_state(eta=eta)
    o.step(1)                                              # forcing_potential of this step is built from eta as loaded
    ref = so.self_gravity_potential(Y, factor, prm["g"], eta)
    got = o.field(6)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
