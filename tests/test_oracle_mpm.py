"""Self-consistency pins for oracle/mpm_ref.c (the reference MPM cannot run here: warp-lang absent,
no golden vectors in the reference — PARITY UNPINNED at the wp boundary, see mpm_ref.c header)."""
import numpy as np
import pytest

from oracle import mpm_ref as R


@pytest.mark.parametrize("prec,tol", [("f64", 1e-12), ("f32", 5e-6)])
def test_svd3_against_numpy(prec, tol):
    rng = np.random.default_rng(0)
    for i in range(500):
        F = np.eye(3) + 0.6 * rng.standard_normal((3, 3))
        if i % 50 == 0:
            F[:, 2] = F[:, 1] * (1 + 1e-3 * i)        # nearly rank deficient
        U, s, V = R.svd3(F, prec)
        assert np.abs(U @ np.diag(s) @ V.T - F).max() < tol * max(1, np.abs(F).max())
        assert np.abs(np.sort(np.abs(s))[::-1] - np.linalg.svd(F, compute_uv=False)).max() < tol * 10
        assert abs(np.linalg.det(U) - 1) < tol * 10 and abs(np.linalg.det(V) - 1) < tol * 10   # proper rotations
        assert abs(s[0]) >= abs(s[1]) >= abs(s[2]) - tol
        if abs(np.linalg.det(F)) > 1e-3:
            assert np.sign(s[2]) == np.sign(np.linalg.det(F))        # wp.svd3 / McAdams convention


def _scene(n=3000, ng=32, prec="f64", materials=(0,), seed=1, **params):
    sc = R.synthetic_scene(n, ng, seed=seed, materials=materials)
    s = R.MpmRef(n, ng, 2.0, prec)
    for k, f in (("x", "X"), ("v", "V"), ("vol", "VOL"), ("density", "DENSITY"), ("E", "E"), ("nu", "NU"), ("material", "MATERIAL")):
        s.set(f, sc[k])
    s.compute_mass()
    s.compute_mu_lam()
    s.set_params(**params)
    return s, sc


def test_momentum_and_mass_conservation_and_clock():
    s, sc = _scene(g=(0.0, 0.0, -9.8))
    m = s.get("MASS")
    p0 = (m[:, None] * s.get("V")).sum(0)
    s.step(40, 1e-4)
    p1 = (m[:, None] * s.get("V")).sum(0)
    assert np.allclose(p1 - p0, [0, 0, -9.8 * 40e-4 * m.sum()], rtol=1e-7, atol=1e-9)   # APIC conserves momentum
    gm, _, _ = s.grid()
    assert abs(gm.sum() - m.sum()) < 1e-10 * m.sum()                                     # partition of unity
    assert abs(s.time - 40e-4) < 1e-15


def test_rest_state_is_a_fixed_point():
    """v = 0, F = I, no gravity: stress is zero and nothing moves."""
    s, sc = _scene()
    s.set("V", np.zeros((3000, 3)))
    x0 = s.get("X").copy()
    s.step(5, 1e-4)
    assert np.abs(s.get("X") - x0).max() == 0.0
    assert np.abs(s.get("F") - np.eye(3)).max() == 0.0
    assert np.abs(s.get("STRESS")).max() == 0.0


def test_fcr_stress_matches_closed_form():
    """tau = 2 mu (F - R) F^T + lam J (J - 1) I with R from scipy's polar decomposition."""
    from scipy.linalg import polar
    s, sc = _scene(n=64)
    rng = np.random.default_rng(5)
    F = np.eye(3)[None] + 0.2 * rng.standard_normal((64, 3, 3))
    s.set("F_TRIAL", F)
    for p in range(64):
        s.stress_of(p)
    tau = s.get("STRESS")
    mu, lam = s.get("MU"), s.get("LAM")
    for p in range(64):
        Rm, _ = polar(F[p])
        J = np.linalg.det(F[p])
        t = 2 * mu[p] * (F[p] - Rm) @ F[p].T + lam[p] * J * (J - 1) * np.eye(3)
        t = 0.5 * (t + t.T)
        assert np.abs(tau[p] - t).max() < 1e-9 * max(1.0, np.abs(t).max())


def test_sand_return_mapping_cases():
    """Drucker-Prager: expansion (tr eps > 0) projects to the rotation; compression inside the cone is
    unchanged (mpm_utils.py:242-279)."""
    s, sc = _scene(n=8, materials=(2,))
    F = np.stack([np.diag([1.2, 1.1, 1.05])] * 4 + [np.diag([0.98, 0.98, 0.98])] * 4)
    s.set("F_TRIAL", F)
    for p in range(8):
        s.stress_of(p)
    Fe = s.get("F")
    assert np.abs(Fe[:4] - np.eye(3)).max() < 1e-12          # F = U V^T = I for a diagonal stretch
    assert np.abs(Fe[4:] - F[4:]).max() < 1e-12              # isotropic compression: delta_gamma <= 0


def test_von_mises_yield_and_hardening():
    s, sc = _scene(n=4, materials=(1,), hardening=1.0, xi=0.5)
    s.set("YIELD", np.full(4, 1e3))
    F = np.stack([np.diag([1.3, 1.0, 0.8])] * 2 + [np.diag([1.0001, 1.0, 1.0])] * 2)
    s.set("F_TRIAL", F)
    y0 = s.get("YIELD").copy()
    for p in range(4):
        s.stress_of(p)
    Fe, y1 = s.get("F"), s.get("YIELD")
    assert np.abs(Fe[:2] - F[:2]).max() > 1e-2 and (y1[:2] > y0[:2]).all()      # yielded + hardened
    # plastic flow is volume preserving in log-strain space: det unchanged
    assert np.allclose(np.linalg.det(Fe[:2]), np.linalg.det(F[:2]), rtol=1e-6)


def test_bounding_box_and_cuboid_bcs():
    s, sc = _scene(g=(0.0, 0.0, -9.8))
    s.add_bc(R.BC_CUBOID, point=[1.0, 1.0, 1.0], size=[2.0, 2.0, 2.0], velocity=[0.25, 0.0, 0.0], end_time=1e-3)
    x0 = s.get("X").copy()
    s.step(10, 1e-4)
    # every node inside the (huge) cuboid moves with the prescribed velocity
    assert np.allclose(s.get("V"), [0.25, 0.0, 0.0], atol=1e-12)
    assert np.allclose(s.get("X") - x0, [0.25 * 1e-3, 0.0, 0.0], atol=1e-12)


def test_f32_tracks_f64():
    a, _ = _scene(prec="f32", g=(0.0, 0.0, -9.8))
    b, _ = _scene(prec="f64", g=(0.0, 0.0, -9.8))
    a.step(100, 1e-4)
    b.step(100, 1e-4)
    assert np.abs(a.get("X") - b.get("X")).max() < 2e-5
