"""CPU tests of the LMMSE oracle (restates LMMSE_ce.m; parity unpinned by execution: SURVEY 8c) and of the
C-ABI's host-side tau_rms restatement."""
import numpy as np
import pytest

from oracle import lmmse
import mamimo_b200 as mm


def test_tau_rms_closed_forms():
    # two equal taps at k=0 and k=3: mean 1.5, E[k^2] = 4.5 -> sqrt(4.5 - 2.25) = 1.5   (LMMSE_ce.m:27-30)
    assert lmmse.tau_rms([1, 0, 0, 1]) == pytest.approx(1.5)
    assert lmmse.tau_rms([2 + 0j]) == 0.0
    rng = np.random.default_rng(5)
    h = rng.standard_normal(37) + 1j * rng.standard_normal(37)
    p = np.abs(h) ** 2 / np.sum(np.abs(h) ** 2)
    k = np.arange(37)
    assert lmmse.tau_rms(h) == pytest.approx(np.sqrt(np.sum(p * k * k) - np.sum(p * k) ** 2), rel=1e-12)


def test_capi_tau_rms_matches_oracle():
    rng = np.random.default_rng(6)
    for h in (rng.standard_normal(100) * 1e-7, rng.standard_normal(16) + 1j * rng.standard_normal(16), np.array([1.0, 0, 0, 1])):
        assert mm.tau_rms(h) == pytest.approx(lmmse.tau_rms(h), rel=1e-13, abs=1e-300)


def test_rpp_is_hermitian_positive_definite_and_literal_equals_solve():
    rng = np.random.default_rng(7)
    n = 48
    Rhp, Rpp = lmmse.correlation_matrices(n, n, 1, 3.7, 10.0)
    assert np.allclose(Rpp, Rpp.conj().T, atol=1e-15)
    assert np.linalg.eigvalsh(Rpp).min() > 0.05            # >= 1/snr = 0.1 up to rounding
    assert np.allclose(Rhp, Rpp - np.eye(n) * 0.1, atol=1e-15)     # Nps = 1: Rhp = rf2
    h = np.exp(-np.arange(8) / 3.0)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    a = lmmse.lmmse_ce(x, n, n, 1, h, 10.0, literal=True)
    b = lmmse.lmmse_ce(x, n, n, 1, h, 10.0, literal=False)
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-12
    # Nps = 1 identity used by the CUDA path: H_mmse = H - (1/snr) inv(Rpp) H
    Rhp, Rpp = lmmse.correlation_matrices(n, n, 1, lmmse.tau_rms(h), 10.0)
    c = x - 0.1 * np.linalg.solve(Rpp, x)
    assert np.linalg.norm(a - c) / np.linalg.norm(a) < 1e-12


def test_high_snr_tends_to_identity():
    # snr -> inf, Nps = 1, Np = Nfft  =>  Rhp inv(Rpp) -> I (only to ~1e-6: Rpp is ill-conditioned)
    rng = np.random.default_rng(8)
    n = 32
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = lmmse.lmmse_ce(x, n, n, 1, np.exp(-np.arange(8) / 2.0), 200.0, literal=False)
    assert np.linalg.norm(y - x) / np.linalg.norm(x) < 1e-5


def test_loop_form_equals_batched_form():
    rng = np.random.default_rng(9)
    nsc, nt, nr = 24, 4, 2
    hD = rng.standard_normal((nsc, nt, nr)) + 1j * rng.standard_normal((nsc, nt, nr))
    tau = np.abs(rng.standard_normal(20)) * 2
    snr = np.array([3.0, 17.0])
    ref = lmmse.helper_mmse_loop(hD, 1, tau, snr)                      # helperMIMOChannelEstimate.m:33-39 literal
    bat = lmmse.lmmse_batched(np.transpose(hD, (2, 1, 0))[None], lmmse.tau_rms(tau), snr[None], 1)[0]
    assert np.linalg.norm(np.transpose(bat, (2, 1, 0)) - ref) / np.linalg.norm(ref) < 1e-12


def _schur_cholesky(t):
    """numpy statement of the generalised Schur algorithm exactly as csrc/lmmse.cuh's lmmse_schur_kernel runs it:
    generator u = t / sqrt(t0), v = u with v0 = 0; step k emits column k of L, shifts u, applies the hyperbolic
    rotation in mixed-downdating form."""
    n = len(t)
    u = (t / np.sqrt(t[0].real)).astype(np.complex128)
    v = u.copy()
    v[0] = 0
    L = np.zeros((n, n), complex)
    for k in range(n):
        L[k:, k] = u[k:]
        if k == n - 1:
            break
        u[k + 1:] = u[k:n - 1].copy()
        rho = v[k + 1] / u[k + 1]
        s = np.sqrt(1 - abs(rho) ** 2)
        un = (u[k + 1:] - np.conj(rho) * v[k + 1:]) / s
        v[k + 1:] = s * v[k + 1:] - rho * un
        u[k + 1:] = un
    return L


@pytest.mark.parametrize("n,t_rms,snr_db,nps", [(234, 3.0, 10.0, 1), (100, 0.5, 30.0, 2), (64, 1e-7, 10.0, 1), (234, 5.0, 60.0, 1)])
def test_schur_factorisation_is_a_backward_stable_cholesky_of_rpp(n, t_rms, snr_db, nps):
    """Rpp of LMMSE_ce.m:35-38 is Hermitian Toeplitz, so the O(n^2) Schur recursion the CUDA path uses reproduces its
    Cholesky factor: ||L L^H - Rpp|| / ||Rpp|| at rounding level even when cond(Rpp) ~ 1e8."""
    _, Rpp = lmmse.correlation_matrices(n, n, nps, t_rms, snr_db)
    assert np.allclose(Rpp[1:, 1:], Rpp[:-1, :-1], atol=0)                 # Toeplitz: constant along diagonals
    L = _schur_cholesky(Rpp[:, 0].copy())
    assert np.linalg.norm(L @ L.conj().T - Rpp) / np.linalg.norm(Rpp) < 1e-13
    assert np.all(np.diag(L).real > 0) and np.allclose(np.triu(L, 1), 0)
