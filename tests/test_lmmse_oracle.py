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
