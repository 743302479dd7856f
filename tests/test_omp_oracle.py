"""oracle.omp against the reference's own ompdecomp loop / omphybweights lines (tests/golden/ref_omp.npz, produced by
tests/golden/make_golden.py::run_omp_lines with mini_matlab executing the reference text)."""
import os

import numpy as np
import pytest

from oracle import omp

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_omp.npz"))


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_oracle_matches_reference_loop(tag):
    nt, ns, nrf, nrays, n = (int(v) for v in G["cfg_" + tag])
    At, F = G["At_" + tag], G["Fopt_" + tag]
    for k in range(n):
        fbb, frf, idx, err = omp.precoder_for_subcarrier(F[k], At, nrf)
        assert np.array_equal(idx + 1, G["idx_" + tag][k])                     # MATLAB indices are 1-based
        assert abs(err - G["err_" + tag][k]) <= 1e-13
        assert np.max(np.abs(frf - G["Frf_" + tag][k])) == 0.0                 # dictionary columns, copied
        assert np.max(np.abs(fbb - G["Fbb_" + tag][k])) <= 1e-12 * np.max(np.abs(G["Fbb_" + tag][k]))
        # the normalisation of omphybweights.m:179
        assert abs(np.sqrt(np.sum(np.abs(frf.T @ fbb.T) ** 2)) - np.sqrt(ns)) <= 1e-12


def test_oracle_early_stop_like_reference():
    """Fopt is exactly one dictionary column: the residual is exactly zero after the first pick and the loop stops
    with one atom although MaxSparsity is 3 (ompdecomp.m:105: Errnorm > eps)."""
    coeff, atoms, idx, err = omp.ompdecomp(G["Fopt_e"], G["At_e"], 3)
    assert np.array_equal(idx + 1, G["idx_e"]) and len(idx) == 1
    assert err == 0.0 == float(G["err_e"])
    assert np.array_equal(coeff, G["coef_e"])


def test_indices_do_not_depend_on_fopt_phases():
    """The selection uses sum_s |a^H r_s|^2: a unitary mix of Fopt's columns changes the coefficients, not the atoms
    nor Frf Fbb Fbb^H Frf^H -- which is why the engine can be compared without fixing the SVD's phases."""
    rng = np.random.default_rng(3)
    At, F = G["At_c"], G["Fopt_c"][0]
    q, _ = np.linalg.qr(rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)))
    fb1, fr1, i1, e1 = omp.precoder_for_subcarrier(F, At, 4)
    fb2, fr2, i2, e2 = omp.precoder_for_subcarrier(F @ q, At, 4)
    assert np.array_equal(i1, i2) and abs(e1 - e2) <= 1e-13
    M1, M2 = fr1.T @ fb1.T, fr2.T @ fb2.T
    assert np.max(np.abs(M1 @ M1.conj().T - M2 @ M2.conj().T)) <= 1e-12


def test_batched_form_layout():
    rng = np.random.default_rng(4)
    H = rng.standard_normal((2, 2, 8, 5)) + 1j * rng.standard_normal((2, 2, 8, 5))
    At = np.exp(2j * np.pi * rng.random((8, 30)))
    idx, Fbb, err, Fopt = omp.omp_precoder(H, At, 2, 3)
    assert idx.shape == (2, 3, 5) and Fbb.shape == (2, 2, 3, 5) and err.shape == (2, 5) and Fopt.shape == (2, 2, 8, 5)
    fb, _, ix, e = omp.precoder_for_subcarrier(Fopt[1, :, :, 3].T, At, 3)
    assert np.array_equal(idx[1, :, 3], ix) and np.array_equal(Fbb[1, :, :, 3], fb) and err[1, 3] == e
    inv = omp.precoder_invariant(Fbb, idx, At)
    M = At[:, ix] @ fb.T
    assert np.allclose(inv[1, 3], M @ M.conj().T)
