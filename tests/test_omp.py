"""OMP hybrid precoder on the GPU (pg/omphybweights.m:178-179 + pg/ompdecomp.m:101-121) against the oracle
that tests/test_omp_oracle.py pins on the reference text: chosen dictionary columns bit-exact (integers), residual
norms and coefficients in FP64, the early stop, and the end-to-end H -> SVD -> OMP chain through its
phase-independent products."""
import os

import numpy as np
import pytest

import mamimo_b200 as mm
from oracle import omp as oomp

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_omp.npz"))


def _engine_F(F):
    """golden Fopt [n, Nt, Ns] (one matrix per subcarrier) -> engine layout [1, Ns, Nt, n]"""
    return np.ascontiguousarray(np.transpose(F, (2, 1, 0))[None])


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_golden_cases_same_fopt(tag):
    """Same Fopt on both sides: indices identical to the reference loop's, Fbb and Errnorm to FP64 accuracy."""
    nt, ns, nrf, nrays, n = (int(v) for v in G["cfg_" + tag])
    with mm.Engine(nt, 1, n, mlp=False) as eng:
        eng.set_steering_dictionary(G["At_" + tag])
        idx, err, Fbb = eng.omp(_engine_F(G["Fopt_" + tag]), ns, nrf)
    assert idx.dtype == np.int32 and idx.shape == (1, nrf, n) and Fbb.shape == (1, ns, nrf, n)
    assert np.array_equal(idx[0].T + 1, G["idx_" + tag])                               # MATLAB is 1-based
    assert np.max(np.abs(err[0, -1] - G["err_" + tag])) <= 1e-6                        # float32 output
    ref = np.transpose(G["Fbb_" + tag], (1, 2, 0))                                     # [n, Ns, NtRF] -> [Ns, NtRF, n]
    assert np.max(np.abs(Fbb[0] - ref)) <= 1e-11 * np.max(np.abs(ref))
    assert np.all(np.diff(err[0], axis=0) <= 1e-7)                                     # the residual never grows


def test_ns1_kernels_agree(monkeypatch):
    """Ns = 1 runs a taller-register-block kernel; the general one must pick the same columns (MAMIMO_OMP_GENERIC)."""
    rng = np.random.default_rng(12)
    nt, nsc = 32, 300
    F = (rng.standard_normal((2, 1, nt, nsc)) + 1j * rng.standard_normal((2, 1, nt, nsc)))
    At = np.exp(2j * np.pi * rng.random((nt, 500)))
    outs = []
    for generic in (False, True):
        if generic:
            monkeypatch.setenv("MAMIMO_OMP_GENERIC", "1")
        with mm.Engine(nt, 1, nsc, mlp=False) as eng:
            eng.set_steering_dictionary(At)
            outs.append(eng.omp(F, 1, 3))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])


def test_early_stop_exact_residual():
    """ompdecomp.m:105: the loop ends when the residual norm is <= eps.  Entries in {1, j, -1, -j}: exact arithmetic."""
    At, Fe = G["At_e"], G["Fopt_e"]
    nt = At.shape[0]
    F = np.zeros((2, 1, nt, 3), dtype=np.complex128)
    F[:, 0, :, :] = (np.random.default_rng(0).standard_normal((2, nt, 3)) + 0.3j)      # ordinary tones around it
    F[1, 0, :, 1] = Fe[:, 0]
    with mm.Engine(nt, 1, 3, mlp=False) as eng:
        eng.set_steering_dictionary(At)
        idx, err, Fbb = eng.omp(F, 1, 3)
    assert idx[1, 0, 1] + 1 == int(G["idx_e"][0]) and idx[1, 1, 1] == -1 and idx[1, 2, 1] == -1
    assert np.all(err[1, :, 1] == 0.0)
    assert np.all(idx[0] >= 0) and np.all(idx[1, :, 0] >= 0) and np.all(idx[1, :, 2] >= 0)
    for p in range(2):
        for k in range(3):
            fb, _, ix, e = oomp.precoder_for_subcarrier(F[p, :, :, k].T, At, 3)
            assert np.array_equal(idx[p, :len(ix), k], ix)
            assert np.max(np.abs(Fbb[p, :, :len(ix), k] - fb)) <= 1e-11
            assert np.all(Fbb[p, :, len(ix):, k] == 0)


@pytest.mark.parametrize("nt,nr,nsc,npkt,ns,nrf,nrays,ctype", [
    (32, 4, 234, 2, 1, 1, 500, np.complex64),        # the reference's own use: numSTS = 1, 500 rays (BER_test :59,372)
    (32, 4, 1024, 2, 2, 4, 500, np.complex128),
    (8, 2, 100, 3, 2, 3, 77, np.complex64),
    (64, 8, 130, 1, 4, 8, 200, np.complex64),
    (6, 3, 70, 2, 3, 2, 64, np.complex128)])
def test_chain_from_channel(nt, nr, nsc, npkt, ns, nrf, nrays, ctype):
    """H -> svd -> omp on the engine against the oracle's own chain.  The two SVDs differ by a phase per vector, so the
    comparison runs on what does not depend on it: the chosen columns, the residual norms, Frf Fbb Fbb^H Frf^H."""
    rng = np.random.default_rng(5 + nt)
    H = (rng.standard_normal((npkt, nr, nt, nsc)) + 1j * rng.standard_normal((npkt, nr, nt, nsc))).astype(ctype)
    At = np.exp(2j * np.pi * rng.random((nt, nrays)))
    with mm.Engine(nt, nr, nsc, mlp=False) as eng:
        eng.set_steering_dictionary(At)
        idx, err, Fbb = eng.omp_precoder(H, ns, nrf)
    r_idx, r_fbb, r_err, _ = oomp.omp_precoder(H, At, ns, nrf)
    same = np.all(idx == r_idx, axis=1)                                # per (pkt, tone)
    if ctype == np.complex128:
        assert same.all()
    else:
        # complex64 vectors out of the engine's SVD carry ~1e-7 of noise: a tone whose two best columns are closer
        # than that may flip; everything else must agree
        assert same.mean() >= 0.995
    assert np.max(np.abs(err[:, -1][same] - r_err[same])) <= (2e-6 if ctype == np.complex64 else 1e-6)
    inv, r_inv = oomp.precoder_invariant(Fbb.astype(np.complex128), idx, At), oomp.precoder_invariant(r_fbb, r_idx, At)
    d = np.linalg.norm((inv - r_inv)[same]) / np.linalg.norm(r_inv[same])
    assert d <= (5e-6 if ctype == np.complex64 else 1e-9)
    # omphybweights.m:179: ||Frf*Fbb||_F = sqrt(Ns) for every tone
    tr = np.real(np.trace(inv, axis1=-2, axis2=-1))
    assert np.max(np.abs(tr - ns)) <= (1e-5 if ctype == np.complex64 else 1e-10)


def test_device_buffers_equal_host_buffers():
    import torch
    rng = np.random.default_rng(9)
    nt, nr, nsc = 32, 4, 234
    H = (rng.standard_normal((3, nr, nt, nsc)) + 1j * rng.standard_normal((3, nr, nt, nsc))).astype(np.complex64)
    At = np.exp(2j * np.pi * rng.random((nt, 300)))
    with mm.Engine(nt, nr, nsc, mlp=False) as eng:
        eng.set_steering_dictionary(At)
        _, V = eng.svd(H)
        a = eng.omp(V, 2, 3)
        b = eng.omp(torch.from_numpy(V).cuda(), 2, 3)
        torch.cuda.synchronize()
    for x, y in zip(a, b):
        assert np.array_equal(x, y.cpu().numpy())


def test_errors():
    with mm.Engine(8, 2, 64, mlp=False) as eng:
        F = np.zeros((1, 2, 8, 64), dtype=np.complex128)
        with pytest.raises(mm.MamimoError):
            eng.omp(F, 1, 1)                                       # no dictionary yet
        with pytest.raises(ValueError):
            eng.set_steering_dictionary(np.ones((7, 10)))          # wrong n_tx
        bad = np.ones((8, 10), dtype=np.complex128)
        bad[3, 4] = np.nan
        with pytest.raises(mm.MamimoError):
            eng.set_steering_dictionary(bad)
        eng.set_steering_dictionary(np.exp(2j * np.pi * np.random.default_rng(0).random((8, 10))))
        with pytest.raises(mm.MamimoError):
            eng.omp(F, 1, 11)                                      # more RF chains than dictionary columns / > 8
        with pytest.raises(ValueError):
            eng.omp(F, 3, 1)                                       # ns > rows of F


def test_full_size_properties_and_sampled_oracle():
    """BASELINE configs[1] shape with the reference's dictionary size (32x4x1024, 500 rays, Ns = 1, NtRF = 4 to exercise
    the rounds): size-independent properties on every tone -- columns never repeat (the residual is orthogonal to the
    chosen ones), the residual norm never grows, ||Frf*Fbb||_F = sqrt(Ns) -- and the oracle on a sample of tones."""
    import torch
    nt, nr, nsc, npkt, ns, nrf, nrays = 32, 4, 1024, 24, 1, 4, 500
    rng = np.random.default_rng(2024)
    _, H = mm.synth.make_packets(75, npkt, nt, nr, nsc, snr_db=10.0, dtype=np.complex128)
    At = np.exp(2j * np.pi * rng.random((nt, nrays)))
    with mm.Engine(nt, nr, nsc, mlp=False, max_pkts=npkt) as eng:
        eng.set_steering_dictionary(At)
        Hd = torch.from_numpy(H).cuda()
        _, V = eng.svd(Hd)
        idx, err, Fbb = (t.cpu().numpy() for t in eng.omp(V, ns, nrf))
        V = V.cpu().numpy()
    assert idx.min() >= 0 and idx.max() < nrays
    srt = np.sort(idx, axis=1)
    assert np.all(np.diff(srt, axis=1) > 0)                                  # four distinct columns per tone
    assert np.all(np.diff(err, axis=1) <= 1e-6)
    A = At[:, idx]                                                           # [nt, npkt, nrf, nsc]
    M = np.einsum("tpjk,psjk->ptsk", A, Fbb)                                 # Frf*Fbb per tone: [npkt, nt, ns, nsc]
    assert np.max(np.abs(np.sqrt(np.sum(np.abs(M) ** 2, axis=(1, 2))) - np.sqrt(ns))) <= 1e-10
    for p, k in zip(rng.integers(0, npkt, 48), rng.integers(0, nsc, 48)):
        fb, _, ix, e = oomp.precoder_for_subcarrier(V[p, :ns, :, k].T, At, nrf)
        assert np.array_equal(idx[p, :, k], ix) and abs(err[p, -1, k] - e) <= 1e-6
        assert np.max(np.abs(Fbb[p, :, :, k] - fb)) <= 1e-11
