"""GPU parity of the LMMSE smoother against oracle/lmmse.py (LMMSE_ce.m restated, FP64).

Tolerance: the CUDA path solves in FP64 (Cholesky) where the oracle evaluates Rhp*inv(Rpp)*H literally with
LAPACK's inverse; they agree to cond(Rpp)*eps.  rel-L2 <= 1e-9 for complex128 I/O at SNR <= 30 dB, <= 2e-7 for
complex64 I/O (output rounding), 1e-5 (the north_star bound) at 60 dB where cond(Rpp) ~ 1e8."""
import os

import numpy as np
import pytest

import mamimo_b200 as mm
from oracle import lmmse, tables
from _util import rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _h(rng, npkt, nr, nt, nsc, dtype):
    return (rng.standard_normal((npkt, nr, nt, nsc)) + 1j * rng.standard_normal((npkt, nr, nt, nsc))).astype(dtype)


@pytest.mark.parametrize("nt,nr,nsc,npkt", [(32, 4, 234, 3), (4, 2, 52, 2), (8, 1, 64, 1), (3, 2, 100, 2), (64, 2, 300, 1)])
@pytest.mark.parametrize("dtype,tol", [(np.complex128, 1e-9), (np.complex64, 2e-7)])
def test_lmmse_parity_per_slab_parameters(nt, nr, nsc, npkt, dtype, tol):
    """every (packet, rx) has its own SNR(i) and every packet its own tau_rms, as in generate_maMIMO_LTF.m:264,342"""
    rng = np.random.default_rng(nsc)
    H = _h(rng, npkt, nr, nt, nsc, dtype)
    t_rms = rng.uniform(0.5, 6.0, npkt)
    snr = rng.uniform(-5.0, 30.0, (npkt, nr))
    with mm.Engine(nt, nr, nsc, mlp=False) as eng:
        out = eng.lmmse(H, t_rms, snr)
    assert out.dtype == dtype and out.shape == H.shape
    ref = lmmse.lmmse_batched(H, t_rms, snr, 1)
    assert rel_l2(ref, out) <= tol


def test_lmmse_matches_the_literal_matlab_loop():
    """helperMIMOChannelEstimate(..., isMMSE=true) drop-in against the literal double loop (:33-39)."""
    rng = np.random.default_rng(3)
    nt, nr = 4, 2
    car = tables.carriers_locations()
    nsc = car.size
    x = tables.ltf_at_carriers().astype(np.float64)
    Y, _ = mm.synth.make_packets(5, 1, nt, nr, nsc, snr_db=10.0, x_tones=x, dtype=np.complex128)
    rxData = np.transpose(Y[0], (2, 1, 0))                   # MATLAB [Nsc, nltf, Nr]
    tau = np.abs(rng.standard_normal(100)) * 3.0             # the `h` argument of LMMSE_ce
    snr = np.array([8.5, 11.25])
    hD, P, ltf_o, hM = mm.helperMIMOChannelEstimate(rxData, {"numSTS": nt, "CarriersLocations": car}, 1, tau, snr, True)
    ref = lmmse.helper_mmse_loop(hD, 1, tau, snr)
    assert hM.shape == hD.shape == (nsc, nt, nr)
    assert rel_l2(ref, hM) <= 1e-9
    # isMMSE = false leaves hDmmse at zeros (:32)
    _, _, _, z = mm.helperMIMOChannelEstimate(rxData, {"numSTS": nt, "CarriersLocations": car}, 1, tau, snr, False)
    assert not z.any()


def test_lmmse_reference_style_tau_in_seconds_and_high_snr():
    """generate_maMIMO_LTF.m:342 passes path delays in SECONDS as `h` (tau_rms ~ 1e-7: Rpp ~ ones + I/snr), and the
    training set uses SNR = 120 dB (full_pipeline_maMIMO_DNNEst.sh:21); 60 dB is checked against the oracle at the
    north_star bound, 120 dB only for finiteness (cond(Rpp) ~ 1e14: the reference's own inv() is noise there)."""
    rng = np.random.default_rng(4)
    nt, nr, nsc = 8, 2, 234
    H = _h(rng, 2, nr, nt, nsc, np.complex128)
    t_rms = mm.tau_rms(np.abs(rng.standard_normal(100)) * 1e-7)
    with mm.Engine(nt, nr, nsc, mlp=False) as eng:
        out = eng.lmmse(H, t_rms, 10.0)
        assert rel_l2(lmmse.lmmse_batched(H, t_rms, 10.0), out) <= 1e-9
        out60 = eng.lmmse(H, 3.0, 60.0)
        assert rel_l2(lmmse.lmmse_batched(H, 3.0, 60.0), out60) <= 1e-5
        out120 = eng.lmmse(H, t_rms, 120.0)
        assert np.isfinite(out120).all()


def test_lmmse_pilot_spacing_and_chunking_and_device_buffers(monkeypatch):
    import torch
    rng = np.random.default_rng(6)
    nt, nr, nsc, npkt = 4, 2, 96, 9
    H = _h(rng, npkt, nr, nt, nsc, np.complex128)
    snr = rng.uniform(0.0, 20.0, (npkt, nr))
    with mm.Engine(nt, nr, nsc, n_ps=2, mlp=False) as eng:         # Nps = 2: general Rhp path (LMMSE_ce.m:33-36)
        out = eng.lmmse(H, 2.0, snr)
    assert rel_l2(lmmse.lmmse_batched(H, 2.0, snr, 2), out) <= 1e-9
    monkeypatch.setenv("MAMIMO_LMMSE_WS_MB", "1")                  # tiny workspace: several chunks
    with mm.Engine(nt, nr, nsc, mlp=False) as eng:
        a = eng.lmmse(H, 2.0, snr)
        Hd = torch.from_numpy(H).cuda()
        b = eng.lmmse(Hd, 2.0, snr)
        eng.synchronize()
    assert rel_l2(lmmse.lmmse_batched(H, 2.0, snr, 1), a) <= 1e-9
    assert np.array_equal(a, b.cpu().numpy())                      # host and device paths are the same kernels


@pytest.mark.parametrize("nt,nr,nsc,npkt", [(1, 1, 31, 2), (1, 3, 33, 1), (2, 1, 32, 3), (40, 1, 65, 1), (64, 8, 64, 2)])
@pytest.mark.parametrize("route", ["1", "0"])
def test_lmmse_edge_shapes_both_routes(nt, nr, nsc, npkt, route, monkeypatch):
    """block-size edges (Nsc below / at / just above a 32-block), a single right-hand side, more right-hand sides than
    one solve CTA holds (40, 64), for the Toeplitz route and the dense-Cholesky cross-check route"""
    monkeypatch.setenv("MAMIMO_LMMSE_SCHUR", route)
    rng = np.random.default_rng(nsc * 7 + nt)
    H = _h(rng, npkt, nr, nt, nsc, np.complex128)
    t_rms = rng.uniform(0.2, 8.0, npkt)
    snr = rng.uniform(-10.0, 35.0, (npkt, nr))
    with mm.Engine(nt, nr, nsc, mlp=False) as eng:
        out = eng.lmmse(H, t_rms, snr)
    assert rel_l2(lmmse.lmmse_batched(H, t_rms, snr, 1), out) <= 1e-9


def test_lmmse_routes_agree_and_empty_batch():
    rng = np.random.default_rng(12)
    nt, nr, nsc = 8, 2, 100
    H = _h(rng, 3, nr, nt, nsc, np.complex128)
    outs = []
    for route in ("1", "0"):
        os.environ["MAMIMO_LMMSE_SCHUR"] = route
        try:
            with mm.Engine(nt, nr, nsc, mlp=False) as eng:
                outs.append(eng.lmmse(H, 2.0, 15.0))
                assert eng.lmmse(H[:0], 2.0, 15.0).shape == (0, nr, nt, nsc)
        finally:
            os.environ.pop("MAMIMO_LMMSE_SCHUR", None)
    assert rel_l2(outs[1], outs[0]) <= 1e-11


def test_lmmse_workspace_grows_with_the_batch():
    """first call small, later call larger on the same engine (the workspace is re-sized, not chunked by the first size)"""
    rng = np.random.default_rng(13)
    nt, nr, nsc = 4, 2, 48
    with mm.Engine(nt, nr, nsc, mlp=False) as eng:
        for npkt in (1, 40, 3):
            H = _h(rng, npkt, nr, nt, nsc, np.complex128)
            snr = rng.uniform(0.0, 20.0, (npkt, nr))
            l0 = eng.stats()["kernel_launches"]
            out = eng.lmmse(H, 1.5, snr)
            assert rel_l2(lmmse.lmmse_batched(H, 1.5, snr, 1), out) <= 1e-9
            assert eng.stats()["kernel_launches"] - l0 <= 3 * 4          # one chunk: (schur, linv, solve) x <= 4 stream groups
