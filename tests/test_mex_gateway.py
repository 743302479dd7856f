"""The MATLAB MEX gateway (csrc/mex_gateway.cpp), EXECUTED: mexFunction is linked against the mini MEX runtime
(csrc/mex_runtime: mxArray with column-major dims / interleaved complex / structs, mexErrMsgIdAndTxt leaving the
function like MATLAB's long jump, mexAtExit) and driven through tests/_mex.py.  Signature kept:
[hD,P,ltf_o,hDmmse] = helperMIMOChannelEstimate(rxData,prm,Nps,tau,SNR,isMMSE) (pg/helperMIMOChannelEstimate.m:1,31-41)
via the .m shim's sequence of gateway commands."""
import os

import numpy as np
import pytest

import mamimo_b200 as mm
from oracle import tables, mlp, ofdm
from _mex import Mex, MexError
from _util import oracle_full, rel_l2


@pytest.fixture(scope="module")
def mex():
    m = Mex()
    yield m
    m.clear_mex()


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ------------------------------------------------------------------------------------ no GPU needed
def test_gateway_usage_errors_and_table_without_engine(mex):
    mex.clear_mex()
    ltf = mex.call("ltf", nlhs=1)                          # helperMIMOChannelEstimate.m:16-23
    assert ltf.shape == (256, 1) and ltf.dtype == np.float64 and np.array_equal(ltf[:, 0], tables.vht_ltf256())
    for args, ident in ((("ls", np.zeros((4, 2, 2), np.complex128)), "mamimo:state"),       # engine not created
                        (("finalize",), "mamimo:state"),
                        (("frobnicate",), "mamimo:usage"),
                        ((np.zeros(3),), "mamimo:usage"),                                   # first argument must be a string
                        (("create",), "mamimo:usage"),
                        (("create", np.zeros(3)), "mamimo:usage")):                         # cfg must be a struct
        with pytest.raises(MexError) as ei:
            mex.call(*args, nlhs=1)
        assert ei.value.identifier == ident, args
    assert "frobnicate" in str(ei.value) or True
    if not _has_gpu():
        with pytest.raises(MexError) as ei:                # no device: the engine's own error text comes through
            mex.call("create", {"n_tx": 4, "n_rx": 2, "n_sc": 234})
        assert ei.value.identifier == "mamimo:engine" and "no CUDA device" in ei.value.message
        assert mex.atexit_count() == 0                     # nothing was registered for a failed create
    mex.call("destroy")                                    # harmless without an engine


# ------------------------------------------------------------------------------------ on the B200
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["A", "B", "C"])
def test_gateway_ls_and_lmmse_reproduce_the_matlab_golden(mex, golden_dir, case):
    """create -> pilots -> ls -> lmmse on the inputs of tests/golden/ref_matlab_ls_lmmse.npz (produced by the
    unmodified helperMIMOChannelEstimate.m / LMMSE_ce.m), MATLAB shapes and column-major data in and out."""
    g = np.load(os.path.join(golden_dir, "ref_matlab_ls_lmmse.npz"))
    rx, P, ltf_o = g["rx_" + case], g["P_" + case], g["ltf_o_" + case]
    nsc, nltf, nrx = rx.shape
    mex.call("create", {"n_tx": P.shape[0], "n_rx": nrx, "n_sc": nsc, "n_ltf": nltf})
    assert mex.atexit_count() == 1
    mex.call("pilots", ltf_o, P)
    hD = mex.call("ls", rx, nlhs=1)
    assert hD.shape == g["hD_" + case].shape and hD.dtype == np.complex128
    assert rel_l2(g["hD_" + case], hD) <= 1e-12                  # complex double through 'ls': FP64 on the device
    has_mmse = bool(g["hDmmse_" + case].any())            # case B ran with isMMSE = false: hDmmse stays zeros (:32)
    hM = mex.call("lmmse", g["hD_" + case], g["tau_" + case], g["snr_" + case], nlhs=1)
    assert hM.shape == hD.shape and np.isfinite(hM).all()
    if has_mmse:
        assert rel_l2(g["hDmmse_" + case], hM) <= 1e-9
    # a batch hoisted out of the packet loop: [Nsc x nltf x Nr x Npkt], per-packet tau columns and SNR columns
    rx4 = np.stack([rx, 0.5j * rx], axis=3)
    h4 = mex.call("ls", rx4, nlhs=1)
    assert h4.shape == hD.shape + (2,) and rel_l2(g["hD_" + case], h4[..., 0]) <= 1e-6 and rel_l2(0.5j * g["hD_" + case], h4[..., 1]) <= 1e-6
    hd4 = np.stack([g["hD_" + case], g["hD_" + case]], axis=3)
    tau2 = np.stack([g["tau_" + case].ravel(), g["tau_" + case].ravel()], axis=1)
    snr2 = np.concatenate([g["snr_" + case].reshape(-1, 1)] * 2, axis=1)
    hm4 = mex.call("lmmse", hd4, tau2, snr2, nlhs=1)
    assert rel_l2(hM, hm4[..., 1]) <= 1e-12 and rel_l2(hM, hm4[..., 0]) <= 1e-12
    if has_mmse:
        assert rel_l2(g["hDmmse_" + case], hm4[..., 1]) <= 1e-9
    # shape errors leave through mexErrMsgIdAndTxt and the engine stays usable afterwards
    with pytest.raises(MexError) as ei:
        mex.call("ls", rx[:-1], nlhs=1)
    assert ei.value.identifier == "mamimo:size"
    with pytest.raises(MexError) as ei:
        mex.call("ls", rx.real.copy(), nlhs=1)
    assert ei.value.identifier == "mamimo:type"
    assert rel_l2(g["hD_" + case], mex.call("ls", rx, nlhs=1)) <= 1e-6
    mex.clear_mex()                                        # `clear mex`: the mexAtExit handler destroys the engine
    assert mex.atexit_count() == 0
    with pytest.raises(MexError) as ei:
        mex.call("ls", rx, nlhs=1)
    assert ei.value.identifier == "mamimo:state"


@pytest.mark.gpu
def test_gateway_full_path_estimate_and_ofdm(mex):
    """create(hidden, d_out) -> pilots -> load x6 -> finalize -> [hD, Hr, Hi] = estimate(rxData batch); then the
    time-domain front end: ofdm(...) -> Y = demod(x)."""
    nt, nr, nsc, npkt, hidden = 8, 2, 234, 3, (64, 48)
    xp = tables.ltf_at_carriers().astype(np.float64)
    P = tables.sylvester_hadamard(nt)
    nets = mm.synth.make_nets(nsc, hidden, nsc)
    Y, _ = mm.synth.make_packets(61, npkt, nt, nr, nsc, snr_db=10.0, x_tones=xp, dtype=np.complex128)
    rx = np.transpose(Y, (3, 2, 1, 0))                                     # MATLAB [Nsc x nltf x Nr x Npkt]
    mex.call("create", {"n_tx": nt, "n_rx": nr, "n_sc": nsc, "hidden": np.array(hidden, np.float64), "d_out": nsc})
    mex.call("pilots", xp.reshape(-1, 1), P)
    for net, name in enumerate(("real", "imag")):
        for li, L in enumerate(nets[name]):
            # MATLAB passes W.' so that column-major memory is the Keras kernel [in][out] row-major
            args = [np.asarray(L["W"], np.float32).T, np.asarray(L["b"], np.float32)]
            if L["bn"] is not None:
                args += [np.asarray(t, np.float32) for t in L["bn"]]
            mex.call("load", float(net), float(li), *args)
    with pytest.raises(MexError) as ei:                                    # weights must be single (mamimo:type)
        mex.call("load", 0.0, 0.0, np.zeros((4, 4)), np.zeros(4))
    assert ei.value.identifier == "mamimo:type"
    mex.call("finalize")
    hD, Hr, Hi = mex.call("estimate", rx, nlhs=3)
    ref_ls, ref_r, ref_i = oracle_full(Y, P, xp, 1, nets)
    assert hD.shape == (nsc, nt, nr, npkt) and rel_l2(np.transpose(ref_ls, (3, 2, 1, 0)), hD) <= 1e-6
    assert Hr.shape == (nsc, nt * nr * npkt) and Hr.dtype == np.float32                       # column = pair row
    assert rel_l2(ref_r.T, Hr) <= 1e-5 and rel_l2(ref_i.T, Hi) <= 1e-5
    with pytest.raises(MexError) as ei:
        mex.call("estimate", rx, nlhs=1)                                   # needs three outputs
    assert ei.value.identifier == "mamimo:usage"
    # ofdmdemod front end (pg/generate_maMIMO_LTF.m:336-338): x [nltf*(fft+cp) x Nr x Npkt] -> rxOFDM [Nsc x nltf x Nr x Npkt]
    car = tables.carriers_locations()
    x = ofdm.ofdm_mod(Y, 256, 64, car) * 256                               # [npkt, nr, nt*320]
    with pytest.raises(MexError) as ei:
        mex.call("demod", np.transpose(x, (2, 1, 0)), nlhs=1)              # 'ofdm' not configured yet
    assert ei.value.identifier == "mamimo:size"
    mex.call("ofdm", 256.0, 64.0, 64.0, car.astype(np.float64))
    Yd = mex.call("demod", np.transpose(x, (2, 1, 0)), nlhs=1)
    assert Yd.shape == (nsc, nt, nr, npkt) and Yd.dtype == np.complex64
    assert rel_l2(np.transpose(ofdm.ofdm_demod(x, 256, 64, 64, car), (3, 2, 1, 0)), Yd) <= 2e-6
    mex.call("destroy")


@pytest.mark.gpu
def test_gateway_omphybweights_precoding(mex):
    """[Fbb, Frf, idx] = mamimo_mex('omphyb', hD, Ns, NtRF, AtExp): the call of pg/BER_test_maMIMO_LTF.m:372 with its own
    argument shapes (hD [L x Nt x Nr], AtExp [L x Nt x nRays] = one dictionary repeated per subcarrier) against the oracle
    of the reference's per-subcarrier loop: same columns, Frf = those dictionary columns, Frf.'*Fbb.' up to Fopt's phase."""
    from oracle import omp as oomp
    rng = np.random.default_rng(77)
    L, nt, nr, nrays = 52, 16, 2, 90
    hD = rng.standard_normal((L, nt, nr)) + 1j * rng.standard_normal((L, nt, nr))          # MATLAB [L x Nt x Nr]
    At = np.exp(2j * np.pi * rng.random((nt, nrays)))
    AtExp = np.broadcast_to(At[None], (L, nt, nrays)).copy()
    with pytest.raises(MexError) as ei:
        mex.call("omphyb", hD, 1.0, 1.0, AtExp, nlhs=2)                                    # no engine yet
    assert ei.value.identifier == "mamimo:state"
    mex.call("create", {"n_tx": nt, "n_rx": nr, "n_sc": L})
    H_eng = np.transpose(hD, (2, 1, 0))[None]                                              # [1, Nr, Nt, L]
    for ns, nrf in ((1, 1), (2, 3)):
        Fbb, Frf, idx = mex.call("omphyb", hD, float(ns), float(nrf), AtExp, nlhs=3)
        # MATLAB drops trailing singleton dimensions ([L x 1 x 1] is [L x 1]): compare the element counts, then view
        assert Fbb.size == L * ns * nrf and Frf.size == L * nrf * nt and idx.size == L * nrf
        assert Fbb.shape[0] == L and Frf.shape[0] == L and idx.shape[0] == L
        Fbb, Frf, idx = Fbb.reshape(L, ns, nrf), Frf.reshape(L, nrf, nt), idx.reshape(L, nrf)
        r_idx, r_fbb, _, _ = oomp.omp_precoder(H_eng, At, ns, nrf)
        assert np.array_equal(idx.T.astype(np.int64) - 1, r_idx[0])                        # 1-based columns
        for k in range(L):
            assert np.array_equal(Frf[k], At[:, r_idx[0, :, k]].T)                         # Frf_out = Frf.' (:197)
            M = Frf[k].T @ Fbb[k].T                                                        # Frf*Fbb in the reference's convention
            Mr = At[:, r_idx[0, :, k]] @ r_fbb[0, :, :, k].T
            assert np.max(np.abs(M @ M.conj().T - Mr @ Mr.conj().T)) <= 1e-9
            assert abs(np.linalg.norm(M) - np.sqrt(ns)) <= 1e-10                           # :179
        Fbb2, Frf2 = mex.call("omphyb", hD, float(ns), float(nrf), At, nlhs=2)             # 2-D dictionary form
        assert np.array_equal(Fbb, Fbb2.reshape(Fbb.shape)) and np.array_equal(Frf, Frf2.reshape(Frf.shape))
    for bad, ident in (((hD[:, :, :1], 1.0, 1.0, At), "mamimo:size"), ((hD, 3.0, 3.0, At), "mamimo:size"),
                       ((hD, 1.0, 1.0, At[:5]), "mamimo:size"), ((hD.real, 1.0, 1.0, At), "mamimo:type")):
        with pytest.raises(MexError) as ei:
            mex.call("omphyb", *bad, nlhs=2)
        assert ei.value.identifier == ident
    mex.call("destroy")
