"""Oracle for the glue around the hot path: CSIPredictor pre/post-processing,
pair ordering, per-pair input assembly, NMSE metric.

PINNED: the functions marked (pinned) are checked in tests/test_golden.py
against vectors produced by running the reference's own Python
(inference.py, massiveMIMO_dataGenerator.py) with a stub tensorflow module --
see tests/golden/make_golden.py.
"""
import numpy as np


# ---- inference.py:48-68 (pinned) -------------------------------------------
def renew_postprocess(out52):
    """[N,52] complex -> [N,64]: pad [6 zeros | 26 | 1 zero | 26 | 5 zeros], ifftshift(axis=1).

    inference.py:56-63.  Raises ValueError where the reference prints
    '[CSIPredictor] ERROR: Output samples must have size 52 ...' and exits(-1) (:64-66).
    """
    out52 = np.asarray(out52)
    if out52.shape[1] != 52:
        raise ValueError("Output samples must have size 52 (assuming FFTLen = 64).")
    n = out52.shape[0]
    tmp = np.zeros((n, 64), dtype=np.result_type(out52.dtype, np.float64))
    tmp[:, 6:32] = out52[:, :26]
    tmp[:, 33:59] = out52[:, 26:]
    return np.fft.ifftshift(tmp, axes=1)


def csi_predictor_inference(X, predict_real, predict_imag):
    """inference.py:24-32 (pinned): planes -> two predicts -> real + 1j*imag -> postprocess."""
    X = np.asarray(X)
    if X.dtype != np.complex128:                     # inference.py:41-43
        raise TypeError("Input batch must be of type np.complex128")
    out = predict_real(X.real) + 1j * predict_imag(X.imag)
    return renew_postprocess(out)


# ---- create_massiveMIMO_CSIest_dnn_dataset.py:62 / BER_test_maMIMO_LTF.m:213-218
def pair_row(p, i_rx, i_tx, n_rx, n_tx):
    """0-based sample row of pair (packet p, rx i_rx, tx i_tx)."""
    return p * (n_rx * n_tx) + i_rx * n_tx + i_tx


def rows_to_csi(pred_rows, n_tx, n_rx):
    """One packet's prediction rows [Nt*Nr, Nsc] -> MATLAB-shaped CSI [Nsc, Nt, Nr].

    BER_test_maMIMO_LTF.m:213-218: CSI(:,iTX,iRX) = pred((iRX-1)*nTX+iTX,:).
    """
    pred_rows = np.asarray(pred_rows)
    nsc = pred_rows.shape[1]
    csi = np.zeros((nsc, n_tx, n_rx), dtype=pred_rows.dtype)
    for irx in range(n_rx):
        for itx in range(n_tx):
            csi[:, itx, irx] = pred_rows[irx * n_tx + itx, :]
    return csi


# ---- massiveMIMO_dataGenerator.py:299-316 (pinned) ---------------------------
def assemble_mode_a(ltf_plane, P, rows, n_rx, n_tx):
    """Per-pair MLP input of the shipped pipeline ('matlab_maMimo', method default).

    ltf_plane : [Npkt, Nr, lenLTF] real (or imag) part of the time-domain preamble
    P         : [Nt, Nt] as the pickle holds it (h5py-transposed MATLAB P); the
                generator feeds column P[:, iTx] (:311)
    rows      : iterable of sample rows (pair_row order)
    returns (Xsig [B, lenLTF], Xp [B, Nt])
    """
    rows = np.asarray(list(rows))
    pkt = rows // (n_rx * n_tx)
    irx = (rows // n_tx) % n_rx
    itx = rows % n_tx
    xsig = np.asarray(ltf_plane)[pkt, irx, :]
    xp = np.asarray(P)[:, itx].T
    return xsig, xp


# ---- BER_test_maMIMO_LTF.m:675-686 -----------------------------------------
def nmse_subk(real, pred):
    """mean over (tx,rx) of ||real(:,t,r)-pred(:,t,r)||^2 / ||real(:,t,r)||^2; inputs [Nsc,Nt,Nr]."""
    real = np.asarray(real)
    pred = np.asarray(pred)
    diff = real - pred
    num = np.sum(np.abs(diff) ** 2, axis=0)
    den = np.sum(np.abs(real) ** 2, axis=0)
    return float(np.mean(num / den))


def rel_l2(ref, x):
    ref = np.asarray(ref)
    x = np.asarray(x)
    return float(np.linalg.norm((x - ref).ravel()) / np.linalg.norm(ref.ravel()))
