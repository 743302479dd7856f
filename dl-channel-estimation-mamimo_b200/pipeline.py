"""Test-mode driver and prediction-file contract of the reference pipeline (SURVEY.md 8(a5), 8(b) CLI row).

Replaces steps 4-5 of full_pipeline_maMIMO_DNNEst.sh:44-55, i.e. the `--test` branch of
massiveMIMO_CSI_prediction_DNN.py:330-346,401-411:

    batch = nTX*nRX rows per packet (:339) -> predict (:346) ->
    per packet savemat('test_csi_predictions_<d>_<pkt>.mat',
                       {'all_pkts_csi_nn_out': {x, y, true_y}}, do_compression=True)   (:404-409)

with pkt ids 1-based (:408).  pg/BER_test_maMIMO_LTF.m:198-207 loads those files and rebuilds
CSI(:,iTX,iRX) = pred((iRX-1)*nTX+iTX,:) (:213-218); read_prediction_files() mirrors that reader.

Only host-side file plumbing lives here; predictions come from the CUDA engine.
"""
import os

import numpy as np
from scipy.io import loadmat, savemat

DIMS = ("real", "imag")


def prediction_path(workdir, d, pkt_id):
    return os.path.join(workdir, "test_csi_predictions_%s_%d.mat" % (d, pkt_id))


def write_prediction_files(workdir, y_real, y_imag, n_tx, n_rx, x_real=None, x_imag=None, true_real=None,
                           true_imag=None, first_pkt_id=1, compress=True):
    """y_* float32 [n_pkt*n_tx*n_rx, n_sc] in pair-row order.  x_* is the net input kept for the MATLAB
    evaluator (it reuses x(:,1:lenIn), BER_test_maMIMO_LTF.m:203,206), true_* the labels (LS @ SNR 120 dB)."""
    if not (os.path.exists(workdir) and os.path.isdir(workdir)):
        # massiveMIMO_CSI_prediction_DNN.py:112-115: message + exit(0) when the directory is missing
        print("Given directory does not exists. Aborting...")
        raise SystemExit(0)
    rows = n_tx * n_rx
    n_pkt = y_real.shape[0] // rows
    if y_real.shape[0] != n_pkt * rows or y_imag.shape != y_real.shape:
        raise ValueError("prediction planes must hold a whole number of nTX*nRX-row packets")
    planes = {"real": (y_real, x_real, true_real), "imag": (y_imag, x_imag, true_imag)}
    written = []
    for d in DIMS:
        y, x, t = planes[d]
        for p in range(n_pkt):
            sl = slice(p * rows, (p + 1) * rows)
            out = {"y": np.asarray(y[sl], dtype=np.float32)}
            out["x"] = np.asarray(x[sl], dtype=np.float64) if x is not None else np.zeros((rows, 0))
            out["true_y"] = np.asarray(t[sl], dtype=np.float64) if t is not None else np.zeros((rows, 0))
            path = prediction_path(workdir, d, first_pkt_id + p)
            savemat(path, {"all_pkts_csi_nn_out": out}, do_compression=compress)
            written.append(path)
    return written


def read_prediction_files(workdir, pkt_id, n_tx, n_rx):
    """What pg/BER_test_maMIMO_LTF.m:198-223 does with one packet's pair of files.
    Returns CSI_dnn complex [n_sc, n_tx, n_rx] (MATLAB shape) and the two x planes."""
    planes, xs = {}, {}
    for d in DIMS:
        s = loadmat(prediction_path(workdir, d, pkt_id), struct_as_record=False, squeeze_me=True)["all_pkts_csi_nn_out"]
        planes[d] = np.asarray(s.y)
        xs[d] = np.asarray(s.x)
    n_sc = planes["real"].shape[1]
    csi = np.zeros((n_sc, n_tx, n_rx), dtype=np.complex128)
    for irx in range(n_rx):
        for itx in range(n_tx):
            r = irx * n_tx + itx                      # (iRX-1)*nTXAnts + iTX, 1-based in MATLAB
            csi[:, itx, irx] = planes["real"][r, :] + 1j * planes["imag"][r, :]
    return csi, xs["real"], xs["imag"]


def run_test_mode(engine, Y, workdir, true_real=None, true_imag=None, first_pkt_id=1, keep_input=True):
    """Mode-C equivalent of `massiveMIMO_CSI_prediction_DNN.py --test`: estimate a batch of packets on the
    GPU and leave one pair of .mat files per packet in workdir.  Returns (H_real, H_imag, H_ls)."""
    Hr, Hi, Hls = engine.estimate(Y, want_ls=True)
    c = engine.cfg
    x_r = x_i = None
    if keep_input:
        flat = np.asarray(Hls).reshape(-1, c.n_sc)
        x_r, x_i = flat.real, flat.imag
    write_prediction_files(workdir, np.asarray(Hr), np.asarray(Hi), c.n_tx, c.n_rx, x_r, x_i, true_real, true_imag,
                           first_pkt_id)
    return Hr, Hi, Hls
