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


def rebuild_rx_signal(x_real, x_imag, len_in, n_tx, n_rx):
    """What pg/BER_test_maMIMO_LTF.m does with the `x` field of one packet's files: keep x(:,1:lenIn) (:203,206) and
    take row (iRx-1)*nTXAnts+1 of each plane as the time-domain preamble of rx antenna iRx (:312-318).
    Returns inputRXSig complex [len_in, n_rx] (MATLAB shape), the input of its ofdmdemod call (:324-326)."""
    xr = np.asarray(x_real)[:, :len_in]
    xi = np.asarray(x_imag)[:, :len_in]
    sig = np.zeros((len_in, n_rx), dtype=np.complex128)
    for irx in range(n_rx):
        sig[:, irx] = xr[irx * n_tx, :] + 1j * xi[irx * n_tx, :]
    return sig


def run_test_mode_time(engine, rx_time, workdir, true_real=None, true_imag=None, first_pkt_id=1):
    """`massiveMIMO_CSI_prediction_DNN.py --test` (:330-346,401-409) from the time-domain preamble the shipped
    pipeline feeds its nets: rx_time complex [n_pkt, n_rx, lenLTF] (MATLAB inputRXSig [lenLTF x Nr] per packet,
    pg/generate_maMIMO_LTF.m:326-327).  The engine is either
      * a mode-A engine (input_mode='time_p': [LTF || P(:,iTx)] -> nets, the pipeline's own network), or
      * a mode-C engine with the OFDM front-end configured (set_ofdm): demod -> LS -> nets (the north_star path).
    Leaves test_csi_predictions_{real,imag}_<pkt>.mat with y = the prediction planes and x = the LTF part of the
    net input (:80,405) -- row (pkt, rx, tx) holds the real / imag plane of rx's time-domain preamble, which is what
    pg/BER_test_maMIMO_LTF.m:203,206,312-318 turns back into inputRXSig.  Returns (H_real, H_imag)."""
    rx_time = np.asarray(rx_time)
    c = engine.cfg
    if rx_time.ndim != 3 or rx_time.shape[1] != c.n_rx or not np.iscomplexobj(rx_time):
        raise ValueError("rx_time must be complex [n_pkt, n_rx=%d, lenLTF]" % c.n_rx)
    if engine.input_mode == "time_p":
        Hr, Hi = engine.predict_time(rx_time.real, rx_time.imag)
    elif engine.input_mode == "ls":
        Hr, Hi = engine.estimate_time(rx_time.astype(np.complex64, copy=False) if rx_time.dtype != np.complex128 else rx_time)
    else:
        raise ValueError("run_test_mode_time needs a mode-A ('time_p') or mode-C ('ls' + set_ofdm) engine")
    # every pair row of (pkt, rx) carries the same LTF (the reference stores it once per rx under a hash,
    # create_massiveMIMO_CSIest_dnn_dataset.py:50-59, and the generator replicates it per tx, :307-309)
    rows = c.n_tx * c.n_rx
    Hr, Hi = np.asarray(Hr), np.asarray(Hi)
    for p in range(rx_time.shape[0]):               # one packet at a time: x is n_tx copies of each rx row (10 MB/plane at 32x4x10240)
        sl = slice(p * rows, (p + 1) * rows)
        x_r = np.repeat(rx_time[p].real.astype(np.float64), c.n_tx, axis=0)
        x_i = np.repeat(rx_time[p].imag.astype(np.float64), c.n_tx, axis=0)
        write_prediction_files(workdir, Hr[sl], Hi[sl], c.n_tx, c.n_rx, x_r, x_i,
                               None if true_real is None else np.asarray(true_real)[sl],
                               None if true_imag is None else np.asarray(true_imag)[sl], first_pkt_id + p)
    return Hr, Hi


def run_test_mode(engine, Y, workdir, true_real=None, true_imag=None, first_pkt_id=1, keep_input=True):
    """Frequency-domain (mode C) variant: estimate a batch of already-demodulated packets Y on the GPU and leave one
    pair of .mat files per packet in workdir.  Its `x` field holds the nets' input here, i.e. the H_LS planes -- NOT
    the time-domain preamble BER_test_maMIMO_LTF.m:312-318 rebuilds its rx signal from; use run_test_mode_time when
    the files feed that evaluator.  Returns (H_real, H_imag, H_ls)."""
    Hr, Hi, Hls = engine.estimate(Y, want_ls=True)
    c = engine.cfg
    x_r = x_i = None
    if keep_input:
        flat = np.asarray(Hls).reshape(-1, c.n_sc)
        x_r, x_i = flat.real, flat.imag
    write_prediction_files(workdir, np.asarray(Hr), np.asarray(Hi), c.n_tx, c.n_rx, x_r, x_i, true_real, true_imag,
                           first_pkt_id)
    return Hr, Hi, Hls
