"""Host-side logic that needs no GPU: prediction-file contract, MEX gateway syntax, MATLAB-shaped argument
checks of the drop-in helper."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import mamimo_b200 as mm
from oracle import postproc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "dl-channel-estimation-mamimo_b200")


def test_prediction_files_round_trip(tmp_path):
    """write_prediction_files -> the reader pg/BER_test_maMIMO_LTF.m:198-223 implements."""
    n_pkt, n_tx, n_rx, n_sc = 3, 4, 2, 10
    rng = np.random.default_rng(0)
    rows = n_pkt * n_tx * n_rx
    yr = rng.standard_normal((rows, n_sc)).astype(np.float32)
    yi = rng.standard_normal((rows, n_sc)).astype(np.float32)
    xr = rng.standard_normal((rows, 7))
    xi = rng.standard_normal((rows, 7))
    files = mm.pipeline.write_prediction_files(str(tmp_path), yr, yi, n_tx, n_rx, xr, xi, yr * 2, yi * 2)
    assert len(files) == 2 * n_pkt
    assert os.path.basename(files[0]) == "test_csi_predictions_real_1.mat"           # 1-based (..._DNN.py:408)
    for p in range(n_pkt):
        csi, x_r, x_i = mm.pipeline.read_prediction_files(str(tmp_path), p + 1, n_tx, n_rx)
        sl = slice(p * n_tx * n_rx, (p + 1) * n_tx * n_rx)
        want = postproc.rows_to_csi(yr[sl].astype(np.complex128) + 1j * yi[sl], n_tx, n_rx)
        assert np.array_equal(csi, want)
        assert np.array_equal(x_r, xr[sl]) and np.array_equal(x_i, xi[sl])


def test_prediction_files_missing_dir_exits_zero(tmp_path):
    with pytest.raises(SystemExit) as ei:          # massiveMIMO_CSI_prediction_DNN.py:112-115
        mm.pipeline.write_prediction_files(str(tmp_path / "nope"), np.zeros((8, 4), np.float32),
                                           np.zeros((8, 4), np.float32), 4, 2)
    assert ei.value.code == 0


def test_mex_gateway_builds_warning_free_and_stays_thin():
    """The gateway is compiled for real into the harness the MEX tests execute (tests/test_mex_gateway.py); here:
    no warnings under -Wall -Wextra against the mini runtime's mex.h, and it stays a thin shim over the C ABI."""
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    src = os.path.join(PKG, "csrc", "mex_gateway.cpp")
    res = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror",
                          "-I" + os.path.join(PKG, "csrc", "mex_runtime"), "-I" + os.path.join(ROOT, "include"), src],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert len(open(src).read().splitlines()) < 300       # ten commands; the longest ('omphyb') re-lays Fbb / Frf out
    assert os.path.exists(mm.build.build_mex_harness())


def test_helper_argument_validation_without_gpu():
    prm = {"numSTS": 4, "CarriersLocations": mm.carriers_locations()}
    with pytest.raises(ValueError):                            # size(rxData,1) must match CarriersLocations
        mm.helperMIMOChannelEstimate(np.zeros((100, 4, 2), np.complex128), prm)
    with pytest.raises(ValueError):                            # nltf should be == numSTS (:10)
        mm.helperMIMOChannelEstimate(np.zeros((234, 3, 2), np.complex128), prm)
    with pytest.raises(ValueError):                            # isMMSE needs tau and SNR (:38)
        mm.helperMIMOChannelEstimate(np.zeros((234, 4, 2), np.complex128), prm, 1, None, 10.0, True)
