"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/mamimo.h declares,
its integer tables are bit-identical to the oracle and to the reference source, and it fails loudly
without a GPU (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import mamimo_b200 as mm
from oracle import tables, postproc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mamimo.h")).read()
    return sorted(set(re.findall(r"MAMIMO_API[^;(]*?\b(mamimo_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from importlib import import_module
    capi = import_module("_mamimo_b200_pkg._capi")
    declared = _header_symbols()
    assert len(declared) >= 20
    assert sorted(capi.SYMBOLS) == declared
    raw = ctypes.CDLL(mm.build.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), "missing export " + name
    assert raw.mamimo_abi_version() == capi.ABI_VERSION == 2


def test_tables_bit_identical(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_tables.npz"))
    assert np.array_equal(mm.vht_ltf256(), tables.vht_ltf256())
    assert np.array_equal(mm.vht_ltf256(), g["ltf256"])
    assert np.array_equal(mm.carriers_locations(), tables.carriers_locations())
    assert np.array_equal(mm.carriers_locations(), g["carriers"])
    for n in (1, 2, 4, 32, 64):
        assert np.array_equal(mm.default_p(n), tables.sylvester_hadamard(n).astype(np.float32))
    with pytest.raises(ValueError):
        mm.default_p(6)


def test_pair_row_matches_oracle():
    for p in (0, 3, 499):
        for irx in range(4):
            for itx in (0, 7, 31):
                assert mm.pair_row(p, irx, itx, 4, 32) == postproc.pair_row(p, irx, itx, 4, 32)


def test_synth_noise_rule_and_shapes():
    Y, H = mm.synth.make_packets(9, 2, 8, 2, 64, snr_db=10.0)
    assert Y.shape == (2, 2, 8, 64) and H.shape == (2, 2, 8, 64) and Y.dtype == np.complex64
    P = mm.synth.sylvester(8)
    sig = np.einsum("prjk,jn->prnk", H.astype(np.complex128), P)
    snr = 10 * np.log10(np.mean(np.abs(sig) ** 2) / np.mean(np.abs(Y - sig) ** 2))
    assert abs(snr - 10.0) < 0.7
    Y2, _ = mm.synth.make_packets(9, 1, 8, 2, 64, snr_db=10.0, first_pkt=1)
    assert np.array_equal(Y2[0], Y[1])          # packet streams are reproducible per (config, packet)


def test_csi_predictor_postprocess_matches_reference(golden_dir):
    """Host-side glue of the drop-in CSIPredictor vs inference.py run for real (golden)."""
    z = np.load(os.path.join(golden_dir, "ref_inference_py.npz"))
    pred = object.__new__(mm.CSIPredictor)          # skip engine construction (needs a GPU)
    pred.experiment = "RICE_RENEW"
    rng = np.random.default_rng(0)
    out52 = rng.standard_normal((5, 52)) + 1j * rng.standard_normal((5, 52))
    assert np.array_equal(pred.postprocess_data(out52), postproc.renew_postprocess(out52))
    with pytest.raises(SystemExit) as ei:            # inference.py:64-66
        pred.postprocess_data(np.zeros((1, 50), dtype=np.complex128))
    assert ei.value.code == -1
    with pytest.raises(SystemExit) as ei:            # inference.py:41-43
        pred.preprocess_data(z["X"].astype(np.complex64))
    assert ei.value.code == -1


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mm.MamimoError) as ei:
        mm.Engine(4, 2, 64, hidden=(32,))
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dl-channel-estimation-mamimo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
