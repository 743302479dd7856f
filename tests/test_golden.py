"""Oracle vs vectors produced by the REAL reference code (tests/golden/make_golden.py)."""
import os

import numpy as np

from oracle import tables, mlp, postproc


def _layers(z, prefix):
    out, i = [], 0
    while "%s_W%d" % (prefix, i) in z:
        L = {"W": z["%s_W%d" % (prefix, i)], "b": z["%s_b%d" % (prefix, i)], "bn": None}
        if "%s_bn%d_gamma" % (prefix, i) in z:
            L["bn"] = tuple(z["%s_bn%d_%s" % (prefix, i, k)] for k in ("gamma", "beta", "mean", "var"))
        out.append(L)
        i += 1
    return out


def test_tables_match_reference_source(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_tables.npz"))
    assert np.array_equal(g["ltf256"], tables.vht_ltf256())
    assert np.array_equal(g["carriers"], tables.carriers_locations())
    assert np.array_equal(g["nulls"], tables.null_carrier_indices())
    assert np.array_equal(g["pilots"], tables.pilot_carrier_indices())


def test_csi_predictor_matches_reference_inference_py(golden_dir):
    z = np.load(os.path.join(golden_dir, "ref_inference_py.npz"))
    nets = {"real": _layers(z, "real"), "imag": _layers(z, "imag")}
    Y = postproc.csi_predictor_inference(z["X"], lambda x: mlp.forward(x, nets["real"]),
                                         lambda x: mlp.forward(x, nets["imag"]))
    assert Y.shape == z["Y"].shape
    assert np.max(np.abs(Y - z["Y"])) < 1e-12


def test_mode_a_assembly_matches_reference_data_generator(golden_dir):
    z = np.load(os.path.join(golden_dir, "ref_data_generator.npz"))
    n_pkt, n_rx, n_tx = int(z["n_pkt"]), int(z["n_rx"]), int(z["n_tx"])
    rows = range(n_pkt * n_rx * n_tx)
    for d, part in (("real", np.real), ("imag", np.imag)):
        xsig, xp = postproc.assemble_mode_a(part(z["ltf"]), z["P"], rows, n_rx, n_tx)
        assert np.array_equal(xsig, z["Xsig_" + d])
        assert np.array_equal(xp, z["Xp_" + d])
        assert np.array_equal(part(z["y"]), z["y_" + d])


def test_ofdm_front_end_matches_reference_numpy_mirror(golden_dir):
    """oracle.ofdm vs massiveMIMO_dataGenerator.py:425-458 (method='reshape') executed for real.
    The mirror FFTs one real plane, keeps the real part, and its bare fftshift also rolls the symbol axis."""
    from oracle import ofdm
    z = np.load(os.path.join(golden_dir, "ref_data_generator_reshape.npz"))
    for tag in "abc":
        fft_len, cp, so, nsym, nrx, npkt = [int(v) for v in z["cfg_" + tag]]
        ltf = z["ltf_" + tag]
        for d, part in (("real", np.real), ("imag", np.imag)):
            X = z["X%s_%s" % (d, tag)]
            Y = ofdm.ofdm_demod(part(ltf), fft_len, cp, so, np.arange(1, fft_len + 1))
            for row in range(X.shape[0]):
                p, irx, itx = row // (nrx * nsym), (row // nsym) % nrx, row % nsym
                sym = (itx - nsym // 2) % nsym
                assert np.allclose(X[row, :fft_len], Y[p, irx, sym].real, rtol=0, atol=1e-12)


def test_ofdm_mod_demod_round_trip():
    from oracle import ofdm
    rng = np.random.default_rng(3)
    car = tables.carriers_locations()
    G = rng.standard_normal((2, 3, 4, car.size)) + 1j * rng.standard_normal((2, 3, 4, car.size))
    x = ofdm.ofdm_mod(G, 256, 64, car)
    assert x.shape == (2, 3, 4 * 320)
    assert np.max(np.abs(ofdm.ofdm_demod(x, 256, 64, 64, car) - G)) < 1e-12
