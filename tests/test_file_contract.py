"""Host-side contract of steps 4-5 of full_pipeline_maMIMO_DNNEst.sh (no GPU needed):

  * the per-packet prediction files are what pg/BER_test_maMIMO_LTF.m:198-226,312-318 reads -- pinned on the reference's
    own lines executed by tests/golden/mini_matlab.py (golden; re-executed live when /root/reference is present);
  * the argv-compatible --test entry (massiveMIMO_CSI_prediction_DNN.py:3-34,330-346,401-411) on a pickle laid out
    as create_massiveMIMO_CSIest_dnn_dataset.py:125 writes it, with a stand-in engine;
  * Keras weight ingestion (save_weights HDF5 layout, loaded-model objects) -> npz the engine loads.
"""
import os
import pickle
import sys

import numpy as np
import pytest

import mamimo_b200 as mm
from oracle import mlp, postproc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MAMIMO_REFERENCE", "/root/reference")


# ------------------------------------------------------------------------------------ BER_test reader
def test_prediction_files_are_what_ber_test_reads(golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, "ref_ber_test_reader.npz"))
    n_tx, n_rx, len_in = int(g["n_tx"]), int(g["n_rx"]), int(g["len_in"])
    mm.pipeline.write_prediction_files(str(tmp_path), g["y_real"], g["y_imag"], n_tx, n_rx, g["x_real"], g["x_imag"])
    for p in (1, 2):
        csi, x_r, x_i = mm.pipeline.read_prediction_files(str(tmp_path), p, n_tx, n_rx)
        assert np.array_equal(csi, g["csi_dnn_%d" % p])                      # MATLAB's CSI_dnn (:213-226), bit for bit
        rx = mm.pipeline.rebuild_rx_signal(x_r, x_i, len_in, n_tx, n_rx)
        assert np.array_equal(rx, g["input_rx_sig_%d" % p])                  # MATLAB's inputRXSig (:312-318)
        assert np.array_equal(csi, postproc.rows_to_csi(g["y_real"][(p - 1) * n_tx * n_rx:p * n_tx * n_rx].astype(np.float64)
                                                        + 1j * g["y_imag"][(p - 1) * n_tx * n_rx:p * n_tx * n_rx], n_tx, n_rx))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_ber_test_reader_lines_run_live_on_run_test_mode_time_files(tmp_path):
    """run_test_mode_time with a stand-in engine -> files -> the reference's reader lines (mini_matlab) recover the
    time-domain rx signal that went in and the prediction planes, packet ids 1-based."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import ber_test_reader_source
    from mini_matlab import MatlabFile
    n_tx, n_rx, len_ltf, n_sc, n_pkt = 4, 2, 40, 12, 3
    rng = np.random.default_rng(5)
    sig = rng.standard_normal((n_pkt, n_rx, len_ltf)) + 1j * rng.standard_normal((n_pkt, n_rx, len_ltf))
    eng = _FakeModeAEngine(n_tx, n_rx, len_ltf, n_sc, mm.synth.make_nets(len_ltf + n_tx, (16,), n_sc))
    eng.set_pilots(None, mm.synth.sylvester(n_tx))
    Hr, Hi = mm.pipeline.run_test_mode_time(eng, sig, str(tmp_path), first_pkt_id=1)
    m = MatlabFile(ber_test_reader_source(REF))
    for p in range(1, n_pkt + 1):
        csi, rx = m.call("ber_reader", [os.path.join(str(tmp_path), "test_csi_predictions_real.mat"),
                                        os.path.join(str(tmp_path), "test_csi_predictions_imag.mat"),
                                        float(p), float(len_ltf), float(n_sc), float(n_tx), float(n_rx)], 2)
        assert np.array_equal(rx, sig[p - 1].T)                               # inputRXSig [lenLTF x Nr], exact (float64 x)
        sl = slice((p - 1) * n_tx * n_rx, p * n_tx * n_rx)
        want = postproc.rows_to_csi(Hr[sl].astype(np.float64) + 1j * Hi[sl], n_tx, n_rx)
        assert np.array_equal(csi, want)


# ------------------------------------------------------------------------------------ CLI with a stand-in engine
class _Cfg:
    pass


class _FakeModeAEngine:
    """Engine-shaped stand-in (mode A) whose predictions come from the oracle: lets the host logic run without a GPU."""
    input_mode = "time_p"

    def __init__(self, n_tx, n_rx, len_ltf, d_out, nets=None):
        self.cfg = _Cfg()
        self.cfg.n_tx, self.cfg.n_rx, self.cfg.len_ltf, self.cfg.d_out = n_tx, n_rx, len_ltf, d_out
        self.nets, self.P, self.closed = nets, None, False

    def set_pilots(self, x, P):
        self.P = np.asarray(P, dtype=np.float64)          # engine orientation: row j = code of tx j

    def load_weights(self, nets):
        self.nets = nets

    def predict_time(self, sr, si):
        c = self.cfg
        rows = np.arange(sr.shape[0] * c.n_rx * c.n_tx)
        out = []
        for part, name in ((sr, "real"), (si, "imag")):
            xsig, xp = postproc.assemble_mode_a(np.asarray(part, np.float32), self.P.T, rows, c.n_rx, c.n_tx)
            out.append(mlp.forward(np.concatenate([xsig, xp], axis=1), self.nets[name]).astype(np.float32))
        return out[0], out[1]

    def close(self):
        self.closed = True


def _make_pickle(path, n_pkt, n_tx, n_rx, len_ltf, n_sc, seed=67):
    """Dataset dict as create_massiveMIMO_CSIest_dnn_dataset.py:125 pickles it."""
    rng = np.random.default_rng(seed)
    ltf = rng.standard_normal((n_pkt, n_rx, len_ltf)) + 1j * rng.standard_normal((n_pkt, n_rx, len_ltf))
    y = rng.standard_normal((n_pkt * n_rx * n_tx, n_sc)) + 1j * rng.standard_normal((n_pkt * n_rx * n_tx, n_sc))
    X = np.zeros((n_pkt * n_rx * n_tx, 2), dtype=np.int64)
    LTF = {}
    for p in range(n_pkt):
        for irx in range(n_rx):
            h = int(rng.integers(1, 2 ** 32))
            LTF[h] = {"real": ltf[p, irx].real.copy(), "imag": ltf[p, irx].imag.copy()}
            for itx in range(n_tx):
                X[p * (n_rx * n_tx) + irx * n_tx + itx] = [h, itx]
    P = mm.synth.sylvester(n_tx) * rng.choice([-1.0, 1.0], size=(1, n_tx))
    ds = {"X": X, "y": {"real": y.real.copy(), "imag": y.imag.copy()}, "LTF": LTF, "P": P.T.copy(),    # h5py-transposed
          "simParams": {"FFTLength": np.float64(8), "CPLen": np.float64(2), "numSym": len_ltf / 10, "symOffset": np.float64(2),
                        "nTX": n_tx, "nRX": n_rx}}
    with open(path, "wb") as f:
        pickle.dump(ds, f)
    return ds, ltf, y, P


def test_cli_test_branch_end_to_end_with_stand_in_engine(tmp_path, capsys):
    n_pkt, n_tx, n_rx, len_ltf, n_sc, hidden = 5, 4, 2, 40, 12, (24, 16)
    pk = str(tmp_path / "testDataset.b")
    ds, ltf, y, P = _make_pickle(pk, n_pkt, n_tx, n_rx, len_ltf, n_sc)
    modeldir, workdir = tmp_path / "model", tmp_path / "model" / "test_results"
    modeldir.mkdir()
    nets = mm.synth.make_nets(len_ltf + n_tx, hidden, n_sc)
    for d in ("real", "imag"):
        mm.weights.save_npz(str(modeldir / (d + "_weights.npz")), nets[d])
    made = []

    def factory(sp, hid, precision, device, max_pkts):
        assert (sp["nTX"], sp["nRX"], sp["lenLTF"], sp["nSubCarr"]) == (n_tx, n_rx, len_ltf, n_sc) and list(hid) == list(hidden)
        made.append(_FakeModeAEngine(n_tx, n_rx, len_ltf, n_sc))
        return made[-1]

    argv = ["--test", "-x", pk, "--nn", "24", "16", "-d", str(workdir), "--modeldir", str(modeldir), "--useGPU", "0",
            "--useBN", "--datasource", "matlab_maMimo", "--valSameTrain", "--chunk-pkts", "2"]
    assert mm.cli.main(argv, engine_factory=factory) == 0                     # the reference: message + exit(0) (:112-115)
    assert "Given directory does not exists" in capsys.readouterr().out and not made
    workdir.mkdir()
    assert mm.cli.main(argv, engine_factory=factory) == 0
    assert made and made[0].closed
    rows = n_tx * n_rx
    allrows = np.arange(n_pkt * rows)
    for d, part in (("real", np.real), ("imag", np.imag)):
        xsig, xp = postproc.assemble_mode_a(part(ltf).astype(np.float32), ds["P"], allrows, n_rx, n_tx)
        ref = mlp.forward(np.concatenate([xsig, xp], axis=1), nets[d])
        for p in range(n_pkt):
            from scipy.io import loadmat
            s = loadmat(str(workdir / ("test_csi_predictions_%s_%d.mat" % (d, p + 1))), struct_as_record=False,
                        squeeze_me=True)["all_pkts_csi_nn_out"]
            sl = slice(p * rows, (p + 1) * rows)
            assert s.y.dtype == np.float32 and postproc.rel_l2(ref[sl], s.y) < 1e-6
            assert np.array_equal(s.x, np.repeat(part(ltf[p]), n_tx, axis=0))          # x = LTF part of the input (:80)
            assert np.array_equal(s.true_y, part(y)[sl])                                # labels ride along (:407)
        assert os.path.exists(str(workdir / (d + "_weights.npz")))
    assert not os.path.exists(str(workdir / ("test_csi_predictions_real_%d.mat" % (n_pkt + 1))))
    # without --valSameTrain the last floor(n_packets * valTrainRatio) packets are the test set (:47-52, :126-128)
    w2 = tmp_path / "w2"
    w2.mkdir()
    argv2 = ["--test", "-x", pk, "--nn", "24", "16", "-d", str(w2), "--modeldir", str(modeldir), "--datasource",
             "matlab_maMimo", "--valTrainRatio", "0.4"]
    assert mm.cli.main(argv2, engine_factory=factory) == 0
    assert sorted(f for f in os.listdir(str(w2)) if f.startswith("test_csi_predictions_real")) == \
        ["test_csi_predictions_real_1.mat", "test_csi_predictions_real_2.mat"]
    from scipy.io import loadmat
    s = loadmat(str(w2 / "test_csi_predictions_real_1.mat"), struct_as_record=False, squeeze_me=True)["all_pkts_csi_nn_out"]
    assert np.array_equal(s.true_y, y.real[3 * rows:4 * rows])                          # packet 3 (0-based) is test packet 1
    # branches this entry point leaves to the reference script
    assert mm.cli.main(["--train", "-x", pk, "--datasource", "matlab_maMimo"], engine_factory=factory) == 2
    assert mm.cli.main(argv[:-2] + ["--nn", "8"], engine_factory=factory) == 2          # --nn must match the stored nets


def test_cli_parser_accepts_the_pipeline_script_line():
    """The literal flag set of full_pipeline_maMIMO_DNNEst.sh:47 and the reference's defaults (..._DNN.py:3-33)."""
    a = mm.cli.build_parser().parse_args(
        "--test -x ds.b --nn 1024 1024 -d out/BS32_SNR10 --modeldir out --useGPU 0 --useBN --datasource matlab_maMimo "
        "--valSameTrain".split())
    assert a.test and not a.train and a.nn == [1024, 1024] and a.workdir == "out/BS32_SNR10" and a.useBN and a.valSameTrain
    d = mm.cli.build_parser().parse_args(["-x", "a", "--datasource", "matlab_maMimo"])
    assert (d.nn, d.bs, d.epochs, d.lr, d.dropout, d.valTrainRatio, d.workdir, d.model) == \
        ([256, 128], 256, 500, 0.0001, 0.15, 0.15, "checkpoint", "FC")
    with pytest.raises(SystemExit):
        mm.cli.build_parser().parse_args(["--train", "--test", "-x", "a", "--datasource", "b"])   # mutually exclusive (:4)


# ------------------------------------------------------------------------------------ Keras weight ingestion
class _Attrs(dict):
    pass


class _FakeH5Group(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.attrs = _Attrs()


def _fake_keras_hdf5(layers, wrap_model_weights=False):
    """Keras 2.x save_weights layout (layer_names / weight_names attributes) of the reference's FC model
    (..._DNN.py:177-227): inputs, drop_test, flatten, concatenate carry no weights."""
    f = _FakeH5Group()
    names = [b"input_1", b"drop_test", b"flatten", b"input_2", b"concatenate"]
    for n in names:
        f[n.decode()] = _FakeH5Group()
        f[n.decode()].attrs["weight_names"] = []
    bn_i = 0
    for i, L in enumerate(layers):
        last = i == len(layers) - 1
        ln = "fc_regressor" if last else "fc_dense%d" % i
        g = _FakeH5Group({ln + "/kernel:0": L["W"], ln + "/bias:0": L["b"]})
        g.attrs["weight_names"] = [(ln + "/kernel:0").encode(), (ln + "/bias:0").encode()]
        f[ln] = g
        names.append(ln.encode())
        if L.get("bn") is not None:
            bn = "batch_normalization" + ("_%d" % bn_i if bn_i else "")
            bn_i += 1
            keys = [bn + "/gamma:0", bn + "/beta:0", bn + "/moving_mean:0", bn + "/moving_variance:0"]
            g = _FakeH5Group(dict(zip(keys, L["bn"])))
            g.attrs["weight_names"] = [k.encode() for k in keys]
            f[bn] = g
            names.append(bn.encode())
            f["drop%d" % i] = _FakeH5Group()
            f["drop%d" % i].attrs["weight_names"] = []
            names.append(("drop%d" % i).encode())
    f.attrs["layer_names"] = names
    if wrap_model_weights:
        top = _FakeH5Group({"model_weights": f})
        return top
    return f


class _FakeKerasLayer:
    def __init__(self, w, name):
        self._w, self.name = w, name

    def get_weights(self):
        return self._w


def _cls(name):
    return type(name, (_FakeKerasLayer,), {})


def test_keras_artifacts_to_engine_layers_and_npz_round_trip(tmp_path):
    nets = mm.synth.make_nets(40, (24, 16), 12)
    want = nets["real"]

    def same(a, b):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert np.array_equal(x["W"], y["W"]) and np.array_equal(x["b"], y["b"])
            assert (x["bn"] is None) == (y["bn"] is None)
            if x["bn"] is not None:
                assert all(np.array_equal(u, v) for u, v in zip(x["bn"], y["bn"]))

    same(mm.weights.layers_from_keras_hdf5(_fake_keras_hdf5(want)), want)
    same(mm.weights.layers_from_keras_hdf5(_fake_keras_hdf5(want, wrap_model_weights=True)), want)   # Model.save .h5
    model = type("M", (), {})()
    model.layers = [_cls("InputLayer")([], "in"), _cls("Dropout")([], "drop_test"), _cls("Flatten")([], "f"), _cls("Concatenate")([], "c")]
    for i, L in enumerate(want):
        model.layers.append(_cls("Dense")([L["W"], L["b"]], "d%d" % i))
        if L["bn"] is not None:
            model.layers += [_cls("BatchNormalization")(list(L["bn"]), "bn%d" % i), _cls("Dropout")([], "drop%d" % i)]
    same(mm.weights.layers_from_keras_model(model), want)
    path = mm.weights.save_npz(str(tmp_path / "real_weights.npz"), want)
    same(mm.weights.load_npz(path), want)
    mm.weights.save_npz(str(tmp_path / "imag_weights.npz"), nets["imag"])
    same(mm.weights.load_nets(str(tmp_path))["imag"], nets["imag"])
    # loud failures
    bad = [dict(L) for L in want]
    bad[-1] = dict(bad[-1], bn=want[0]["bn"])
    with pytest.raises(ValueError):
        mm.weights.validate(bad)
    with pytest.raises(ValueError):
        mm.weights.validate([want[0], want[2]])                                   # widths do not chain
    model.layers.append(_cls("Conv1D")([np.zeros((3, 1, 4))], "conv"))
    with pytest.raises(ValueError):
        mm.weights.layers_from_keras_model(model)
    with pytest.raises(FileNotFoundError):
        mm.weights.load_nets(str(tmp_path / "nowhere"))
    (tmp_path / "h5only").mkdir()
    (tmp_path / "h5only" / "real_weights-improvement.hdf5").write_bytes(b"")
    with pytest.raises(ImportError, match="keras_to_npz"):                        # no h5py here: the error names the way out
        mm.weights.find_net(str(tmp_path / "h5only"), "real")


def test_keras_to_npz_tool_cli(tmp_path, capsys):
    import importlib.util
    spec = importlib.util.spec_from_file_location("keras_to_npz", os.path.join(ROOT, "tools", "keras_to_npz.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    assert tool.main(["--modeldir", str(tmp_path)]) == 1                          # sources missing: message + exit code
    assert "missing" in capsys.readouterr().out
    with pytest.raises(SystemExit):
        tool.main([])
