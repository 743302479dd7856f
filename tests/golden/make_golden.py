#!/usr/bin/env python
"""Generate tests/golden/*.npz by running / parsing the REAL reference in the
build container (where /root/reference exists).  The vectors are committed; the
GPU box never needs the reference.

What is pinned here:
  1. ltf256 / carriers: parsed numerically out of the MATLAB source text of
     helperMIMOChannelEstimate.m:16-23 and generate_maMIMO_LTF.m:99-102.
  2. inference.py's CSIPredictor.inference() executed unmodified, with a stub
     ``tensorflow.keras`` whose load_model() returns a deterministic numpy MLP
     (TensorFlow itself is not installed).  Pins pre/post-processing glue.
  3. massiveMIMO_dataGenerator.DataGenerator executed unmodified (stub
     ``tensorflow.keras.utils.Sequence``) on a synthetic pickle-shaped dataset
     built with create_massiveMIMO_CSIest_dnn_dataset.py:62's row formula.
     Pins per-pair input assembly and pair ordering.
  4. helperMIMOChannelEstimate.m and LMMSE_ce.m (MATLAB) executed UNMODIFIED by the MATLAB-subset interpreter
     tests/golden/mini_matlab.py (MATLAB / Octave are not installed): LS estimate, ltf(ind), the isMMSE branch with
     per-rx SNR, the data-phase case numSTS = 1, and LMMSE_ce called directly with Nps = 2.  helperGetP (a MathWorks
     example helper that is not in the reference repo) is supplied as Sylvester-Hadamard.  Pins oracle.ls and
     oracle.lmmse, and through them the CUDA path.  Also NMSE_subk and the CSI(:,iTX,iRX) rebuild loop of
     BER_test_maMIMO_LTF.m (:675-686, :213-218).

Usage:  python tests/golden/make_golden.py   (writes next to this file)
"""
import os
import re
import sys
import types
import numpy as np

REF = os.environ.get("MAMIMO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------- 1. tables
def _matlab_vec(expr, env):
    """Evaluate a MATLAB column-vector literal made of numbers, a:b ranges,
    zeros(n,1) and identifiers in env.  Separators: ';', ',', whitespace."""
    expr = expr.replace("...", " ")
    expr = re.sub(r"zeros\((\d+),1\)", lambda m: "Z%s" % m.group(1), expr)
    out = []
    for tok in re.split(r"[;,\s]+", expr.strip("[] \n\t'")):
        if not tok:
            continue
        if tok.startswith("Z"):
            out += [0] * int(tok[1:])
        elif tok in env:
            out += list(env[tok])
        elif ":" in tok:
            a, b = tok.split(":")
            out += list(range(int(eval(a)), int(eval(b)) + 1))
        else:
            out.append(int(eval(tok)))
    return out


def parse_tables():
    src = open(os.path.join(REF, "packet_generation/phased_arr/helperMIMOChannelEstimate.m")).read()
    env = {}
    for name in ("ltfLeft", "ltfRight", "ltf"):
        m = re.search(r"^%s\s*=\s*\[(.*?)\];" % name, src, re.S | re.M)
        env[name] = _matlab_vec(m.group(1), env)
    gen = open(os.path.join(REF, "packet_generation/phased_arr/generate_maMIMO_LTF.m")).read()
    nulls = _matlab_vec(re.search(r"prm\.NullCarrierIndices\s*=\s*\[(.*?)\]", gen).group(1), {})
    pilots = _matlab_vec(re.search(r"prm\.PilotCarrierIndices\s*=\s*\[(.*?)\]", gen).group(1), {})
    fft_len = int(re.search(r"prm\.FFTLength\s*=\s*(\d+)", gen).group(1))
    carriers = sorted(set(range(1, fft_len + 1)) - set(nulls) - set(pilots))   # setdiff, :102
    return dict(ltf256=np.asarray(env["ltf"], np.int8), nulls=np.asarray(nulls, np.int32),
                pilots=np.asarray(pilots, np.int32), carriers=np.asarray(carriers, np.int32))


# ---------------------------------------------------------------- stub TF
class _NumpyKerasModel:
    """Keras-like model: Dense(relu)->BN->Dense(relu)->BN->Dense(linear), float64."""

    def __init__(self, seed, d_in, hidden, d_out):
        rng = np.random.default_rng(seed)
        dims = [d_in] + list(hidden) + [d_out]
        self.layers_ = []
        for i in range(len(dims) - 1):
            lim = np.sqrt(6.0 / (dims[i] + dims[i + 1]))
            L = dict(W=rng.uniform(-lim, lim, (dims[i], dims[i + 1])),
                     b=rng.uniform(-0.1, 0.1, dims[i + 1]))
            if i < len(dims) - 2:
                n = dims[i + 1]
                L["bn"] = (rng.uniform(0.5, 1.5, n), rng.uniform(-0.1, 0.1, n),
                           rng.uniform(-0.1, 0.1, n), rng.uniform(0.5, 1.5, n))
            self.layers_.append(L)

    def predict(self, x, batch_size=None):
        h = np.asarray(x, np.float64)
        for i, L in enumerate(self.layers_):
            h = h @ L["W"] + L["b"]
            if i < len(self.layers_) - 1:
                h = np.maximum(h, 0.0)
                g, be, mu, var = L["bn"]
                h = g * (h - mu) / np.sqrt(var + 1e-3) + be
        return h

    def summary(self):
        pass


def install_stub_tf(model_factory):
    tf = types.ModuleType("tensorflow")
    keras = types.ModuleType("tensorflow.keras")
    models = types.ModuleType("tensorflow.keras.models")
    utils = types.ModuleType("tensorflow.keras.utils")
    models.load_model = model_factory
    utils.Sequence = object
    keras.models = models
    keras.utils = utils
    tf.keras = keras
    sys.modules.update({"tensorflow": tf, "tensorflow.keras": keras,
                        "tensorflow.keras.models": models, "tensorflow.keras.utils": utils})


def flatten_layers(prefix, layers, out):
    for i, L in enumerate(layers):
        out["%s_W%d" % (prefix, i)] = L["W"]
        out["%s_b%d" % (prefix, i)] = L["b"]
        if "bn" in L:
            for nm, t in zip(("gamma", "beta", "mean", "var"), L["bn"]):
                out["%s_bn%d_%s" % (prefix, i, nm)] = t


# ---------------------------------------------------------------- 2. inference.py
def run_inference_py():
    d_in, hidden, d_out = 52, (48, 40), 52
    nets = {"real": _NumpyKerasModel(6701, d_in, hidden, d_out),
            "imag": _NumpyKerasModel(6702, d_in, hidden, d_out)}

    def load_model(path):
        return nets["real"] if os.path.basename(path).startswith("real") else nets["imag"]

    install_stub_tf(load_model)
    sys.path.insert(0, REF)
    import inference as ref_inference                      # the unmodified reference file
    pred = ref_inference.CSIPredictor("/nonexistent/model_dir")
    rng = np.random.default_rng(67)
    X = (rng.standard_normal((9, d_in)) + 1j * rng.standard_normal((9, d_in))).astype(np.complex128)
    Y = pred.inference(X)
    out = dict(X=X, Y=Y)
    flatten_layers("real", nets["real"].layers_, out)
    flatten_layers("imag", nets["imag"].layers_, out)
    return out


# ---------------------------------------------------------------- 3. DataGenerator
def run_data_generator():
    sys.path.insert(0, REF)
    import massiveMIMO_dataGenerator as ref_dg             # unmodified reference file
    n_pkt, n_rx, n_tx, len_ltf, nsc = 3, 2, 4, 24, 10
    rng = np.random.default_rng(67)
    ltf = rng.standard_normal((n_pkt, n_rx, len_ltf)) + 1j * rng.standard_normal((n_pkt, n_rx, len_ltf))
    P = rng.choice([-1.0, 1.0], size=(n_tx, n_tx))
    y = rng.standard_normal((n_pkt * n_rx * n_tx, nsc)) + 1j * rng.standard_normal((n_pkt * n_rx * n_tx, nsc))
    X = np.zeros((n_pkt * n_rx * n_tx, 2), dtype=np.int64)
    LTF = {}
    for p in range(n_pkt):
        for irx in range(n_rx):
            h = 1000 + p * n_rx + irx                       # stands in for the random 32-bit hash (:52-59)
            LTF[h] = {"real": ltf[p, irx].real.copy(), "imag": ltf[p, irx].imag.copy()}
            for itx in range(n_tx):
                samp_ix = p * (n_rx * n_tx) + irx * n_tx + itx   # create_..._dataset.py:62
                X[samp_ix] = [h, itx]
    dataset = {"X": X, "LTF": LTF, "P": P, "y": {"real": y.real.copy(), "imag": y.imag.copy()}}
    prm = {"lenLTF": len_ltf, "nTX": n_tx, "nRX": n_rx, "nSubCarr": nsc}
    out = dict(ltf=ltf, P=P, y=y, n_pkt=n_pkt, n_rx=n_rx, n_tx=n_tx)
    for d in ("real", "imag"):
        gen = ref_dg.DataGenerator(list(range(X.shape[0])), dataset, d, prm,
                                   datasource="matlab_maMimo", method="default", batch_size=n_tx * n_rx)
        gen.reorder_indexes()                               # test-mode order, ..._DNN.py:337
        xs, xp, ys = [], [], []
        for b in range(len(gen)):
            (Xsig, Xp), yb, _ = gen[b]
            xs.append(Xsig[:, :, 0]); xp.append(Xp); ys.append(yb)
        out["Xsig_" + d] = np.concatenate(xs); out["Xp_" + d] = np.concatenate(xp)
        out["y_" + d] = np.concatenate(ys)
    return out


# ---------------------------------------------------------------- 4. DataGenerator method='reshape' (OFDM demod mirror)
def run_data_generator_reshape():
    """massiveMIMO_dataGenerator.py:425-458 executed unmodified: the author's numpy mirror of ofdmdemod
    (F-order reshape, CP removal with symOffset, FFT, fftshift).  Pins the oracle's OFDM front-end."""
    sys.path.insert(0, REF)
    import massiveMIMO_dataGenerator as ref_dg
    out = {}
    for tag, (fft_len, cp_len, sym_off, n_sym) in {"a": (16, 4, 4, 4), "b": (32, 8, 3, 2), "c": (64, 16, 16, 1)}.items():
        n_tx, n_rx, n_pkt, nsc = n_sym, 2, 2, 5
        rng = np.random.default_rng([67, fft_len])
        len_ltf = (fft_len + cp_len) * n_sym
        ltf = rng.standard_normal((n_pkt, n_rx, len_ltf)) + 1j * rng.standard_normal((n_pkt, n_rx, len_ltf))
        P = rng.choice([-1.0, 1.0], size=(n_tx, n_tx))
        ltf_freq = rng.choice([-1.0, 1.0], size=nsc)
        X = np.zeros((n_pkt * n_rx * n_tx, 2), dtype=np.int64)
        LTF = {}
        for p in range(n_pkt):
            for irx in range(n_rx):
                h = 5000 + p * n_rx + irx
                LTF[h] = {"real": ltf[p, irx].real.copy(), "imag": ltf[p, irx].imag.copy()}
                for itx in range(n_tx):
                    X[p * (n_rx * n_tx) + irx * n_tx + itx] = [h, itx]
        y = np.zeros((X.shape[0], nsc))
        dataset = {"X": X, "LTF": LTF, "P": P, "y": {"real": y, "imag": y}}
        prm = {"lenLTF": len_ltf, "nTX": n_tx, "nRX": n_rx, "nSubCarr": nsc, "FFTLength": fft_len, "CPLen": cp_len,
               "numSym": n_sym, "symOffset": sym_off, "ltf_freqdom": ltf_freq}
        res = {}
        for d in ("real", "imag"):
            gen = ref_dg.DataGenerator(list(range(X.shape[0])), dataset, d, prm, datasource="matlab_maMimo",
                                       method="reshape", batch_size=n_tx * n_rx)
            gen.reorder_indexes()
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")            # the reference stores a complex FFT into a real array
                res[d] = np.concatenate([gen[b][0] for b in range(len(gen))])
        out["ltf_" + tag] = ltf
        out["cfg_" + tag] = np.asarray([fft_len, cp_len, sym_off, n_sym, n_rx, n_pkt])
        out["Xreal_" + tag] = res["real"]
        out["Ximag_" + tag] = res["imag"]
    return out


# ---------------------------------------------------------------- 4. the MATLAB hot path, interpreted
def run_matlab_hot_path():
    sys.path.insert(0, HERE)
    from mini_matlab import MatlabFile
    pg = os.path.join(REF, "packet_generation", "phased_arr")
    src_h = open(os.path.join(pg, "helperMIMOChannelEstimate.m"), encoding="latin-1").read()
    src_l = open(os.path.join(pg, "LMMSE_ce.m"), encoding="latin-1").read()

    def hadamard(n):
        n = int(np.asarray(n).item())
        H = np.ones((1, 1))
        while H.shape[0] < n:
            H = np.block([[H, H], [H, -H]])
        return H

    m = MatlabFile(src_h, src_l, externals={"helperGetP": hadamard})
    tabs = parse_tables()
    car = tabs["carriers"].astype(np.float64).reshape(-1, 1)
    rng = np.random.default_rng(67)
    out = {}
    # A: sounding phase, 4 x 2, isMMSE = true with SNR(i) per rx and a delay vector as LMMSE_ce's `h`
    # B: 8 x 3, isMMSE = false        C: data phase numSTS = nltf = 1 (generate_maMIMO_LTF.m:578)
    for tag, nt, nr, mmse in (("A", 4, 2, 1.0), ("B", 8, 3, 0.0), ("C", 1, 2, 0.0)):
        rx = rng.standard_normal((car.size, nt, nr)) + 1j * rng.standard_normal((car.size, nt, nr))
        tau = np.abs(rng.standard_normal((1, 100))) * 2.5
        snr = rng.uniform(0.0, 25.0, (nr, 1))
        prm = {"numSTS": float(nt), "CarriersLocations": car}
        hD, P, ltf_o, hM = m.call("helperMIMOChannelEstimate", [rx, prm, 1.0, tau, snr, mmse], 4)
        out.update({"rx_" + tag: rx, "tau_" + tag: tau, "snr_" + tag: snr, "hD_" + tag: hD, "P_" + tag: P,
                    "ltf_o_" + tag: ltf_o, "hDmmse_" + tag: hM})
    # LMMSE_ce called directly: comb pilots Nps = 2, and the reference-style tiny delays (seconds) at 30 dB
    for tag, n, nps, h, snr in (("nps2", 48, 2.0, np.exp(-np.arange(8) / 3.0).reshape(1, -1), 12.0),
                                ("sec", 64, 1.0, np.abs(rng.standard_normal((1, 100))) * 1e-7, 30.0)):
        x = rng.standard_normal((n, 1)) + 1j * rng.standard_normal((n, 1))
        y = m.call("LMMSE_ce", [x, float(n), float(n), nps, h, snr], 1)[0]
        out.update({"ce_x_" + tag: x, "ce_h_" + tag: h, "ce_y_" + tag: y, "ce_par_" + tag: np.array([n, nps, snr])})
    out["carriers"] = tabs["carriers"]
    # NMSE_subk (sub-function at the end of BER_test_maMIMO_LTF.m, :675-686) and the loop that rebuilds
    # CSI(:,iTX,iRX) from the prediction rows (:213-218; the literal lines wrapped in a function header)
    ber = open(os.path.join(pg, "BER_test_maMIMO_LTF.m"), encoding="latin-1").read()
    sub = ber[ber.index("function [out] = NMSE_subk"):]
    lines = ber.splitlines()
    i0 = next(i for i, l in enumerate(lines) if l.strip().startswith("for iRX = 1:nRXAnts"))
    loop = "\n".join(lines[i0:i0 + 6])
    assert "CSI_dnn_imag(:,iTX,iRX)" in loop and loop.rstrip().endswith("end")
    wrap = ("function [CSI_dnn_real, CSI_dnn_imag] = rebuild(predicted_CSI_real, predicted_CSI_imag, nRXAnts, nTXAnts)\n"
            + loop + "\nend\n")
    m2 = MatlabFile(sub, wrap)
    Href = rng.standard_normal((52, 4, 3)) + 1j * rng.standard_normal((52, 4, 3))
    Hest = Href + 0.1 * (rng.standard_normal((52, 4, 3)) + 1j * rng.standard_normal((52, 4, 3)))
    out.update(nmse_ref=Href, nmse_est=Hest, nmse_val=m2.call("NMSE_subk", [Href, Hest], 1)[0],
               nmse_zero=m2.call("NMSE_subk", [Href, Href], 1)[0], nmse_one=m2.call("NMSE_subk", [Href, 0 * Href], 1)[0])
    pr, pi = rng.standard_normal((12, 52)), rng.standard_normal((12, 52))
    cr, ci = m2.call("rebuild", [pr, pi, 3.0, 4.0], 2)
    out.update(rebuild_pred_real=pr, rebuild_pred_imag=pi, rebuild_csi_real=cr, rebuild_csi_imag=ci)
    return out


# ---------------------------------------------------------------- 6. BER_test_maMIMO_LTF.m reader of the prediction files
def ber_test_reader_source(ref=REF):
    """The literal lines of pg/BER_test_maMIMO_LTF.m that consume one packet's pair of prediction files -- the
    isSeparateFiles branch (:198-221: file names, load, .y / .x(:,1:lenIn), CSI rebuild loop), `CSI_dnn = complex(...)`
    (:226) and the rebuild of the time-domain rx signal from the x planes (:312-318) -- wrapped in a function header so
    mini_matlab can run them.  Nothing but the header / trailer is written here."""
    ber = open(os.path.join(ref, "packet_generation/phased_arr/BER_test_maMIMO_LTF.m"), encoding="latin-1").read()
    lines = ber.splitlines()
    a0 = next(i for i, l in enumerate(lines) if l.strip().startswith("real_file_pkt = insertBefore("))
    a1 = next(i for i in range(a0, len(lines)) if lines[i].strip() == "end" and lines[i + 1].strip() == "end"
              and "CSI_dnn_imag(:,iTX,iRX)" in lines[i - 1]) + 1            # closes `for iTX`, then `for iRX`
    part_a = lines[a0:a1 + 1]
    assert "load(real_file_pkt);" in "".join(part_a) and "inputLTF_IMAG = inputLTF_IMAG(:,1:lenIn);" in "\n".join(part_a)
    b0 = next(i for i, l in enumerate(lines) if l.strip().startswith("CSI_dnn = complex(CSI_dnn_real,CSI_dnn_imag)"))
    c0 = next(i for i, l in enumerate(lines) if l.strip().startswith("inputRXSig_real = zeros(lenIn,nRXAnts);"))
    c1 = next(i for i in range(c0, len(lines)) if lines[i].strip().startswith("inputRXSig = complex(inputRXSig_real,inputRXSig_imag);"))
    body = "\n".join(part_a + [lines[b0]] + lines[c0:c1 + 1])
    return ("function [CSI_dnn, inputRXSig] = ber_reader(real_csi_pred, imag_csi_pred, p, lenIn, nSubCar, nTXAnts, nRXAnts)\n"
            "CSI_dnn_real = zeros(nSubCar,nTXAnts,nRXAnts);\nCSI_dnn_imag = zeros(nSubCar,nTXAnts,nRXAnts);\n"   # :182-183
            + body + "\nend\n")


def run_ber_test_reader():
    """Files written by this repo's writer (pipeline.write_prediction_files, no GPU involved), read back by the
    reference's own lines.  Stores the planes that went in and what MATLAB's side makes of them."""
    import tempfile
    from mini_matlab import MatlabFile
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import mamimo_b200 as mm
    n_tx, n_rx, n_sc, len_in, n_pkt = 4, 3, 10, 24, 2
    rng = np.random.default_rng(6705)
    rows = n_tx * n_rx
    y_r = rng.standard_normal((n_pkt * rows, n_sc)).astype(np.float32)
    y_i = rng.standard_normal((n_pkt * rows, n_sc)).astype(np.float32)
    sig = rng.standard_normal((n_pkt, n_rx, len_in)) + 1j * rng.standard_normal((n_pkt, n_rx, len_in))
    x_r = np.repeat(sig.real, n_tx, axis=1).reshape(-1, len_in)
    x_i = np.repeat(sig.imag, n_tx, axis=1).reshape(-1, len_in)
    m = MatlabFile(ber_test_reader_source())
    out = dict(n_tx=n_tx, n_rx=n_rx, len_in=len_in, y_real=y_r, y_imag=y_i, x_real=x_r, x_imag=x_i)
    with tempfile.TemporaryDirectory() as td:
        mm.pipeline.write_prediction_files(td, y_r, y_i, n_tx, n_rx, x_r, x_i, first_pkt_id=1)
        for p in range(1, n_pkt + 1):
            csi, rx = m.call("ber_reader", [os.path.join(td, "test_csi_predictions_real.mat"),
                                            os.path.join(td, "test_csi_predictions_imag.mat"),
                                            float(p), float(len_in), float(n_sc), float(n_tx), float(n_rx)], 2)
            out["csi_dnn_%d" % p] = csi
            out["input_rx_sig_%d" % p] = rx
    return out


# ---------------------------------------------------------------- 7. omphybweights.m: the per-subcarrier SVD lines
def run_omp_svd_lines():
    """getWeightsForSubcarrier's first statements (pg/omphybweights.m:174-176), literally:
    H = Hin.'; [~,~,v] = svd(H); Fopt = v(:,1:Ns);  -- svd itself is LAPACK (numpy here, MATLAB there)."""
    from mini_matlab import MatlabFile
    src = open(os.path.join(REF, "packet_generation/phased_arr/omphybweights.m"), encoding="latin-1").read()
    lines = src.splitlines()
    i0 = next(i for i, l in enumerate(lines) if l.strip() == "H = Hin.';")
    body = lines[i0:i0 + 3]
    assert body[1].strip() == "[~,~,v] = svd(H);" and body[2].strip() == "Fopt = v(:,1:Ns);"
    wrap = "function [Fopt, H] = topsvd(Hin, Ns)\n" + "\n".join(body) + "\nend\n"

    def svd(A):
        U, S, Vh = np.linalg.svd(np.asarray(A), full_matrices=True)
        Sm = np.zeros(A.shape)
        Sm[:len(S), :len(S)] = np.diag(S)
        return U, Sm, Vh.conj().T

    m = MatlabFile(wrap, externals={"svd": svd})
    rng = np.random.default_rng(6706)
    out = {}
    for tag, (nt, nr, n) in {"a": (8, 2, 5), "b": (32, 4, 6), "c": (4, 4, 3)}.items():
        Hin = rng.standard_normal((n, nt, nr)) + 1j * rng.standard_normal((n, nt, nr))      # n "subcarriers" of [Nt x Nr]
        if tag == "b":
            Hin[2] *= 1e-3                                                                   # amplitude spread
            Hin[4, :, 3] = Hin[4, :, 2] * (0.5 - 0.2j)                                        # one rank-deficient matrix
        F = np.stack([m.call("topsvd", [Hin[k], float(nt)], 1)[0] for k in range(n)])
        out["Hin_" + tag], out["Fopt_" + tag] = Hin, F
    return out


# ---------------------------------------------------------------- 8. ompdecomp.m loop + omphybweights.m weights lines
def run_omp_lines():
    """ompdecomp's loop (pg/ompdecomp.m:98-121, default identity weight) and getWeightsForSubcarrier's precoding lines
    (pg/omphybweights.m:178-179,196-197), taken from the reference text and executed by mini_matlab.  The argument
    parsing / validation around them (inputParser, validateattributes) is replaced by a plain signature."""
    from mini_matlab import MatlabFile
    base = os.path.join(REF, "packet_generation/phased_arr")
    omp = open(os.path.join(base, "ompdecomp.m"), encoding="latin-1").read().splitlines()
    i0 = next(i for i, l in enumerate(omp) if l.strip() == "Watom_temp = complex(zeros(Nelem,Nsparsity));")
    i1 = next(i for i, l in enumerate(omp) if l.strip() == "WatomIdx = WatomIdx_temp(1:Ns);")
    size_line = next(l for l in omp if l.strip() == "[Nelem,Nw] = size(Wopt);")
    loop = omp[i0:i1 + 1]
    assert any(l.strip() == "while m <= Nsparsity && Errnorm > eps" for l in loop)
    hyb = open(os.path.join(base, "omphybweights.m"), encoding="latin-1").read().splitlines()
    j0 = next(i for i, l in enumerate(hyb) if l.strip() == "[Fbb,Frf] = ompdecomp(Fopt,At,'MaxSparsity',NtRF);")
    assert hyb[j0 + 1].strip() == "Fbb = sqrt(Ns)*Fbb/norm(Frf*Fbb,'fro');"
    outs = [l for l in hyb if l.strip() in ("Fbb_out = Fbb.';", "Frf_out = Frf.';")][:2]
    assert len(outs) == 2
    wrap = ("function [Wcoeff,Watom,WatomIdx,Errnorm] = ompdecomp(Wopt,Adict,pname,Nsparsity)\n" + size_line + "\n"
            "W = eye(Nelem);\n" + "\n".join(loop) + "\nend\n"
            "function [Fbb_out,Frf_out] = weights(Fopt,Ns,NtRF,At)\n" + hyb[j0] + "\n" + hyb[j0 + 1] + "\n" +
            "\n".join(outs) + "\nend\n")
    m = MatlabFile(wrap)
    rng = np.random.default_rng(6707)
    out = {}

    def steer(nt, nrays):                                   # unit-modulus columns, like steervec's output
        return np.exp(2j * np.pi * rng.random((nt, nrays)))

    def orthonormal(nt, ns):
        q, _ = np.linalg.qr(rng.standard_normal((nt, ns)) + 1j * rng.standard_normal((nt, ns)))
        return q

    cases = {"a": (8, 2, 3, 40, 6), "b": (32, 1, 1, 500, 6), "c": (16, 4, 4, 64, 4), "d": (32, 2, 8, 120, 3)}
    for tag, (nt, ns, nrf, nrays, n) in cases.items():
        At = steer(nt, nrays)
        F = np.stack([orthonormal(nt, ns) for _ in range(n)])
        fbb, frf, idx, err = [], [], [], []
        for k in range(n):
            c, a, ix, e = m.call("ompdecomp", [F[k], At, "MaxSparsity", float(nrf)], 4)
            fo, ro = m.call("weights", [F[k], float(ns), float(nrf), At], 2)
            fbb.append(fo); frf.append(ro); idx.append(np.asarray(ix).ravel()); err.append(float(np.asarray(e).real.item()))
        out["At_" + tag], out["Fopt_" + tag] = At, F
        out["Fbb_" + tag], out["Frf_" + tag] = np.stack(fbb), np.stack(frf)
        out["idx_" + tag], out["err_" + tag] = np.stack(idx).astype(np.int64), np.asarray(err)
        out["cfg_" + tag] = np.asarray([nt, ns, nrf, nrays, n])
    # early stop: Fopt is exactly one dictionary column (entries in {1, j, -1, -j}: every product is exact)
    nt, nrays = 16, 24
    At = (1j) ** rng.integers(0, 4, size=(nt, nrays))
    Fe = (At[:, 7] / np.sqrt(nt)).reshape(nt, 1)
    c, a, ix, e = m.call("ompdecomp", [Fe, At, "MaxSparsity", 3.0], 4)
    out["At_e"], out["Fopt_e"], out["coef_e"] = At, Fe, np.asarray(c)
    out["idx_e"], out["err_e"] = np.asarray(ix).ravel().astype(np.int64), np.asarray(float(np.asarray(e).real.item()))
    return out


def main():
    np.savez_compressed(os.path.join(HERE, "ref_omp.npz"), **run_omp_lines())
    np.savez_compressed(os.path.join(HERE, "ref_svd.npz"), **run_omp_svd_lines())
    np.savez_compressed(os.path.join(HERE, "ref_ber_test_reader.npz"), **run_ber_test_reader())
    np.savez_compressed(os.path.join(HERE, "ref_matlab_ls_lmmse.npz"), **run_matlab_hot_path())
    install_stub_tf(lambda path: None)
    np.savez_compressed(os.path.join(HERE, "ref_data_generator_reshape.npz"), **run_data_generator_reshape())
    np.savez_compressed(os.path.join(HERE, "ref_tables.npz"), **parse_tables())
    np.savez_compressed(os.path.join(HERE, "ref_inference_py.npz"), **run_inference_py())
    np.savez_compressed(os.path.join(HERE, "ref_data_generator.npz"), **run_data_generator())
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
