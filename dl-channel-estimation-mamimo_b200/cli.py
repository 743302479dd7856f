"""argv-compatible stand-in for step 4 of full_pipeline_maMIMO_DNNEst.sh:44-48,

    massiveMIMO_CSI_prediction_DNN.py --test -x <pickle> --nn 1024 1024 -d <workdir> --modeldir <dir>
                                      --useGPU 0 --useBN --datasource matlab_maMimo --valSameTrain

Same flags, defaults and exit behaviour as the reference parser (massiveMIMO_CSI_prediction_DNN.py:3-34,104-115);
only the --test branch (:330-346,401-409) exists here -- training, the CONV1D variant and --input_opt stay with the
reference script.  What it does with them:

  -x pickle {X, y, LTF, P, simParams} written by create_massiveMIMO_CSIest_dnn_dataset.py:125
  sample order restored (reorder_indexes, :337), batch = nTX*nRX rows = one packet (:339)
  nets [LTF (lenLTF) || P(:, iTx) (nTX)] -> --nn ... -> nSubCarr on the CUDA engine (mode A, de-duplicated first layer)
  weights from --modeldir: <d>_weights.npz, else <d>_weights-improvement.hdf5 (:278-281; needs h5py)
  -> <workdir>/test_csi_predictions_<d>_<pkt>.mat {all_pkts_csi_nn_out: x, y, true_y} (:404-409), pkt ids 1-based,
     and <workdir>/<d>_weights.npz in place of the Keras SavedModel of :411 (what this repo's CSIPredictor loads).

Host-side plumbing only: predictions come from the engine passed in by `engine_factory` (default: the CUDA engine).
"""
import argparse
import os
import pickle
import sys

import numpy as np

from . import weights as _weights


def build_parser():
    """The reference's parser, flag for flag (massiveMIMO_CSI_prediction_DNN.py:3-33)."""
    p = argparse.ArgumentParser(description="Train-test CSI prediction network")
    g = p.add_mutually_exclusive_group()
    g.add_argument("--train", action="store_true", help="Trigger training of the network on train data")
    g.add_argument("--test", action="store_true", help="Trigger testing of the network on test data")
    g.add_argument("--input_opt")
    p.add_argument("--model", default="FC")
    p.add_argument("-x", required=True, help="Input datafile used for train/test")
    p.add_argument("-y", default="")
    p.add_argument("--datasource", required=True, default="matlab_maMimo")
    p.add_argument("-d", "--workdir", default="checkpoint")
    p.add_argument("--modeldir", default="")
    p.add_argument("--epochs", default=500, type=int)
    p.add_argument("--lr", default=0.0001, type=float)
    p.add_argument("--bs", default=256, type=int)
    p.add_argument("--nn", default=[256, 128], type=int, nargs="+")
    p.add_argument("--dropout", default=0.15, type=float)
    p.add_argument("--useBN", action="store_true")
    p.add_argument("--method", default="default")
    p.add_argument("--excludeBER", action="store_true")
    p.add_argument("--useKeras", action="store_true")
    p.add_argument("--useGPU", default="0")
    p.add_argument("--valTrainRatio", default=0.15, type=float)
    p.add_argument("--valSameTrain", action="store_true")
    p.add_argument("--execTime", action="store_true")
    p.add_argument("--testDropInput", action="store_true")
    p.add_argument("--inFraction", default=1)
    p.add_argument("--decimate_max", action="store_true")
    p.add_argument("--decimate_avg", action="store_true")
    p.add_argument("--onlyReal", action="store_true")
    p.add_argument("--onlyImag", action="store_true")
    # engine knobs (not in the reference)
    p.add_argument("--precision", default="fp16x3", choices=["tf32x3", "fp16x3", "fp32_simt"])
    p.add_argument("--chunk-pkts", default=64, type=int, help="packets per engine call")
    return p


def load_dataset(path):
    """loadDataset, matlab_maMimo branch (massiveMIMO_dataGenerator.py:20-54)."""
    with open(path, "rb") as f:
        ds = pickle.load(f)
    sp = dict(ds["simParams"])
    n_samples = int(ds["X"].shape[0])
    sp["lenLTF"] = int(np.asarray(ds["LTF"][ds["X"][0, 0]]["real"]).shape[0])
    sp["nSubCarr"] = int(ds["y"]["real"].shape[1])
    sp["nTX"], sp["nRX"] = int(sp["nTX"]), int(sp["nRX"])
    n_packets = n_samples / (sp["nTX"] * sp["nRX"])
    if not float(n_packets).is_integer():
        print("Num. of packets is not an integer. Please double check --nTX and --nRX arguments to match the provided "
              "dataset. Aborting...")
        raise SystemExit(-1)
    return ds, sp, int(n_packets)


def packet_signals(ds, sp, pkts):
    """Time-domain preamble [len(pkts), nRX, lenLTF] complex of the given packets: the LTF of sample
    p*(nRX*nTX) + iRx*nTX (create_massiveMIMO_CSIest_dnn_dataset.py:62) looked up by its hash (:50-59)."""
    n_tx, n_rx, L = sp["nTX"], sp["nRX"], sp["lenLTF"]
    out = np.empty((len(pkts), n_rx, L), dtype=np.complex128)
    for i, p in enumerate(pkts):
        for irx in range(n_rx):
            e = ds["LTF"][ds["X"][p * n_rx * n_tx + irx * n_tx, 0]]
            out[i, irx] = np.asarray(e["real"])[:L] + 1j * np.asarray(e["imag"])[:L]
    return out


def _default_engine_factory(sp, hidden, precision, device, max_pkts):
    from .engine import Engine
    return Engine(sp["nTX"], sp["nRX"], 8, hidden=hidden, d_in=sp["lenLTF"] + sp["nTX"], d_out=sp["nSubCarr"],
                  input_mode="time_p", len_ltf=sp["lenLTF"], precision=precision, device=device, max_pkts=max_pkts)


def main(argv=None, engine_factory=None):
    args = build_parser().parse_args(argv)
    if args.train or args.input_opt:
        print("This entry point replaces the --test branch only; run training with the reference script.")
        return 2
    if not args.test:
        return 0                                   # the reference does nothing without --train / --test either
    if not (os.path.exists(args.workdir) and os.path.isdir(args.workdir)):
        print("Given directory does not exists. Aborting...")            # :112-115
        return 0
    if args.datasource != "matlab_maMimo" or args.model != "FC":
        print("Only --datasource matlab_maMimo with --model FC is wired to the CUDA engine.")
        return 2
    if args.testDropInput or args.decimate_max or args.decimate_avg or str(args.inFraction) != "1":
        print("--testDropInput / --decimate_* / --inFraction change the network input; not supported by this entry point.")
        return 2
    from .pipeline import run_test_mode_time

    ds, sp, n_packets = load_dataset(args.y if args.y else args.x)
    if args.y or args.valSameTrain:                                        # :124-140
        pkts = list(range(n_packets))
    else:
        n_test = int(np.floor(n_packets * args.valTrainRatio))
        pkts = list(range(n_packets - n_test, n_packets))
    model_dir = args.modeldir if args.modeldir else args.workdir          # :278-281
    nets = _weights.load_nets(model_dir)
    hidden = [int(np.asarray(L["W"]).shape[1]) for L in nets["real"][:-1]]
    if hidden != list(args.nn):
        print("--nn %s does not match the stored network %s. Aborting..." % (args.nn, hidden))
        return 2
    d_in = int(np.asarray(nets["real"][0]["W"]).shape[0])
    if d_in != sp["lenLTF"] + sp["nTX"] or int(np.asarray(nets["real"][-1]["W"]).shape[1]) != sp["nSubCarr"]:
        print("stored network (%d inputs) does not match the dataset (lenLTF %d + nTX %d). Aborting..."
              % (d_in, sp["lenLTF"], sp["nTX"]))
        return 2
    has_bn = any(L.get("bn") is not None for L in nets["real"])
    if has_bn != bool(args.useBN):
        print("WARNING: --useBN %s but the stored network %s BatchNormalization; using the stored network."
              % (args.useBN, "has" if has_bn else "has no"))
    chunk = max(1, int(args.chunk_pkts))
    factory = engine_factory or _default_engine_factory
    eng = factory(sp, hidden, args.precision, int(str(args.useGPU).split(",")[0] or 0), chunk)
    try:
        # the pickle holds P as h5py read it (MATLAB P transposed; the generator feeds P[:, iTx], :311): the engine
        # wants row j = code of tx j
        eng.set_pilots(None, np.asarray(ds["P"], dtype=np.float64).T)
        eng.load_weights(nets)
        rows = sp["nTX"] * sp["nRX"]
        dims_done = 0
        for c0 in range(0, len(pkts), chunk):
            sel = pkts[c0:c0 + chunk]
            sig = packet_signals(ds, sp, sel)
            ridx = np.concatenate([np.arange(p * rows, (p + 1) * rows) for p in sel])
            run_test_mode_time(eng, sig, args.workdir, true_real=ds["y"]["real"][ridx], true_imag=ds["y"]["imag"][ridx],
                               first_pkt_id=c0 + 1)                         # file ids count test packets from 1 (:408)
            dims_done += len(sel)
    finally:
        if hasattr(eng, "close"):
            eng.close()
    if args.onlyReal or args.onlyImag:              # the reference would have written one plane only (:166-170)
        drop = "imag" if args.onlyReal else "real"
        for i in range(len(pkts)):
            os.remove(os.path.join(args.workdir, "test_csi_predictions_%s_%d.mat" % (drop, i + 1)))
    for d in _weights.DIMS:                         # stands in for CSI_predictor.save(<d>_keras_model), :411
        _weights.save_npz(os.path.join(args.workdir, d + "_weights.npz"), nets[d])
    print("wrote %d packets x 2 planes to %s" % (dims_done, args.workdir))
    return 0


if __name__ == "__main__":
    sys.exit(main())
