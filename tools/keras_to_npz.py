#!/usr/bin/env python
"""Convert the reference's Keras artefacts into the TensorFlow-free npz files the engine loads
(`<d>_weights.npz`: W0,b0[,bn0_gamma,bn0_beta,bn0_mean,bn0_var],W1,b1,...).  Run it once, on a box that has the
reference's Python environment (README.md:26-30); the GPU box then needs neither TensorFlow nor h5py.

    # weights saved by training (massiveMIMO_CSI_prediction_DNN.py:278-281,328): needs h5py only
    python tools/keras_to_npz.py --modeldir <MODEL_DIR>/BS32_denoise_..._SNR120
    # SavedModel dirs written by --test (:411), the ones inference.py:15-16 loads: needs tensorflow
    python tools/keras_to_npz.py --modeldir <workdir> --saved-model
    # explicit files
    python tools/keras_to_npz.py --real a.hdf5 --imag b.hdf5 -o out_dir

Writes real_weights.npz / imag_weights.npz next to the sources (or into -o) and prints the layer shapes.
"""
import argparse
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _weights_module():
    # weights.py is plain numpy (+ lazy h5py / tensorflow): load it standalone so this tool never needs the CUDA library
    spec = importlib.util.spec_from_file_location(
        "mamimo_b200_weights", os.path.join(ROOT, "dl-channel-estimation-mamimo_b200", "weights.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--modeldir", help="directory holding {real,imag}_weights-improvement.hdf5 or {real,imag}_keras_model")
    ap.add_argument("--saved-model", action="store_true", help="read the <d>_keras_model SavedModel dirs (needs tensorflow)")
    ap.add_argument("--real", help="explicit source of the real net (.hdf5/.h5 file or SavedModel dir)")
    ap.add_argument("--imag", help="explicit source of the imaginary net")
    ap.add_argument("-o", "--outdir", help="where to write the npz files (default: next to the sources)")
    args = ap.parse_args(argv)
    w = _weights_module()
    src = {}
    for d in w.DIMS:
        path = getattr(args, d)
        if not path:
            if not args.modeldir:
                ap.error("give --modeldir or both --real and --imag")
            path = os.path.join(args.modeldir, d + ("_keras_model" if args.saved_model else "_weights-improvement.hdf5"))
        src[d] = path
    outdir = args.outdir or args.modeldir or os.path.dirname(os.path.abspath(src["real"]))
    os.makedirs(outdir, exist_ok=True)
    for d, path in src.items():
        if not os.path.exists(path):
            print("missing: %s" % path)
            return 1
        layers = w.load_saved_model(path) if os.path.isdir(path) else w.load_keras_hdf5(path)
        out = w.save_npz(os.path.join(outdir, d + "_weights.npz"), layers)
        print("%s -> %s" % (path, out))
        for i, L in enumerate(layers):
            print("   dense%d %s%s" % (i, tuple(L["W"].shape), "  + BatchNormalization" if L["bn"] is not None else ""))
    return 0


if __name__ == "__main__":
    sys.exit(main())
