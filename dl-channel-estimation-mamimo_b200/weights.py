"""Weight ingestion for the FC nets: Keras artefacts of the reference -> the flat layer list the engine loads.

The reference leaves three kinds of artefact per net d in {real, imag}:
  <modeldir>/<d>_weights-improvement.hdf5   Model.save_weights after training (massiveMIMO_CSI_prediction_DNN.py:278-281,
                                            loaded by the --test branch at :334)
  <workdir>/<d>_keras_model                 Model.save (SavedModel dir) at the end of --test (:411), what
                                            inference.py:15-16 loads
and this repo adds the TensorFlow-free form the engine reads directly:
  <dir>/<d>_weights.npz                     W0,b0[,bn0_gamma,bn0_beta,bn0_mean,bn0_var],W1,b1,...

Layer list = [dict(W [in,out] Keras kernel, b [out], bn = None | (gamma, beta, moving_mean, moving_var)), ...]:
Dense(relu) [-> BatchNormalization] ... -> Dense(linear)  (:211-227; Dropout / Flatten / Concatenate / Input /
GaussianNoise carry no weights and are skipped).  Everything here is host-side file plumbing: no math.

h5py / tensorflow are imported lazily and only by the functions that need them (neither is needed to RUN the
engine; tools/keras_to_npz.py converts on a box that has them).
"""
import os

import numpy as np

DIMS = ("real", "imag")
BN_KEYS = ("gamma", "beta", "mean", "var")


# ------------------------------------------------------------------------------------------ npz (engine-native)
def save_npz(path, layers):
    out = {}
    for i, L in enumerate(layers):
        out["W%d" % i] = np.asarray(L["W"], np.float32)
        out["b%d" % i] = np.asarray(L["b"], np.float32)
        if L.get("bn") is not None:
            for k, t in zip(BN_KEYS, L["bn"]):
                out["bn%d_%s" % (i, k)] = np.asarray(t, np.float32)
    np.savez(path, **out)
    return path


def load_npz(path):
    z = np.load(path)
    layers, i = [], 0
    while "W%d" % i in z:
        L = {"W": z["W%d" % i], "b": z["b%d" % i], "bn": None}
        if "bn%d_gamma" % i in z:
            L["bn"] = tuple(z["bn%d_%s" % (i, k)] for k in BN_KEYS)
        layers.append(L)
        i += 1
    if not layers:
        raise ValueError("%s holds no W0/b0 arrays" % path)
    validate(layers)
    return layers


def validate(layers):
    """shape chain of a Dense stack; BN only after hidden layers (the final Dense(linear) has none, :227)"""
    for i, L in enumerate(layers):
        W, b = np.asarray(L["W"]), np.asarray(L["b"])
        if W.ndim != 2 or b.shape != (W.shape[1],):
            raise ValueError("layer %d: kernel %s / bias %s do not form a Dense layer" % (i, W.shape, b.shape))
        if i and np.asarray(layers[i - 1]["W"]).shape[1] != W.shape[0]:
            raise ValueError("layer %d: input width %d does not match the previous layer's %d outputs"
                             % (i, W.shape[0], np.asarray(layers[i - 1]["W"]).shape[1]))
        if L.get("bn") is not None:
            if i == len(layers) - 1:
                raise ValueError("the final Dense(linear) layer has no BatchNormalization")
            if any(np.asarray(t).shape != b.shape for t in L["bn"]):
                raise ValueError("layer %d: BatchNormalization vectors must have %d entries" % (i, b.size))
    return layers


# ------------------------------------------------------------------------------------------ Keras objects
def layers_from_keras_model(model):
    """A loaded tf.keras Model (or anything with .layers whose items have .get_weights() and a class name):
    Dense -> [kernel, bias]; BatchNormalization -> [gamma, beta, moving_mean, moving_variance] of the Dense before."""
    layers = []
    for lyr in model.layers:
        kind = type(lyr).__name__
        w = lyr.get_weights()
        if kind == "Dense":
            if len(w) != 2:
                raise ValueError("Dense layer %r without a bias is not supported" % getattr(lyr, "name", "?"))
            layers.append({"W": np.asarray(w[0]), "b": np.asarray(w[1]), "bn": None})
        elif kind == "BatchNormalization":
            if not layers or len(w) != 4:
                raise ValueError("BatchNormalization must follow a Dense layer and carry gamma, beta, mean, variance")
            layers[-1]["bn"] = tuple(np.asarray(t) for t in w)
        elif w:
            raise ValueError("layer %r (%s) carries weights this engine has no kernel for" % (getattr(lyr, "name", "?"), kind))
    return validate(layers)


# ------------------------------------------------------------------------------------------ Keras HDF5 (save_weights / .h5)
def _names(attr):
    return [n.decode() if isinstance(n, bytes) else str(n) for n in attr]


def layers_from_keras_hdf5(f):
    """f: an open h5py.File (or the 'model_weights' group of a full-model .h5) written by Keras 2.x
    Model.save_weights: attrs['layer_names'] in model order; each layer group has attrs['weight_names'] such as
    'fc_dense0/kernel:0', 'fc_dense0/bias:0', 'batch_normalization/gamma:0', '.../moving_variance:0'."""
    if "layer_names" not in f.attrs and "model_weights" in f:
        f = f["model_weights"]
    layers = []
    for lname in _names(f.attrs["layer_names"]):
        g = f[lname]
        wn = _names(g.attrs["weight_names"])
        if not wn:
            continue
        tensors = {n.split("/")[-1].split(":")[0]: np.asarray(g[n]) for n in wn}
        if set(tensors) == {"kernel", "bias"}:
            layers.append({"W": tensors["kernel"], "b": tensors["bias"], "bn": None})
        elif {"gamma", "beta", "moving_mean", "moving_variance"} <= set(tensors):
            if not layers:
                raise ValueError("BatchNormalization %r precedes every Dense layer" % lname)
            layers[-1]["bn"] = (tensors["gamma"], tensors["beta"], tensors["moving_mean"], tensors["moving_variance"])
        else:
            raise ValueError("layer %r carries weights %s this engine has no kernel for" % (lname, sorted(tensors)))
    return validate(layers)


def load_keras_hdf5(path):
    try:
        import h5py
    except ImportError as ex:                       # loud, with the way out
        raise ImportError("%s is a Keras HDF5 file and h5py is not installed here: convert it once with "
                          "tools/keras_to_npz.py on a box that has h5py (no TensorFlow needed)" % path) from ex
    with h5py.File(path, "r") as f:
        return layers_from_keras_hdf5(f)


def load_saved_model(path):
    try:
        from tensorflow.keras.models import load_model
    except ImportError as ex:
        raise ImportError("%s is a Keras SavedModel and TensorFlow is not installed here: convert it once with "
                          "tools/keras_to_npz.py --saved-model on a box that has it" % path) from ex
    return layers_from_keras_model(load_model(path))


# ------------------------------------------------------------------------------------------ directory lookup
def find_net(model_dir, d):
    """Resolution order inside a model directory: engine-native npz, then the reference's artefacts."""
    cands = [(os.path.join(model_dir, d + "_weights.npz"), load_npz),
             (os.path.join(model_dir, d + "_weights-improvement.hdf5"), load_keras_hdf5),
             (os.path.join(model_dir, d + "_keras_model"), load_saved_model)]
    for path, loader in cands:
        if os.path.exists(path):
            return loader(path)
    raise FileNotFoundError("no weights for the %r net in %s (looked for %s)"
                            % (d, model_dir, ", ".join(os.path.basename(c[0]) for c in cands)))


def load_nets(model_dir):
    return {d: find_net(model_dir, d) for d in DIMS}
