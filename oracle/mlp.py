"""Oracle forward of the reference's FC CSI-prediction network.

Reference graph (massiveMIMO_CSI_prediction_DNN.py:173-234 with the pipeline
flags ``--nn 1024 1024 --useBN`` of full_pipeline_maMIMO_DNNEst.sh:47):

    z0 = [Flatten(x) || p]                         (:207-208, matlab_maMimo only)
    for each hidden size n:  Dense(n, relu)        (:211-214)
                             BatchNormalization()  (:215-218; Keras defaults eps=1e-3)
                             Dropout (identity at inference, :222-226)
    Dense(nSubCarr, linear)                        (:227)

Two independent nets ('real', 'imag') with separate weights (:173).  ReLU comes
BEFORE BatchNorm, so BN folds into the NEXT Dense layer (fold_bn below).

Keras Dense kernel layout is [in, out]; y = x @ W + b.

TensorFlow is unavailable here, so the FP64 numpy forward with
unfused BN is "truth"; forward(..., dtype=float32) stands in for TF's FP32
result.  Parity unpinned by execution of the reference.
"""
import numpy as np

BN_EPS = 1e-3   # keras.layers.BatchNormalization default epsilon


def forward(x, layers, dtype=np.float64):
    """x [rows, D_in]; layers = list of dict(W [in,out], b [out], bn=None|(gamma,beta,mean,var)).

    All layers but the last are Dense(relu) (+BN); the last is linear.
    """
    h = np.asarray(x, dtype=dtype)
    n = len(layers)
    for li, L in enumerate(layers):
        W = np.asarray(L["W"], dtype=dtype)
        b = np.asarray(L["b"], dtype=dtype)
        h = h @ W + b
        if li < n - 1:
            h = np.maximum(h, dtype(0))
            bn = L.get("bn")
            if bn is not None:
                g, beta, mu, var = (np.asarray(t, dtype=dtype) for t in bn)
                h = g * (h - mu) / np.sqrt(var + dtype(BN_EPS)) + beta
    return h


def fold_bn(layers):
    """Fold each hidden layer's BN into the following Dense (float64).

    W'_{l+1} = diag(s_l) W_{l+1},  b'_{l+1} = b_{l+1} + (beta_l - mu_l s_l) W_{l+1},
    s_l = gamma_l / sqrt(var_l + eps).  Returns layers without BN.
    """
    out = []
    scale = shift = None
    for L in layers:
        W = np.asarray(L["W"], dtype=np.float64)
        b = np.asarray(L["b"], dtype=np.float64)
        if scale is not None:
            b = b + shift @ W
            W = scale[:, None] * W
        out.append({"W": W, "b": b, "bn": None})
        bn = L.get("bn")
        if bn is not None:
            g, beta, mu, var = (np.asarray(t, dtype=np.float64) for t in bn)
            scale = g / np.sqrt(var + BN_EPS)
            shift = beta - mu * scale
        else:
            scale = shift = None
    return out


def predict_pair_nets(x_real, x_imag, nets, dtype=np.float64):
    """Both nets: returns (y_real, y_imag).  nets = {'real': layers, 'imag': layers}."""
    return forward(x_real, nets["real"], dtype), forward(x_imag, nets["imag"], dtype)
