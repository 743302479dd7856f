"""Synthetic workload of SURVEY.md 8(d): no dataset or trained weights ship with the reference.

Static tables come from default_rng(67) (67 is the reference's own fixed seed,
pg/generate_maMIMO_LTF.m:43); packet p of config c from default_rng([67, c, p]).
This is the workload generator shared by tests and bench.py -- it is not a checker and
contains no estimator math.
"""
import numpy as np

SEED = 67
N_TAPS = 16


def sylvester(n):
    h = np.ones((1, 1))
    while h.shape[0] < n:
        h = np.block([[h, h], [h, -h]])
    return h


def make_pilots(n_sc, n_ps=1):
    """X_pilot in {+1,-1}^n_pil."""
    rng = np.random.default_rng([SEED, 1, n_sc, n_ps])
    n_pil = (n_sc + n_ps - 1) // n_ps
    return rng.choice([-1.0, 1.0], size=n_pil)


def make_nets(d_in, hidden, d_out, use_bn=True):
    """Glorot-uniform kernels as in ..._DNN.py:213,227; non-trivial BN so folding is exercised."""
    nets = {}
    for ni, name in enumerate(("real", "imag")):
        rng = np.random.default_rng([SEED, 2, ni, d_in, d_out] + list(hidden))
        dims = [d_in] + list(hidden) + [d_out]
        layers = []
        for i in range(len(dims) - 1):
            lim = np.sqrt(6.0 / (dims[i] + dims[i + 1]))
            L = {"W": rng.uniform(-lim, lim, (dims[i], dims[i + 1])).astype(np.float32),
                 "b": rng.uniform(-0.1, 0.1, dims[i + 1]).astype(np.float32), "bn": None}
            if use_bn and i < len(dims) - 2:
                n = dims[i + 1]
                L["bn"] = (rng.uniform(0.5, 1.5, n).astype(np.float32), rng.uniform(-0.1, 0.1, n).astype(np.float32),
                           rng.uniform(-0.1, 0.1, n).astype(np.float32), rng.uniform(0.5, 1.5, n).astype(np.float32))
            layers.append(L)
        nets[name] = layers
    return nets


def make_channel(rng, n_rx, n_tx, n_sc):
    """H [n_rx, n_tx, n_sc]: FFT of a 16-tap complex-Gaussian exponential-PDP impulse response, unit mean power."""
    pdp = np.exp(-np.arange(N_TAPS) / 4.0)
    pdp /= pdp.sum()
    taps = (rng.standard_normal((n_rx, n_tx, N_TAPS)) + 1j * rng.standard_normal((n_rx, n_tx, N_TAPS))) * np.sqrt(pdp / 2)
    return np.fft.fft(taps, n=n_sc, axis=-1)


def make_packets(cfg_id, n_pkt, n_tx, n_rx, n_sc, snr_db, P=None, x_tones=None, first_pkt=0, dtype=np.complex64):
    """Y [n_pkt, n_rx, n_ltf, n_sc] and the true channel H [n_pkt, n_rx, n_tx, n_sc].

    Y[k,n,i] = x[k] * sum_j H[k,j,i] P[j,n] + w,  w ~ CN(0, mean|signal|^2 * 10^(-SNR/10))
    (noise rule mirrors pg/generate_maMIMO_LTF.m:241-245).  x_tones is the full-grid tone sequence [n_sc].
    snr_db may be a scalar or an array [n_pkt].
    """
    P = sylvester(n_tx) if P is None else np.asarray(P)
    x = np.ones(n_sc) if x_tones is None else np.asarray(x_tones)
    snr = np.broadcast_to(np.asarray(snr_db, dtype=np.float64), (n_pkt,))
    Y = np.empty((n_pkt, n_rx, P.shape[1], n_sc), dtype=dtype)
    Ht = np.empty((n_pkt, n_rx, n_tx, n_sc), dtype=dtype)
    for p in range(n_pkt):
        rng = np.random.default_rng([SEED, cfg_id, first_pkt + p])
        H = make_channel(rng, n_rx, n_tx, n_sc)
        sig = np.einsum("rjk,jn->rnk", H, P) * x[None, None, :]
        sigma2 = np.mean(np.abs(sig) ** 2) * 10.0 ** (-snr[p] / 10.0)
        w = (rng.standard_normal(sig.shape) + 1j * rng.standard_normal(sig.shape)) * np.sqrt(sigma2 / 2)
        Y[p] = sig + w
        Ht[p] = H
    return Y, Ht
