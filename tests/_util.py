"""Shared helpers for the parity tests (oracle side)."""
import numpy as np

from oracle import ls, interp, mlp, postproc


def oracle_ls(Y, P, x_pilot, n_ps=1):
    """FP64 LS (+interp) on the given inputs: [Npkt,Nr,Nltf,Nsc] -> [Npkt,Nr,Nt,Nsc]."""
    Y = np.asarray(Y).astype(np.complex128)
    n_sc = Y.shape[-1]
    Hp = ls.ls_estimate(Y[..., ::n_ps], P, x_pilot)
    return interp.interp_linear(Hp, n_sc, n_ps)


def oracle_full(Y, P, x_pilot, n_ps, nets, dtype=np.float64):
    H = oracle_ls(Y, P, x_pilot, n_ps)
    X = H.reshape(-1, H.shape[-1])
    return H, mlp.forward(X.real, nets["real"], dtype), mlp.forward(X.imag, nets["imag"], dtype)


def rel_l2(ref, x):
    return postproc.rel_l2(ref, x)


def nmse_per_packet(ref_r, ref_i, r, i, n_pkt, n_rx, n_tx):
    """NMSE_subk (BER_test_maMIMO_LTF.m:675-686) averaged over packets, rows in pair_row order."""
    ref = (np.asarray(ref_r) + 1j * np.asarray(ref_i)).reshape(n_pkt, n_rx, n_tx, -1)
    got = (np.asarray(r) + 1j * np.asarray(i)).reshape(n_pkt, n_rx, n_tx, -1)
    vals = [postproc.nmse_subk(np.transpose(ref[p], (2, 1, 0)), np.transpose(got[p], (2, 1, 0))) for p in range(n_pkt)]
    return float(np.mean(vals))
