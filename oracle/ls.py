"""Oracle LS channel estimate (restates helperMIMOChannelEstimate.m).

Reference: packet_generation/phased_arr/helperMIMOChannelEstimate.m:8-9,24-36

    Puse  = P(1:numSTS,1:numSTS)'            (conjugate transpose, :24)
    denom = nltf .* ltf(ind)                 (:27)
    hD(:,j,i) = rxsym*Puse(:,j) ./ denom     (:33-36), rxsym = rxData(:,1:nltf,i)

i.e.  H[k,j,i] = ( sum_n Y[k,n,i] * conj(P[j,n]) ) / ( nltf * ltf[ind k] ).

MATLAB arrays are column-major, so rxData [Nsc x nltf x Nr] has memory order
[rx][sym][k] and hD [Nsc x numSTS x Nr] has memory order [rx][tx][k]; the
batched forms below use exactly those C-order layouts with a leading packet
axis: Y[pkt, rx, sym, k] -> H[pkt, rx, tx, k].

PINNED: tests/golden/ref_matlab_ls_lmmse.npz holds the outputs of the reference's own
helperMIMOChannelEstimate.m source text, executed unmodified by tests/golden/mini_matlab.py
(MATLAB / Octave are not installed); tests/test_golden_matlab.py checks this module against
them to 1e-15.  Also anchored by the round-trip identity of SURVEY.md 8(c)(ii) in tests/test_oracle.py.
"""
import numpy as np


def ls_estimate_loop(rx_data, P, ltf_ind, num_sts=None):
    """Literal double-loop form, MATLAB shapes.

    rx_data : complex [Nsc, nltf, Nr]   (helperMIMOChannelEstimate.m:1,8)
    P       : [>=numSTS, >=numSTS] mapping matrix (helperGetP output, :13)
    ltf_ind : [Nsc] pilot tone values ltf(ind) (:27,29)
    returns hD complex128 [Nsc, numSTS, Nr]     (:31)
    """
    rx_data = np.asarray(rx_data, dtype=np.complex128)
    nsc, nltf, nrx = rx_data.shape
    if num_sts is None:
        num_sts = nltf
    P = np.asarray(P, dtype=np.complex128)
    puse = P[:num_sts, :num_sts].conj().T                 # :24
    denom = nltf * np.asarray(ltf_ind, dtype=np.float64)   # :27
    hD = np.zeros((nsc, num_sts, nrx), dtype=np.complex128)  # :31
    for i in range(nrx):                                   # :33
        rxsym = rx_data[:, :nltf, i]                       # :34
        for j in range(num_sts):                           # :35
            hD[:, j, i] = (rxsym @ puse[:, j]) / denom     # :36
    return hD


def ls_estimate(Y, P, x_pilot, dtype=np.complex128):
    """Vectorised batched form.

    Y       : complex [Npkt, Nr, nltf, Nsc]  (memory order of MATLAB rxData per packet)
    P       : [Nt, nltf]
    x_pilot : [Nsc] (ltf(ind); +/-1 in the reference, any non-zero complex here)
    returns H [Npkt, Nr, Nt, Nsc]
    """
    Y = np.asarray(Y, dtype=dtype)
    P = np.asarray(P, dtype=dtype)
    nltf = Y.shape[2]
    # sum_n Y[p,r,n,k] conj(P[j,n])
    H = np.einsum("prnk,jn->prjk", Y, P.conj(), optimize=True)
    return H / (nltf * np.asarray(x_pilot, dtype=dtype))[None, None, None, :]


def mat_to_batched(rx_data):
    """[Nsc, nltf, Nr] (MATLAB logical shape) -> [1, Nr, nltf, Nsc] (C order == MATLAB memory)."""
    return np.ascontiguousarray(np.transpose(rx_data, (2, 1, 0)))[None]


def batched_to_mat(H):
    """[Nr, Nt, Nsc] -> MATLAB logical [Nsc, Nt, Nr]."""
    return np.transpose(H, (2, 1, 0))


def ls_estimate_torch(Y, P, x_pilot):
    """Same arithmetic as ls_estimate in torch CPU complex128 (one BLAS thread pool shared with the
    torch FC restatement; used by bench.py's CPU-baseline leg).  Y may be a numpy array or a tensor."""
    import torch
    Yt = torch.as_tensor(Y).to(torch.complex128)
    Pc = torch.as_tensor(np.asarray(P, dtype=np.complex128).conj().copy())
    nltf = Yt.shape[2]
    den = torch.as_tensor(nltf * np.asarray(x_pilot, dtype=np.complex128))
    return torch.matmul(Pc, Yt) / den        # [Nt,nltf] @ [p,r,nltf,k] -> [p,r,Nt,k]
