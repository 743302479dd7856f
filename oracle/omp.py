"""Oracle for the transmit side of the hybrid-precoder consumer (orthogonal matching pursuit over a steering dictionary).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference: packet_generation/phased_arr/BER_test_maMIMO_LTF.m:364-376 draws one dictionary At [Nt x nRays] per packet,
repeats it for every subcarrier and calls `[Fbb,Frf] = omphybweights(hDp{1},numSTS,numSTS,AtExp)`; per subcarrier that
is getWeightsForSubcarrier (packet_generation/phased_arr/omphybweights.m:171-178, precoding-only outputs :195-198):

    Fopt = first Ns right singular vectors of H = Hin.'                         (:174-176, oracle/svd.py)
    [Fbb, Frf] = ompdecomp(Fopt, At, 'MaxSparsity', NtRF)                       (:178)
    Fbb = sqrt(Ns) * Fbb / ||Frf*Fbb||_F                                        (:179)
    returned: Fbb.' [Ns x NtRF], Frf.' [NtRF x Nt]                              (:195-197)

and ompdecomp's loop with the default identity weight (packet_generation/phased_arr/ompdecomp.m:101-121):
the residual starts as Fopt; each round picks the dictionary column with the largest correlation energy
sum_s |a_k^H r_s|^2 (first index on ties, MATLAB `max`), refits ALL chosen columns to Fopt by least squares, and
renormalises the residual by its Frobenius norm; it stops after NtRF columns or when that norm falls to eps.

What is basis-independent (Fopt's columns are defined up to a unitary mix inside the dominant subspace when singular
values coincide, and up to a phase each otherwise): the chosen column INDICES (integers), the residual norm, and
Frf*Fbb*Fbb^H*Frf^H.  The coefficient matrix itself inherits Fopt's phases: compare it through those products, or
feed both sides the same Fopt.

PINNED on the reference's own loop text executed by tests/golden/mini_matlab.py: tests/golden/ref_omp.npz
(tests/test_omp_oracle.py).
"""
import numpy as np

EPS = np.finfo(np.float64).eps


def ompdecomp(Wopt, Adict, max_sparsity=1):
    """Wopt [N x Nw], Adict [N x Nd] complex -> (coeff [ns x Nw], atoms [N x ns], idx [ns] 0-based, errnorm)."""
    Wopt = np.asarray(Wopt, dtype=np.complex128)
    Adict = np.asarray(Adict, dtype=np.complex128)
    res = Wopt.copy()
    err = 1.0
    idx = []
    coeff = np.zeros((0, Wopt.shape[1]), dtype=np.complex128)
    while len(idx) < max_sparsity and err > EPS:
        psi = Adict.conj().T @ res
        energy = np.sum(np.abs(psi) ** 2, axis=1)
        idx.append(int(np.argmax(energy)))                      # first maximum, like MATLAB's max
        A = Adict[:, idx]
        coeff = np.linalg.solve(A.conj().T @ A, A.conj().T @ Wopt)
        diff = Wopt - A @ coeff
        err = float(np.sqrt(np.sum(np.abs(diff) ** 2)))
        res = diff / err if err > 0 else diff
    return coeff, Adict[:, idx], np.asarray(idx, dtype=np.int64), err


def precoder_for_subcarrier(Fopt, At, n_rf):
    """getWeightsForSubcarrier after the SVD, precoding only.  Fopt [Nt x Ns], At [Nt x nRays].
    Returns (Fbb_out [Ns x ns], Frf_out [ns x Nt], idx [ns] 0-based, errnorm): the transposed forms the reference
    returns."""
    ns = Fopt.shape[1]
    coeff, atoms, idx, err = ompdecomp(Fopt, At, n_rf)
    coeff = np.sqrt(ns) * coeff / np.sqrt(np.sum(np.abs(atoms @ coeff) ** 2))
    return coeff.T, atoms.T, idx, err


def omp_precoder(H, At, ns, n_rf):
    """Batched form over the engine's layout.  H [n_pkt, n_rx, n_tx, n_sc] complex, At [n_tx, n_rays] ->
    idx int32 [n_pkt, n_rf, n_sc] (-1 where the loop stopped early), Fbb complex128 [n_pkt, ns, n_rf, n_sc]
    (Fbb[p, s, j, k] = the reference's Fbb(k, s, j)), errnorm [n_pkt, n_sc], Fopt [n_pkt, ns, n_tx, n_sc]."""
    from . import svd as _svd
    H = np.asarray(H)
    n_pkt, n_rx, n_tx, n_sc = H.shape
    _, V1 = _svd.svd_invariants(H)
    Fopt = V1[:, :ns]
    idx = np.full((n_pkt, n_rf, n_sc), -1, dtype=np.int32)
    Fbb = np.zeros((n_pkt, ns, n_rf, n_sc), dtype=np.complex128)
    err = np.zeros((n_pkt, n_sc))
    for p in range(n_pkt):
        for k in range(n_sc):
            fb, _, ix, e = precoder_for_subcarrier(Fopt[p, :, :, k].T, At, n_rf)
            idx[p, :len(ix), k] = ix
            Fbb[p, :, :len(ix), k] = fb
            err[p, k] = e
    return idx, Fbb, err, Fopt


def precoder_invariant(Fbb, idx, At):
    """Frf*Fbb*Fbb^H*Frf^H [n_pkt, n_sc, n_tx, n_tx] from omp_precoder's outputs (independent of Fopt's phases)."""
    At = np.asarray(At, dtype=np.complex128)
    n_pkt, ns, n_rf, n_sc = Fbb.shape
    out = np.zeros((n_pkt, n_sc, At.shape[0], At.shape[0]), dtype=np.complex128)
    for p in range(n_pkt):
        for k in range(n_sc):
            sel = idx[p, :, k]
            m = int(np.sum(sel >= 0))
            M = At[:, sel[:m]] @ Fbb[p, :, :m, k].T           # Frf [Nt x m] * Fbb [m x Ns]
            out[p, k] = M @ M.conj().T
    return out
