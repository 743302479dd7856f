"""Oracle for the per-subcarrier SVD of H-hat (restates the first lines of getWeightsForSubcarrier).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference: packet_generation/phased_arr/omphybweights.m:169-176, called once per subcarrier (:160-163) with
Hin = squeeze(Hchann_in(m,:,:)) = [Nt x Nr] from packet_generation/phased_arr/BER_test_maMIMO_LTF.m:372

    H = Hin.';                 % [Nr x Nt], plain transpose: H(i,j) = hD(k,j,i)           (:174)
    [~,~,v] = svd(H);          % v [Nt x Nt]                                               (:175)
    Fopt = v(:,1:Ns);          % call site passes Ns = NtRF = numSTS (BER_test :71 = sum(numSTSVec): 1 as shipped)  (:176)

H has rank <= Nr < Nt, so columns Nr+1..Nt of v are an arbitrary (LAPACK-build-defined) basis of the null space (only
Ns <= Nr is meaningful) and every column is only defined up to a phase.  The engine (mamimo_svd) therefore returns, and this oracle defines, the
basis-independent content: the Nr singular values and the Nr dominant right singular vectors, compared through the
projector V1 V1^H (and through sigma_r = ||H v_r||).

PINNED on the reference's own lines: tests/golden/make_golden.py runs :174-176 through tests/golden/mini_matlab.py
(svd supplied by numpy/LAPACK, which is what MATLAB calls too) -> tests/golden/ref_svd.npz; tests/test_svd_oracle.py
checks this module's projector and singular values against what those lines returned.
"""
import numpy as np


def svd_invariants(H):
    """H [n_pkt, n_rx, n_tx, n_sc] complex -> sigma [n_pkt, n_rx, n_sc] (descending), V1 [n_pkt, n_rx, n_tx, n_sc]
    with V1[p, r, :, k] = r-th right singular vector of the [n_rx x n_tx] matrix H[p, :, :, k]."""
    H = np.asarray(H).astype(np.complex128)
    Hm = np.transpose(H, (0, 3, 1, 2))                      # [pkt, k, rx, tx]
    _, S, Vh = np.linalg.svd(Hm, full_matrices=False)       # Vh [pkt, k, r, tx]: rows are v_r^H
    sigma = np.transpose(S, (0, 2, 1))
    V1 = np.transpose(np.conj(Vh), (0, 2, 3, 1))            # [pkt, r, tx, k]
    return sigma, V1


def projector(V1):
    """V1 [n_pkt, n_rx, n_tx, n_sc] -> P [n_pkt, n_sc, n_tx, n_tx] = sum_r v_r v_r^H (phase-invariant)."""
    V = np.transpose(np.asarray(V1).astype(np.complex128), (0, 3, 2, 1))     # [pkt, k, tx, r]
    return V @ np.conj(np.swapaxes(V, -1, -2))


def projector_from_fopt(Fopt, n_rx):
    """The same projector from the reference's Fopt = v(:,1:Ns) of ONE matrix (its first n_rx columns)."""
    F = np.asarray(Fopt)[:, :n_rx]
    return F @ F.conj().T
