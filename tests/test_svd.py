"""Per-subcarrier SVD of H-hat (pg/omphybweights.m:174-176): oracle pinned on the reference's own lines
(golden), CUDA path against the oracle through the basis-independent invariants -- singular values, the projector
V1 V1^H onto the row space, orthonormality and H V1 V1^H = H."""
import os

import numpy as np
import pytest

import mamimo_b200 as mm
from oracle import svd as osvd
from _util import rel_l2


def _to_engine_layout(Hin):
    """reference Hin [n, Nt, Nr] (one [Nt x Nr] matrix per subcarrier) -> engine H [1, Nr, Nt, n]"""
    return np.ascontiguousarray(np.transpose(Hin, (2, 1, 0))[None])


# ------------------------------------------------------------------------------------ oracle vs the reference lines (CPU)
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_oracle_matches_reference_svd_lines(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "ref_svd.npz"))
    Hin, Fopt = g["Hin_" + tag], g["Fopt_" + tag]
    n, nt, nr = Hin.shape
    sigma, V1 = osvd.svd_invariants(_to_engine_layout(Hin))
    P = osvd.projector(V1)[0]                                                  # [n, nt, nt]
    for k in range(n):
        H = Hin[k].T                                                            # the reference's H = Hin.'  (:174)
        s_ref = np.array([np.linalg.norm(H @ Fopt[k][:, r]) for r in range(nr)])   # sigma_r = ||H v_r||
        assert np.allclose(sigma[0, :, k], s_ref, rtol=1e-10, atol=1e-12 * s_ref[0])
        rank = int(np.sum(s_ref > 1e-9 * s_ref[0]))
        Pref = osvd.projector_from_fopt(Fopt[k], rank)
        Pk = osvd.projector(V1[:, :rank])[0][k] if rank < nr else P[k]
        assert np.linalg.norm(Pk - Pref) <= 1e-9 * np.sqrt(rank)


# ------------------------------------------------------------------------------------ CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("nt,nr,nsc,npkt,ctype", [(32, 4, 1024, 3, np.complex64), (32, 4, 234, 2, np.complex128),
                                                  (8, 2, 100, 4, np.complex64), (64, 8, 256, 2, np.complex64),
                                                  (4, 1, 52, 2, np.complex128), (6, 3, 77, 3, np.complex64)])
def test_svd_invariants_parity(nt, nr, nsc, npkt, ctype):
    if nt & (nt - 1):                 # the synthetic link needs a power-of-two preamble; any matrix will do here
        rng = np.random.default_rng(71)
        H = (rng.standard_normal((npkt, nr, nt, nsc)) + 1j * rng.standard_normal((npkt, nr, nt, nsc))).astype(ctype)
    else:
        _, H = mm.synth.make_packets(71, npkt, nt, nr, nsc, snr_db=10.0, dtype=ctype)
    H = H * np.linspace(1.0, 1e-3, npkt)[:, None, None, None].astype(H.real.dtype)       # packets 60 dB apart
    with mm.Engine(nt, nr, nsc, mlp=False) as eng:
        sigma, V1 = eng.svd(H)
        s_only = eng.svd(H, want_vectors=False)
    assert sigma.shape == (npkt, nr, nsc) and V1.shape == H.shape and V1.dtype == ctype
    assert np.array_equal(sigma, s_only)
    s_ref, V_ref = osvd.svd_invariants(H)
    tol = 2e-6 if ctype == np.complex64 else 1e-10
    for p in range(npkt):                                                               # per packet: relative to ITS scale
        assert rel_l2(s_ref[p], sigma[p]) <= tol
        assert np.all(np.diff(sigma[p], axis=0) <= 1e-6 * sigma[p, 0])                   # descending
    P, Pref = osvd.projector(V1), osvd.projector(V_ref)
    assert np.linalg.norm(P - Pref) / np.linalg.norm(Pref) <= (2e-5 if ctype == np.complex64 else 1e-8)
    # orthonormal columns and H V1 V1^H = H (rank <= n_rx): size-independent properties
    V = np.transpose(V1.astype(np.complex128), (0, 3, 2, 1))                             # [pkt, k, tx, r]
    G = np.conj(np.swapaxes(V, -1, -2)) @ V
    assert np.max(np.abs(G - np.eye(nr))) <= (5e-6 if ctype == np.complex64 else 1e-10)
    Hm = np.transpose(H.astype(np.complex128), (0, 3, 1, 2))                             # [pkt, k, rx, tx]
    for p in range(npkt):
        assert rel_l2(Hm[p], Hm[p] @ P[p]) <= (5e-6 if ctype == np.complex64 else 1e-9)


@pytest.mark.gpu
def test_svd_reference_lines_golden_on_gpu(golden_dir):
    """the golden matrices of the reference's own lines, incl. the rank-deficient one and the 1e-3-scaled one"""
    g = np.load(os.path.join(golden_dir, "ref_svd.npz"))
    for tag in ("a", "b", "c"):
        Hin, Fopt = g["Hin_" + tag], g["Fopt_" + tag]
        n, nt, nr = Hin.shape
        with mm.Engine(nt, nr, n, mlp=False) as eng:
            sigma, V1 = eng.svd(_to_engine_layout(Hin))
        for k in range(n):
            H = Hin[k].T
            s_ref = np.array([np.linalg.norm(H @ Fopt[k][:, r]) for r in range(nr)])
            # absolute floor: the Gram route resolves sigma_r only down to ~sqrt(eps) * sigma_1 (svd.cuh header); the
            # rank-deficient golden matrix has sigma_4 = 1e-15 and comes back as rounding noise of that size
            assert np.allclose(sigma[0, :, k], s_ref, rtol=1e-9, atol=1e-7 * s_ref[0]), (tag, k)
            rank = int(np.sum(s_ref > 1e-6 * s_ref[0]))
            Pk = osvd.projector(V1[:, :rank])[0][k]
            assert np.linalg.norm(Pk - osvd.projector_from_fopt(Fopt[k], rank)) <= 1e-7 * np.sqrt(rank), (tag, k)
            if rank < nr:                                     # vectors of vanishing singular values come back as zeros
                assert not V1[0, rank:, :, k].any()


@pytest.mark.gpu
def test_svd_device_buffers_and_full_batch_properties():
    """configs[1] batch size on device buffers: chunked launch == per-packet results, and the estimator's own output
    feeds it (LS -> SVD without leaving the device)."""
    import torch
    nt, nr, nsc, npkt = 32, 4, 1024, 500
    x = mm.synth.make_pilots(nsc)
    Yg, _ = mm.synth.make_packets(72, 10, nt, nr, nsc, snr_db=10.0, x_tones=x)
    Y = torch.from_numpy(np.concatenate([Yg] * 50)).cuda()
    with mm.Engine(nt, nr, nsc, mlp=False) as eng:
        eng.set_pilots(x, None)
        H = eng.ls_estimate(Y)
        sigma, V1 = eng.svd(H)
        torch.cuda.synchronize()
        assert torch.equal(sigma[:10], sigma[490:]) and torch.equal(V1[:10], V1[250:260])
        s1, v1 = eng.svd(H[137:138].contiguous())
        assert torch.equal(s1[0], sigma[137]) and torch.equal(v1[0], V1[137])
        s_ref, _ = osvd.svd_invariants(H[:2].cpu().numpy())
    assert rel_l2(s_ref, sigma[:2].cpu().numpy()) <= 2e-6
