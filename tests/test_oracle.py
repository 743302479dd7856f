"""Oracle self-checks: the known-answer material of SURVEY.md 8(c) (no GPU)."""
import numpy as np
import pytest

from oracle import tables, ls, interp, mlp, postproc


def test_ltf_table_known_answers():
    ltf = tables.vht_ltf256()
    ind = tables.carriers_locations()
    assert ltf.shape == (256,) and set(np.unique(ltf)) == {-1, 0, 1}
    zeros = np.nonzero(ltf == 0)[0] + 1
    assert zeros.tolist() == list(range(1, 8)) + [129] + list(range(251, 257))
    assert ind.size == 234 and np.all(np.diff(ind) > 0)
    assert ind[:8].tolist() == list(range(8, 16)) and ind[-1] == 250
    assert np.all(ltf[ind - 1] != 0) and int(ltf[ind - 1].sum()) == 46
    assert ltf[tables.pilot_carrier_indices() - 1].tolist() == [1, -1, 1, -1, -1, -1, -1, -1]


@pytest.mark.parametrize("nt,nr,nsc", [(4, 2, 16), (32, 4, 234), (8, 3, 52)])
def test_ls_round_trip_identity(nt, nr, nsc):
    """Y[k,n,i] = ltf[k] sum_j H[k,j,i] P[j,n]  =>  LS returns H exactly (FP64, +/-1 P)."""
    rng = np.random.default_rng(67)
    P = tables.sylvester_hadamard(nt)
    x = tables.ltf_at_carriers()[:nsc] if nsc <= 234 else rng.choice([-1.0, 1.0], nsc)
    H = rng.standard_normal((nsc, nt, nr)) + 1j * rng.standard_normal((nsc, nt, nr))
    Y = np.einsum("kji,jn->kni", H, P) * x[:, None, None]
    hD = ls.ls_estimate_loop(Y, P, x)
    assert np.max(np.abs(hD - H)) < 1e-13
    Hb = ls.ls_estimate(ls.mat_to_batched(Y), P, x)
    assert np.max(np.abs(ls.batched_to_mat(Hb[0]) - hD)) < 1e-13


def test_ls_data_phase_is_plain_divide():
    rng = np.random.default_rng(1)
    Y = rng.standard_normal((234, 1, 4)) + 1j * rng.standard_normal((234, 1, 4))
    x = tables.ltf_at_carriers()
    hD = ls.ls_estimate_loop(Y, np.ones((1, 1)), x)
    assert np.allclose(hD[:, 0, :], Y[:, 0, :] / x[:, None], rtol=0, atol=1e-15)


def test_ls_complex_p_conjugation():
    rng = np.random.default_rng(2)
    n = 4
    F = np.fft.fft(np.eye(n))                         # complex orthogonal: F F^H = n I
    H = rng.standard_normal((8, n, 2)) + 1j * rng.standard_normal((8, n, 2))
    x = rng.choice([-1.0, 1.0], 8)
    Y = np.einsum("kji,jn->kni", H, F) * x[:, None, None]
    assert np.max(np.abs(ls.ls_estimate_loop(Y, F, x) - H)) < 1e-12


def test_interp_identity_and_linear():
    rng = np.random.default_rng(3)
    H = rng.standard_normal((3, 40)) + 1j * rng.standard_normal((3, 40))
    assert np.array_equal(interp.interp_linear(H, 40, 1), H)
    k = np.arange(37)
    lin = (2.0 - 0.25 * k) + 1j * (0.5 + 0.1 * k)      # exactly linear: interp + extrapolation reproduce it
    out = interp.interp_linear(lin[::4][None], 37, 4)
    assert np.allclose(out[0], lin, atol=1e-12)
    assert np.allclose(interp.interp_linear(np.array([[1 + 2j]]), 3, 4), (1 + 2j) * np.ones((1, 3)))


def test_bn_folding_matches_unfused():
    rng = np.random.default_rng(4)
    dims = [12, 16, 9, 7]
    layers = []
    for i in range(3):
        L = {"W": rng.standard_normal((dims[i], dims[i + 1])), "b": rng.standard_normal(dims[i + 1]), "bn": None}
        if i < 2:
            n = dims[i + 1]
            L["bn"] = (rng.uniform(0.5, 1.5, n), rng.standard_normal(n), rng.standard_normal(n), rng.uniform(0.5, 1.5, n))
        layers.append(L)
    x = rng.standard_normal((5, 12))
    assert np.allclose(mlp.forward(x, layers), mlp.forward(x, mlp.fold_bn(layers)), rtol=1e-12, atol=1e-12)


def test_pair_row_and_inverse():
    n_rx, n_tx, nsc = 3, 5, 4
    rows = np.arange(2 * n_rx * n_tx * nsc, dtype=np.float64).reshape(2 * n_rx * n_tx, nsc)
    seen = set()
    for p in range(2):
        csi = postproc.rows_to_csi(rows[p * n_rx * n_tx:(p + 1) * n_rx * n_tx], n_tx, n_rx)
        for irx in range(n_rx):
            for itx in range(n_tx):
                r = postproc.pair_row(p, irx, itx, n_rx, n_tx)
                seen.add(r)
                assert np.array_equal(csi[:, itx, irx], rows[r])
    assert seen == set(range(2 * n_rx * n_tx))


def test_nmse_known_answers():
    rng = np.random.default_rng(5)
    H = rng.standard_normal((10, 4, 2)) + 1j * rng.standard_normal((10, 4, 2))
    assert postproc.nmse_subk(H, H) == 0.0
    assert abs(postproc.nmse_subk(H, np.zeros_like(H)) - 1.0) < 1e-15


def test_renew_layout_round_trip():
    """64-FFT non-zero bins 6..31, 33..58 after fftshift <-> pad [6,26,1,26,5] + ifftshift."""
    rng = np.random.default_rng(6)
    v = np.zeros((2, 64), dtype=np.complex128)
    sup = list(range(6, 32)) + list(range(33, 59))
    shifted = np.zeros_like(v)
    shifted[:, sup] = rng.standard_normal((2, 52)) + 1j * rng.standard_normal((2, 52))
    v = np.fft.ifftshift(shifted, axes=1)
    sel = np.fft.fftshift(v, axes=1)[:, sup]
    assert np.array_equal(postproc.renew_postprocess(sel), v)
    with pytest.raises(ValueError):
        postproc.renew_postprocess(np.zeros((1, 51)))


def test_mlp_oracle_matches_independent_torch_layers():
    """TensorFlow cannot run here, so the Keras layer arithmetic of oracle.mlp (Dense: x @ W + b, relu, then
    BatchNormalization at inference: gamma * (x - mean) / sqrt(var + 1e-3) + beta, ..._DNN.py:211-227) is cross-checked
    against an independent library implementing the same documented layers: torch.nn.Linear / ReLU /
    BatchNorm1d(eps=1e-3).eval() in float64."""
    import torch
    import mamimo_b200 as mm
    nets = mm.synth.make_nets(40, (64, 48), 30)
    rng = np.random.default_rng(11)
    x = rng.standard_normal((17, 40))
    for name in ("real", "imag"):
        mods = []
        for li, L in enumerate(nets[name]):
            lin = torch.nn.Linear(L["W"].shape[0], L["W"].shape[1]).double()
            lin.weight.data = torch.from_numpy(np.asarray(L["W"], np.float64).T.copy())     # Keras kernel is [in, out]
            lin.bias.data = torch.from_numpy(np.asarray(L["b"], np.float64).copy())
            mods.append(lin)
            if li < len(nets[name]) - 1:
                mods.append(torch.nn.ReLU())
                g, be, mu, var = (torch.from_numpy(np.asarray(t, np.float64).copy()) for t in L["bn"])
                bn = torch.nn.BatchNorm1d(g.numel(), eps=1e-3).double()
                bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var = g, be, mu, var
                mods.append(bn)
        net = torch.nn.Sequential(*mods).eval()
        with torch.no_grad():
            ref = net(torch.from_numpy(x)).numpy()
        got = mlp.forward(x, nets[name])
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-13
        assert np.linalg.norm(mlp.forward(x, mlp.fold_bn(nets[name])) - ref) / np.linalg.norm(ref) < 1e-12
