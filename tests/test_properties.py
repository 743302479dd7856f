"""Property tests (hypothesis): size-independent identities of the path, on the oracle (CPU) and on the CUDA path (GPU)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

import mamimo_b200 as mm
from oracle import tables, ls, interp, mlp, ofdm

POW2 = st.sampled_from([1, 2, 4, 8, 16])


@settings(max_examples=25, deadline=None)
@given(nt=POW2, nr=st.integers(1, 4), nsc=st.integers(1, 96), seed=st.integers(0, 2 ** 16))
def test_oracle_ls_round_trip_any_shape(nt, nr, nsc, seed):
    rng = np.random.default_rng(seed)
    P = tables.sylvester_hadamard(nt)
    x = rng.choice([-1.0, 1.0], nsc)
    H = rng.standard_normal((2, nr, nt, nsc)) + 1j * rng.standard_normal((2, nr, nt, nsc))
    Y = np.einsum("prjk,jn->prnk", H, P) * x
    assert np.max(np.abs(ls.ls_estimate(Y, P, x) - H)) < 1e-12


@settings(max_examples=25, deadline=None)
@given(nsc=st.integers(2, 200), nps=st.integers(1, 12), seed=st.integers(0, 2 ** 16))
def test_oracle_interp_reproduces_affine_and_pilots(nsc, nps, seed):
    rng = np.random.default_rng(seed)
    k = np.arange(nsc)
    a, b = rng.standard_normal(2) + 1j * rng.standard_normal(2)
    lin = a + b * k
    out = interp.interp_linear(lin[::nps][None], nsc, nps)[0]
    npil = len(lin[::nps])
    if npil >= 2:
        assert np.allclose(out, lin, atol=1e-9 * (1 + abs(b) * nsc))       # affine in, affine out (incl. extrapolation)
    assert np.allclose(out[::nps], lin[::nps])                              # pilots are kept


@settings(max_examples=15, deadline=None)
@given(log2n=st.integers(2, 8), cp_frac=st.sampled_from([0, 4, 8]), seed=st.integers(0, 2 ** 16))
def test_oracle_ofdm_round_trip(log2n, cp_frac, seed):
    n = 1 << log2n
    cp = n // cp_frac if cp_frac else 0
    rng = np.random.default_rng(seed)
    car = np.sort(rng.choice(np.arange(1, n + 1), size=max(1, n // 2), replace=False))
    G = rng.standard_normal((1, 2, 3, car.size)) + 1j * rng.standard_normal((1, 2, 3, car.size))
    x = ofdm.ofdm_mod(G, n, cp, car)
    for off in {cp, cp // 2}:
        got = ofdm.ofdm_demod(x, n, cp, off, car)
        if off == cp:
            assert np.max(np.abs(got - G)) < 1e-10
        else:   # an early window is a pure per-carrier phase ramp: magnitudes are preserved
            assert np.allclose(np.abs(got), np.abs(G), atol=1e-10)


@settings(max_examples=10, deadline=None)
@given(seed=st.integers(0, 2 ** 16), d=st.integers(3, 40), h=st.integers(2, 30))
def test_oracle_bn_fold_equivalence(seed, d, h):
    rng = np.random.default_rng(seed)
    dims = [d, h, h + 1, d]
    layers = []
    for i in range(3):
        L = {"W": rng.standard_normal((dims[i], dims[i + 1])), "b": rng.standard_normal(dims[i + 1]), "bn": None}
        if i < 2:
            n = dims[i + 1]
            L["bn"] = (rng.uniform(0.5, 1.5, n), rng.standard_normal(n), rng.standard_normal(n), rng.uniform(0.2, 2, n))
        layers.append(L)
    x = rng.standard_normal((4, d))
    assert np.allclose(mlp.forward(x, layers), mlp.forward(x, mlp.fold_bn(layers)), rtol=1e-10, atol=1e-10)


# ------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(nt=st.sampled_from([1, 2, 4, 8, 16, 32]), nr=st.integers(1, 3), nsc=st.integers(1, 300), npkt=st.integers(1, 4),
       nps=st.sampled_from([1, 1, 2, 5]), seed=st.integers(0, 2 ** 16))
def test_gpu_ls_matches_oracle_any_shape(nt, nr, nsc, npkt, nps, seed):
    """Ragged sizes (odd Nsc, Nsc < tile, Nsc not a multiple of Nps) against the FP64 oracle; plus homogeneity."""
    from _util import oracle_ls, rel_l2
    rng = np.random.default_rng(seed)
    npil = (nsc + nps - 1) // nps
    xp = rng.choice([-1.0, 1.0], npil)
    Y = (rng.standard_normal((npkt, nr, nt, nsc)) + 1j * rng.standard_normal((npkt, nr, nt, nsc))).astype(np.complex64)
    with mm.Engine(nt, nr, nsc, n_ps=nps, mlp=False) as eng:
        eng.set_pilots(xp, None)
        H = eng.ls_estimate(Y)
        H3 = eng.ls_estimate((Y * np.complex64(3.0)).astype(np.complex64))
    assert rel_l2(oracle_ls(Y, tables.sylvester_hadamard(nt), xp, nps), H) <= 1e-6
    assert rel_l2(3.0 * H.astype(np.complex128), H3) <= 1e-6


@pytest.mark.gpu
@settings(max_examples=8, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(rows=st.integers(1, 600), d_in=st.integers(1, 300), h=st.integers(1, 300), d_out=st.integers(1, 300),
       prec=st.sampled_from(["fp16x3", "tf32x3"]), seed=st.integers(0, 2 ** 16))
def test_gpu_fc_matches_oracle_any_shape(rows, d_in, h, d_out, prec, seed):
    """Ragged M / K / N (none a multiple of the 256 x 256 x 64 tile) through the tcgen05 kernels."""
    from _util import rel_l2
    nets = mm.synth.make_nets(d_in, (h,), d_out)
    rng = np.random.default_rng(seed)
    Xr = rng.standard_normal((rows, d_in)).astype(np.float32)
    Xi = rng.standard_normal((rows, d_in)).astype(np.float32)
    with mm.Engine(1, 1, 1, n_ltf=1, hidden=(h,), d_in=d_in, d_out=d_out, input_mode="planes", precision=prec) as eng:
        eng.load_weights(nets)
        Yr, Yi = eng.predict_planes(Xr, Xi)
    ref_r, ref_i = mlp.forward(Xr, nets["real"]), mlp.forward(Xi, nets["imag"])
    scale = max(np.linalg.norm(ref_r), np.linalg.norm(ref_i), 1e-30)
    assert np.linalg.norm(Yr - ref_r) / scale <= 1e-5 and np.linalg.norm(Yi - ref_i) / scale <= 1e-5
