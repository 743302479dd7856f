"""Golden vectors produced by running the reference's own MATLAB source text (helperMIMOChannelEstimate.m,
LMMSE_ce.m) through tests/golden/mini_matlab.py in the build container (tests/golden/make_golden.py section 4;
MATLAB / Octave are not installed).  CPU part: the numpy oracle reproduces them (this PINS oracle.ls and
oracle.lmmse); GPU part: the CUDA path reproduces them through the C ABI."""
import os
import sys

import numpy as np
import pytest

import mamimo_b200 as mm
from oracle import ls, lmmse, tables
from _util import rel_l2

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_matlab_ls_lmmse.npz"))


# ------------------------------------------------------------------------------------------------ interpreter sanity
def test_mini_matlab_semantics():
    from mini_matlab import MatlabFile
    src = """
function [a,b,c,d,e] = probe(x, s)
v = [1; 2;-3; 4];            % column literal, '-3' is an element
r = [0:3].';                 % range in brackets, plain transpose
a = v(2:end) .* r(1:3) - 1./[2 4 8].';
M = repmat([1 2 3], 2, 1);   b = M(:, [1 3])' * [1; 1i];
c = zeros(2,3,2);  for k = 1:2, c(:,k,2) = [k; 10*k]; end
q = x*x';  d = sum(x.*conj(x))/q + 10^(s*0.1) + numel(M) + length(v) + size(M,2);
[~, n2] = size(M);  if n2 == 3 && ~(n2 < 1), e = x(end)^2; else, e = 0; end
end
"""
    m = MatlabFile(src)
    a, b, c, d, e = m.call("probe", [np.array([[1 + 2j, 3.0]]), 20.0], 5)
    assert np.allclose(a.ravel(), [2 * 0 - 0.5, -3 * 1 - 0.25, 4 * 2 - 0.125])
    assert np.allclose(b, np.array([[1 + 1j], [3 + 3j]]))          # M(:,[1 3])' is 2x2 [[1,1],[3,3]]
    assert c.shape == (2, 3, 2) and np.allclose(c[:, :, 1], [[1, 2, 0], [10, 20, 0]]) and not c[:, :, 0].any()
    assert np.allclose(d, 1.0 + 100.0 + 6 + 4 + 3)
    assert np.allclose(e, 9.0)


def test_mini_matlab_array_semantics_against_numpy():
    """column-major reshape / linear indexing, N-d slicing with end and colon, implicit expansion, matrix vs
    element-wise operators, right division, conjugate vs plain transpose, struct fields, nested functions"""
    from mini_matlab import MatlabFile
    src = """
function [a,b,c,d,e,f,g,h] = probe(X, prm)
A = reshape(1:24, [2 3 4]);
a = A(2, end, 3) + A(end) + numel(A(:, 2:end, [1 4]));
B = squeeze(A(2, :, :));            b = B(:, 2).' * B(:, 3);
c = X .* [1; 2] + [10 20 30];       % implicit expansion 2x3 .* 2x1 + 1x3
d = X * X' / (X * X.' + eye(2));    % mrdivide
e = helper(prm.gain, X(:));         f = prm.idx(end:-1:1);
M = zeros(3);  M(2, :) = 1:3;  M(:, 3) = M(:, 3) + [7; 8; 9];  M(end, end) = -M(2, 3);
g = M;  h = sum(sum(M > 0)) + any_neg(M);
end
function y = helper(gain, v)
y = gain * (v' * v) ^ 0.5 + length(v);
end
function t = any_neg(M)
t = 0;
for k = 1:numel(M), if M(k) < 0, t = t + 1; end, end
end
"""
    rng = np.random.default_rng(3)
    X = rng.standard_normal((2, 3)) + 1j * rng.standard_normal((2, 3))
    prm = {"gain": 2.5, "idx": np.array([[4.0, 5.0, 6.0]])}
    a, b, c, d, e, f, g, h = MatlabFile(src).call("probe", [X, prm], 8)
    A = np.arange(1, 25).reshape((2, 3, 4), order="F")
    assert a.item() == A[1, 2, 2] + A.ravel(order="F")[-1] + A[:, 1:, [0, 3]].size
    B = A[1, :, :]
    assert b.item() == B[:, 1] @ B[:, 2]
    assert np.allclose(c, X * np.array([[1], [2]]) + np.array([[10, 20, 30]]))
    assert np.allclose(d, (X @ X.conj().T) @ np.linalg.inv(X @ X.T + np.eye(2)))
    v = X.ravel(order="F")
    assert np.allclose(e, 2.5 * np.sqrt(np.vdot(v, v)) + 6)
    assert np.array_equal(f, [[6.0, 5.0, 4.0]])
    M = np.zeros((3, 3)); M[1, :] = [1, 2, 3]; M[:, 2] += [7, 8, 9]; M[2, 2] = -M[1, 2]
    assert np.array_equal(g, M) and h.item() == (M > 0).sum() + 1


# ------------------------------------------------------------------------------------------------ oracle pinned
@pytest.mark.parametrize("tag", ["A", "B", "C"])
def test_oracle_ls_reproduces_interpreted_matlab(g, tag):
    rx, P, hD = g["rx_" + tag], g["P_" + tag], g["hD_" + tag]
    assert np.array_equal(g["ltf_o_" + tag].ravel(), tables.ltf_at_carriers())        # ltf(ind), :29
    assert np.array_equal(g["carriers"], tables.carriers_locations())
    ours = ls.ls_estimate_loop(rx, P, tables.ltf_at_carriers())
    assert rel_l2(hD, ours) <= 1e-15
    batched = ls.ls_estimate(ls.mat_to_batched(rx), P, tables.ltf_at_carriers())[0]
    assert rel_l2(hD, ls.batched_to_mat(batched)) <= 1e-15
    if tag != "A":
        assert not g["hDmmse_" + tag].any()                                           # isMMSE = false: zeros (:32)


def test_oracle_lmmse_reproduces_interpreted_matlab(g):
    hD, tau, snr = g["hD_A"], g["tau_A"].ravel(), g["snr_A"].ravel()
    assert rel_l2(g["hDmmse_A"], lmmse.helper_mmse_loop(hD, 1, tau, snr)) <= 1e-12
    bat = lmmse.lmmse_batched(np.transpose(hD, (2, 1, 0))[None], lmmse.tau_rms(tau), snr[None], 1)[0]
    assert rel_l2(g["hDmmse_A"], np.transpose(bat, (2, 1, 0))) <= 1e-12
    for tag in ("nps2", "sec"):
        n, nps, s = g["ce_par_" + tag]
        y = lmmse.lmmse_ce(g["ce_x_" + tag].ravel(), int(n), int(n), int(nps), g["ce_h_" + tag].ravel(), float(s))
        assert rel_l2(g["ce_y_" + tag].ravel(), y) <= 1e-12
        assert mm.tau_rms(g["ce_h_" + tag].ravel()) == pytest.approx(lmmse.tau_rms(g["ce_h_" + tag].ravel()), rel=1e-13)


def test_oracle_nmse_and_pair_rebuild_reproduce_interpreted_matlab(g):
    """NMSE_subk (BER_test_maMIMO_LTF.m:675-686) and the loop that rebuilds CSI(:,iTX,iRX) from prediction rows
    (:213-218), both executed from the reference's text."""
    from oracle import postproc
    assert postproc.nmse_subk(g["nmse_ref"], g["nmse_est"]) == pytest.approx(g["nmse_val"].item(), rel=1e-13)
    assert g["nmse_zero"].item() == 0.0 and g["nmse_one"].item() == pytest.approx(1.0, rel=1e-15)
    pr, pi = g["rebuild_pred_real"], g["rebuild_pred_imag"]
    n_tx, n_rx = g["rebuild_csi_real"].shape[1:]
    csi = postproc.rows_to_csi(pr + 1j * pi, n_tx, n_rx)
    assert np.array_equal(csi.real, g["rebuild_csi_real"]) and np.array_equal(csi.imag, g["rebuild_csi_imag"])
    for i_rx in range(n_rx):                                   # the C ABI's row formula is the inverse of that loop
        for i_tx in range(n_tx):
            row = mm.pair_row(0, i_rx, i_tx, n_rx, n_tx)
            assert np.array_equal(g["rebuild_csi_real"][:, i_tx, i_rx], pr[row])


# ------------------------------------------------------------------------------------------------ CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["A", "B", "C"])
def test_cuda_ls_reproduces_interpreted_matlab(g, tag):
    rx, P = g["rx_" + tag], g["P_" + tag]
    prm = {"numSTS": P.shape[0], "CarriersLocations": g["carriers"]}
    hD, Pm, ltf_o, hM = mm.helperMIMOChannelEstimate(rx, prm, 1, None, None, False, P=P)
    assert rel_l2(g["hD_" + tag], hD) <= 1e-12                     # complex double in/out: FP64 on the device, like MATLAB
    assert np.array_equal(ltf_o, g["ltf_o_" + tag]) and not hM.any()


@pytest.mark.gpu
def test_cuda_lmmse_reproduces_interpreted_matlab(g):
    rx, P = g["rx_A"], g["P_A"]
    prm = {"numSTS": P.shape[0], "CarriersLocations": g["carriers"]}
    hD, _, _, hM = mm.helperMIMOChannelEstimate(rx, prm, 1, g["tau_A"].ravel(), g["snr_A"].ravel(), True, P=P)
    assert rel_l2(g["hDmmse_A"], hM) <= 1e-9                       # FP64 LS feeds the FP64 smoother
    # the smoother alone, fed MATLAB's own hD: FP64 end to end
    nt, nr = P.shape[0], rx.shape[2]
    with mm.Engine(nt, nr, rx.shape[0], mlp=False) as eng:
        out = eng.lmmse(np.ascontiguousarray(np.transpose(g["hD_A"], (2, 1, 0))[None]), lmmse.tau_rms(g["tau_A"].ravel()),
                        g["snr_A"].ravel()[None])
    assert rel_l2(g["hDmmse_A"], np.transpose(out[0], (2, 1, 0))) <= 1e-9
    for tag in ("nps2", "sec"):
        n, nps, s = g["ce_par_" + tag]
        with mm.Engine(1, 1, int(n), n_ltf=1, n_ps=int(nps), mlp=False) as eng:
            y = eng.lmmse(g["ce_x_" + tag].reshape(1, 1, 1, -1), lmmse.tau_rms(g["ce_h_" + tag].ravel()), float(s))
        assert rel_l2(g["ce_y_" + tag].ravel(), y.ravel()) <= 1e-9
