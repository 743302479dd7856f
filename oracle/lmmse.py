"""Oracle LMMSE smoother (restates LMMSE_ce.m and its call site in helperMIMOChannelEstimate.m).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference: packet_generation/phased_arr/LMMSE_ce.m:23-39

    snr = 10^(SNR*0.1)                                                  (:23)
    k = 0:length(h)-1;  hh = h*h';  tmp = h.*conj(h).*k                 (:27-28)
    r = sum(tmp)/hh;  r2 = tmp*k.'/hh;  tau_rms = sqrt(r2 - r^2)        (:29-30)
    df = 1/Nfft;  j2pi_tau_df = j*2*pi*tau_rms*df                       (:31-32)
    rf  = 1./(1 + j2pi_tau_df*(K1 - K2*Nps)),  K1 = row 0..Nfft-1, K2 = col 0..Np-1     (:33-34)
    rf2 = 1./(1 + j2pi_tau_df*Nps*(K3 - K4)),  K3 = row 0..Np-1,   K4 = col 0..Np-1     (:35-36)
    Rhp = rf;  Rpp = rf2 + eye(Np)/snr                                   (:37-38)
    H_MMSE = transpose(Rhp*inv(Rpp)*H_tilde)                             (:39)

and the caller, packet_generation/phased_arr/helperMIMOChannelEstimate.m:37-39:

    hDmmse(:,j,i) = LMMSE_ce(hD(:,j,i), Nsc, Nsc, Nps, tau, SNR(i))     (Nfft = Np = numel(CarriersLocations))

`h` is whatever the caller passes as `tau` (generate_maMIMO_LTF.m:342 passes the path-delay vector of the
scattering channel); LMMSE_ce only uses it through tau_rms, so the engine's API takes tau_rms per packet
(`tau_rms()` below is the restated formula) and SNR in dB per (packet, rx).

PINNED: tests/golden/ref_matlab_ls_lmmse.npz holds the outputs of the reference's own LMMSE_ce.m /
helperMIMOChannelEstimate.m source text, executed unmodified by tests/golden/mini_matlab.py (MATLAB / Octave are
not installed); tests/test_golden_matlab.py checks this module against them to 1e-12 (per-rx SNR, Nps = 2,
delays in seconds).  Further anchors (tests/test_lmmse_oracle.py): the literal form equals
the solve form; Rpp is Hermitian positive definite; tau_rms of a two-tap profile has the closed form; with
Nps = 1 the result equals H - (1/snr) * inv(Rpp) * H; snr -> inf gives the identity to ~1e-6.
"""
import numpy as np


def tau_rms(h):
    """LMMSE_ce.m:27-30 for a row vector h (real or complex)."""
    h = np.asarray(h).ravel().astype(np.complex128)
    k = np.arange(h.size, dtype=np.float64)
    hh = np.real(np.vdot(h, h))                   # h*h'
    tmp = (h * np.conj(h)).real * k               # h.*conj(h).*k
    r = tmp.sum() / hh
    r2 = (tmp * k).sum() / hh
    return float(np.sqrt(r2 - r * r))


def correlation_matrices(n_fft, n_p, n_ps, t_rms, snr_db):
    """(Rhp [n_fft, n_p], Rpp [n_p, n_p]) exactly as LMMSE_ce.m:23,31-38."""
    snr = 10.0 ** (snr_db * 0.1)
    df = 1.0 / n_fft
    j2pi_tau_df = 1j * 2.0 * np.pi * t_rms * df
    K1 = np.arange(n_fft, dtype=np.float64)[:, None]
    K2 = np.arange(n_p, dtype=np.float64)[None, :]
    rf = 1.0 / (1.0 + j2pi_tau_df * (K1 - K2 * n_ps))
    K3 = np.arange(n_p, dtype=np.float64)[:, None]
    K4 = np.arange(n_p, dtype=np.float64)[None, :]
    rf2 = 1.0 / (1.0 + j2pi_tau_df * n_ps * (K3 - K4))
    return rf, rf2 + np.eye(n_p) / snr


def lmmse_ce(h_tilde, n_fft, n_p, n_ps, h, snr_db, literal=True):
    """One pair: h_tilde [n_p] -> H_MMSE [n_fft] (LMMSE_ce.m:39).  literal=True evaluates Rhp*inv(Rpp)*H
    with an explicit inverse like MATLAB's inv(); literal=False uses a solve (better conditioned)."""
    h_tilde = np.asarray(h_tilde, dtype=np.complex128).ravel()
    Rhp, Rpp = correlation_matrices(n_fft, n_p, n_ps, tau_rms(h), snr_db)
    if literal:
        return Rhp @ np.linalg.inv(Rpp) @ h_tilde
    return Rhp @ np.linalg.solve(Rpp, h_tilde)


def helper_mmse_loop(hD, n_ps, tau, snr_db):
    """helperMIMOChannelEstimate.m:33-39, MATLAB shapes: hD [Nsc, numSTS, Nr], snr_db [Nr] -> hDmmse like hD."""
    hD = np.asarray(hD, dtype=np.complex128)
    nsc, nsts, nrx = hD.shape
    out = np.zeros_like(hD)
    for i in range(nrx):
        for j in range(nsts):
            out[:, j, i] = lmmse_ce(hD[:, j, i], nsc, nsc, n_ps, tau, snr_db[i])
    return out


def lmmse_batched(H_ls, t_rms, snr_db, n_ps=1):
    """Batched form in the engine's layout: H_ls [Npkt, Nr, Nt, Nsc], t_rms [Npkt], snr_db [Npkt, Nr]
    -> H_mmse [Npkt, Nr, Nt, Nsc] complex128 (one solve per (pkt, rx), all Nt pairs as right-hand sides)."""
    H_ls = np.asarray(H_ls, dtype=np.complex128)
    npkt, nrx, nt, nsc = H_ls.shape
    t_rms = np.broadcast_to(np.asarray(t_rms, dtype=np.float64), (npkt,))
    snr_db = np.broadcast_to(np.asarray(snr_db, dtype=np.float64), (npkt, nrx))
    out = np.empty_like(H_ls)
    cache = {}
    for p in range(npkt):
        for i in range(nrx):
            key = (float(t_rms[p]), float(snr_db[p, i]))
            if key not in cache:
                Rhp, Rpp = correlation_matrices(nsc, nsc, n_ps, key[0], key[1])
                cache[key] = Rhp @ np.linalg.inv(Rpp)
            out[p, i] = (cache[key] @ H_ls[p, i].T).T
    return out
