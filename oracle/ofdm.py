"""Oracle OFDM demodulation front-end.

Reference call (MATLAB Communications Toolbox, not in the repo):
    rxOFDM = ofdmdemod(inputRXSig, prm.FFTLength, prm.CyclicPrefixLength, prm.CyclicPrefixLength,
                       prm.NullCarrierIndices, prm.PilotCarrierIndices)      generate_maMIMO_LTF.m:336-338
and the author's own numpy mirror of it, massiveMIMO_dataGenerator.py:437-453:
    input2D  = reshape(ltfSig, (FFT+CP, numSym), order='F')                  (:437-439)
    noCP_ix  = range(CP, FFT + symOffset) ++ range(symOffset, CP)            (:442)
    afterFFT = fft(input2D[noCP_ix, :], n=FFT, axis=0)                       (:451)
    fftshift on the frequency axis                                           (:453 -- the mirror's bare
        np.fft.fftshift shifts BOTH axes, which is a bug for numSym > 1; ofdmdemod shifts frequency only)
then the null and pilot carriers are removed, leaving prm.CarriersLocations (generate_maMIMO_LTF.m:101-102).

Shapes: x [Npkt, Nr, numSym*(FFT+CP)] (MATLAB inputRXSig [lenLTF x Nr] per packet, column-major) ->
        Y [Npkt, Nr, numSym, Nsc]     (MATLAB rxOFDM [Nsc x numSym x Nr] per packet) -- the LS stage's input.

ofdmdemod itself cannot run here (no MATLAB): parity unpinned by execution; pinned only by the round trip
against an ofdmmod-style modulator (tests) and by agreement with the author's numpy mirror for numSym = 1.
"""
import numpy as np


def window_indices(fft_len, cp_len, sym_offset):
    """massiveMIMO_dataGenerator.py:442."""
    return np.asarray(list(range(cp_len, fft_len + sym_offset)) + list(range(sym_offset, cp_len)), dtype=np.int64)


def ofdm_demod(x, fft_len, cp_len, sym_offset, carriers_1based, n_sym=None):
    x = np.asarray(x)
    sym_len = fft_len + cp_len
    if n_sym is None:
        n_sym = x.shape[-1] // sym_len
    xs = x[..., : n_sym * sym_len].reshape(x.shape[:-1] + (n_sym, sym_len))      # F-order reshape of a column
    win = xs[..., window_indices(fft_len, cp_len, sym_offset)]
    spec = np.fft.fftshift(np.fft.fft(win.astype(np.complex128), n=fft_len, axis=-1), axes=-1)
    return spec[..., np.asarray(carriers_1based, dtype=np.int64) - 1]


def ofdm_mod(grid, fft_len, cp_len, carriers_1based):
    """Inverse used by the round-trip test: data carriers -> time-domain symbols with cyclic prefix
    (what helperGenPreamble / ofdmmod produce; unnormalised so that demod(mod(G)) == G)."""
    grid = np.asarray(grid, dtype=np.complex128)                                  # [..., n_sym, n_sc]
    full = np.zeros(grid.shape[:-1] + (fft_len,), dtype=np.complex128)
    full[..., np.asarray(carriers_1based, dtype=np.int64) - 1] = grid
    t = np.fft.ifft(np.fft.ifftshift(full, axes=-1), axis=-1)
    t = np.concatenate([t[..., fft_len - cp_len:], t], axis=-1)                   # prepend CP
    return t.reshape(t.shape[:-2] + (-1,))
