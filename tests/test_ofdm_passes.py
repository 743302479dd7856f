"""numpy statement of the three register passes of csrc/ofdm.cuh (ofdm_r16_kernel / ofdm_r16_tma_kernel): N = 16 x 16 x R3,
Stockham autosort passes with radices (16, 16, R3) and stage sizes ns = (1, 16, 256), checked against numpy's FFT.
Documents the index algebra the CUDA kernel implements (pass-2 twiddle exp(-2 pi i r k / 256), pass-3 twiddle
exp(-2 pi i r k / N), output of pass p at (j - k) R + k + r ns)."""
import numpy as np
import pytest


def stockham_pass(x, radix, ns):
    """one Stockham pass: for j < N/R, k = j mod ns: out[(j - k) R + k + r ns] = DFT_R over q of in[j + q N/R] w^(q k)"""
    n = x.size
    out = np.empty_like(x)
    for j in range(n // radix):
        k = j % ns
        v = np.array([x[j + q * (n // radix)] * np.exp(-2j * np.pi * q * k / (radix * ns)) for q in range(radix)])
        y = np.fft.fft(v)                                   # the in-register DFT-R, natural order in and out
        for r in range(radix):
            out[(j - k) * radix + k + r * ns] = y[r]
    return out


@pytest.mark.parametrize("log2n", [8, 9, 10, 11, 12])
def test_three_register_passes_equal_the_fft(log2n):
    n = 1 << log2n
    r3 = n // 256
    rng = np.random.default_rng(log2n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = stockham_pass(x, 16, 1)
    y = stockham_pass(y, 16, 16)
    if r3 > 1:
        y = stockham_pass(y, r3, 256)
    assert np.linalg.norm(y - np.fft.fft(x)) / np.linalg.norm(y) < 1e-12


def test_window_rotation_matches_the_reference_mirror():
    """window = x[cp : N + off] ++ x[off : cp]  (massiveMIMO_dataGenerator.py:442) is the symbol with its first `off`
    samples moved to the end: a cyclic rotation of the periodic extension.  With off == cp the window is the body itself
    and the cyclic prefix is never read (what the bulk-copy kernel exploits)."""
    n, cp = 64, 16
    rng = np.random.default_rng(1)
    body = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    sym = np.concatenate([body[-cp:], body])                 # CP + body
    for off in (0, 5, cp):
        first = n + off - cp
        window = np.concatenate([sym[cp:cp + first], sym[off:off + (n - first)]])
        assert window.size == n
        if off == cp:
            assert np.array_equal(window, body)
        # the window is a cyclic rotation of samples of the periodic extension: its spectrum has the body's magnitudes
        assert np.allclose(np.abs(np.fft.fft(window)), np.abs(np.fft.fft(body)))
    assert np.array_equal(np.concatenate([sym[cp:cp + n]]), body)
