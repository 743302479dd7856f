"""Integer tables of the reference LS estimator (oracle side).

Follows packet_generation/phased_arr/helperMIMOChannelEstimate.m:16-23 (the
256-tone VHT-LTF pattern with a single DC null) and
packet_generation/phased_arr/generate_maMIMO_LTF.m:96-102 (null / pilot carrier
sets and CarriersLocations = setdiff(1:256, nulls U pilots)).

The tone table is rebuilt here from its 802.11 structure (two 26-tone L-LTF
halves glued by fixed connector runs) rather than stored as a flat literal, so
that the product's own flat table (csrc/tables.h) and this one are two
independent encodings that the tests compare bit for bit.
"""
import numpy as np

FFT_LEN = 256
CP_LEN = 64

# 802.11 L-LTF halves (26 tones each), as +/-1.
_LEFT = "++--++-+-++++++--++-+-++++"
_RIGHT = "+--++-+-+----" "-++--+-+-++++"


def _pm(s):
    return [1 if c == "+" else -1 for c in s]


def vht_ltf256():
    """ltf[0..255] in {-1,0,+1}; index i here is MATLAB index i+1.

    helperMIMOChannelEstimate.m:20-23: 7 guard zeros, then four
    (left, +1, right) blocks separated by connector runs, one DC zero in the
    middle, 6 guard zeros at the end.
    """
    blk = _pm(_LEFT) + [1] + _pm(_RIGHT)          # 53 tones
    conn_a = _pm("---++-+-++-")                    # 11 tones (line 20/22 tail)
    mid_l = _pm("+-+-")                            # before DC
    mid_r = _pm("+--+")                            # after DC
    seq = ([0] * 7 + blk + conn_a + blk + mid_l + [0] + mid_r
           + blk + conn_a + blk + [0] * 6)
    out = np.asarray(seq, dtype=np.int8)
    assert out.shape == (FFT_LEN,)
    return out


def null_carrier_indices():
    """1-based, generate_maMIMO_LTF.m:99: [1:7 129 256-5:256]."""
    return np.asarray(list(range(1, 8)) + [129] + list(range(251, 257)), dtype=np.int32)


def pilot_carrier_indices():
    """1-based, generate_maMIMO_LTF.m:100."""
    return np.asarray([26, 54, 90, 118, 140, 168, 204, 232], dtype=np.int32)


def carriers_locations():
    """1-based ascending data-carrier indices (234 of them), generate_maMIMO_LTF.m:101-102."""
    non_data = set(null_carrier_indices().tolist()) | set(pilot_carrier_indices().tolist())
    return np.asarray([i for i in range(1, FFT_LEN + 1) if i not in non_data], dtype=np.int32)


def ltf_at_carriers():
    """ltf_o = ltf(ind), helperMIMOChannelEstimate.m:29 (234 values, all +/-1)."""
    return vht_ltf256()[carriers_locations() - 1]


def sylvester_hadamard(n):
    """Default stand-in for helperGetP(numSTS) (helperMIMOChannelEstimate.m:13).

    helperGetP is a MathWorks example helper that is NOT in the reference repo;
    the reference persists P in every .mat and treats it as data.  We use the
    Sylvester construction (natural order) as the default +/-1 orthogonal
    mapping matrix: P P^H = n I.  Parity unpinned (documented assumption).
    """
    assert n >= 1 and (n & (n - 1)) == 0, "numSTS must be a power of two (generate_maMIMO_LTF.m:17,24)"
    h = np.ones((1, 1), dtype=np.float64)
    while h.shape[0] < n:
        h = np.block([[h, h], [h, -h]])
    return h
