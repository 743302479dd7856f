"""Oracle subcarrier interpolation (north_star addition; NOT in the reference).

Every call site of the reference passes pilot spacing Nps = 1
(generate_maMIMO_LTF.m:342,578; BER_test_maMIMO_LTF.m:288,331,536) and Nps is
only consumed by LMMSE_ce.m:34,36, so nothing in the reference pins this
function: **parity unpinned**; the definition below IS the specification.

Definition (comb pilots at k = 0, Nps, 2*Nps, ...):
  * LS is evaluated at pilot tones only.
  * Between two pilots, real and imaginary parts are interpolated linearly
    and independently:  H[k] = H[p0] + (k - p0)/Nps * (H[p1] - H[p0]).
  * Past the last pilot the last segment's slope is extended (linear
    extrapolation); with a single pilot the value is held.
  * Nps = 1 is the identity (every tone is a pilot), bit for bit.

To make the GPU and oracle agree on rounding, the weights are formed as
w = (k - p0) / Nps in the working precision and the result as
(1 - w) * H[p0] + w * H[p1] evaluated as  H[p0] + w * (H[p1] - H[p0]).
"""
import numpy as np


def pilot_positions(nsc, nps):
    return np.arange(0, nsc, nps)


def interp_linear(H_pilot, nsc, nps):
    """H_pilot [..., Np] at k = 0, Nps, ... -> H [..., Nsc]."""
    H_pilot = np.asarray(H_pilot)
    if nps == 1:
        assert H_pilot.shape[-1] == nsc
        return H_pilot.copy()
    npil = H_pilot.shape[-1]
    assert npil == len(pilot_positions(nsc, nps))
    k = np.arange(nsc)
    if npil == 1:
        return np.repeat(H_pilot, nsc, axis=-1)
    seg = np.minimum(k // nps, npil - 2)           # last segment extended
    w = (k - seg * nps).astype(H_pilot.real.dtype) / H_pilot.real.dtype.type(nps)
    h0 = H_pilot[..., seg]
    h1 = H_pilot[..., seg + 1]
    return h0 + w * (h1 - h0)
