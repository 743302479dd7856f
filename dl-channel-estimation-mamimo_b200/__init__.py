"""B200-native massive-MIMO OFDM channel-estimation engine (hot path only).

Importing this package loads libmamimo_b200.so (hand-written sm_100a kernels behind the C ABI of
include/mamimo.h).  There is no CPU fallback: a missing library raises ImportError.

The directory name contains hyphens, so import it through the root-level shim:

    import mamimo_b200 as mm
    eng = mm.Engine(n_tx=32, n_rx=4, n_sc=1024, hidden=(1024, 1024))
"""
from . import build, synth, sharding, pipeline, weights, cli  # noqa: F401
from ._capi import MamimoError, PRECISIONS, INPUT_MODES  # noqa: F401
from .engine import (Engine, CSIPredictor, helperMIMOChannelEstimate, vht_ltf256, carriers_locations,  # noqa: F401
                     default_p, pair_row, pinned_empty, tau_rms)

__all__ = ["Engine", "CSIPredictor", "helperMIMOChannelEstimate", "vht_ltf256", "carriers_locations",
           "default_p", "pair_row", "pinned_empty", "tau_rms", "MamimoError", "synth", "build", "sharding", "pipeline", "weights", "cli"]
