"""CPU oracle: a numpy restatement of the reference's hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU reference -- never as the thing shipped.  The product path
(``dl-channel-estimation-mamimo_b200``) never imports this package and fails
loudly when its CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):

* ``oracle.postproc`` (CSIPredictor glue, pair ordering, per-pair input
  assembly) is PINNED against the reference's own Python code executed in the
  build container with a stub ``tensorflow`` module
  (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).
* ``oracle.tables`` (VHT-LTF tone table, carrier index set) is pinned by the
  integer known-answers listed in SURVEY.md section 8(a2)/(c).
* ``oracle.ls`` (helperMIMOChannelEstimate) and ``oracle.mlp`` (Keras FC
  graph) restate MATLAB / TensorFlow code that cannot run here (no MATLAB,
  Octave, TensorFlow, h5py): **parity unpinned** by execution of the
  reference; they are anchored on algebraic known-answer identities only.
* ``oracle.lmmse`` (LMMSE_ce.m via helperMIMOChannelEstimate.m:37-39) restates MATLAB code: **parity
  unpinned** by execution; anchored on the identities listed in its header.
* ``oracle.interp`` has no reference counterpart at all (the reference always
  uses Nps = 1): **parity unpinned**, defined here.
"""

from . import tables, ls, mlp, postproc, interp, lmmse  # noqa: F401
