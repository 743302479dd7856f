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
* ``oracle.ls`` (helperMIMOChannelEstimate.m) and ``oracle.lmmse`` (LMMSE_ce.m and its call site) are PINNED
  against the reference's own MATLAB source text, executed unmodified by the MATLAB-subset interpreter
  ``tests/golden/mini_matlab.py`` (MATLAB / Octave are not installed): ``tests/golden/ref_matlab_ls_lmmse.npz``,
  checked in ``tests/test_golden_matlab.py`` (LS 1e-16, LMMSE 1e-12 incl. per-rx SNR, Nps = 2 and the data-phase
  numSTS = 1 case); the same file pins ``postproc.nmse_subk`` / ``rows_to_csi`` on BER_test_maMIMO_LTF.m's own
  text.  ``helperGetP`` (a MathWorks example helper, not in the reference repo) is input data there.
* ``oracle.mlp`` (Keras FC graph) restates TensorFlow layer semantics that cannot run here (no TensorFlow / h5py):
  **parity unpinned** by execution of TensorFlow itself; the glue around it (``CSIPredictor.inference`` end to end
  with a numpy Keras stand-in) is pinned, and the layer arithmetic is anchored on BN-fold == unfused identities.
* ``oracle.svd`` (first lines of omphybweights.m's getWeightsForSubcarrier) is PINNED on those lines executed by
  ``mini_matlab`` (``tests/golden/ref_svd.npz``), through the basis-independent invariants only.
* ``oracle.omp`` (ompdecomp.m's loop and the precoding lines of omphybweights.m's getWeightsForSubcarrier) is PINNED on
  that text executed by ``mini_matlab`` (``tests/golden/ref_omp.npz``): atom indices exactly, coefficients to 1e-12,
  the early stop on an exactly representable residual.
* ``oracle.interp`` has no reference counterpart at all (the reference always
  uses Nps = 1): **parity unpinned**, defined here.
"""

from . import tables, ls, mlp, postproc, interp, lmmse, svd, omp  # noqa: F401
