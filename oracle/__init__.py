"""CPU oracle for the hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Everything under ``oracle/`` is a CPU restatement (numpy fp64 / plain C) of the
reference's algorithm for the path named in BASELINE.json:north_star.  It is the
*checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import, link or execute it.  The
product (``jmd-diversity-in-bayesian-optimization_b200/``) never imports it and
has no CPU fallback.

Pinning status:
* DPP scoring / ranked sampling, combinadic un-ranking, simplex noise and the
  Wildcat Wells surface are pinned against outputs of the reference's own
  importable code (``tests/golden/*.npz``, produced by
  ``tests/golden/make_golden.py`` through ``oracle/refshim.py``).
* GP posterior / EI / marginal likelihood: the arithmetic lives in botorch 0.8.0 /
  gpytorch 1.9.0 (environment.yml:26,47), which are neither vendored under
  /root/reference nor installed here, and the reference has no tests
  => **parity unpinned** for those functions.  The restatement follows the
  published algorithms and is cross-checked against
  scikit-learn's independent exact-GP implementation instead.
"""
