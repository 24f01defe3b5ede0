"""CPU oracle for the b200med hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This package restates, on the CPU, the algorithm of the reference's train / inference hot path
(GonzaloPlaaza/Multimodal-Error-Detection, ``MED/dataset`` + ``MED/modeling``).  Each function
cites the reference file:line it follows.  Integer / index work is restated in numpy loops and
in plain C (``oracle/c/med_oracle.c``); floating-point layers call the same third-party
arithmetic the reference calls (``torch`` CPU fp32 ops, ``sklearn.metrics``), because that is
where the reference's arithmetic lives.

Pinning: the reference has no tests and no golden vectors of its own, so the oracle is pinned
against outputs of the *reference itself*, executed in the build container by
``tests/golden/make_golden.py`` (which imports ``/root/reference/MED`` unmodified, with empty
stubs for the absent ``mlflow``/``clip`` packages) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every oracle function against those fixtures.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker / the timed CPU
baseline.  Nothing under ``multimodal_error_detection_b200/`` imports it.
"""
