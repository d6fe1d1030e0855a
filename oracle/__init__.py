"""CPU oracle for the hot path (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED: the reference is TensorFlow-1.15 graph code; TensorFlow is not
installable in this image (Python 3.12, no network) and the reference ships no
tests, golden vectors or fixtures.  Everything in this package is therefore a
*restatement* of the reference algorithm from its call sites plus the published
TF-1.15 semantics, cross-checked against float64
naive-loop micro-oracles (``naive64``) and pinned by the committed fixtures in
``tests/golden`` that ``oracle/make_golden.py`` generates.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
(``unsupervised_anomaly_detection_brain_mri_b200``) never does.
"""
