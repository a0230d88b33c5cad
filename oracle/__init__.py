"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference denoiser path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or the
reported CPU baseline -- never as the thing measured or shipped.

PARITY UNPINNED: the reference's arithmetic lives in TensorFlow 1.x
(tf.contrib.slim / tf.layers / tf.image / tf.nn, version not pinned anywhere
in the reference tree; era evidence says TF 1.5-1.8), which is not installed
here and cannot be (TF1 contrib does not exist for Python 3.12, no network).
The reference tree holds no golden vectors, known-answer tests or fixtures
for this path.  The oracle therefore restates the
published TF op semantics and is anchored on closed-form
mini-cases, adjoint identities and the integer tile-plan goldens of
SURVEY.md App. D -- not on outputs of the reference itself.  ``naive.py`` is a
second, independently written restatement (plain numpy); the CPU suite checks
that the two agree to rounding.
"""
