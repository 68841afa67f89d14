"""CPU oracle for the madrona-learn learner hot path.

TEST INFRASTRUCTURE ONLY.  This package is a NumPy/C restatement of the reference's
algorithms (each function cites the reference file:line it follows).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import, call, link or execute anything under ``oracle/`` -- and there only as the checker
or as the timed CPU baseline, never as part of the product path.  The product
(``madrona-learn_b200``) never imports this package and fails loudly when its CUDA library is
missing.

Parity pinning (see DESIGN.md "Oracle"):
  * jax/flax/optax are NOT installable in this image, so the reference cannot run natively.
  * ``oracle/jax_shim`` is a NumPy stand-in for the small slice of the jax API the reference's
    hot-path functions use; ``tests/golden/make_golden.py`` imports the *unmodified reference
    source files* from /root/reference under that shim and records their outputs as
    fixtures.  The restatement below is checked against those fixtures
    (``tests/test_oracle_golden.py``).  Functions marked ``PARITY UNPINNED`` in their
    docstring (threefry PRNG bits, flax Dense/LayerNorm/LSTM numerics, optax Adam) depend on
    third-party packages absent from /root/reference and are restated from their published
    algorithms; the known-answer vectors we hold for them are listed in DESIGN.md.
"""

from . import algo_common, moving_avg, metrics, prng, layouts  # noqa: F401
