"""CPU oracle for the B200 variogram / ordinary-kriging hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or as the
CPU arm being timed), never as the thing shipped.  The product path
(``scikit-gstat_b200/``) never imports this package and fails loudly when the
CUDA library is missing.

The oracle is a numpy/scipy restatement of scikit-gstat 1.0.23's algorithm for
the path ``distance -> lag bin -> estimator`` (+ directional mask) and
``OrdinaryKriging.transform``.  The arithmetic of that path lives in third
party wheels that the reference does not pin (``requirements.txt:1-9``):
scipy ``pdist``/``cdist`` (1.18.1 here), numpy ``digitize``/``median``/
``percentile``/``linalg.inv`` (2.3.5 here).  The oracle calls those same
functions at the reference's own call sites, so it is the reference's
arithmetic, minus the Python classes around it.

Parity pinning: every function here is checked (tests/test_oracle_*.py)
against golden vectors produced by importing the unmodified reference from
``/root/reference`` in the build container (``oracle/make_golden.py``; vectors
committed under ``tests/golden/``) and against the known-answer values that
the reference's own tests hold (``skgstat/tests/test_variogram.py:677-685``,
``test_binning.py:17-79``, ``test_estimator.py:16-59``,
``test_directionalvariogram.py:109-135``).
"""
