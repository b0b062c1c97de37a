"""CPU oracle for the radiobear_b200 hot paths -- TEST INFRASTRUCTURE ONLY.

This package is a plain numpy/scipy restatement of the reference algorithms
(david-deboer/radiobear v2.0.1) for the two hot paths:

* ``alpha_oracle``  -- constituents/<gas>/<formalism>.alpha + Alpha.get_layers
* ``ray_oracle``    -- raypath.compute_ds / findEdge / Shape._calcEllipse
* ``rt_oracle``     -- Brightness.single

It is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product package ``radiobear_b200`` never imports ``oracle`` and fails loudly when the
CUDA library is missing.

Parity pin: every function here is checked in ``tests/test_oracle_golden.py`` against
golden vectors produced by importing the unmodified Python reference in the build
container (``tests/golden/make_golden.py``; the generating script is committed) and against the
reference's own known-answer table ``scripts/benchmark.py:20-24``.
"""
