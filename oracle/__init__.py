"""CPU oracle: a NumPy/C restatement of the reference (patrick-kidger/optimistix 0.1.0) hot path.

THIS IS TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import or execute anything in
this directory.  The product (``optimistix_b200``) never does, and raises if its CUDA library
is missing rather than falling back here.

Pinning status
--------------
The reference is pure Python on JAX; ``jax``/``equinox``/``lineax`` are not installed in this
image, so the reference itself cannot be executed here or on the GPU box.  The restatement is
therefore pinned against every *known answer* the reference's own tests hold for this path
(``tests/test_oracle_known_answers.py``: 0.96118069, Beale [3, 0.5], exact 3.0, the
``nonfinite`` / ``nonlinear_max_steps_reached`` result codes, the Newton bounded roots, the
least-squares / root-find / minimise zoo at the reference tolerances), all of which are
tolerance checks.  The reference's tests never pin iteration counts or exact iterates
(SURVEY.md F6), so **iteration-count parity is unpinned**: it is checked CUDA-vs-oracle only.
The linear algebra of the reference lives in ``lineax`` (unpinned ``>=0.0.6``, not vendored);
its published behaviour (Householder QR least squares, LU with partial pivoting, "non-finite
solution => singular") is restated with LAPACK through NumPy.
"""
