"""CPU oracle for the MPC solve hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker / the CPU arm being timed.
The product path (``autompc_b200``) never imports this package and fails loudly
when its CUDA library is missing.

Parity status
-------------
The reference's own test-suite does not pin this path (SURVEY.md section 4 / 8c:
no test imports MPPI, MLP or runs iLQR).  The oracle is therefore pinned against
outputs of the *unmodified reference code itself*, imported in the build
container from ``/root/reference`` through ``oracle/ref_loader.py`` and frozen
as fixtures in ``tests/golden/`` by ``oracle/make_golden.py`` (committed), plus
the three QuadCost known answers of ``tests/test_costs.py:192-205``.

For ``ctrl_dim > 1`` the unmodified reference MPPI raises
(``autompc/control/mppi.py:139``); the oracle is the documented one-line
restatement (noise drawn with trailing dim ``ctrl_dim`` -- the intent recorded
in the commented line ``mppi.py:125``), everything else unchanged.
"""
