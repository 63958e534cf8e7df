"""CPU oracle for the linear-MPC / structured-NN hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product package
(``industrial_nnmpc_2021_b200``); only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and there only
as the checker or as the timed CPU baseline.

It restates, in plain NumPy/SciPy float64, the algorithm of the reference
(``/root/reference/lib/linearMPC.py``, ``lib/LinearMPCLayers.py``,
``lib/controller_evaluation.py``); each function cites the file:line it follows.

Parity pinning status
---------------------
The reference ships no tests, golden vectors or saved datasets, and its QP solver
(``cvxopt.solvers.qp``, version unpinned, not installable here) cannot run in this
container: **solver parity is unpinned** by reference artefacts.  What IS pinned, by fixtures
generated from the reference's own code imported from ``/root/reference`` (see
``tests/golden/make_golden.py``): the QP *formulation* (``P, tq, G, h, tA, tB, Pf, Krep`` of
``DenseQPRegulator``; ``P, G, h, tA, tb, q, b`` of ``TargetSelector``), the augmented
matrices, ``dlqr``, ``sample_prbs_like`` and ``get_updated_average_stage_cost``.  The regulator
QP is strictly convex, so its minimiser is unique and solver independent; the oracle solves it
to KKT <= 1e-12, tighter than cvxopt's own stopping tolerances.
"""
