"""CPU oracle for the mpopt collocation-transcription hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.

What it restates
----------------
``/root/reference/mpopt/mpopt.py`` (pure Python over CasADi + SciPy), lines

* ``CollocationRoots``  4134-4276  -> :mod:`oracle.collocation`
* ``Collocation``       3706-4131  -> :mod:`oracle.collocation`
* ``mpopt`` transcription 95-639, initial guess 641-708 -> :mod:`oracle.nlp`
* CasADi's forward AD + structural sparsity (third-party, casadi==3.6.0 pinned
  in the reference's requirements.txt:4, absent from /root/reference and from
  this image) -> :mod:`oracle.dual`, a vectorised dual-number restatement.

PARITY UNPINNED (values of g / jac_g / grad_f): the reference's own tests hold
no golden vector for ``g(z)``, ``jac_g(z)`` or the Jacobian pattern
 and CasADi/IPOPT cannot be installed here, so the
reference cannot be run.  What *is* pinned, and is checked in
``tests/test_oracle_*.py``:

* the p=1 known-answer tests of tests/test_mpopt.py:927-1086 (nodes, Lagrange
  basis, D = [[-1/h, 1/h], [-1/h, 1/h]], second-order D = 0),
* tau0/tau1 == first/last root (tests/test_mpopt.py:627-634),
* the composite shapes (tests/test_mpopt.py:333-346),
* the seven (n_vars, n_eq, n_ineq) triples printed by IPOPT in the stored
  notebook outputs,
* the hand-derived golden G0 of SURVEY.md Appendix A,
* mpmath 50-digit tables, finite-difference / complex-step Jacobians.
"""
