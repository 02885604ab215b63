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

PARITY PINNED TO THE REFERENCE ITSELF (round 2).  CasADi / IPOPT cannot be
installed here and the reference's tests hold no golden vector for ``g(z)``,
``jac_g(z)`` or the Jacobian pattern -- but the
reference's transcription is pure Python over a SMALL slice of CasADi's API.
``oracle/refrun`` provides that slice (a scalar-expression engine with SX's
construction-time simplifications, ``Function``, sparse forward AD, an exact
``integrator``, an interior-point ``nlpsol``), imports
``/root/reference/mpopt/mpopt.py`` UNMODIFIED and runs its own
``create_nlp()``, residual helpers and solves:

* ``tests/golden/ref_*.npz`` (33 NLPs: every problem, scheme, row block, both
  adaptive classes) hold f, g, grad_f, CSR Jacobian, Lagrangian Hessian, bounds
  and initial guess computed BY THE REFERENCE'S CODE; this oracle reproduces
  them with identical index arrays and values to 1e-12
  (``tests/test_reference_golden.py``), and so does the CUDA path on the GPU;
* ``tests/golden/refres_*.npz``: the same for the interpolation / residual path;
* the reference's own 45 unit tests pass on the stand-in, and its solves land on
  the optima stored in its notebooks (real CasADi + IPOPT output) to 8e-8 ..
  3e-6 (``tests/test_reference_suite.py``), which qualifies the stand-in.

What the stand-in cannot reproduce is CasADi's floating-point operation ORDER
inside AD (values agree to rounding, not bit for bit) and IDAS's integration
error in the quadrature weights (quirk Q2).  Also pinned, in
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
