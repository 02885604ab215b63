"""Generates tests/golden/delta3_drag_kkt.npz: a Karush-Kuhn-Tucker point (z, lam_g, lam_x) of the Delta III
transcription (4 phases, one segment of degree 11 per phase, LGR, aerodynamic drag on) found by the interior-point
solver on the ORACLE's evaluators: cold start without drag, then continued with drag, exactly as the reference's
notebook does (docs/source/notebooks/multi_stage_launch_vehicle_ascent.ipynb).  This problem is hard for any solver
(IPOPT takes 320 + 119 iterations), so the anchor tests do not re-solve it: they CERTIFY the stored point with the
evaluators under test -- feasibility, stationarity, complementarity and multiplier signs to 1e-8 -- and compare the
objective THEY evaluate there with the number the notebook stores.
Run: python tests/golden/make_anchor_start.py   (about two minutes on CPU)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mpopt_b200.ipm import solve_nlp  # noqa: E402
from mpopt_b200.problems import delta3_launch_vehicle  # noqa: E402
from oracle.hessian import hess_l  # noqa: E402
from oracle.nlp import OracleNLP  # noqa: E402


def solve(ora, z0, **kw):
    lbx, ubx, lbg, ubg = ora.bounds()
    p = ora.seg_width_params()
    return solve_nlp(lambda z: ora.f(z, p), lambda z: ora.grad_f(z, p), lambda z: ora.g(z, p), lambda z: ora.jac_g(z, p),
                     lambda z, lf, lg: hess_l(ora, z, p, lf, lg), z0, lbx, ubx, lbg, ubg, **kw)


a = OracleNLP(delta3_launch_vehicle(0.0), 1, 11, "LGR")
r0 = solve(a, a.initialize_solution(), tol=1e-9, max_iter=300)
b = OracleNLP(delta3_launch_vehicle(1.0), 1, 11, "LGR")
r1 = solve(b, r0.x, tol=1e-10, max_iter=500)
assert r1.success, r1
print("objective with drag:", repr(r1.f), "iterations", r1.iter, "error", r1.err)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "delta3_drag_kkt.npz"), z=r1.x, lam_g=r1.lam_g, lam_x=r1.lam_x)
