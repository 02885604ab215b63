"""Reference-held numbers for this path: the optimal objectives the reference's executed notebooks store (IPOPT's
17-digit "Objective" line), each at the notebook's own (n_segments, poly_orders, scheme).

They are the only VALUES the reference holds for the transcription (no g / jac_g vector is stored
anywhere), and they depend on every piece of it: nodes, differentiation matrices, quadrature weights, row / variable
layout, bounds, the node functions and their first and second derivatives.  An optimum is found here with the
interior-point solver of ``mpopt_b200.ipm`` (IPOPT is not installable) driven by a set of evaluators -- the numpy
oracle's on CPU (pins the ORACLE to the reference), the CUDA path's through the C ABI on the GPU (pins the PRODUCT).

Tolerances.  IPOPT stops at a scaled error of 1e-8, so ~1e-7 relative agreement is the best possible; the Mayer-only
problem (Delta III: final mass) reaches it.  Objectives with a running cost go through the quadrature weights, which
the reference integrates with SUNDIALS-IDAS at CasADi's default tolerances (reltol 1e-6, abstol 1e-8;
mpopt.py:3869-3877, quirk Q2) while this package integrates them exactly: re-solving with weights produced by an ODE
solver at those tolerances (tests/test_anchor_oracle.py::test_q2_band) moves the optimum by 3e-6 .. 5e-5 relative, the
same size as the gaps below -- and by far the largest relative weight errors sit on the tiny end-point weights that
dominate the hyper-sensitive objective (boundary layers).  ``rtol`` is about twice the observed gap of each case.
"""
import numpy as np

NB = "docs/source/notebooks/"
# (problem, n_segments, poly_orders, scheme, stored objective, reference file:line, rtol, observed gap)
ANCHORS = [
    ("moon_lander", 10, 6, "LGR", 8.2477255075783038, NB + "moon_lander.ipynb:185", 6e-6, 3.1e-6),
    ("moon_lander", 2, 30, "CGL", 8.2457172048588543, NB + "moon_lander.ipynb:294", 3e-6, 1.2e-6),
    ("moon_lander", 2, 30, "LGL", 8.2425586640613506, NB + "moon_lander.ipynb:387", 5e-7, 8.0e-8),
    ("van_der_pol", 1, 25, "LGR", 2.8734932991287625, NB + "vanderpol.ipynb:191", 6e-6, 2.9e-6),
    ("van_der_pol", 1, 25, "CGL", 2.8734060736207896, NB + "vanderpol.ipynb:294", 3e-6, 1.1e-6),
    ("van_der_pol", 1, 25, "LGL", 2.8734849959084205, NB + "vanderpol.ipynb:387", 3e-6, 1.3e-6),
    ("hyper_sensitive", 5, 50, "LGR", 1.1498050755273090, NB + "hypersensitive.ipynb:179", 1e-5, 4.2e-6),
    ("hyper_sensitive", 5, 50, "LGL", 1.1502075893651909, NB + "hypersensitive.ipynb:374", 5e-5, 2.3e-5),
    ("hyper_sensitive", 5, 50, "CGL", 1.1406025536022588, NB + "hypersensitive.ipynb:281", 2.5e-4, 1.2e-4),
]
# two-phase Schwartz, one segment of degree 20 per phase: the stored optimum is zero to rounding (6.4e-22, 6.7e-22,
# 6.7e-22 for LGR / CGL / LGL; twophaseschwartz.ipynb:209, :312, :405) -- an absolute check
ANCHORS_ZERO = [("two_phase_schwartz", 1, 20, s, 1e-12, NB + "twophaseschwartz.ipynb") for s in ("LGR", "CGL", "LGL")]
# Delta III ascent with drag, 4 phases x one segment of degree 11 (mp.mpopt(ocp, 1, 11)): objective = - final mass /
# lift-off mass, no quadrature involved.  multi_stage_launch_vehicle_ascent.ipynb:545 ("Optimal Solution Found").
DELTA3 = ("delta3_launch_vehicle", 1, 11, "LGR", -2.4977981075384650e-02, NB + "multi_stage_launch_vehicle_ascent.ipynb:545", 3e-7, 6.1e-8)


class Evaluators:
    """f, grad_f, g, jac_g, hess_l, bounds and start of one transcription, from the oracle or from the CUDA path."""

    def __init__(self, kind, ocp, K, p, scheme):
        self.kind = kind
        if kind == "oracle":
            from oracle.hessian import hess_l
            from oracle.nlp import OracleNLP

            o = OracleNLP(ocp, K, p, scheme)
            w = o.seg_width_params()
            self.f, self.grad_f = (lambda z: o.f(z, w)), (lambda z: o.grad_f(z, w))
            self.g, self.jac_g = (lambda z: o.g(z, w)), (lambda z: o.jac_g(z, w))
            self.hess_l = lambda z, lf, lg: hess_l(o, z, w, lf, lg)
            self.bounds, self.start = o.bounds(), o.initialize_solution()
        else:
            from mpopt_b200.nlp import Transcription

            t = Transcription(ocp, K, p, scheme)
            w = t.seg_width_params()
            self.f, self.grad_f = (lambda z: t.f(z, w)), (lambda z: t.grad_f(z, w))
            self.g, self.jac_g = (lambda z: t.g(z, w)), (lambda z: t.jac_g(z, w))
            self.hess_l = lambda z, lf, lg: t.hess_l(z, w, lf, lg)
            self.bounds, self.start = t.bounds(), t.initial_guess()
            self.tr = t

    def solve(self, z0=None, **kw):
        from mpopt_b200.ipm import solve_nlp

        lbx, ubx, lbg, ubg = self.bounds
        return solve_nlp(self.f, self.grad_f, self.g, self.jac_g, self.hess_l, self.start if z0 is None else z0,
                         lbx, ubx, lbg, ubg, **kw)


def check_anchor(kind, problem, K, p, scheme, ref, rtol):
    from mpopt_b200.problems import REGISTRY

    r = Evaluators(kind, REGISTRY[problem](), K, p, scheme).solve(tol=1e-10)
    assert r.success, f"{problem} {scheme}: the interior-point solve did not converge (error {r.err:.1e} after {r.iter})"
    gap = abs(r.f - ref) / abs(ref)
    assert gap <= rtol, f"{problem} K={K} p={p} {scheme}: optimum {r.f!r} vs stored {ref!r}: relative gap {gap:.2e} > {rtol:.0e}"
    return r


def kkt_certificate(ev, z, lam_g, lam_x):
    """Largest violation of the first-order optimality conditions of the NLP at (z, lam_g, lam_x), IPOPT's
    convention grad f + J^T lam_g + lam_x = 0 with lam > 0 on an active upper bound: (feasibility, stationarity,
    complementarity), each scaled like IPOPT's error measure."""
    lbx, ubx, lbg, ubg = ev.bounds
    g, J = ev.g(z), ev.jac_g(z)
    feas = max(np.max(np.maximum(lbx - z, 0)), np.max(np.maximum(z - ubx, 0)), np.max(np.maximum(lbg - g, 0)),
               np.max(np.maximum(g - ubg, 0)))
    sd = max(100.0, (np.abs(lam_g).sum() + np.abs(lam_x).sum()) / (lam_g.size + lam_x.size)) / 100.0
    stat = np.max(np.abs(ev.grad_f(z) + J.T @ lam_g + lam_x)) / sd

    def comp(lam, v, lo, hi):
        free = lo < hi  # a fixed variable / equality row takes any multiplier
        with np.errstate(invalid="ignore"):
            return _comp(lam, v, lo, hi, free)

    def _comp(lam, v, lo, hi, free):
        up = np.where(np.isfinite(hi), np.maximum(lam, 0) * (hi - v), np.maximum(lam, 0) * 1e300)
        dn = np.where(np.isfinite(lo), np.maximum(-lam, 0) * (v - lo), np.maximum(-lam, 0) * 1e300)
        return float(np.max(np.where(free, np.maximum(up, dn), 0.0), initial=0.0))

    return float(feas), float(stat), max(comp(lam_x, z, lbx, ubx), comp(lam_g, g, lbg, ubg)) / sd


def check_delta3(kind):
    """Certify the stored KKT point with the evaluators under test and compare ITS objective with the notebook's."""
    import os

    from mpopt_b200.problems import delta3_launch_vehicle

    problem, K, p, scheme, ref, _, rtol, _ = DELTA3
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "delta3_drag_kkt.npz"))
    ev = Evaluators(kind, delta3_launch_vehicle(1.0), K, p, scheme)
    feas, stat, comp = kkt_certificate(ev, d["z"], d["lam_g"], d["lam_x"])
    assert feas <= 1e-9 and stat <= 1e-8 and comp <= 1e-8, f"Delta III: not a KKT point of this NLP: {(feas, stat, comp)}"
    # ... and not of a perturbed one: the certificate is sharp
    assert kkt_certificate(ev, d["z"] * (1 + 1e-6), d["lam_g"], d["lam_x"])[0] > 1e-8
    f = ev.f(d["z"])
    gap = abs(f - ref) / abs(ref)
    assert gap <= rtol, f"Delta III: optimum {f!r} vs stored {ref!r}: relative gap {gap:.2e}"
    return f
