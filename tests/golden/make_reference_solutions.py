"""Optimal objectives of the reference's own ``solve()`` (tests/golden/ref_solutions.json), from the UNMODIFIED reference
module run on the stand-ins of oracle/refrun (CasADi slice + interior-point ``nlpsol``).

    python tests/golden/make_reference_solutions.py

BASELINE.json's config 1 (moon-lander, 20 segments of degree 3, LGR: the reference's mp.solve path) and the other
problems of the reference's test-suite, with the fixed-width and the widths-as-variables driver.  The GPU test
(tests/test_reference_golden.py::test_mp_solve_matches_reference_solve) solves the same problems through
``mpopt_b200.mp`` and compares the optimum.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.dirname(HERE), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

#: (driver class, problem, n_segments, poly_orders, scheme)
CASES = [
    ("mpopt", "moon_lander", 20, 3, "LGR"),          # BASELINE.json configs[0]
    ("mpopt", "moon_lander", 5, 6, "LGL"),
    ("mpopt", "van_der_pol", 3, 8, "CGL"),
    ("mpopt", "hyper_sensitive", 10, 8, "LGR"),
    ("mpopt", "two_phase_schwartz", 2, 8, "LGR"),
    ("mpopt_adaptive", "moon_lander", 3, 3, "LGR"),   # tests/test_mpopt.py:258-259
    ("mpopt_adaptive", "van_der_pol", 3, 5, "LGR"),
]


def main():
    from mpopt_b200.problems import REGISTRY
    from oracle.refrun import run_reference as rr

    ref = rr.load_reference()
    out = []
    for cls, problem, K, p, scheme in CASES:
        mpo = getattr(ref, cls)(rr.reference_ocp(ref, REGISTRY[problem]), K, p, scheme)
        mpo._MUTE_ = True
        if cls == "mpopt_adaptive":
            import builtins, contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                sol = mpo.solve(nlp_solver_options={"ipopt.tol": 1e-10})
        else:
            sol = mpo.solve(nlp_solver_options={"ipopt.tol": 1e-10})
        ok = mpo.nlp_solver.stats()["success"]
        rec = dict(cls=cls, problem=problem, n_segments=K, poly_orders=p, scheme=scheme, f=float(sol["f"]), success=bool(ok),
                   iterations=int(mpo.nlp_solver.stats()["iter_count"]))
        if cls == "mpopt_adaptive":
            rec["widths"] = [float(v) for v in mpo._nlp_sw_params]
        print(rec)
        out.append(rec)
    with open(os.path.join(HERE, "ref_solutions.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
