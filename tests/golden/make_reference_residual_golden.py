"""Reference-generated fixtures of the interpolation / residual path: tests/golden/refres_*.npz.

    python tests/golden/make_reference_residual_golden.py [--all]

The reference's own ``get_residual_grid_taus`` (mpopt.py:1152-1203), ``interpolate_single_phase`` (:1489-1542),
``get_dynamics_residuals_single_phase`` (:1428-1487), ``get_state_second_derivative_single_phase`` (:1285-1358) and
``compute_states_from_solution_dynamics`` (:989-1076) are run on the unmodified reference module (oracle/refrun) at a
seeded ``(z, p)`` for the three grid types; the fixture keeps the target points and what came back, per phase, rows =
points segment by segment.  Uniform degrees only: with mixed degrees the reference's "mid-points" / "spectral" grids
build a ragged ``np.array`` (:1188), which numpy >= 1.24 refuses.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.dirname(HERE), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

#: name -> (problem, n_segments, degree, scheme, dirichlet widths)
CASES = {
    "vdp_K3_p4_LGR": ("van_der_pol", 3, 4, "LGR", True),
    "sink_K3_p5_LGL": ("kitchen_sink", 3, 5, "LGL", True),       # 2 phases, parameters, explicit time, scaling
    "syn63_K3_p5_CGL": ("synthetic_6_3", 3, 5, "CGL", False),
    "hyper_K4_p7_LGR": ("hyper_sensitive", 4, 7, "LGR", True),   # scale_t
}
GRIDS = ("fixed", "mid-points", "spectral")


def _rows(lst):
    out = [np.atleast_2d(np.asarray(a, float)) for a in lst if a is not None and len(a)]
    return np.concatenate(out) if out else np.zeros((0, 0))


def point(problem, K, deg, scheme, dirichlet):
    from helpers import random_point
    from mpopt_b200.problems import REGISTRY
    from oracle.nlp import OracleNLP

    ora = OracleNLP(REGISTRY[problem](), K, deg, scheme)
    return random_point(ora, dirichlet=dirichlet)


def run(name, ref=None):
    import casadi as ca

    from mpopt_b200.problems import REGISTRY
    from oracle.refrun import run_reference as rr

    ref = ref or rr.load_reference()
    problem, K, deg, scheme, dirichlet = CASES[name]
    z, p = point(problem, K, deg, scheme, dirichlet)
    mpo = ref.mpopt(rr.reference_ocp(ref, REGISTRY[problem]), K, deg, scheme)
    mpo._MUTE_ = True
    mpo.create_solver()
    mpo._nlp_sw_params = p  # what solve() sets before it calls the solver (mpopt.py:793-797)
    sol = {"x": ca.DM(z)}
    out = dict(z=z, p=p, n_phases=mpo._ocp.n_phases)
    for ph in range(mpo._ocp.n_phases):
        for grid in GRIDS:
            taus = [np.asarray(t, float) for t in mpo.get_residual_grid_taus(ph, grid)]
            key = f"ph{ph}_{grid}_"
            out[key + "counts"] = np.array([len(t) for t in taus], np.int64)
            out[key + "taus"] = np.concatenate(taus)
            xi, ui, ti, a, Dxi, Dui, _, t0, tf = mpo.interpolate_single_phase(sol, phase=ph, target_nodes=taus, options={})
            out[key + "xi"], out[key + "ui"] = np.array(xi), np.array(ui)
            out[key + "dxi"], out[key + "dui"] = np.array(Dxi), np.array(Dui)
            out[key + "ti"] = np.asarray(ti, float).ravel()
            _, res, _ = mpo.get_dynamics_residuals_single_phase(sol, ph, target_nodes=taus)
            out[key + "res"] = _rows(res)
            _, ddx, ddu = mpo.get_state_second_derivative_single_phase(sol, ph, nodes=taus)
            n = int(out[key + "counts"].sum())
            out[key + "ddxi"], out[key + "ddui"] = _rows(ddx).reshape(n, -1), _rows(ddu).reshape(n, -1)
            xint, _, _, resx = mpo.compute_states_from_solution_dynamics(sol, ph, nodes=taus)
            out[key + "xint"] = _rows(xint)
            out[key + "res_x"] = _rows([np.array(r) for r in resx if r is not None])
    return out


def main():
    from oracle.refrun import run_reference as rr

    ref = rr.load_reference()
    for name in CASES:
        path = os.path.join(HERE, "refres_" + name + ".npz")
        if os.path.exists(path) and "--all" not in sys.argv:
            continue
        out = run(name, ref)
        np.savez_compressed(path, **out)
        print("refres_" + name, {k: v.shape for k, v in out.items() if k.startswith("ph0_fixed")})


if __name__ == "__main__":
    main()
