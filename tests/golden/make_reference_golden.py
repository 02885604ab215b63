"""Golden vectors produced by RUNNING THE REFERENCE ITSELF (tests/golden/ref_*.npz).

    python tests/golden/make_reference_golden.py [--all] [case ...]

Needs /root/reference (read-only, not present on the GPU box -- hence committed fixtures).  The reference module
/root/reference/mpopt/mpopt.py is imported UNMODIFIED through ``oracle/refrun`` (a stand-in for the slice of CasADi's API
it calls: the reference's own ``OCP / Collocation / mpopt / mpopt_adaptive`` classes build the NLP ``{f, x, g, p}`` with
their own formulas, index logic, orderings, bounds and initial guess; the stand-in only carries out the arithmetic and the
derivatives).  For every case the file holds, at a seeded point ``(z, p)``:

    f, g, grad_f, the CSR Jacobian (rowptr, colind, values) with the pattern the reference's graph has (exact-zero table
    entries folded away as SX does), the variable / constraint bounds, the initial guess, and the lower triangle of the
    Hessian of ``lam_f f + lam_g . g`` (what ``ca.nlpsol`` differentiates, mpopt.py:757).

The cases of ``make_golden.py`` reuse that file's points, so the two sets of fixtures can be compared entry by entry.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.dirname(HERE), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def _moon(**flags):
    def make():
        from mpopt_b200.problems import moon_lander

        ocp = moon_lander()
        for k, v in flags.items():
            getattr(ocp, k)[0] = v
        return ocp

    return make


def _registry(name):
    def make():
        from mpopt_b200.problems import EXAMPLES, REGISTRY

        return {**REGISTRY, **EXAMPLES}[name]()

    return make


#: name -> (OCP factory, n_segments, poly_orders, scheme, dirichlet widths, adaptive)
EXTRA = {
    # BASELINE.json's degrees at sizes the expression engine handles in seconds
    "syn63_K4_p15_LGR": (_registry("synthetic_6_3"), 4, 15, "LGR", True, False),       # headline dynamics, headline degree
    "syn63_K2_p20_LGL": (_registry("synthetic_6_3"), 2, 20, "LGL", False, False),      # config 4's degree and scheme
    "moon_K4_p15_LGR": (_registry("moon_lander"), 4, 15, "LGR", True, False),          # config 2's degree
    "vdp_K6_mixed_CGL": (_registry("van_der_pol"), 6, [3, 30, 3, 3, 30, 3], "CGL", True, False),  # config 3
    "schwartz_K5_p10_LGR": (_registry("two_phase_schwartz"), 5, 10, "LGR", True, False),  # config 5: 2 phases + links, path rows
    # optional row blocks: control slope (DU), slope continuity (dU); mid-point control rows (mU, on by default) off
    "moon_diffu_K6_p4_LGL": (_moon(diff_u=1, du_continuity=1), 6, 4, "LGL", True, False),
    "moon_nomidu_K5_p3_LGR": (_moon(midu=0), 5, 3, "LGR", True, False),
    "moon_all_K3_mixed_CGL": (_moon(diff_u=1, du_continuity=1), 3, [4, 2, 5], "CGL", True, False),
    # remaining problems of the reference's tests / examples
    "hyper_K15_p15_LGR": (_registry("hyper_sensitive"), 15, 15, "LGR", False, False),  # scale_t = 1e-3
    "hyper_K5_p50_LGR": (_registry("hyper_sensitive"), 5, 50, "LGR", False, False),    # hypersensitive.ipynb:165-170
    "vdp_K1_p25_LGL": (_registry("van_der_pol"), 1, 25, "LGL", False, False),          # vanderpol.ipynb
    "chachuat_K3_p6_CGL": (_registry("chachuat_3_10"), 3, 6, "CGL", True, False),
    "generic2_K3_mixed_CGL": (_registry("generic_two_phase"), 3, [2, 5, 3], "CGL", True, False),
    "robot_K8_p4_LGR": (_registry("robot_arm"), 8, 4, "LGR", False, False),
    "sink_K7_mixed_LGL": (_registry("kitchen_sink"), 7, [3, 4, 6, 2, 5, 4, 1], "LGL", True, False),
    "sink_K4_p1_LGR": (_registry("kitchen_sink"), 4, 1, "LGR", False, False),
    "moon_K3_p40_mixed_LGL": (_registry("moon_lander"), 3, [40, 6, 33], "LGL", True, False),   # degrees above 31
    "delta3_K4_mixed_LGL": (_registry("delta3_launch_vehicle"), 4, [4, 6, 5, 3], "LGL", True, False),
    "falcon9_K3_mixed_CGL": (_registry("falcon9_launcher"), 3, [7, 2, 5], "CGL", True, False),
    "alp_rider_K6_p5_LGR": (_registry("alp_rider"), 6, 5, "LGR", True, False),         # tests/test_examples.py
    "mine_opt_K4_p6_LGL": (_registry("mine_opt"), 4, 6, "LGL", False, False),
    "dae_vdp_K5_p4_CGL": (_registry("dae_van_der_pol"), 5, 4, "CGL", True, False),
    # mpopt_adaptive (widths as variables)
    "adaptive_hyper_K5_p15_LGR": (_registry("hyper_sensitive"), 5, 15, "LGR", False, True),    # tests/test_mpopt.py:268-277
    "adaptive_syn63_K3_p5_LGL": (_registry("synthetic_6_3"), 3, 5, "LGL", False, True),
    "adaptive_schwartz_K3_p4_LGR": (_registry("two_phase_schwartz"), 3, 4, "LGR", False, True),
}


def cases():
    """name -> (factory, K, poly_orders, scheme, adaptive, point) with ``point`` = None (seeded here) or the npz that
    already fixes ``(z, p)``."""
    import make_golden as M

    out = {}
    for name, case in M.CASES.items():
        out[name] = (_registry(case[0]), case[1], case[2], case[3], len(case) > 4, os.path.join(HERE, name + ".npz"), False)
    for name, (fac, K, po, scheme, dirichlet, adaptive) in EXTRA.items():
        out[name] = (fac, K, po, scheme, adaptive, None, dirichlet)
    return out


def point(name, fac, K, po, scheme, adaptive, npz, dirichlet):
    """Seeded evaluation point (the oracle is used for the LAYOUT of z only: which columns are times / parameters / widths)."""
    if npz is not None:
        G = np.load(npz)
        return G["z"], G["p"]
    from helpers import random_point
    from oracle.adaptive import OracleAdaptiveNLP
    from oracle.nlp import OracleNLP

    ora = (OracleAdaptiveNLP if adaptive else OracleNLP)(fac(), K, po, scheme)
    z, p = random_point(ora, dirichlet=dirichlet)
    if name.startswith("robot"):
        z = np.abs(z) + 0.5
    if name.startswith(("delta3", "falcon9")):
        z = ora.initialize_solution() * (1.0 + 0.01 * np.random.default_rng(7).standard_normal(ora.n_z))
    if adaptive:
        rng = np.random.default_rng(12)
        for ph in range(ora.P):
            z[ora.colW(ph, np.arange(K))] = rng.dirichlet(np.ones(K)) * 0.8 + 0.2 / K
            z[ora.colT0(ph)] = 0.3 + 0.25 * ph
        p = np.zeros(0)
    return z, p


def run(name, case, ref=None):
    """Evaluate one case with the reference; dict of arrays as stored in ref_<name>.npz."""
    from oracle.refrun import run_reference as rr

    ref = ref or rr.load_reference()
    fac, K, po, scheme, adaptive, npz, dirichlet = case
    z, p = point(name, fac, K, po, scheme, adaptive, npz, dirichlet)
    R = rr.ReferenceNLP(ref, rr.reference_ocp(ref, fac), K, po, scheme, adaptive)
    f, g, grad, J = R.evaluate(z, p)
    lam_g = np.random.default_rng(21).uniform(-1, 1, R.n_g)
    H = R.hess_l(z, p, 0.75, lam_g)
    H.sort_indices()
    zmin, zmax, gmin, gmax = R.all_bounds()
    return dict(z=z, p=p, f=f, g=g, grad=grad, rowptr=J.indptr.astype(np.int64), colind=J.indices.astype(np.int64),
                values=J.data, zmin=zmin, zmax=zmax, gmin=gmin, gmax=gmax, z0=R.initial_guess(), lam_f=0.75, lam_g=lam_g,
                hrowptr=H.indptr.astype(np.int64), hcolind=H.indices.astype(np.int64), hvalues=H.data)


def main():
    from oracle.refrun import run_reference as rr

    ref = rr.load_reference()
    want = [a for a in sys.argv[1:] if not a.startswith("--")]
    for name, case in cases().items():
        path = os.path.join(HERE, "ref_" + name + ".npz")
        if want and name not in want:
            continue
        if os.path.exists(path) and "--all" not in sys.argv and not want:
            continue
        out = run(name, case, ref)
        np.savez_compressed(path, **out)
        print("ref_" + name, "n_z", out["z"].size, "n_g", out["g"].size, "nnz", out["values"].size, "nnz_H", out["hvalues"].size)


if __name__ == "__main__":
    main()
