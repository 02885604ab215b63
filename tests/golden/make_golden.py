"""Generate the committed golden vectors from the CPU oracle (NOT from a run of the reference: CasADi/IPOPT are
not installable here -- SURVEY.md 8c).  They freeze the oracle's answers so that (i) an accidental change of the
oracle is caught on CPU and (ii) the GPU tests have fixed files to compare with.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

CASES = {
    "moon_lander_K20_p3_LGR": ("moon_lander", 20, 3, "LGR"),                # BASELINE config 1
    "van_der_pol_K9_mixed_CGL": ("van_der_pol", 9, [3, 30, 3] * 3, "CGL"),  # BASELINE config 3 in small
    "kitchen_sink_K4_LGR": ("kitchen_sink", 4, [3, 2, 4, 3], "LGR"),        # 2 phases, params, time, scaling
    "synthetic63_K3_p5_LGR": ("synthetic_6_3", 3, 5, "LGR"),                # headline dynamics in small
    # the two launch-vehicle NLPs at the sizes of their stored IPOPT banners (multi_stage_launch_vehicle_ascent.ipynb:466-471,
    # falcon9_to_orbit.ipynb:480-485), evaluated near the examples' initial guesses
    "delta3_K1_p11_LGR": ("delta3_launch_vehicle", 1, 11, "LGR"),
    "falcon9_K5_p6_LGR": ("falcon9_launcher", 5, 6, "LGR"),
    # widths-as-variables NLP of mpopt_adaptive; a 5th entry marks the adaptive transcription
    "adaptive_moon_K3_p3_LGR": ("moon_lander", 3, 3, "LGR", "adaptive"),    # tests/test_mpopt.py:258-259
    "adaptive_sink_K3_LGR": ("kitchen_sink", 3, [3, 2, 4], "LGR", "adaptive"),  # time-dependent: dense width columns
}


def build(case):
    """(oracle, z, p) of one case."""
    from helpers import random_point
    from mpopt_b200.problems import REGISTRY
    from oracle.adaptive import OracleAdaptiveNLP
    from oracle.nlp import OracleNLP

    problem, K, po, scheme = case[:4]
    adaptive = len(case) > 4
    ora = (OracleAdaptiveNLP if adaptive else OracleNLP)(REGISTRY[problem](), K, po, scheme, drop_exact_zeros=False)
    z, p = random_point(ora, dirichlet=True)
    if problem in ("delta3_launch_vehicle", "falcon9_launcher"):
        z = ora.initialize_solution() * (1.0 + 0.01 * np.random.default_rng(7).standard_normal(ora.n_z))
    if adaptive:
        rng = np.random.default_rng(12)
        for ph in range(ora.P):
            z[ora.colW(ph, np.arange(K))] = rng.dirichlet(np.ones(K)) * 0.8 + 0.2 / K
            z[ora.colT0(ph)] = 0.3 + 0.25 * ph
        p = np.zeros(0)
    return ora, z, p


def main():
    for name, case in CASES.items():
        if os.path.exists(os.path.join(HERE, name + ".npz")) and "--all" not in sys.argv:
            continue  # committed vectors stay byte-identical unless asked
        ora, z, p = build(case)
        f, g, grad, J = ora._eval(z, p)
        zmin, zmax, gmin, gmax = ora.bounds()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), z=z, p=p, f=f, g=g, grad=grad, rowptr=J.indptr.astype(np.int64),
                            colind=J.indices.astype(np.int64), values=J.data, zmin=zmin, zmax=zmax, gmin=gmin, gmax=gmax,
                            z0=ora.initialize_solution())
        print(name, "n_z", ora.n_z, "n_g", ora.n_g, "nnz", J.nnz)
    # Hessian of the Lagrangian (nlp_hess_l): separate files, so the vectors above stay byte-identical
    from oracle.hessian import hess_l

    for name, case in CASES.items():
        path = os.path.join(HERE, name + "_hess.npz")
        if os.path.exists(path) and "--all" not in sys.argv:
            continue
        ora, z, p = build(case)
        lam = np.random.default_rng(21).uniform(-1, 1, ora.n_g)
        try:
            H = hess_l(ora, z, p, 0.75, lam)
        except NotImplementedError:  # adaptive NLP with explicit time dependence
            continue
        np.savez_compressed(path, z=z, p=p, lam_f=0.75, lam_g=lam, rowptr=H.indptr.astype(np.int64),
                            colind=H.indices.astype(np.int64), values=H.data)
        print(name + "_hess", "nnz", H.nnz)


if __name__ == "__main__":
    main()
