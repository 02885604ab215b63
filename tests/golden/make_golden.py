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
}


def main():
    from helpers import random_point
    from mpopt_b200.problems import REGISTRY
    from oracle.nlp import OracleNLP

    for name, (problem, K, po, scheme) in CASES.items():
        ora = OracleNLP(REGISTRY[problem](), K, po, scheme, drop_exact_zeros=False)
        z, p = random_point(ora, dirichlet=True)
        f, g, grad, J = ora._eval(z, p)
        zmin, zmax, gmin, gmax = ora.bounds()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), z=z, p=p, f=f, g=g, grad=grad, rowptr=J.indptr.astype(np.int64),
                            colind=J.indices.astype(np.int64), values=J.data, zmin=zmin, zmax=zmax, gmin=gmin, gmax=gmax,
                            z0=ora.initialize_solution())
        print(name, "n_z", ora.n_z, "n_g", ora.n_g, "nnz", J.nnz)


if __name__ == "__main__":
    main()
