"""Device-resident g + jac_g timing of every BASELINE.json configuration (same method as bench.py's timed region:
back-to-back launches, outputs rotating over buffer sets larger than L2).  Not the bench: context for profiles/README.md."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mpopt_b200.nlp import Transcription  # noqa: E402
from mpopt_b200.problems import REGISTRY  # noqa: E402

CONFIGS = [
    ("config 2: moon-lander K=4096 p=15 LGR", "moon_lander", 4096, 15, "LGR"),
    ("headline: synthetic 6/3 K=4096 p=15 LGR", "synthetic_6_3", 4096, 15, "LGR"),
    ("config 3: van-der-Pol K=2048 p=[3,30,3..] CGL", "van_der_pol", 2048, [30 if k % 3 == 1 else 3 for k in range(2048)], "CGL"),
    ("config 4: synthetic 6/3 K=8192 p=20 LGL", "synthetic_6_3", 8192, 20, "LGL"),
    ("config 5: two-phase Schwartz K=1024/phase p=10 LGR", "two_phase_schwartz", 1024, 10, "LGR"),
]
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
for name, prob, K, po, scheme in CONFIGS:
    tr = Transcription(REGISTRY[prob](), K, po, scheme)
    B = 8 * (tr.n_z + tr.n_p + tr.n_g + tr.nnz)
    R = max(2, int(np.ceil(300e6 / B)))
    rng = np.random.default_rng(1)
    z = rng.uniform(-1, 1, tr.n_z)
    nvar = tr.n_z // tr.P
    for ph in range(tr.P):
        z[(ph + 1) * nvar - 2 - tr.na], z[(ph + 1) * nvar - 1 - tr.na] = 0.2 * ph, 2.0 + ph
    p = np.concatenate([rng.dirichlet(np.ones(K)) for _ in range(tr.P)])
    zd = [torch.from_numpy(z + 1e-3 * i).to(dev) for i in range(R)]
    pd = torch.from_numpy(p).to(dev)
    gd = [torch.empty(tr.n_g, dtype=torch.float64, device=dev) for _ in range(R)]
    vd = [torch.empty(tr.nnz, dtype=torch.float64, device=dev) for _ in range(R)]
    step = lambda i: tr.g_jac_dev(zd[i % R].data_ptr(), pd.data_ptr(), gd[i % R].data_ptr(), vd[i % R].data_ptr(), stream.cuda_stream)
    for i in range(10):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record(stream)
    for i in range(n):
        step(i)
    e1.record(stream)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(json.dumps({"config": name, "program": tr.program_origin, "n_z": tr.n_z, "nnz": tr.nnz, "MB": round(B / 1e6, 2),
                      "us_per_eval": round(us, 2), "evals_per_s": round(1e6 / us), "GBs": round(B / us / 1e3, 1),
                      "frac_of_measured_peak": round(B / us / 1e3 / peak, 3), "launches_per_eval": tr.launches // (n + 10)}))
    del tr, zd, gd, vd
