"""Device time of the evaluators of the widths-as-variables NLP (mp.mpopt_adaptive) at the headline size
(synthetic 6/3, K=4096, p=15, LGR) and at the reference's own sizes, CUDA events around back-to-back calls with
device-resident inputs.  Context for profiles/README.md; the bench metric is the fixed-width g + jac_g."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mpopt_b200.nlp import Transcription  # noqa: E402
from mpopt_b200.problems import hyper_sensitive, moon_lander, synthetic_6_3  # noqa: E402

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sp = stream.cuda_stream
for name, make, K, p in (("synthetic_6_3", synthetic_6_3, 4096, 15), ("moon_lander", moon_lander, 4096, 15),
                         ("hyper_sensitive (tests:276-277)", hyper_sensitive, 5, 15)):
    tr = Transcription(make(), K, p, "LGR", adaptive=True)
    rng = np.random.default_rng(0)
    z = rng.uniform(-1, 1, tr.n_z)
    L = tr.layout
    z[L.colT0(0)], z[L.colTF(0)] = 0.0, 1.0
    z[L.colW(0, 0): L.colW(0, 0) + K] = rng.dirichlet(np.ones(K))
    zd = torch.from_numpy(z).to(dev)
    g = torch.empty(tr.n_g, dtype=torch.float64, device=dev)
    v = torch.empty(tr.nnz, dtype=torch.float64, device=dev)
    grad = torch.empty(tr.n_z, dtype=torch.float64, device=dev)
    f = torch.empty(1, dtype=torch.float64, device=dev)
    calls = {
        "g + jac_g": (lambda: tr.g_jac_dev(zd.data_ptr(), None, g.data_ptr(), v.data_ptr(), sp), 8 * (tr.n_z + tr.n_g + tr.nnz)),
        "g only": (lambda: tr.g_jac_dev(zd.data_ptr(), None, g.data_ptr(), None, sp), 8 * (tr.n_z + tr.n_g)),
        "f + grad_f": (lambda: tr.f_grad_dev(zd.data_ptr(), None, f.data_ptr(), grad.data_ptr(), sp), 8 * 2 * tr.n_z),
    }
    for ev, (fn, nbytes) in calls.items():
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = tr.launches
        e0.record(stream)
        for _ in range(50):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 20
        print(json.dumps({"problem": name, "K": K, "p": p, "evaluator": ev, "us": round(us, 2),
                          "launches_per_eval": (tr.launches - l0) // 50, "n_z": tr.n_z, "n_g": tr.n_g, "nnz": tr.nnz,
                          "algorithmic_MB": round(nbytes / 1e6, 2), "GBs": round(nbytes / us / 1e3, 1)}))
    del tr
