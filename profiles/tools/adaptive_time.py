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
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _timing import time_call  # noqa: E402
from mpopt_b200 import _lib  # noqa: E402
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
    nh = len(tr.hess_structure()[1])
    hv = torch.empty(nh, dtype=torch.float64, device=dev)
    lam = torch.from_numpy(rng.uniform(-1, 1, tr.n_g)).to(dev)
    import ctypes as C
    Lb = _lib.lib()
    calls = {
        "hess_l": (lambda: _lib.check(Lb.mpx_eval_hess_l_dev(tr._plan, zd.data_ptr(), None, C.c_double(0.7), lam.data_ptr(), hv.data_ptr(), sp)), 8 * (tr.n_z + tr.n_g + nh)),
        "g + jac_g": (lambda: tr.g_jac_dev(zd.data_ptr(), None, g.data_ptr(), v.data_ptr(), sp), 8 * (tr.n_z + tr.n_g + tr.nnz)),
        "g only": (lambda: tr.g_jac_dev(zd.data_ptr(), None, g.data_ptr(), None, sp), 8 * (tr.n_z + tr.n_g)),
        "f + grad_f": (lambda: tr.f_grad_dev(zd.data_ptr(), None, f.data_ptr(), grad.data_ptr(), sp), 8 * 2 * tr.n_z),
    }
    for ev, (fn, nbytes) in calls.items():
        l0 = tr.launches
        host_us, graph_us = time_call(fn, stream, n=20, warm=4)
        lpe = (tr.launches - l0) // (44 if graph_us is not None else 24)
        us = graph_us if graph_us is not None else host_us
        print(json.dumps({"problem": name, "K": K, "p": p, "evaluator": ev, "us": round(us, 2),
                          "host_issued_us": round(host_us, 2), "launches_per_eval": lpe, "n_z": tr.n_z, "n_g": tr.n_g, "nnz": tr.nnz,
                          "algorithmic_MB": round(nbytes / 1e6, 2), "GBs": round(nbytes / us / 1e3, 1)}))
    del tr
