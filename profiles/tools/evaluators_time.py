"""Device time of every evaluator of the headline NLP (synthetic 6/3, K=4096, p=15, LGR), CUDA events around 100
back-to-back calls with device-resident inputs.  Context for profiles/README.md; the bench metric is g + jac_g only."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from mpopt_b200 import _lib  # noqa: E402
from mpopt_b200.nlp import Transcription  # noqa: E402
from mpopt_b200.problems import synthetic_6_3  # noqa: E402

tr = Transcription(synthetic_6_3(), 4096, 15, "LGR")
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sp = stream.cuda_stream
rng = np.random.default_rng(0)
z = rng.uniform(-1, 1, tr.n_z)
z[-2:] = [0.0, 1.0]
zd = torch.from_numpy(z).to(dev)
pd = torch.from_numpy(rng.dirichlet(np.ones(4096))).to(dev)
lam = torch.from_numpy(rng.uniform(-1, 1, tr.n_g)).to(dev)
g = torch.empty(tr.n_g, dtype=torch.float64, device=dev)
v = torch.empty(tr.nnz, dtype=torch.float64, device=dev)
grad = torch.empty(tr.n_z, dtype=torch.float64, device=dev)
f = torch.empty(1, dtype=torch.float64, device=dev)
nh = len(tr.hess_structure()[1])
hv = torch.empty(nh, dtype=torch.float64, device=dev)
L = _lib.lib()
calls = {
    "g + jac_g (mpx_eval_g_jac_dev)": (lambda: tr.g_jac_dev(zd.data_ptr(), pd.data_ptr(), g.data_ptr(), v.data_ptr(), sp), 8 * (tr.n_z + tr.n_p + tr.n_g + tr.nnz)),
    "g only": (lambda: tr.g_jac_dev(zd.data_ptr(), pd.data_ptr(), g.data_ptr(), None, sp), 8 * (tr.n_z + tr.n_p + tr.n_g)),
    "f + grad_f (mpx_eval_f_grad_dev)": (lambda: tr.f_grad_dev(zd.data_ptr(), pd.data_ptr(), f.data_ptr(), grad.data_ptr(), sp), 8 * (2 * tr.n_z + tr.n_p)),
    "hess_l (mpx_eval_hess_l_dev)": (lambda: _lib.check(L.mpx_eval_hess_l_dev(tr._plan, zd.data_ptr(), pd.data_ptr(), C.c_double(0.7), lam.data_ptr(), hv.data_ptr(), sp)), 8 * (tr.n_z + tr.n_p + tr.n_g + nh)),  # z, p, multipliers read, every entry written once
}
from _timing import time_call  # noqa: E402

for name, (fn, nbytes) in calls.items():
    host_us, graph_us = time_call(fn, stream, n=50)
    us = graph_us if graph_us is not None else host_us
    print(json.dumps({"evaluator": name, "us": round(us, 2), "host_issued_us": round(host_us, 2), "algorithmic_MB": round(nbytes / 1e6, 2),
                      "GBs": round(nbytes / us / 1e3, 1), "timing": "CUDA-graph replay of 50 calls" if graph_us is not None else "host-issued"}))
print(json.dumps({"nnz_jac": tr.nnz, "nnz_hess_lower": nh}))

# the inner step of every h-adaptive pass: dynamics residual at the mid points of all segments (host-pointer entry
# point: upload of z, kernel, download of six arrays), wall clock
import time

zh = z.copy()
ph = rng.dirichlet(np.ones(4096))
taus = [0.5 * (tr.tables(15)[0][:-1] + tr.tables(15)[0][1:])] * 4096
for _ in range(3):
    out = tr.residuals(zh, ph, 0, taus)
t0 = time.perf_counter()
for _ in range(20):
    out = tr.residuals(zh, ph, 0, taus)
ms = (time.perf_counter() - t0) / 20 * 1e3
print(json.dumps({"evaluator": "dynamics residuals at 61 440 mid points (mpx_eval_residuals, host pointers, wall clock, "
                               "per-segment lists as the reference passes them)",
                  "ms": round(ms, 3), "points": int(sum(out["counts"]))}))
taus2 = np.ascontiguousarray(np.stack(taus))  # [K, 15]: packed without a Python loop over the segments
for _ in range(3):
    out = tr.residuals(zh, ph, 0, taus2)
t0 = time.perf_counter()
for _ in range(20):
    out = tr.residuals(zh, ph, 0, taus2)
ms2 = (time.perf_counter() - t0) / 20 * 1e3
print(json.dumps({"evaluator": "the same with taus as one [K, 15] array", "ms": round(ms2, 3), "points": int(sum(out["counts"]))}))
