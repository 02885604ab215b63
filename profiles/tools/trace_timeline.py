"""Per-warp timeline of the g + jac_g kernel (MPX_TRACE=1): where the microseconds of one launch go.

    MPX_TRACE=1 python profiles/tools/trace_timeline.py [--mode isolated|chain] [--K 4096] [--deg 15]

Stamps per warp (SM cycle counter, lane 0), see mpx_gjac2_kernel: 0 entry, 1 after griddepcontrol.wait, 2 tables landed,
3 constant blocks issued, 4..9 row block s handed to the copy engine, 10 everything issued, 11 images read by the
engine; 13/14 global nanosecond timer at entry / exit, 15 SM id.  `chain` traces the LAST of n back-to-back launches
(steady state, programmatic dependent launch); `isolated` one launch after an L2-evicting read.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
os.environ.setdefault("MPX_TRACE", "1")
import torch  # noqa: E402

from mpopt_b200.nlp import Transcription  # noqa: E402
from mpopt_b200.problems import REGISTRY  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="chain")
ap.add_argument("--K", type=int, default=4096)
ap.add_argument("--deg", type=int, default=15)
ap.add_argument("--scheme", default="LGR")
ap.add_argument("--problem", default="synthetic_6_3")
ap.add_argument("--n", type=int, default=12)
args = ap.parse_args()

dev = torch.device("cuda", 0)
tr = Transcription(REGISTRY[args.problem](), args.K, args.deg, args.scheme, device=0)
rng = np.random.default_rng(1)
z = rng.uniform(-1, 1, tr.n_z)
z[-2:] = [0.0, 1.0]
R = 4
zd = [torch.from_numpy(z + 1e-3 * i).to(dev) for i in range(R)]
pd = torch.from_numpy(np.full(tr.n_p, 1.0 / args.K)).to(dev)
gd = [torch.zeros(tr.n_g, dtype=torch.float64, device=dev) for _ in range(R)]
vd = [torch.zeros(tr.nnz, dtype=torch.float64, device=dev) for _ in range(R)]
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
sp = st.cuda_stream
flush = torch.zeros(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)


def step(i):
    k = i % R
    tr.g_jac_dev(zd[k].data_ptr(), pd.data_ptr(), gd[k].data_ptr(), vd[k].data_ptr(), sp)


for i in range(5):
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if args.mode == "chain":
    tr._L.mpx_gate(sp, 400.0)
    e0.record(st)
    for i in range(args.n):
        step(i)
    e1.record(st)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / args.n
else:
    torch.sum(flush)
    e0.record(st)
    step(0)
    e1.record(st)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3
RING = 8
raw = tr.trace().astype(np.int64)
W = raw.shape[0] // RING
raw = raw.reshape(RING, W, -1)
np.save(os.path.join(os.environ.get("MPX_TRACE_OUT", "."), f"trace_{args.mode}_K{args.K}_d{args.deg}.npy"), raw)
order = np.argsort([r[:, 13][r[:, 13] > 0].min() if (r[:, 13] > 0).any() else 1 << 62 for r in raw])
g0 = min(r[:, 13][r[:, 13] > 0].min() for r in raw if (r[:, 13] > 0).any())
ghz = 1.965
names = ["entry", "pdl_wait done", "tables+inputs", "const issued"] + [f"F{s} out" for s in range(6)] + \
        ["all issued", "images drained"]
out = {"mode": args.mode, "us_per_launch_events": us, "warps_per_launch": int(W), "launches": []}
for li in order:
    T = raw[li]
    T = T[T[:, 13] > 0]
    if not len(T):
        continue
    tg = (T[:, [13]] - g0) + (T[:, :12] - T[:, [0]]) / ghz  # ns on the global axis
    rec = {}
    for i, nm in enumerate(names):
        col = tg[:, i][T[:, i] > 0] / 1e3
        if len(col):
            rec[nm] = [round(float(v), 2) for v in (col.min(), np.percentile(col, 10), np.median(col), np.percentile(col, 90), col.max())]
    ex = (T[:, 14] - g0) / 1e3
    rec["exit"] = [round(float(v), 2) for v in (ex.min(), np.percentile(ex, 10), np.median(ex), np.percentile(ex, 90), ex.max())]
    sm = T[:, 15]
    last = np.array([ex[sm == s_].max() for s_ in np.unique(sm)])
    first = np.array([((T[:, 13] - g0) / 1e3)[sm == s_].min() for s_ in np.unique(sm)])
    rec["per_sm_last_exit"] = [round(float(v), 2) for v in (last.min(), np.percentile(last, 10), np.median(last), np.percentile(last, 90), last.max())]
    rec["per_sm_busy_us"] = [round(float(v), 2) for v in ((last - first).min(), np.median(last - first), (last - first).max())]
    out["launches"].append(rec)
print(json.dumps(out))
