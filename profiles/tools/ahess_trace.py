"""Hessian of the widths-as-variables NLP at the headline size: device time (CUDA-graph replay) and, with MPX_TRACE=1,
the per-segment timeline of mpx_adapt_hess_kernel (thread 0 of every CTA stamps the SM cycle counter at the phase
boundaries, the global nanosecond timer at entry / exit and the SM id)."""
import ctypes as C, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _timing import time_call
from mpopt_b200 import _lib
from mpopt_b200.nlp import Transcription
from mpopt_b200.problems import synthetic_6_3
K = int(os.environ.get("AH_K", 4096))
tr = Transcription(synthetic_6_3(), K, 15, "LGR", adaptive=True)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sp = stream.cuda_stream
rng = np.random.default_rng(0)
z = rng.uniform(-1, 1, tr.n_z)
L = tr.layout
z[L.colT0(0)], z[L.colTF(0)] = 0.0, 1.0
z[L.colW(0, 0): L.colW(0, 0) + K] = rng.dirichlet(np.ones(K))
zd = torch.from_numpy(z).to(dev)
nh = len(tr.hess_structure()[1])
hv = torch.empty(nh, dtype=torch.float64, device=dev)
lam = torch.from_numpy(rng.uniform(-1, 1, tr.n_g)).to(dev)
lib = _lib.lib()
fn = lambda: _lib.check(lib.mpx_eval_hess_l_dev(tr._plan, zd.data_ptr(), None, C.c_double(0.7), lam.data_ptr(), hv.data_ptr(), sp))
if os.environ.get("AH_ONCE"):  # under ncu: a few plain calls
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    sys.exit(0)
host_us, graph_us = time_call(fn, stream, n=20, warm=4)
nbytes = 8 * (tr.n_z + tr.n_g + nh)
print(json.dumps({"evaluator": "hess_l of the adaptive NLP", "K": K, "us": round(graph_us or host_us, 2), "host_issued_us": round(host_us, 2),
                  "nnz_hess_lower": nh, "algorithmic_MB": round(nbytes / 1e6, 1), "GBs": round(nbytes / (graph_us or host_us) / 1e3, 1)}))
after_graph = hv.clone()
fn(); torch.cuda.synchronize()
print(json.dumps({"graph replays leave the same bits as a plain call": bool(torch.equal(after_graph, hv)),
                  "zero_fill": tr.hess_zero_fill}))
if os.environ.get("MPX_TRACE") != "1":
    sys.exit(0)
fn(); torch.cuda.synchronize()
nw, sl = C.c_int64(), C.c_int64()
lib.mpx_trace_read(tr._plan, C.byref(nw), C.byref(sl), None)
buf = np.zeros(nw.value * sl.value, dtype=np.uint64)
lib.mpx_trace_read(tr._plan, C.byref(nw), C.byref(sl), buf.ctypes.data_as(C.POINTER(C.c_ulong)))
base = tr.N // 32 + 1
T = buf.reshape(-1, sl.value)[base: base + K].astype(np.int64)
ghz = 1.965
q = lambda a: [round(float(x), 2) for x in np.percentile(a, [0, 10, 50, 90, 100])]
us = lambda a, b: (T[:, b] - T[:, a]) / ghz / 1e3
print("all segments: span us", (T[:, 10].max() - T[:, 0].min()) / 1e3, " odd segments start (global us)", q((T[1::2, 0] - T[:, 0].min()) / 1e3))
for par in (0, 1):
    P = T[par::2]
    g0 = P[:, 0].min()
    print(f"parity {par}: launch span us", (P[:, 10].max() - g0) / 1e3, "CTAs", len(P), " CTA start (global us)", q((P[:, 0] - g0) / 1e3))
print("percentiles 0/10/50/90/100 over all segments, us:")
print("tables + nodes -> shared (1->2)      ", q(us(1, 2)))
print("interpolation (2->3)                 ", q(us(2, 3)))
print("functors: mid points + nodes (3->4)  ", q(us(3, 4)))
print("positions arrived + barrier (4->5)   ", q(us(4, 5)))
print("(a) rows w_k / T0 / TF (5->6)        ", q(us(5, 6)))
print("(b) scalars (6->7)                   ", q(us(6, 7)))
print("(c) blocks, warp 0's share (7->8)    ", q(us(7, 8)))
print("diagonal pass (8->9)                 ", q(us(8, 9)))
print("segment total (1->9)                 ", q(us(1, 9)))
cta = T[:, 14]
gaps = []
for c in np.unique(cta):
    R = T[cta == c]
    R = R[np.argsort(R[:, 0])]
    gaps.extend(((R[1:, 0] - R[:-1, 10]) / 1e3).tolist())
if gaps:
    print("gap between a CTA's segments, global us (end stamp -> next start stamp)", q(np.array(gaps)), " segments per CTA", q(np.bincount(cta.astype(int))))
print("end of work -> index stored (10->11), end barrier (11->13) us", q((T[:, 11] - T[:, 10]) / 1e3), q((T[:, 13] - T[:, 11]) / 1e3))
print("loop top -> start stamp (12->0): even, odd us", q((T[0::2, 0] - T[0::2, 12]) / 1e3), q((T[1::2, 0] - T[1::2, 12]) / 1e3))
print("segment, global timer (0->10) us     ", q((T[:, 10] - T[:, 0]) / 1e3))
sm = T[:, 15]
print("segments per SM", q(np.bincount(sm.astype(int))[np.bincount(sm.astype(int)) > 0]))
