"""Per-warp timeline of mpx_hess_kernel at the headline NLP (MPX_TRACE=1: lane 0 of every warp stamps the SM cycle
counter after its loads, the node functor, the entry stores, the staged row stores, the ticket and the final step, plus
the global nanosecond timer at entry / exit).  Summarised in profiles/r02/hess_timeline.md."""
import ctypes as C, os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["MPX_TRACE"] = "1"
from mpopt_b200 import _lib
from mpopt_b200.nlp import Transcription
from mpopt_b200.problems import synthetic_6_3
tr = Transcription(synthetic_6_3(), 4096, 15, "LGR")
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sp = stream.cuda_stream
rng = np.random.default_rng(0)
z = rng.uniform(-1, 1, tr.n_z); z[-2:] = [0.0, 1.0]
zd = torch.from_numpy(z).to(dev); pd = torch.from_numpy(rng.dirichlet(np.ones(4096))).to(dev)
lam = torch.from_numpy(rng.uniform(-1, 1, tr.n_g)).to(dev)
nh = len(tr.hess_structure()[1]); hv = torch.empty(nh, dtype=torch.float64, device=dev)
L = _lib.lib()
for _ in range(20):
    _lib.check(L.mpx_eval_hess_l_dev(tr._plan, zd.data_ptr(), pd.data_ptr(), C.c_double(0.7), lam.data_ptr(), hv.data_ptr(), sp))
torch.cuda.synchronize()
nw, sl = C.c_int64(), C.c_int64()
L.mpx_trace_read.restype = C.c_int
L.mpx_trace_read(tr._plan, C.byref(nw), C.byref(sl), None)
buf = np.zeros(nw.value * sl.value, dtype=np.uint64)
L.mpx_trace_read(tr._plan, C.byref(nw), C.byref(sl), buf.ctypes.data_as(C.POINTER(C.c_ulong)))
W = (tr.N + 31) // 32
T = buf.reshape(-1, sl.value)[:W].astype(np.int64)
g0 = T[:, 0].min()
ghz = 1.965
start = (T[:, 0] - g0) / 1e3
def q(a): return [round(float(x), 2) for x in np.percentile(a, [0, 10, 50, 90, 100])]
print("warps", W, "kernel span us (globaltimer)", (T[:, 8].max() - g0) / 1e3)
print("warp start us          ", q(start))
print("loads  (1->2) us             ", q((T[:, 2] - T[:, 1]) / ghz / 1e3))
print("hess_node (2->3) us          ", q((T[:, 3] - T[:, 2]) / ghz / 1e3))
print("corner partials, ticket (3->4)", q((T[:, 4] - T[:, 3]) / ghz / 1e3))
print("entry stores (4->5)          ", q((T[:, 5] - T[:, 4]) / ghz / 1e3))
print("row stores (5->6)            ", q((T[:, 6] - T[:, 5]) / ghz / 1e3))
print("terminal entries (6->7)      ", q((T[:, 7] - T[:, 6]) / ghz / 1e3))
print("warp total (1->7) us   ", q((T[:, 7] - T[:, 1]) / ghz / 1e3))
print("warp end us (global)   ", q((T[:, 8] - g0) / 1e3))
end = (T[:, 8] - g0) / 1e3
order = np.argsort(-end)[:12]
print("slowest warps: warp, cta, start, loads, node, ticket, entries, rows, term, end")
for w in order:
    t = T[w]
    print(w, t[9], round(start[w], 2), *[round((t[b] - t[a]) / ghz / 1e3, 2) for a, b in ((1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7))], round(end[w], 2))
byc = {}
for w in range(W):
    byc.setdefault(int(T[w, 9]), []).append(end[w])
ce = np.array([max(v) for v in byc.values()])
print("CTA end us percentiles", q(ce))
