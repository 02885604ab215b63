"""Per-CTA timeline of mpx_adapt_kernel at the headline size (MPX_TRACE=1: thread 0 of every CTA stamps the SM cycle
counter at the phase boundaries, plus the global nanosecond timer at entry / exit and the SM id)."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["MPX_TRACE"] = "1"
from mpopt_b200 import _lib
from mpopt_b200.nlp import Transcription
from mpopt_b200.problems import synthetic_6_3
K = 4096
tr = Transcription(synthetic_6_3(), K, 15, "LGR", adaptive=True)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sp = stream.cuda_stream
rng = np.random.default_rng(0)
z = rng.uniform(-1, 1, tr.n_z)
L = tr.layout
z[L.colT0(0)], z[L.colTF(0)] = 0.0, 1.0
z[L.colW(0, 0): L.colW(0, 0) + K] = rng.dirichlet(np.ones(K))
zd = torch.from_numpy(z).to(dev)
g = torch.empty(tr.n_g, dtype=torch.float64, device=dev)
v = torch.empty(tr.nnz, dtype=torch.float64, device=dev)
for _ in range(6):
    tr.g_jac_dev(zd.data_ptr(), None, g.data_ptr(), v.data_ptr(), sp)
torch.cuda.synchronize()
lib = _lib.lib()
nw, sl = C.c_int64(), C.c_int64()
lib.mpx_trace_read(tr._plan, C.byref(nw), C.byref(sl), None)
buf = np.zeros(nw.value * sl.value, dtype=np.uint64)
lib.mpx_trace_read(tr._plan, C.byref(nw), C.byref(sl), buf.ctypes.data_as(C.POINTER(C.c_ulong)))
T = buf.reshape(-1, sl.value)[:K].astype(np.int64)
ghz = 1.965
g0 = T[:, 0].min()
def q(a): return [round(float(x), 2) for x in np.percentile(a, [0, 10, 50, 90, 100])]
us = lambda a, b: (T[:, b] - T[:, a]) / ghz / 1e3
print("kernel span us", (T[:, 10].max() - g0) / 1e3, " CTAs", K)
print("CTA start us (global)         ", q((T[:, 0] - g0) / 1e3))
print("tables + nodes loaded (1->2)  ", q(us(1, 2)))
print("interpolation (2->3)          ", q(us(2, 3)))
print("functor at mid points (3->4)  ", q(us(3, 4)))
print("constant blocks issued (4->5) ", q(us(4, 5)))
print("six row images (5->6)         ", q(us(5, 6)))
print("d/dw part, barrier (6->9)     ", q(us(6, 9)))
print("segment total (1->9)          ", q(us(1, 9)))
sm = T[:, 15]
per = {}
for k in range(K):
    per.setdefault(int(sm[k]), []).append(k)
print("segments per SM", q(np.array([len(v_) for v_ in per.values()])))
