"""One adaptive g + jac_g evaluation at the headline size (for ncu)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mpopt_b200.nlp import Transcription  # noqa: E402
from mpopt_b200.problems import synthetic_6_3  # noqa: E402

K = 4096
tr = Transcription(synthetic_6_3(), K, 15, "LGR", adaptive=True)
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
z = rng.uniform(-1, 1, tr.n_z)
L = tr.layout
z[L.colT0(0)], z[L.colTF(0)] = 0.0, 1.0
z[L.colW(0, 0): L.colW(0, 0) + K] = rng.dirichlet(np.ones(K))
zd = torch.from_numpy(z).to(dev)
g = torch.empty(tr.n_g, dtype=torch.float64, device=dev)
v = torch.empty(tr.nnz, dtype=torch.float64, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    tr.g_jac_dev(zd.data_ptr(), None, g.data_ptr(), v.data_ptr(), None)
tr.sync()
torch.cuda.synchronize()
