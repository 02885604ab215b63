"""Loader of the compiled CPU baseline ``cpu_ref.cpp`` (TEST INFRASTRUCTURE -- see the header of that file).

Only ``tests/``, ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs and ``__graft_entry__.build()`` may use
this module; ``mpopt_b200/`` never does.  The library is compiled with ``-O3 -march=native`` ON THE MACHINE THAT RUNS
IT (the build box and the GPU box have different host CPUs): the file name carries a hash of the source and of the
CPU model, so a copy that travelled from another machine is ignored and rebuilt (g++ is part of the image).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
SRC = os.path.join(HERE, "cpu_ref.cpp")
FLAGS = ["-O3", "-march=native", "-fopenmp", "-std=c++17", "-fPIC", "-shared"]
PROBLEMS = {"synthetic_6_3": 0, "moon_lander": 1, "van_der_pol": 2}


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith(("model name", "flags")):
                yield line
                if line.startswith("flags"):
                    return
    except OSError:
        yield "unknown"


def lib_path():
    h = hashlib.sha256(open(SRC, "rb").read() + "".join(_cpu_model()).encode() + " ".join(FLAGS).encode()).hexdigest()[:12]
    return os.path.join(OUT, f"libcpu_ref_{h}.so")


def build():
    path = lib_path()
    if not os.path.exists(path):
        os.makedirs(OUT, exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        tmp = path + f".{os.getpid()}.tmp"
        subprocess.run([cxx] + FLAGS + ["-o", tmp, SRC], check=True, capture_output=True)
        os.replace(tmp, path)
    return path


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.cpu_ref_create.restype = C.c_void_p
        L.cpu_ref_create.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.cpu_ref_destroy.argtypes = [C.c_void_p]
        L.cpu_ref_sizes.argtypes = [C.c_void_p] * 4
        L.cpu_ref_structure.argtypes = [C.c_void_p] * 3
        L.cpu_ref_eval.argtypes = [C.c_void_p] * 5
        L.cpu_ref_threads.restype = C.c_int
        L.cpu_ref_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


class CpuRef:
    """Fused g + jac_g of one of the supported single-phase problems on the host cores (C++ / OpenMP).

    Tables (D, mid-point interpolation) come from the numpy oracle's ``Collocation`` -- table construction is setup,
    not part of the per-iteration path that is timed."""

    def __init__(self, problem, n_segments, poly_orders, scheme, midu=True, params=None):
        from oracle.collocation import diff_matrix, interpolation_matrix, roots  # the oracle's own tables

        L = lib()
        self.K = int(n_segments)
        po = np.asarray([poly_orders] * self.K if np.isscalar(poly_orders) else poly_orders, dtype=np.int32)
        degs = np.unique(po).astype(np.int32)
        Ds, Cs = [], []
        for d in degs:
            r = roots(scheme, int(d))
            Ds.append(np.ascontiguousarray(diff_matrix(r)))
            mid = 0.5 * (r[:-1] + r[1:])
            Cs.append(np.ascontiguousarray(interpolation_matrix(r, mid)))
        self._keep = (po, degs, Ds, Cs, None if params is None else np.ascontiguousarray(params, dtype=float))
        Dp = (C.c_void_p * len(degs))(*[a.ctypes.data for a in Ds])
        Cp = (C.c_void_p * len(degs))(*[a.ctypes.data for a in Cs])
        par = self._keep[4].ctypes.data if params is not None else None
        self._h = L.cpu_ref_create(PROBLEMS[problem], par, self.K, po.ctypes.data, len(degs), degs.ctypes.data, Dp, Cp, int(midu))
        if not self._h:
            raise ValueError(f"cpu_ref: unknown problem {problem!r}")
        s = [C.c_int64() for _ in range(3)]
        L.cpu_ref_sizes(self._h, *[C.byref(v) for v in s])
        self.n_z, self.n_g, self.nnz = (int(v.value) for v in s)
        self._L = L

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.cpu_ref_destroy(self._h)
            self._h = None

    def structure(self):
        rp, ci = np.empty(self.n_g + 1, np.int64), np.empty(self.nnz, np.int64)
        self._L.cpu_ref_structure(self._h, rp.ctypes.data, ci.ctypes.data)
        return rp, ci

    def eval(self, z, w, g=None, vals=None):
        g = np.empty(self.n_g) if g is None else g
        vals = np.empty(self.nnz) if vals is None else vals
        self._L.cpu_ref_eval(self._h, z.ctypes.data, w.ctypes.data, g.ctypes.data, vals.ctypes.data)
        return g, vals

    @property
    def threads(self):
        return int(self._L.cpu_ref_threads())

    def set_threads(self, n):
        self._L.cpu_ref_set_threads(int(n))


def synthetic_params():
    """A | B | C of the seeded synthetic 6/3 dynamics (same draws as mpopt_b200.problems.synthetic_6_3)."""
    rng = np.random.default_rng(6)
    A = rng.uniform(-1, 1, (6, 6))
    B = rng.uniform(-1, 1, (6, 3))
    Cm = rng.uniform(-1, 1, (6, 6))
    return np.concatenate([A.ravel(), B.ravel(), Cm.ravel()])
