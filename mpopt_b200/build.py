"""Build libmpx.so: the C ABI (csrc/mpx_plan.cu) plus the ahead-of-time kernel instantiations.

``generate()`` traces every problem of :data:`mpopt_b200.problems.REGISTRY`, writes one
``csrc/gen/mpx_aot_<key>.cu`` per distinct generated source (the hand-written kernels of
``csrc/mpx_kernels.cuh`` instantiated for those node functors), and ``build()`` compiles
everything for sm_100a with nvcc (cross-compiles without a GPU).  Problems that are not in
the registry are compiled at run time through NVRTC by the library itself.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
GEN = os.path.join(CSRC, "gen")
LIB = os.path.join(HERE, "libmpx.so")
OBJ = os.path.join(HERE, "build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def aot_source(key: str, program, degrees=()) -> str:
    P = program.n_phases
    lines = [
        f"// AOT instantiation of the collocation kernels for program {key} (generated; see mpopt_b200/build.py)",
        '#include "../mpx_program.h"',
        f"#define MPX_PHASE_NAME(k) MpxPh_{key}_##k",
        program.cuda_source(),
        "namespace {",
    ]
    for k in range(P):
        lines.append(f"const MpxAotPhase<MpxPh_{key}_{k}" + "".join(f", {int(d)}" for d in degrees) + f"> k_{k};")
    lines.append("const MpxPhaseKernels* const phases[] = {" + ", ".join(f"&k_{k}" for k in range(P)) + "};")
    if P > 1:  # one g + jac_g launch for all phases (mpx_gjac2_multi_kernel)
        lines.append("const MpxAotProgram<MpxDegs<" + ", ".join(str(int(d)) for d in degrees) + ">, "
                     + ", ".join(f"MpxPh_{key}_{k}" for k in range(P)) + "> k_all;")
        lines.append(f'MpxProgramEntry entry = {{"{key}", {P}, phases, nullptr, &k_all}};')
    else:
        lines.append(f'MpxProgramEntry entry = {{"{key}", {P}, phases, nullptr, nullptr}};')
    lines.append("struct Reg { Reg() { mpx_register_program(&entry); } } reg;")
    lines.append("}  // namespace")
    return "\n".join(lines) + "\n"


def generate(verbose=False):
    from .problems import AOT_DEGREES, REGISTRY
    from .program import Program

    os.makedirs(GEN, exist_ok=True)
    wanted, degs = {}, {}
    for name, make in REGISTRY.items():
        prog = Program(make())
        wanted.setdefault(prog.key(), (name, prog))
        degs[prog.key()] = sorted(set(degs.get(prog.key(), ())) | set(AOT_DEGREES.get(name, ())))
    for fn in os.listdir(GEN):
        if fn.startswith("mpx_aot_") and fn[8:-3] not in wanted:
            os.remove(os.path.join(GEN, fn))
    paths = []
    for key, (name, prog) in wanted.items():
        path = os.path.join(GEN, f"mpx_aot_{key}.cu")
        src = f"// problem: {name}\n" + aot_source(key, prog, degs[key])
        if not os.path.exists(path) or open(path).read() != src:
            with open(path, "w") as f:
                f.write(src)
        paths.append(path)
        if verbose:
            print(f"  {name:22s} -> {os.path.relpath(path, HERE)}")
    return sorted(paths)


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libmpx.so")
    return nvcc


def _digest(src):
    """Content hash of a source and of every header it can include (mtimes do not survive a snapshot copy)."""
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    deps = [src] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    deps.append(os.path.join(HERE, "..", "include", "mpx.h"))
    for d in deps:
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _compile(src, force):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    stamp = obj + ".sha"
    dig = _digest(src)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, None
    cmd = [_nvcc()] + NVCC_FLAGS + ["-Xptxas", "-v", "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return obj, r.stderr


def build(force=False, verbose=False):
    """Generate + compile + link.  Returns the path of libmpx.so."""
    os.makedirs(OBJ, exist_ok=True)
    srcs = [os.path.join(CSRC, "mpx_plan.cu"), os.path.join(CSRC, "mpx_shims.cu")] + generate(verbose)
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        results = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in results]
    keep = {os.path.basename(o) for o in objs}
    for fn in os.listdir(OBJ):  # objects of programs that are no longer generated (their hash changed)
        if fn.endswith((".o", ".o.sha")) and fn.replace(".sha", "") not in keep:
            os.remove(os.path.join(OBJ, fn))
    rebuilt = [l for _, l in results if l is not None]
    if rebuilt:
        with open(os.path.join(OBJ, "ptxas.log"), "a" if len(rebuilt) < len(results) else "w") as f:
            f.write("".join(rebuilt))
    if force or rebuilt or not os.path.exists(LIB):
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
