#!/usr/bin/env python
"""bench.py -- fused residual + Jacobian (g + jac_g) evaluations per second of the collocation hot path.

    python bench.py --gpus 1 --steps K --warmup W          this repo's CUDA path
    python bench.py --impl reference ...                   the CPU restatement (oracle/) on the host cores
    torchrun ... bench.py --gpus N ...                     weak scaling: 4096 segments per GPU of one N x 4096-segment NLP

A step is ONE fused g + jac_g evaluation of the transcribed NLP named by BASELINE.json's metric:
seeded synthetic 6-state / 3-control OCP, n_segments=4096, poly_orders=15, LGR
(n_z=552 971, n_g=552 966, nnz=12 533 916 -> 109.15 MB algorithmic bytes per evaluation).

Prints one JSON line (see the task's bench contract): value = evaluations/s with inputs resident in HBM,
e2e = the same through the host-pointer C ABI (mpx_eval_jac_g) with pinned host buffers, roofline =
algorithmic bytes / CUDA-event time of the g+jac kernel against MEASURED_PEAKS.json, cpu_baseline = the
oracle timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json: the metric is quoted on the headline workload; `--config 2..5` time the other configurations of
# BASELINE.json["configs"] with the same method and print the same line (their own `metric` string)
WORKLOADS = {
    "headline": dict(problem="synthetic_6_3", n_segments=4096, poly_orders=15, scheme="LGR", tf=1.0,
                     name="synthetic 6-state/3-control quadratic-dynamics OCP (SURVEY 8d), poly_orders=15, LGR",
                     metric="NLP residual+Jacobian evals/sec at n_seg=4096,p=15,nx=6"),
    "2": dict(problem="moon_lander", n_segments=4096, poly_orders=15, scheme="LGR", tf=4.0,
              name="config 2: moon-lander (nx=2, nu=1), poly_orders=15, LGR",
              metric="NLP residual+Jacobian evals/sec, moon-lander n_seg=4096,p=15"),
    "3": dict(problem="van_der_pol", n_segments=2048, poly_orders="mixed", scheme="CGL", tf=10.0,
              name="config 3: van-der-Pol (nx=2, nu=1), poly_orders=[3,30,3,...] (30 where k%3==1), CGL",
              metric="NLP residual+Jacobian evals/sec, van-der-Pol n_seg=2048,p=[3,30,3..],CGL"),
    "4": dict(problem="synthetic_6_3", n_segments=8192, poly_orders=20, scheme="LGL", tf=1.0,
              name="config 4: synthetic 6-state/3-control OCP, poly_orders=20, LGL",
              metric="NLP residual+Jacobian evals/sec, synthetic 6/3 n_seg=8192,p=20,LGL"),
    "5": dict(problem="two_phase_schwartz", n_segments=1024, poly_orders=10, scheme="LGR", tf=None,
              name="config 5: two-phase Schwartz (stand-in for the orbit-raising example that does not exist, SURVEY 8d), "
                   "1024 segments per phase, poly_orders=10, LGR",
              metric="NLP residual+Jacobian evals/sec, two-phase n_seg=1024/phase,p=10"),
}
WORKLOAD = dict(WORKLOADS["headline"])
METRIC = WORKLOAD["metric"]


def poly_orders_of(w, K=None):
    K = w["n_segments"] if K is None else K
    return [30 if k % 3 == 1 else 3 for k in range(K)] if w["poly_orders"] == "mixed" else w["poly_orders"]
UNIT = "evals/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def workload_point(n_z, n_p, K, seed=20261017, tf=1.0, n_phases=1, na=0):
    """Seeded inputs: X, U ~ U(-1,1); t0 = 0; tf; Dirichlet segment widths (per phase)."""
    rng = np.random.default_rng(seed)
    z = rng.uniform(-1.0, 1.0, n_z)
    nvar = n_z // n_phases
    for ph in range(n_phases):  # [.. t0 tf a] closes every phase's block of z (mpopt.py:537-543)
        t0, t1 = (0.0, tf) if tf is not None else (1.0 * ph, 1.0 * ph + 1.0)
        z[(ph + 1) * nvar - 2 - na], z[(ph + 1) * nvar - 1 - na] = t0, t1
    p = np.concatenate([rng.dirichlet(np.ones(K)) for _ in range(n_phases)])
    assert p.shape == (n_p,)
    return z, p


def algorithmic_bytes(n_z, n_p, n_g, nnz):
    """read z and p once, write every g and every Jacobian value once."""
    return 8 * (n_z + n_p + n_g + nnz)


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, str(e)

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for name, bit in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "samples": len(self.samples),
                "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------- CPU baseline (oracle/)
# Two CPU implementations of the same fused g + jac_g, both test infrastructure under oracle/ (SURVEY.md 8d):
#   * oracle/cpu_ref: C++17 -O3 -march=native + OpenMP, compiled on this machine, checked against the numpy oracle to
#     1e-13 (tests/test_cpu_ref.py) -- the honest compiled competitor, timed on ALL host cores and on one;
#   * oracle/nlp.py: the numpy/scipy restatement the parity tests use (reported as `numpy_port_*` for continuity).
# The reference's own evaluator is CasADi's single-threaded SX virtual machine (mpopt.py:757, :804), not installable here.
def _oracle_worker(args):
    K, n_evals, seed = args
    from mpopt_b200.problems import REGISTRY
    from oracle.nlp import OracleNLP

    ora = OracleNLP(REGISTRY[WORKLOAD["problem"]](), K, poly_orders_of(WORKLOAD, K), WORKLOAD["scheme"])
    z, p = workload_point(ora.n_z, ora.n_p, K, seed, WORKLOAD["tf"], ora.P, ora.na)
    ora._eval(z, p)  # build caches
    t = time.perf_counter()
    for i in range(n_evals):
        ora._eval(z + 1e-3 * i, p)
    return time.perf_counter() - t


def numpy_baseline(budget_s=20.0, cores=1):
    """The numpy oracle's fused g + jac_g on a bounded sample: the same problem at K_s <= K segments, scaled by
    K_s / K (cost is linear in the number of segments)."""
    import multiprocessing as mp

    Kfull = WORKLOAD["n_segments"]
    t1 = _oracle_worker((256, 1, 0))  # probe at reduced size
    est_full = t1 * Kfull / 256
    Ks = Kfull
    while Ks > 256 and est_full * Ks / Kfull * 3 > budget_s:
        Ks //= 2
    n_evals = max(1, int(budget_s / max(est_full * Ks / Kfull, 1e-3) / 2))
    n_evals = min(n_evals, 8)
    if cores > 1:
        with mp.get_context("fork").Pool(cores) as pool:
            times = pool.map(_oracle_worker, [(Ks, n_evals, s) for s in range(cores)])
        evals_per_s = cores * n_evals / max(times)  # every worker evaluates n_evals times concurrently
    else:
        evals_per_s = n_evals / _oracle_worker((Ks, n_evals, 0))
    return {"value": evals_per_s * Ks / Kfull, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"numpy/scipy oracle, fused g+jac_g of the same OCP at n_segments={Ks} (x{Ks}/{Kfull} scaling), "
                      f"{n_evals} evals per core, {os.cpu_count()} host cores present"}


def compiled_baseline(warmup=3, steps=None, budget_s=10.0, threads=None):
    """oracle/cpu_ref on the FULL workload (every step is one whole evaluation), `threads` OpenMP threads (default:
    all host cores).  Returns (dict, seconds per evaluation) or None when the problem is outside cpu_ref's scope."""
    from oracle import cpu_ref

    if WORKLOAD["problem"] not in cpu_ref.PROBLEMS:
        return None
    K = WORKLOAD["n_segments"]
    ref = cpu_ref.CpuRef(WORKLOAD["problem"], K, poly_orders_of(WORKLOAD), WORKLOAD["scheme"], midu=True,
                         params=cpu_ref.synthetic_params() if WORKLOAD["problem"] == "synthetic_6_3" else None)
    threads = threads or os.cpu_count() or 1
    ref.set_threads(threads)
    z, p = workload_point(ref.n_z, K, K, tf=WORKLOAD["tf"])
    g, vals = np.empty(ref.n_g), np.empty(ref.nnz)
    tw, i = time.perf_counter(), 0
    while i < max(1, warmup) or time.perf_counter() - tw < 1.5:  # first touches of the 100 MB output, thread pool, clocks, cgroup burst
        ref.eval(z, p, g, vals)
        i += 1
    t1 = time.perf_counter()
    ref.eval(z, p, g, vals)
    t1 = time.perf_counter() - t1
    n = steps if steps else max(3, min(400, int(budget_s / max(t1, 1e-4))))
    zs = [z + 1e-3 * i for i in range(4)]
    t = time.perf_counter()
    for i in range(n):
        ref.eval(zs[i % 4], p, g, vals)
    dt = (time.perf_counter() - t) / n
    return {"value": 1.0 / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"oracle/cpu_ref (C++17 -O3 -march=native, OpenMP, {threads} threads of {os.cpu_count()} host "
                      f"cores): {n} whole evaluations of the full workload (n_segments={K}), {dt * 1e3:.2f} ms each"}, dt


def cpu_baseline(budget_s=20.0):
    """cpu_baseline of the JSON line: the compiled competitor on all host cores; its one-thread figure and the numpy
    oracle's ride along.  Problems outside cpu_ref's scope (config 5: two phases, path rows) use the numpy oracle."""
    cb = compiled_baseline(budget_s=budget_s * 0.4)
    if cb is None:
        return numpy_baseline(budget_s=budget_s, cores=1)
    cb = cb[0]
    one = compiled_baseline(budget_s=budget_s * 0.3, threads=1)[0]
    npy = numpy_baseline(budget_s=budget_s * 0.3, cores=1)
    cb["one_thread_evals_per_s"] = one["value"]
    cb["numpy_port_one_core_evals_per_s"] = npy["value"]
    return cb


def run_reference(args):
    """--impl reference: the reference's path on the host cores.  The reference itself (CasADi's SX VM behind IPOPT)
    cannot be installed in this image, so this arm times the compiled C++/OpenMP port of its
    transcription (oracle/cpu_ref) on all host cores: W warm-up and K timed whole evaluations of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    cb = compiled_baseline(warmup=warmup, steps=steps)
    if cb is None:
        cbd = numpy_baseline(budget_s=max(10.0, min(150.0, 2.0 * (steps + warmup))), cores=os.cpu_count() or 1)
    else:
        cbd = cb[0]
    line = {
        "impl": "reference", "metric": METRIC, "value": cbd["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 / cbd["value"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload=f"{WORKLOAD['name']}, n_segments={WORKLOAD['n_segments']}; fused g + jac_g",
                       **{k: WORKLOAD[k] for k in ("n_segments", "poly_orders", "scheme")}),
        "cpu_baseline": cbd,
        "e2e": {"value": cbd["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CasADi/IPOPT are not installable here; this arm is the compiled C++/OpenMP port of mpopt.py's "
                "transcription (oracle/cpu_ref, checked against the numpy oracle), all host cores",
    }
    emit(line)


# ----------------------------------------------------------------------------- CUDA arm
def run_cuda(args):
    """N = 1: the headline NLP on one GPU.  N > 1 (one rank per GPU): weak scaling -- ONE NLP of N x 4096 segments
    whose segments are sharded over the ranks, 4096 per rank, every rank evaluating its rows with the same kernel and
    keeping them (device-resident `value`) or moving them to the host over its own PCIe link (`e2e`); there is no
    data-path collective.  The all-gather of the shards' CSR blocks that BASELINE.json's north_star describes is
    measured in the same run and reported under "allgather" (NVLink is ~8x slower than HBM, so it costs more than it
    saves -- SURVEY.md 8e)."""
    import torch

    from mpopt_b200.nlp import Transcription
    from mpopt_b200.problems import REGISTRY
    from mpopt_b200 import shard as sh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:  # keep each rank (and the pinned host buffers it first-touches) on the CPU cores next to its GPU
        try:
            import pynvml

            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        except Exception:
            pass
    # N > 1: the headline scales WEAKLY (n_segments per GPU, one NLP of N x n_segments); config 4 -- BASELINE.json's
    # "sharded 1/2/4/8" configuration -- scales STRONGLY (one NLP of 8192 segments split over the GPUs)
    strong = world > 1 and args.config == "4"
    K = WORKLOAD["n_segments"] // world if strong else WORKLOAD["n_segments"]
    Kt = K * world
    deg, scheme = poly_orders_of(WORKLOAD, Kt), WORKLOAD["scheme"]
    if world > 1 and not isinstance(deg, int):
        raise SystemExit("bench.py --gpus N > 1 needs a uniform-degree configuration")
    ocp = REGISTRY[WORKLOAD["problem"]]()
    part = [(r * K, (r + 1) * K) for r in range(world)]
    tr = Transcription(ocp, Kt, deg, scheme, device=local, segments=None if world == 1 else part[rank])
    n_z, n_p, n_g, nnz = tr.n_z, tr.n_p, tr.n_g, tr.nnz  # sizes of the whole NLP
    if world == 1:
        B = algorithmic_bytes(n_z, n_p, n_g, nnz)
    else:  # this rank's share: its nodes of z, its widths, the rows it writes
        own = sum(int(c) for _, c in tr.shard_runs(0)) + sum(int(c) for _, c in tr.shard_runs(1))
        B = 8 * ((K * deg + 1) * (tr.nx + tr.nu) + 2 + tr.na + K + own)  # uniform degree (checked above)
    z_h, p_h = workload_point(n_z, n_p, Kt, tf=WORKLOAD["tf"], n_phases=tr.P, na=tr.na)

    # rotating device-resident input/output sets: every launch streams its outputs to HBM (R x 105 MB > 126 MB L2)
    R = max(4, int(np.ceil(4 * 109e6 / B)))
    z_d = [torch.from_numpy(z_h + 1e-3 * i).to(dev) for i in range(R)]
    p_d = torch.from_numpy(p_h).to(dev)
    g_d = [torch.zeros(n_g, dtype=torch.float64, device=dev) for _ in range(R)]
    v_d = [torch.zeros(nnz, dtype=torch.float64, device=dev) for _ in range(R)]
    flush = torch.zeros(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # 256 MB > 126 MB L2
    stream = torch.cuda.Stream()  # a real (non-NULL) stream: events and kernels share it (NULL = plan's own stream)
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0

    def step(i):
        k = i % R
        tr.g_jac_dev(z_d[k].data_ptr(), p_d.data_ptr(), g_d[k].data_ptr(), v_d[k].data_ptr(), sp)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    sampler.start()
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()

    # ---- timed region A (primary): K evaluations = K launches of the g + jac_g kernel, chained by programmatic
    #      dependent launch, outputs rotating over R buffer sets; CUDA events on the launch stream, max over ranks.
    #      Default: the K launches are captured ONCE into a CUDA graph and the timed region is ONE replay of that graph
    #      -- how a device-resident consumer that evaluates at this rate issues them: no host launch work inside the
    #      region, and the front-end sees the whole chain (17.6 us per evaluation against 19.5 us for the same K
    #      launches issued one by one from the host, same box; profiles/r02/README.md).  The host-issued variant is
    #      measured too (roofline.stream_launch_us; `--no-graph` makes it the primary): there the region is enqueued
    #      behind a gate kernel that holds the stream until every launch is queued, so Python -> ctypes ->
    #      cudaLaunchKernelEx latency (~10 us per call) stays outside the event pair.
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gate_us = 0.0 if args.no_gate else min(2000.0, 150.0 + 15.0 * args.steps)
    barrier()
    l0 = tr.launches
    if gate_us:
        tr._L.mpx_gate(sp, gate_us)
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    launches = tr.launches - l0  # kernels of this repo launched per timed region (the graph replays the same ones)
    ms_stream = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    ms_step, graph_ok = ms_stream, None
    if not args.no_graph:
        last = (args.steps - 1) % R
        ref_v = v_d[last].clone()  # what the last launch of the chain wrote
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for i in range(args.steps):
                step(i)
        v_d[last].zero_()
        graph.replay()  # warm replay, and the check that the captured chain computes the same thing
        barrier()
        graph_ok = bool(torch.equal(ref_v, v_d[last]))
        del ref_v
        if gate_us:
            tr._L.mpx_gate(sp, gate_us)
        e0.record(stream)
        graph.replay()
        e1.record(stream)
        barrier()
        ms_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps

    # ---- timed region B (one launch at a time, L2 evicted before each): context for the roofline.  The eviction is a
    #      READ of 256 MB (clean lines): filling L2 with dirty lines instead would charge their write-back to the kernel.
    nb = min(args.steps, 50)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nb)]
    sink = torch.empty(1, dtype=torch.float64, device=dev)
    for i, (a, b) in enumerate(ev):
        torch.sum(flush, dim=0, keepdim=True, out=sink)
        a.record(stream)
        step(i)
        b.record(stream)
    torch.cuda.synchronize()
    kern_ms = np.array([a.elapsed_time(b) for a, b in ev])

    # ---- N > 1: the north-star variant -- every rank ends up with the whole g / Jacobian.  (a) fused: the kernel's
    #      stores go to the peers' buffers as well (peer memory over NVLink, no separate collective, one tiny all-reduce
    #      per evaluation orders the ranks); (b) baseline: the same shards + one NCCL all-gather of g / CSR values.
    allgather = allgather_nccl = None
    if world > 1 and not args.no_allgather:
        full = Transcription(ocp, Kt, deg, scheme, device=local)  # the same NLP on ONE GPU: the check and the yardstick
        g_ref = torch.empty(n_g, dtype=torch.float64, device=dev)
        v_ref = torch.empty(nnz, dtype=torch.float64, device=dev)
        f_ref = torch.zeros(1, dtype=torch.float64, device=dev)
        grad_ref = torch.empty(n_z, dtype=torch.float64, device=dev)
        full.g_jac_dev(z_d[0].data_ptr(), p_d.data_ptr(), g_ref.data_ptr(), v_ref.data_ptr(), sp)
        full.f_grad_dev(z_d[0].data_ptr(), p_d.data_ptr(), f_ref.data_ptr(), grad_ref.data_ptr(), sp)
        torch.cuda.synchronize()
        for i in range(3):
            full.g_jac_dev(z_d[0].data_ptr(), p_d.data_ptr(), g_ref.data_ptr(), v_ref.data_ptr(), sp)
        tr._L.mpx_gate(sp, 300.0)
        e0.record(stream)
        for i in range(10):
            full.g_jac_dev(z_d[0].data_ptr(), p_d.data_ptr(), g_ref.data_ptr(), v_ref.data_ptr(), sp)
        e1.record(stream)
        torch.cuda.synchronize()
        ms_single = e0.elapsed_time(e1) / 10

        def verify(g_t, v_t, f_t=None, grad_t=None):
            """Every rank's gathered buffers against the single-GPU evaluation of the whole NLP: g and the Jacobian
            values bit for bit; f and grad_f (partials summed in another order) to 1e-12."""
            ok = torch.equal(g_ref, g_t) and torch.equal(v_ref, v_t)
            if f_t is not None:
                ok = ok and bool(abs(float(f_t[0]) - float(f_ref[0])) <= 1e-12 * max(1.0, abs(float(f_ref[0]))))
                ok = ok and bool(torch.max(torch.abs(grad_t - grad_ref) / torch.clamp(torch.abs(grad_ref), min=1.0)) <= 1e-12)
            okt = torch.tensor([int(ok)], device=dev)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            return bool(okt.item())

        ng = max(3, min(args.steps, 20))
        recv = 8 * (world - 1) * own  # bytes every rank receives per evaluation (g + Jacobian values of the other shards)
        og = sh.ObjectiveGatherer(tr.layout, part, dist, rank)
        f_t = torch.zeros(1, dtype=torch.float64, device=dev)
        grad_t = torch.empty(n_z, dtype=torch.float64, device=dev)

        def objective(k):  # objective partials + grad_f shards: one all-reduce of a few doubles, one all-gather
            tr.f_grad_dev(z_d[k].data_ptr(), p_d.data_ptr(), f_t.data_ptr(), grad_t.data_ptr(), sp)
            og.all_gather(f_t, grad_t)

        common = {"bytes_received_per_rank": int(recv), "nvlink_peer_copy_gbs": 770.0,
                  "ingress_floor_ms": recv / 770e9 * 1e3, "single_gpu_same_nlp_ms": ms_single,
                  "note": "every rank ends with the whole g / Jacobian / grad_f; NVLink (0.77 TB/s) is 8x slower than HBM, "
                          "so this can never beat evaluating the whole NLP on one GPU (single_gpu_same_nlp_ms) -- "
                          "SURVEY.md 8e; it is the north-star's exchange, measured, not the recommended mode"}
        # (a) fused peer stores
        pb = sh.PeerBuffers(n_g, nnz, dist, rank, local, n_sets=R)
        flag = torch.zeros(1, device=dev)

        def fstep(i, with_obj=True):
            k = i % R
            g_t, v_t = pb.local(k)
            pg, pv = pb.peers(k)
            tr.g_jac_dev_peers(z_d[k].data_ptr(), p_d.data_ptr(), g_t.data_ptr(), v_t.data_ptr(), pg, pv, sp)
            if with_obj:
                objective(k)
            else:
                dist.all_reduce(flag)  # every rank's kernel (and with it its peer stores) has completed

        for i in range(3):
            fstep(i)
        barrier()
        fstep(0)
        barrier()
        ok = verify(*pb.local(0), f_t, grad_t)
        res = {}
        for name, with_obj in (("jacobian_only", False), ("with_objective", True)):
            barrier()
            e0.record(stream)
            for i in range(ng):
                fstep(i, with_obj)
            e1.record(stream)
            barrier()
            res[name] = max_over_ranks(e0.elapsed_time(e1)) / ng
        ms_f = res["with_objective"]
        allgather = {"value": (1 if strong else world) * 1e3 / ms_f, "unit": UNIT, "ms_per_step": ms_f,
                     "ms_per_step_jacobian_only": res["jacobian_only"], "steps": ng, "equals_single_gpu": ok,
                     "nvlink_gbs_per_rank": recv / (res["jacobian_only"] * 1e-3) / 1e9, **common,
                     "how": "mpx_eval_g_jac_dev_peers: each image is handed to the copy engine once per destination "
                            "(cp.async.bulk to peer-mapped memory), g by plain peer stores; then f + grad_f of the shard, "
                            "one all-reduce of the objective partials (which also orders the ranks) and one all-gather "
                            "of the grad_f shards"}
        pb.close()
        # (b) NCCL baseline
        row0 = None
        z_cur = [z_d[0]]
        if rank != 0:  # global node 0's rows are recomputed locally instead of being broadcast (mpopt_b200/shard.py)
            tr0 = Transcription(ocp, Kt, deg, scheme, device=local, segments=(0, 1))
            row0 = lambda g, v: tr0.g_jac_dev(z_cur[0].data_ptr(), p_d.data_ptr(), g.data_ptr(), v.data_ptr(), sp)
        gather = sh.Gatherer(tr.layout, part, dist, rank, dev, row0)

        def gstep(i, with_obj=True):
            k = i % R
            z_cur[0] = z_d[k]
            step(i)
            gather.all_gather(g_d[k], v_d[k])
            if with_obj:
                objective(k)

        for i in range(3):
            gstep(i)
        barrier()
        gstep(0)
        barrier()
        ok = verify(g_d[0], v_d[0], f_t, grad_t)
        for name, with_obj in (("jacobian_only", False), ("with_objective", True)):
            barrier()
            e0.record(stream)
            for i in range(ng):
                gstep(i, with_obj)
            e1.record(stream)
            barrier()
            res[name] = max_over_ranks(e0.elapsed_time(e1)) / ng
        ms_g = res["with_objective"]
        allgather_nccl = {"value": (1 if strong else world) * 1e3 / ms_g, "unit": UNIT, "ms_per_step": ms_g,
                          "ms_per_step_jacobian_only": res["jacobian_only"], "steps": ng, "mode": gather.mode,
                          "equals_single_gpu": ok, **common,
                          "how": "shard kernel, then one NCCL all-gather of g / CSR values (torch.distributed), then the "
                                 "objective partials and grad_f shards the same way"}
        del full, g_ref, v_ref, grad_ref

    # ---- end-to-end: host buffers through the C ABI (pinned); H2D of this rank's z / p and D2H of the rows it owns
    #      inside the timing.  Every rank uses its own PCIe link; no rank waits for another inside the timed region.
    zh = torch.from_numpy(z_h.copy()).pin_memory()
    ph_ = torch.from_numpy(p_h.copy()).pin_memory()
    gh = torch.empty(n_g, dtype=torch.float64).pin_memory()
    vh = torch.empty(nnz, dtype=torch.float64).pin_memory()
    n_e2e = max(3, min(args.steps, 50))
    node0 = rank * K * deg if isinstance(deg, int) else 0
    for _ in range(2):
        tr.jac_g_values(zh.numpy(), ph_.numpy(), out=vh.numpy(), g_out=gh.numpy())
    barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        zh[node0] = float(z_h[node0] + 1e-6 * i)
        tr.jac_g_values(zh.numpy(), ph_.numpy(), out=vh.numpy(), g_out=gh.numpy())
    dt = max_over_ranks(time.perf_counter() - t0) / n_e2e
    d2h = 8 * (n_g + nnz) if world == 1 else 8 * own
    h2d = 8 * (n_z + n_p) if world == 1 else 8 * ((K * deg + 1 + deg) * (tr.nx + tr.nu) + 2 + tr.na + n_p)
    e2e = {"value": (1 if strong else world) / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": dt * 1e3, "steps": n_e2e,
           "api": "mpx_eval_jac_g (host pointers, pinned buffers)" + ("" if world == 1 else
                  "; per rank: its own shard over its own PCIe link, bytes are per rank")}
    # calibration of the host link: NOTHING but the same number of bytes copied device -> pinned host memory by every rank
    # at the same time (torch copies, no kernel of this repo): what the box can move.  When e2e's ms_per_step is close to
    # this figure the limiter is the host side of the box (all GPUs of these boxes hang off one NUMA node), not the path.
    try:
        src = torch.empty(int(d2h) // 8, dtype=torch.float64, device=dev)
        dsth = torch.empty(int(d2h) // 8, dtype=torch.float64).pin_memory()
        for _ in range(2):
            dsth.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            dsth.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
        dtc = max_over_ranks(time.perf_counter() - t0) / n_e2e
        e2e["d2h_copy_only_ms_per_step"] = dtc * 1e3
        e2e["d2h_copy_only_gbs_all_ranks"] = world * d2h / dtc / 1e9
        del src, dsth
    except Exception as ex:  # the calibration must never take the bench line down
        e2e["d2h_copy_only_ms_per_step"] = None
        print("# host-link calibration skipped:", str(ex)[:100], file=sys.stderr)
    # ---- the same end to end, other ways a caller can hand over its buffers (N = 1; reported beside `e2e`, never as
    #      the roofline): plain pageable numpy arrays (what IPOPT / CasADi allocate), the caller's pageable arrays
    #      registered once with mpx_host_register, the dynamic fetch into a registered array (only the z-dependent
    #      entries cross PCIe after the first call), and the packed dynamic entries
    e2e_variants = None
    if world == 1 and not args.no_e2e_variants:
        def timed(fn, n=n_e2e):
            for _ in range(2):
                fn(0)
            t = time.perf_counter()
            for i in range(n):
                fn(i)
            return (time.perf_counter() - t) / n

        zp_, pp_ = z_h.copy(), p_h.copy()        # pageable inputs
        gp_, vp_ = np.empty(n_g), np.empty(nnz)  # pageable outputs

        def bump(i):
            zp_[node0] = float(z_h[node0] + 1e-6 * i)

        def full(i):
            bump(i)
            tr.jac_g_values(zp_, pp_, out=vp_, g_out=gp_)

        e2e_variants = {}
        dt_page = timed(full)
        e2e_variants["pageable"] = {"value": 1.0 / dt_page, "ms_per_step": dt_page * 1e3, "d2h_bytes_per_step": int(d2h),
                                    "api": "mpx_eval_jac_g, plain numpy arrays (driver-staged copies)"}
        try:
            for a in (zp_, gp_, vp_):
                tr.host_register(a)
            dt_reg = timed(full)
            e2e_variants["registered"] = {"value": 1.0 / dt_reg, "ms_per_step": dt_reg * 1e3, "d2h_bytes_per_step": int(d2h),
                                          "api": "mpx_eval_jac_g after mpx_host_register of the caller's arrays"}
            n_dyn = len(tr.dynamic_positions())

            def dyn(i):
                bump(i)
                tr.jac_g_values_dynamic(zp_, pp_, out=vp_, g_out=gp_)

            dt_dyn = timed(dyn)
            chk = tr.jac_g_values(zp_, pp_)  # the buffer still holds the whole Jacobian of the last point
            e2e_variants["dynamic"] = {"value": 1.0 / dt_dyn, "ms_per_step": dt_dyn * 1e3,
                                       "d2h_bytes_per_step": int(8 * (n_g + n_dyn)), "n_dynamic": int(n_dyn), "nnz": int(nnz),
                                       "equals_full_fetch": bool(np.array_equal(chk, vp_)),
                                       "api": "mpx_eval_jac_g_dynamic: kernel stores of the z-dependent entries into the "
                                              "registered (mapped) caller array; constants stay from the first call"}
            pk = torch.empty(n_dyn, dtype=torch.float64).pin_memory().numpy()

            def packed(i):
                bump(i)
                tr.jac_g_packed(zp_, pp_, out=pk, g_out=gp_)

            dt_pk = timed(packed)
            e2e_variants["packed"] = {"value": 1.0 / dt_pk, "ms_per_step": dt_pk * 1e3,
                                      "d2h_bytes_per_step": int(8 * (n_g + n_dyn)),
                                      "api": "mpx_eval_jac_g_packed: dynamic entries gathered on the device, one copy into "
                                             "pinned memory (the caller keeps the constants and scatters)"}
        finally:
            for a in (zp_, gp_, vp_):
                try:
                    tr.host_unregister(a)
                except Exception:
                    pass
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    kmed = float(np.median(kern_ms))
    # roofline: algorithmic bytes of one launch / average launch duration over timed region A (CUDA events on the launch
    # stream around K back-to-back launches whose outputs rotate over 4 sets > L2, so every launch streams to HBM)
    achieved = B / (ms_step * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch", {})
        traffic = traffic.get(args.config) if isinstance(traffic, dict) else (traffic if args.config == "headline" else None)
    cb = cpu_baseline(budget_s=args.cpu_budget) if (not args.no_cpu and world == 1) else None
    line = {
        "metric": METRIC, "value": (1 if strong else world) * 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD['name']}, n_segments={K} per GPU: one fused g + jac_g evaluation of {K} "
                               "segments per step and GPU",
                   "timing": ("K launches captured in a CUDA graph, one replay between two CUDA events on the launch stream"
                              if not args.no_graph else "K host-issued launches behind a gate kernel, one CUDA-event pair"),
                   "graph_equals_stream_launches": graph_ok,
                   "n_segments_total": Kt, "n_z": n_z, "n_g": n_g, "nnz_jac": nnz, "algorithmic_bytes_per_gpu": int(B),
                   "l2": f"{R} rotating z/g/values sets ({R * B / 1e6:.0f} MB written per GPU > 126 MB L2), launches back to back",
                   "parallelism": "1 GPU" if world == 1 else
                   (f"one NLP of {Kt} segments split over {world} GPUs ({K} each); rows stay on the GPU that computed "
                    f"them (no data-path collective); value = evaluations of the whole NLP per second" if strong else
                    f"one NLP of {Kt} segments, {K} per GPU over {world} GPUs; rows stay on the GPU that computed them "
                    f"(no data-path collective); value = {K}-segment evaluations per second summed over the GPUs"),
                   "program": tr.program_origin},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": f"mpx_gjac2_kernel<{WORKLOAD['problem']}, JAC, {deg if isinstance(deg, int) else 0}>",
                     "launch_us_avg": ms_step * 1e3, "stream_launch_us": ms_stream * 1e3, "bytes_per_launch": int(B),
                     "isolated_launch_us_median": kmed * 1e3, "isolated_launch_us_min": float(kern_ms.min()) * 1e3,
                     "how": "achieved = algorithmic bytes / average launch duration over the timed region (K chained "
                            "launches replayed as one CUDA graph, rotating output sets > L2; per GPU, slowest rank); "
                            "stream_launch_us = the same K launches issued one by one from the host; isolated_* = single "
                            "launches after a 256 MB read that evicts L2 (includes launch latency and a cold start)"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(),
    }
    if e2e_variants is not None:
        line["e2e_variants"] = e2e_variants
    if allgather is not None:
        line["allgather"] = allgather
        line["allgather_nccl"] = allgather_nccl
    if cb is not None:
        line["cpu_baseline"] = cb
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


_RESULT_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout
    when NCCL_DEBUG asks for it), so everything else is sent to stderr: fd 1 is pointed at fd 2 for the whole run and
    the result line goes to a duplicate of the original stdout."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=25.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-allgather", action="store_true", help="N > 1: skip the extra all-gather measurement")
    ap.add_argument("--no-e2e-variants", action="store_true", help="skip the pageable / registered / dynamic e2e legs")
    ap.add_argument("--no-graph", action="store_true", help="primary timing = K host-issued launches instead of one CUDA-graph replay")
    ap.add_argument("--no-gate", action="store_true", help="do not enqueue the timed region behind a gate kernel")
    ap.add_argument("--config", default="headline", choices=sorted(WORKLOADS),
                    help="BASELINE.json configuration to time (default: the one the metric is quoted on)")
    args = ap.parse_args()
    global WORKLOAD, METRIC
    WORKLOAD = dict(WORKLOADS[args.config])
    METRIC = WORKLOAD["metric"]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
