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

WORKLOAD = dict(problem="synthetic_6_3", n_segments=4096, poly_orders=15, scheme="LGR", tf=1.0)
METRIC = "NLP residual+Jacobian evals/sec at n_seg=4096,p=15,nx=6"
UNIT = "evals/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def workload_point(n_z, n_p, K, seed=20261017, tf=1.0):
    """Seeded inputs: X, U ~ U(-1,1); t0 = 0; tf; Dirichlet segment widths."""
    rng = np.random.default_rng(seed)
    z = rng.uniform(-1.0, 1.0, n_z)
    z[-2:] = [0.0, tf]
    p = rng.dirichlet(np.ones(K))
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


# ----------------------------------------------------------------------------- CPU baseline (oracle)
def _oracle_worker(args):
    K, n_evals, seed = args
    from mpopt_b200.problems import REGISTRY
    from oracle.nlp import OracleNLP

    ora = OracleNLP(REGISTRY[WORKLOAD["problem"]](), K, WORKLOAD["poly_orders"], WORKLOAD["scheme"])
    z, p = workload_point(ora.n_z, ora.n_p, K, seed, WORKLOAD["tf"])
    ora._eval(z, p)  # build caches
    t = time.perf_counter()
    for i in range(n_evals):
        ora._eval(z + 1e-3 * i, p)
    return time.perf_counter() - t


def cpu_baseline(budget_s=20.0, cores=1):
    """Time the oracle's fused g + jac_g on a bounded sample: the same problem at K_s <= 4096 segments, scaled
    by K_s / 4096 (cost is linear in the number of segments).  Returns the cpu_baseline dict."""
    import multiprocessing as mp

    Kfull = WORKLOAD["n_segments"]
    t1 = _oracle_worker((256, 1, 0))  # probe at 1/16 size
    est_full = t1 * Kfull / 256
    Ks = Kfull
    while Ks > 256 and est_full * Ks / Kfull * 3 > budget_s:
        Ks //= 2
    n_evals = max(1, int(budget_s / max(est_full * Ks / Kfull, 1e-3) / 2))
    n_evals = min(n_evals, 8)
    if cores > 1:
        with mp.get_context("fork").Pool(cores) as pool:
            times = pool.map(_oracle_worker, [(Ks, n_evals, s) for s in range(cores)])
        # every worker evaluates n_evals times concurrently; building the oracle is not counted
        evals_per_s = cores * n_evals / max(times)
    else:
        per = _oracle_worker((Ks, n_evals, 0)) / n_evals
        evals_per_s = 1.0 / per
    value = evals_per_s * Ks / Kfull
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"numpy/scipy oracle, fused g+jac_g of the same OCP at n_segments={Ks} (x{Ks}/{Kfull} scaling), "
                      f"{n_evals} evals per core, {os.cpu_count()} host cores present"}


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's path on all host cores (the reference itself,
    CasADi + IPOPT, cannot be installed in this image -- SURVEY.md 8c)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps, warmup = args.steps, args.warmup
    budget = min(150.0, 2.0 * (steps + warmup))
    cb = cpu_baseline(budget_s=max(10.0, budget), cores=cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload="synthetic 6-state/3-control OCP, n_segments=4096, poly_orders=15, LGR; "
                                "fused g + jac_g", **{k: WORKLOAD[k] for k in ("n_segments", "poly_orders", "scheme")}),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CasADi/IPOPT are not installable here; this is the numpy/scipy oracle port of mpopt.py's transcription",
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
    K = WORKLOAD["n_segments"]
    Kt = K * world  # weak scaling: 4096 segments per GPU
    deg, scheme = WORKLOAD["poly_orders"], WORKLOAD["scheme"]
    ocp = REGISTRY[WORKLOAD["problem"]]()
    part = [(r * K, (r + 1) * K) for r in range(world)]
    tr = Transcription(ocp, Kt, deg, scheme, device=local, segments=None if world == 1 else part[rank])
    n_z, n_p, n_g, nnz = tr.n_z, tr.n_p, tr.n_g, tr.nnz  # sizes of the whole NLP
    if world == 1:
        B = algorithmic_bytes(n_z, n_p, n_g, nnz)
    else:  # this rank's share: its nodes of z, its widths, the rows it writes
        own = sum(int(c) for _, c in tr.shard_runs(0)) + sum(int(c) for _, c in tr.shard_runs(1))
        B = 8 * ((K * deg + 1) * (tr.nx + tr.nu) + 2 + tr.na + K + own)
    z_h, p_h = workload_point(n_z, n_p, Kt, tf=WORKLOAD["tf"])

    # rotating device-resident input/output sets: every launch streams its outputs to HBM (R x 105 MB > 126 MB L2)
    R = 4
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

    # ---- timed region A (primary): K back-to-back steps over rotating buffer sets, one event pair, max over ranks
    l0 = tr.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    launches = tr.launches - l0
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
        def verify(g_t, v_t):  # equal to a plain single-GPU evaluation of the whole NLP, bit for bit
            full = Transcription(ocp, Kt, deg, scheme, device=local)
            g_ref = torch.empty(n_g, dtype=torch.float64, device=dev)
            v_ref = torch.empty(nnz, dtype=torch.float64, device=dev)
            full.g_jac_dev(z_d[0].data_ptr(), p_d.data_ptr(), g_ref.data_ptr(), v_ref.data_ptr(), sp)
            torch.cuda.synchronize()
            okt = torch.tensor([int(torch.equal(g_ref, g_t) and torch.equal(v_ref, v_t))], device=dev)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            return bool(okt.item())

        ng = max(3, min(args.steps, 20))
        recv = 8 * (world - 1) * own  # bytes every rank receives per evaluation
        # (a) fused peer stores
        pb = sh.PeerBuffers(n_g, nnz, dist, rank, local, n_sets=R)
        flag = torch.zeros(1, device=dev)

        def fstep(i):
            k = i % R
            g_t, v_t = pb.local(k)
            pg, pv = pb.peers(k)
            tr.g_jac_dev_peers(z_d[k].data_ptr(), p_d.data_ptr(), g_t.data_ptr(), v_t.data_ptr(), pg, pv, sp)
            dist.all_reduce(flag)  # every rank's kernel (and with it its peer stores) has completed

        for i in range(3):
            fstep(i)
        barrier()
        fstep(0)
        torch.cuda.synchronize()
        ok = verify(*pb.local(0)) if world <= 2 else None
        barrier()
        e0.record(stream)
        for i in range(ng):
            fstep(i)
        e1.record(stream)
        barrier()
        ms_f = max_over_ranks(e0.elapsed_time(e1)) / ng
        allgather = {"value": world * 1e3 / ms_f, "unit": UNIT, "ms_per_step": ms_f, "steps": ng, "equals_single_gpu": ok,
                     "bytes_sent_per_rank": int(recv), "nvlink_gbs_per_rank": recv / (ms_f * 1e-3) / 1e9,
                     "how": "mpx_eval_g_jac_dev_peers: each image is handed to the copy engine once per destination "
                            "(cp.async.bulk to peer-mapped memory), g by plain peer stores; + one 4-byte all-reduce"}
        pb.close()
        # (b) NCCL baseline
        row0 = None
        z_cur = [z_d[0]]
        if rank != 0:  # global node 0's rows are recomputed locally instead of being broadcast (mpopt_b200/shard.py)
            tr0 = Transcription(ocp, Kt, deg, scheme, device=local, segments=(0, 1))
            row0 = lambda g, v: tr0.g_jac_dev(z_cur[0].data_ptr(), p_d.data_ptr(), g.data_ptr(), v.data_ptr(), sp)
        gather = sh.Gatherer(tr.layout, part, dist, rank, dev, row0)

        def gstep(i):
            k = i % R
            z_cur[0] = z_d[k]
            step(i)
            gather.all_gather(g_d[k], v_d[k])

        for i in range(3):
            gstep(i)
        barrier()
        gstep(0)
        torch.cuda.synchronize()
        ok = verify(g_d[0], v_d[0]) if world <= 2 else None
        barrier()
        e0.record(stream)
        for i in range(ng):
            gstep(i)
        e1.record(stream)
        barrier()
        ms_g = max_over_ranks(e0.elapsed_time(e1)) / ng
        allgather_nccl = {"value": world * 1e3 / ms_g, "unit": UNIT, "ms_per_step": ms_g, "steps": ng, "mode": gather.mode,
                          "equals_single_gpu": ok, "bytes_received_per_rank": int(recv),
                          "how": "shard kernel, then one NCCL all-gather of g / CSR values (torch.distributed)"}

    # ---- end-to-end: host buffers through the C ABI (pinned); H2D of this rank's z / p and D2H of the rows it owns
    #      inside the timing.  Every rank uses its own PCIe link; no rank waits for another inside the timed region.
    zh = torch.from_numpy(z_h.copy()).pin_memory()
    ph_ = torch.from_numpy(p_h.copy()).pin_memory()
    gh = torch.empty(n_g, dtype=torch.float64).pin_memory()
    vh = torch.empty(nnz, dtype=torch.float64).pin_memory()
    n_e2e = max(3, min(args.steps, 50))
    for _ in range(2):
        tr.jac_g_values(zh.numpy(), ph_.numpy(), out=vh.numpy(), g_out=gh.numpy())
    barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        zh[rank * K * deg] = float(z_h[rank * K * deg] + 1e-6 * i)
        tr.jac_g_values(zh.numpy(), ph_.numpy(), out=vh.numpy(), g_out=gh.numpy())
    dt = max_over_ranks(time.perf_counter() - t0) / n_e2e
    d2h = 8 * (n_g + nnz) if world == 1 else 8 * own
    h2d = 8 * (n_z + n_p) if world == 1 else 8 * ((K * deg + 1 + deg) * (tr.nx + tr.nu) + 2 + tr.na + n_p)
    e2e = {"value": world / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": dt * 1e3, "steps": n_e2e,
           "api": "mpx_eval_jac_g (host pointers, pinned buffers)" + ("" if world == 1 else
                  "; per rank: its own shard over its own PCIe link, bytes are per rank")}
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
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    cb = cpu_baseline(budget_s=args.cpu_budget, cores=1) if (not args.no_cpu and world == 1) else None
    line = {
        "metric": METRIC, "value": world * 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic 6-state/3-control quadratic-dynamics OCP (SURVEY 8d), n_segments=4096 per GPU, "
                               "poly_orders=15, LGR: one fused g + jac_g evaluation of 4096 segments per step and GPU",
                   "n_segments_total": Kt, "n_z": n_z, "n_g": n_g, "nnz_jac": nnz, "algorithmic_bytes_per_gpu": int(B),
                   "l2": f"{R} rotating z/g/values sets ({R * B / 1e6:.0f} MB written per GPU > 126 MB L2), launches back to back",
                   "parallelism": "1 GPU" if world == 1 else
                   f"one NLP of {Kt} segments, {K} per GPU over {world} GPUs; rows stay on the GPU that computed them "
                   "(no data-path collective); value = 4096-segment evaluations per second summed over the GPUs",
                   "program": tr.program_origin},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "mpx_gjac2_kernel<synthetic_6_3, JAC, 15>",
                     "launch_us_avg": ms_step * 1e3, "bytes_per_launch": int(B),
                     "isolated_launch_us_median": kmed * 1e3, "isolated_launch_us_min": float(kern_ms.min()) * 1e3,
                     "how": "achieved = algorithmic bytes / average launch duration over the timed region (back-to-back "
                            "launches, rotating output sets > L2; per GPU, slowest rank); isolated_* = single launches "
                            "after a 256 MB read that evicts L2 (includes launch latency and a cold start)"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(),
    }
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
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-allgather", action="store_true", help="N > 1: skip the extra all-gather measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
