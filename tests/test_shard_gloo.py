"""Multi-process (world_size 2 and 3, gloo, CPU) test of the segment sharding + all-gather logic.

Each rank fills ONLY the runs its segment range owns (values taken from the CPU oracle), gathers, and must end
up with the complete g / Jacobian-value vectors."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpopt_b200 import shard as sh
from mpopt_b200.layout import Layout
from mpopt_b200.problems import REGISTRY
from mpopt_b200.program import Program
from oracle.nlp import OracleNLP


def _setup(problem, K, po):
    ocp = REGISTRY[problem]()
    pol = [po] * K if isinstance(po, int) else list(po)
    ora = OracleNLP(ocp, K, po, "LGR", drop_exact_zeros=False)
    lay = Layout(Program(ocp), pol, [bool(v) for v in ocp.diff_u], [r["has_mU"] for r in ora._rows],
                 [bool(v) for v in ocp.du_continuity], ora.n_links)
    z = np.random.default_rng(0).uniform(0.5, 1.5, ora.n_z)
    return ora, lay, pol, z


def _worker(rank, world, port, problem, K, po, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ora, lay, pol, z = _setup(problem, K, po)
        J = ora.jac_g(z)
        g_full, v_full = torch.from_numpy(ora.g(z)), torch.from_numpy(J.data.copy())
        assert lay.nnz_full == len(v_full)
        part = sh.partition(pol, world)
        g = torch.full_like(g_full, float("nan"))
        v = torch.full_like(v_full, float("nan"))
        for kind, buf, full in ((0, g, g_full), (1, v, v_full)):
            for off, cnt in lay.shard_runs(kind, *part[rank]):
                buf[off:off + cnt] = full[off:off + cnt]

        def row0(gb, vb):  # what a one-segment plan on this rank would write
            for kind, buf, full in ((0, gb, g_full), (1, vb, v_full)):
                for off, cnt in lay.shard_runs(kind, 0, 1):
                    if off + cnt <= len(buf):
                        buf[off:off + cnt] = full[off:off + cnt]

        gat = sh.Gatherer(lay, part, dist, rank, None, row0)
        gat.all_gather(g, v)
        ok = bool(torch.equal(g, g_full) and torch.equal(v, v_full))
        q.put((rank, gat.mode, ok))
    finally:
        dist.destroy_process_group()


def _objective_worker(rank, world, port, problem, K, po, q):
    """Every rank holds its shard of grad_f (node entries of the nodes it owns) and partial sums of J and of the
    entries all nodes contribute to; after ObjectiveGatherer.all_gather every rank holds the oracle's f and grad_f."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ora, lay, pol, z = _setup(problem, K, po)
        f_full, grad_full = float(ora.f(z)), torch.from_numpy(ora.grad_f(z).copy())
        part = sh.partition(pol, world)
        og = sh.ObjectiveGatherer(lay, part, dist, rank)
        grad = torch.full_like(grad_full, float("nan"))
        for off, cnt in lay.shard_runs(2, *part[rank]):
            grad[off:off + cnt] = grad_full[off:off + cnt]
        # partial sums: a fixed, rank-dependent split of every summed entry that adds up to the full value
        wts = torch.tensor([(r + 1.0) for r in range(world)], dtype=torch.float64)
        share = float(wts[rank] / wts.sum())
        idx = torch.as_tensor(og.glob)
        nx, per = lay.nx, len(og.glob) // lay.P
        x0 = torch.zeros(len(og.glob), dtype=torch.bool)
        for ph in range(lay.P):
            x0[ph * per + per - nx:(ph + 1) * per] = True
        grad[idx[~x0]] = grad_full[idx[~x0]] * share
        # x0 entries: the owner of node 0 holds the running-cost part, the last shard the Mayer part (here: a split)
        own0, tail = part[rank][0] == 0, part[rank][1] == lay.K
        if own0 and tail:
            grad[idx[x0]] = grad_full[idx[x0]]
        elif own0:
            grad[idx[x0]] = 0.25 * grad_full[idx[x0]]
        elif tail:
            grad[idx[x0]] = 0.75 * grad_full[idx[x0]]
        f = torch.tensor([f_full * share], dtype=torch.float64)
        og.all_gather(f, grad)
        ok = bool(torch.allclose(grad, grad_full, rtol=1e-14, atol=1e-14) and abs(float(f[0]) - f_full) <= 1e-13 * max(1, abs(f_full)))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("problem,K,po,world", [
    ("moon_lander", 8, 3, 2), ("kitchen_sink", 5, [3, 4, 2, 6, 3], 3), ("two_phase_schwartz", 4, 3, 2)])
def test_objective_partials_and_grad_shards_gather(problem, K, po, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_objective_worker, args=(r, world, port, problem, K, po, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert all(r[1] for r in res), res


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("problem,K,po,world,mode", [
    ("moon_lander", 8, 3, 2, "inplace"),
    ("synthetic_6_3", 6, 4, 3, "inplace"),
    ("van_der_pol", 7, [3, 6, 3, 2, 6, 3, 4], 2, "packed"),
    ("two_phase_schwartz", 5, 3, 2, "packed"),
])
def test_sharded_gather_reconstructs_full_outputs(problem, K, po, world, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, problem, K, po, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert [r[0] for r in res] == list(range(world))
    assert all(r[1] == mode for r in res), res
    assert all(r[2] for r in res), res


def test_partition_balances_mixed_degrees():
    po = [30 if k % 3 == 1 else 3 for k in range(2048)]
    for world in (2, 4, 8):
        part = sh.partition(po, world)
        assert part[0][0] == 0 and part[-1][1] == 2048
        assert all(a[1] == b[0] for a, b in zip(part, part[1:]))
        cost = [sum(p * (p + 1) for p in po[b:e]) for b, e in part]
        assert max(cost) / min(cost) < 1.02
    assert sh.partition([15] * 4096, 8)[3] == (1536, 2048)


def test_shard_runs_cover_everything_once():
    for problem, K, po in (("kitchen_sink", 5, [3, 4, 2, 6, 3]), ("synthetic_6_3", 8, 5)):
        ora, lay, pol, _ = _setup(problem, K, po)
        for world in (1, 2, 4):
            part = sh.partition(pol, world)
            for kind, total in ((0, lay.n_g), (1, lay.nnz_full)):
                cover = np.zeros(total, int)
                for kb, ke in part:
                    for off, cnt in lay.shard_runs(kind, kb, ke):
                        cover[off:off + cnt] += 1
                assert (cover == 1).all()
