"""Shared timing helper of the profiles/tools scripts: device time of one evaluator call, two ways.

host_us  -- n back-to-back calls issued from Python between two CUDA events (includes whatever the host needs per call:
            small kernels are HOST-bound this way, ~6-9 us per ctypes call + launch);
graph_us -- the same calls captured once into a CUDA graph, one replay between two events: what the GPU needs."""
import torch


def time_call(fn, stream, n=50, warm=5, graph=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    host_us = e0.elapsed_time(e1) * 1e3 / n
    graph_us = None
    if graph:
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for _ in range(n):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            e0.record(stream)
            g.replay()
            e1.record(stream)
            torch.cuda.synchronize()
            graph_us = e0.elapsed_time(e1) * 1e3 / n
        except Exception as ex:  # an evaluator that synchronises or allocates cannot be captured
            graph_us = None
            print("# graph capture failed:", str(ex)[:120])
            torch.cuda.synchronize()
    return host_us, graph_us
