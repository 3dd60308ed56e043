"""Probe: what does the boundary between two CUDA-graph replays of the step cost?
  a) nodes of the captured step (debug dump)
  b) replay rate of graphs with 1 / 4 trivial kernels
  c) the step captured once vs twice per graph (inside one graph consecutive steps are chained by
     programmatic dependent launches, across replays they are not)
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench

dev = torch.device('cuda:0')
torch.cuda.set_device(0)


def timed(fn, n):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


x = torch.zeros(32, device=dev)
for k in (1, 4):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        x.add_(1)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(k):
            x.add_(1)
    print(f'graph of {k} trivial kernel(s): {timed(g.replay, 2000):.2f} us per replay')

w = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'nyuv2'])
arm = bench.Arm(w, w['B'], dev, 0, fused=True, graph=False)
for n_steps in (1, 2, 4):
    for _ in range(3):
        arm.eager_step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    if n_steps == 1:
        g.enable_debug_mode()
    with torch.cuda.graph(g):
        for _ in range(n_steps):
            r = arm.eager_step()
    if n_steps == 1:
        os.makedirs('gpurun_out', exist_ok=True)
        g.debug_dump('gpurun_out/step_graph.dot')
    us = timed(g.replay, 500)
    print(f'{w["name"]}: graph of {n_steps} step(s): {us / n_steps:.1f} us per step')
arm.pq.check_status()
