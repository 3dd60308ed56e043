#!/usr/bin/env python
"""Pinned-host <-> device copy bandwidth of the box (the ceiling of bench.py's `e2e` number).

    python scripts/pcie_bw.py
"""
import json

import torch


def main():
    dev = torch.device('cuda', 0)
    out = {}
    for mb in (64, 512, 2048):
        n = mb << 20
        h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        d = torch.empty(n, dtype=torch.uint8, device=dev)
        h2 = torch.empty(n // 16, dtype=torch.uint8, pin_memory=True)
        d2 = torch.empty(n // 16, dtype=torch.uint8, device=dev)
        s2 = torch.cuda.Stream(dev)
        for name, fn in (('h2d', lambda: d.copy_(h, non_blocking=True)),
                         ('d2h', lambda: h.copy_(d, non_blocking=True))):
            for _ in range(2):
                fn()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(3, (8 << 30) // n)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            out[f'{name}_{mb}MB_GBs'] = round(n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
        # h2d with a concurrent small d2h stream (what the pipeline does)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, (8 << 30) // n)
        e0.record()
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
        e1.record()
        torch.cuda.synchronize(dev)
        out[f'h2d_with_d2h_{mb}MB_GBs'] = round(n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
