#!/usr/bin/env python
"""Host cost of the MAAE part of a validation step (64 frames x 20 matched instances with an
orientation): the reference's form -- six scalar tensor operations and one update of the
(device) state per matched pair, mae.py:157-162 -- against the batched form of metric/mae.py."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import random
    import torch
    from nicr_mt_scene_analysis_b200.metric import MeanAbsoluteAngularError
    from nicr_mt_scene_analysis_b200.metric.mae import abs_angle_error_rad
    dev = torch.device('cuda', 0)
    rnd = random.Random(1)
    preds = [{i: rnd.uniform(-3, 3) for i in range(20)} for _ in range(64)]
    targets = [{i: rnd.uniform(-3, 3) for i in range(20)} for _ in range(64)]
    m = MeanAbsoluteAngularError(device=dev)

    def per_pair():
        for p, t in zip(preds, targets):
            for k, a in p.items():
                err = abs_angle_error_rad(torch.tensor(a), torch.tensor(t[k]))
                m.sum_angular_error += err.to(dev)
                m.n_elements += 1

    def batched():
        m.update(preds, targets)

    for name, fn in (('per pair (reference form)', per_pair), ('batched', batched)):
        fn()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize(dev)
        print(f'{name}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per 64-frame step (1280 pairs)')


if __name__ == '__main__':
    main()
