# -*- coding: utf-8 -*-
"""Synthetic decoder outputs for the panoptic post-processing / evaluation path.

Shapes and statistics follow SURVEY.md section 8(d) (which mirrors the reference's GT
encoders, `data/preprocessing/instance.py:143-150, 223-256`, and the sigma = 8 centre
Gaussians of `tests/test_metrics.py:55`): blocky label map, N(0,1) logits with a +4
boost on the region's class, Gaussian centre heat-map with peaks of exactly 1.0,
normalised offsets to the nearest centre plus small noise, per-instance orientation.

Pure torch; runs on CPU (parity tests / golden vectors) and on CUDA (bench).
"""
from typing import Dict, Optional

import torch


def default_is_thing(n_classes_without_void: int):
    """odd class indices are things (about 50 %), SURVEY.md section 8(d)."""
    return tuple(bool(c % 2) for c in range(n_classes_without_void))


def make_frame(C: int, H: int, W: int, K: int, seed: int, with_orientation: bool = True,
               device='cpu', quantize: Optional[str] = 'q10') -> Dict[str, torch.Tensor]:
    """One synthetic frame. `quantize`: None (raw f32), 'q10' (logits on a 2^-10 grid:
    no sub-2^-23 gaps) or 'tie' (logits on a 0.25 grid, heat
    on a 1/16 grid, integer pixel offsets: exact ties everywhere)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)

    def randn(*shape):
        return torch.randn(*shape, generator=g, device=dev, dtype=torch.float32)

    # blocky label map -> logits
    hl, wl = (H + 31) // 32, (W + 31) // 32
    low = torch.randint(0, C, (hl, wl), generator=g, device=dev)
    labels = low.repeat_interleave(32, 0).repeat_interleave(32, 1)[:H, :W]
    logits = randn(C, H, W)
    logits.scatter_add_(0, labels[None], torch.full((1, H, W), 4.0, device=dev))
    if quantize == 'q10':
        logits = torch.round(logits * 1024.0) / 1024.0
    elif quantize == 'tie':
        logits = torch.round(logits * 4.0) / 4.0

    # centres >= 8 px away from the borders
    cy = torch.randint(8, max(H - 8, 9), (K,), generator=g, device=dev)
    cx = torch.randint(8, max(W - 8, 9), (K,), generator=g, device=dev)
    ys = torch.arange(H, device=dev, dtype=torch.float32)[:, None]
    xs = torch.arange(W, device=dev, dtype=torch.float32)[None, :]
    heat = torch.zeros(H, W, device=dev)
    best_d2 = torch.full((H, W), float('inf'), device=dev)
    near_y = torch.zeros(H, W, device=dev)
    near_x = torch.zeros(H, W, device=dev)
    near_i = torch.zeros(H, W, device=dev, dtype=torch.long)
    for i in range(K):
        d2 = (ys - cy[i].float()) ** 2 + (xs - cx[i].float()) ** 2
        heat = torch.maximum(heat, torch.exp(-d2 / (2.0 * 8.0 ** 2)))
        closer = d2 < best_d2
        best_d2 = torch.where(closer, d2, best_d2)
        near_y = torch.where(closer, cy[i].float().expand(H, W), near_y)
        near_x = torch.where(closer, cx[i].float().expand(H, W), near_x)
        near_i = torch.where(closer, torch.full_like(near_i, i), near_i)
    if quantize == 'tie':
        heat = torch.round(heat * 16.0) / 16.0

    off_y = near_y - ys
    off_x = near_x - xs
    if quantize == 'tie':
        offset = torch.stack((off_y / H, off_x / W))       # exact pixel targets -> ties
    else:
        offset = torch.stack((off_y / H + 0.002 * randn(H, W),
                              off_x / W + 0.002 * randn(H, W)))

    out = {'logits': logits.contiguous(), 'heat': heat[None].contiguous(),
           'offset': offset.contiguous()}
    if with_orientation:
        base = torch.rand(K, generator=g, device=dev) * 6.2831853
        ang = base[near_i] + 0.1 * randn(H, W)
        out['orientation'] = torch.stack((torch.cos(ang), torch.sin(ang))).contiguous()
    return out


def make_batch(B: int, C: int, H: int, W: int, K: int, seed: int = 0,
               with_orientation: bool = True, device='cpu',
               quantize: Optional[str] = 'q10') -> Dict[str, torch.Tensor]:
    """Batch of synthetic frames: logits (B,C,H,W), heat (B,1,H,W), offset (B,2,H,W),
    [orientation (B,2,H,W)], all f32 contiguous NCHW."""
    frames = [make_frame(C, H, W, K, seed * 1000 + b, with_orientation, device, quantize)
              for b in range(B)]
    return {k: torch.stack([f[k] for f in frames]).contiguous() for k in frames[0]}


def poison_logits(logits: torch.Tensor, fraction: float, seed: int = 0) -> torch.Tensor:
    """Overwrite the logits of about `fraction` of the pixels (in place) with the non-finite
    patterns that matter for `softmax -> max` (semantic.py:52-53): a NaN, a +Inf, +Inf next to
    -Inf, nothing but -Inf (all of them: every probability NaN, index 0), and -Inf next to finite
    logits (harmless: probability 0), each at a random class (or the first / the winning one)."""
    B, C, H, W = logits.shape
    g = torch.Generator()
    g.manual_seed(10_000 + seed)
    n = max(6, int(fraction * B * H * W))
    bs = torch.randint(0, B, (n,), generator=g)
    ys = torch.randint(0, H, (n,), generator=g)
    xs = torch.randint(0, W, (n,), generator=g)
    cs = torch.randint(0, C, (n,), generator=g)
    kinds = torch.arange(n) % 6
    inf = float('inf')
    for b, y, x, c, kind in zip(bs.tolist(), ys.tolist(), xs.tolist(), cs.tolist(), kinds.tolist()):
        px = logits[b, :, y, x]
        if kind == 0:
            px[c] = float('nan')
        elif kind == 1:
            px[c] = inf
        elif kind == 2:
            px[c] = inf
            px[(c + 1) % C] = -inf
        elif kind == 3:
            px[:] = -inf
        elif kind == 4:
            px[c] = -inf
        else:       # the winner itself becomes -Inf: the runner-up wins
            px[int(torch.argmax(torch.nan_to_num(px, nan=-inf)))] = -inf
    return logits


def saturate_heat(heat: torch.Tensor, step: int = 4, value: float = 1.0,
                  frames=None) -> torch.Tensor:
    """Overwrite the centre heat-maps (B,1,H,W) of `frames` (default: all) in place with a lattice
    of exactly tied peaks, one every `step` pixels: every peak equals the k-th value, so ALL of
    them become centres (instance.py:152-155) -- more than 255 for any frame larger than
    16 * step pixels squared, the regime in which the reference's uint8 ids wrap
    (instance.py:236)."""
    B = heat.shape[0]
    for b in (range(B) if frames is None else frames):
        heat[b].fill_(0.0)
        heat[b, 0, step // 2::step, step // 2::step] = value
    return heat


def make_eval_targets(panoptic: torch.Tensor, max_instances_per_category: int = 1 << 16,
                      shift: int = 5):
    """Evaluation targets: panoptic target = prediction rolled
    by `shift` px along W (gives a TP/FP/FN mix), semantic target uint8 = target // L."""
    tgt = torch.roll(panoptic, shifts=shift, dims=-1).contiguous()
    sem = (tgt // max_instances_per_category).to(torch.uint8)
    return tgt, sem


def make_batch_dict(B: int, H: int, W: int):
    """The minimal `batch` the reference's postprocess() needs (resize.py:30-71)."""
    return {
        'semantic_fullres': torch.zeros(B, H, W),
        'instance_fullres': torch.zeros(B, H, W),
        '_applied_preprocessing': [[{'type': 'Resize',
                                     'valid_region_slice_y': slice(0, H),
                                     'valid_region_slice_x': slice(0, W)}]] * B,
    }
