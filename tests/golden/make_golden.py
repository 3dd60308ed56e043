# -*- coding: utf-8 -*-
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Runs only in the authoring container (needs /root/reference); the fixtures it writes
are committed and are what pins the oracle (and the CUDA path) on machines without the
reference.  The reference is imported through the stub packages in oracle/ref_stubs/
(stand-ins for the absent `torchmetrics` / `nicr_scene_analysis_datasets`, no arithmetic).

    python tests/golden/make_golden.py

Every fixture stores the inputs and the reference's outputs; python dict outputs are
stored as JSON strings.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/src')
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ref_stubs'))
os.environ.setdefault('OMP_NUM_THREADS', '1')

from nicr_mt_scene_analysis.metric import MeanIntersectionOverUnion  # noqa: E402
from nicr_mt_scene_analysis.metric.pq import compare_and_accumulate  # noqa: E402
from nicr_mt_scene_analysis.model.postprocessing import get_postprocessing_class  # noqa: E402
from nicr_mt_scene_analysis.utils.panoptic_merge import deeplab_merge_batch  # noqa: E402

from nicr_mt_scene_analysis_b200 import testing  # noqa: E402


def jdump(obj):
    return np.array(json.dumps(obj))


def dicts_to_json(list_of_dicts):
    return jdump([{str(k): v for k, v in d.items()} for d in list_of_dicts])


def meta_to_json(meta):
    return jdump([{str(k): {kk: (list(vv) if isinstance(vv, tuple) else vv)
                            for kk, vv in v.items()} for k, v in d.items()} for d in meta])


def run_postprocess(name, B, C, H, W, K, seed, quantize, with_orientation=True, top_k=64,
                    ks=3, thr=0.1, apply_fg=False, normalized=True, dist_thr=None,
                    compute_scores=False, poison=0.0, saturate=()):
    data = testing.make_batch(B, C, H, W, K, seed=seed, with_orientation=with_orientation,
                              quantize=quantize)
    if saturate:    # frames with hundreds of exactly tied peaks: > 255 centres, uint8 ids wrap
        testing.saturate_heat(data['heat'], step=4, frames=saturate)
        # the last of them with (nearly) zero offsets: every thing pixel joins the lattice centre
        # next to it, so all centres own pixels and the wrapped ids collide everywhere
        data['offset'][saturate[-1]] *= 0.01
    if poison:      # non-finite logits at a fraction of the pixels (semantic.py:52-53 on NaN / Inf)
        testing.poison_logits(data['logits'], poison, seed)
    if not normalized:      # offsets in pixels
        data['offset'][:, 0] *= H
        data['offset'][:, 1] *= W
    is_thing = testing.default_is_thing(C)
    has_ori = tuple(bool(t and (c % 4 == 1)) for c, t in enumerate(is_thing))
    sem = get_postprocessing_class('semantic')()
    ins = get_postprocessing_class('instance', heatmap_threshold=thr,
                                   heatmap_nms_kernel_size=ks, top_k_instances=top_k,
                                   heatmap_apply_foreground_mask=apply_fg,
                                   normalized_offset=normalized,
                                   offset_distance_threshold=dist_thr)()
    pan = get_postprocessing_class('panoptic', semantic_postprocessing=sem,
                                   instance_postprocessing=ins,
                                   semantic_classes_is_thing=is_thing,
                                   semantic_class_has_orientation=has_ori,
                                   normalized_offset=normalized,
                                   compute_scores=compute_scores)()
    inst_out = (data['heat'], data['offset']) + ((data['orientation'],) if with_orientation else ())
    batch = testing.make_batch_dict(B, H, W)
    r = pan.postprocess(((data['logits'], inst_out), (None, None)), batch, is_training=False)
    out = {
        'logits': data['logits'].numpy(), 'heat': data['heat'].numpy(),
        'offset': data['offset'].numpy(),
        'is_thing': np.array(is_thing), 'has_orientation': np.array(has_ori),
        'cfg': jdump(dict(top_k=top_k, ks=ks, thr=thr, apply_fg=apply_fg, normalized=normalized,
                          dist_thr=dist_thr, compute_scores=compute_scores)),
        'semantic_segmentation_idx': r['semantic_segmentation_idx'].numpy().astype(np.uint8),
        'semantic_segmentation_score': r['semantic_segmentation_score'].numpy(),
        'panoptic_foreground_mask': r['panoptic_foreground_mask'].numpy(),
        'panoptic_segmentation_deeplab': r['panoptic_segmentation_deeplab'].numpy(),
        'panoptic_segmentation_deeplab_semantic_idx':
            r['panoptic_segmentation_deeplab_semantic_idx'].numpy().astype(np.uint8),
        'panoptic_segmentation_deeplab_instance_idx':
            r['panoptic_segmentation_deeplab_instance_idx'].numpy(),
        'ids': dicts_to_json(r['panoptic_segmentation_deeplab_ids']),
        'meta': meta_to_json(r['panoptic_segmentation_deeplab_instance_meta']),
    }
    if with_orientation:
        out['orientation'] = data['orientation'].numpy()
        out['orientations'] = dicts_to_json(r['orientations_panoptic_segmentation_deeplab_instance'])
    if compute_scores:
        for k in ('semantic_score', 'instance_score', 'panoptic_score'):
            out[k] = r[f'panoptic_segmentation_deeplab_{k}'].numpy()
    # centres as the instance post-processing itself reports them
    fg = r['panoptic_foreground_mask']
    cmask, clist = ins._get_instance_centers(data['heat'], fg)
    out['center_mask'] = cmask.numpy()
    out['centers'] = jdump([c.tolist() for c in clist])
    np.savez_compressed(os.path.join(HERE, f'post_{name}.npz'), **out)
    print(name, 'centres/frame', [len(c) for c in clist],
          'instances', [len(d) for d in r['panoptic_segmentation_deeplab_ids']])


def run_fullres():
    """network resolution != dataset resolution: crop to the valid region, then resize
    (dense_base.py:15-58; nearest for index maps, bilinear on the logits before a second
    softmax/arg-max, semantic.py:63-72 and panoptic.py:242-291)."""
    B, C, H, W, K = 2, 6, 96, 128, 4
    FH, FW = 150, 190
    data = testing.make_batch(B, C, H, W, K, seed=6, with_orientation=True, quantize='q10')
    is_thing = testing.default_is_thing(C)
    has_ori = tuple(bool(t and (c % 4 == 1)) for c, t in enumerate(is_thing))
    sem = get_postprocessing_class('semantic')()
    ins = get_postprocessing_class('instance')()
    pan = get_postprocessing_class('panoptic', semantic_postprocessing=sem,
                                   instance_postprocessing=ins,
                                   semantic_classes_is_thing=is_thing,
                                   semantic_class_has_orientation=has_ori)()
    sl_y, sl_x = slice(0, 90), slice(4, 124)
    batch = {'semantic_fullres': torch.zeros(B, FH, FW), 'instance_fullres': torch.zeros(B, FH, FW),
             '_applied_preprocessing': [[{'type': 'Resize', 'valid_region_slice_y': sl_y,
                                          'valid_region_slice_x': sl_x}]] * B}
    r = pan.postprocess(((data['logits'], (data['heat'], data['offset'], data['orientation'])),
                         (None, None)), batch, is_training=False)
    out = {'logits': data['logits'].numpy(), 'heat': data['heat'].numpy(),
           'offset': data['offset'].numpy(), 'orientation': data['orientation'].numpy(),
           'is_thing': np.array(is_thing), 'has_orientation': np.array(has_ori),
           'fullres_shape': np.array([FH, FW]), 'valid_y': np.array([0, 90]),
           'valid_x': np.array([4, 124])}
    for k in ('semantic_segmentation_idx_fullres', 'panoptic_segmentation_deeplab_fullres',
              'panoptic_segmentation_deeplab_instance_idx_fullres',
              'panoptic_segmentation_deeplab_semantic_idx_fullres',
              'semantic_segmentation_score_fullres', 'semantic_output_fullres'):
        out[k] = r[k].numpy()
    np.savez_compressed(os.path.join(HERE, 'fullres.npz'), **out)
    print('fullres', {k: v.shape for k, v in out.items() if k.endswith('fullres')})


def run_centers():
    """tie-heavy heat-maps straight into _get_instance_centers (instance.py:78-168)."""
    g = torch.Generator().manual_seed(7)
    H, W = 40, 52
    cases = []
    for ks, k, thr, q in ((3, 7, 0.1, 8.0), (5, 4, 0.3, 4.0), (3, 64, 0.1, 1024.0),
                          (1, 5, 0.5, 8.0), (3, 3, -0.5, 2.0), (7, 2, 0.05, 16.0)):
        heat = torch.rand(3, 1, H, W, generator=g)
        heat = torch.round(heat * q) / q
        if thr < 0:
            heat[:, :, 0, 0] = 0.0           # the (0,0) corner case
            heat[1] = -heat[1]
        fgm = torch.rand(3, H, W, generator=g) > 0.3
        for apply_fg in (False, True):
            post = get_postprocessing_class('instance', heatmap_threshold=thr,
                                            heatmap_nms_kernel_size=ks, top_k_instances=k,
                                            heatmap_apply_foreground_mask=apply_fg)()
            cmask, clist = post._get_instance_centers(heat.clone(), fgm)
            cases.append(dict(ks=ks, k=k, thr=thr, apply_fg=apply_fg, heat=heat.numpy().tolist(),
                              fg=fgm.numpy().astype(int).tolist(),
                              mask=cmask.numpy().astype(int).tolist(),
                              centers=[c.tolist() for c in clist]))
            print('centers', ks, k, thr, apply_fg, [len(c) for c in clist])
    np.savez_compressed(os.path.join(HERE, 'centers.npz'), cases=jdump(cases))


def blocky(g, B, H, W, n_values, block):
    low = torch.randint(0, n_values, (B, (H + block - 1) // block, (W + block - 1) // block),
                        generator=g)
    return low.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :H, :W].contiguous()


def run_merge():
    """deeplab_merge_batch as a stand-alone API (panoptic_merge.py:18-40, 172-225), with
    void (0) in the semantic map and a foreground mask that disagrees with it."""
    g = torch.Generator().manual_seed(11)
    B, H, W = 4, 48, 64
    sem = blocky(g, B, H, W, 7, 8)                       # 0 = void, 1..6
    # sprinkle noise so majority votes have close calls / exact ties
    noise = torch.randint(0, 7, (B, H, W), generator=g)
    sem = torch.where(torch.rand(B, H, W, generator=g) < 0.3, noise, sem)
    ins = blocky(g, B, H, W, 9, 12).to(torch.uint8)
    fg = blocky(g, B, H, W, 4, 6) > 0
    thing_ids = np.array([2, 3, 5])
    L = 1 << 16
    pan, ids = deeplab_merge_batch(sem, ins, fg, L, thing_ids, 0)
    pan2, ids2 = deeplab_merge_batch(sem, ins, fg, 1000, thing_ids, 3)
    np.savez_compressed(os.path.join(HERE, 'merge.npz'), sem=sem.numpy(), ins=ins.numpy(),
                        fg=fg.numpy(), thing_ids=thing_ids, L=L, pan=pan.numpy(),
                        ids=dicts_to_json(ids), pan_L1000_void3=pan2.numpy(),
                        ids_L1000_void3=dicts_to_json(ids2))
    print('merge', [len(d) for d in ids])


def run_naive_merge():
    """ground-truth targets: naive_merge_semantic_and_instance_np (panoptic_merge.py:43-107),
    the body of PanopticTargetGenerator; uint16 instance ids up to 65535, instances that span
    several classes, void inside instances."""
    from nicr_mt_scene_analysis.utils.panoptic_merge import naive_merge_semantic_and_instance_np
    g = torch.Generator().manual_seed(23)
    B, H, W = 3, 60, 84
    sem = blocky(g, B, H, W, 9, 9).numpy().astype(np.uint8)
    ins = blocky(g, B, H, W, 12, 14).numpy().astype(np.uint16)
    ins[ins == 7] = 40000
    ins[ins == 11] = 65535
    thing_ids = np.array([2, 3, 5, 8])
    pans, dicts = [], []
    for b in range(B):
        pan, d = naive_merge_semantic_and_instance_np(sem[b], ins[b], 1 << 16, thing_ids, 0)
        pans.append(pan.astype(np.int64))
        dicts.append(d)
    np.savez_compressed(os.path.join(HERE, 'naive_merge.npz'), sem=sem, ins=ins.astype(np.int32),
                        thing_ids=thing_ids, pan=np.stack(pans), ids=dicts_to_json(dicts))
    print('naive merge', [len(d) for d in dicts])


def run_instance_targets():
    """InstanceTargetGenerator (data/preprocessing/instance.py:97-286) on random blocky GT:
    uint16 instance ids, instances whose majority class is stuff (skipped), exact bincount ties,
    centres near the borders (clipped stamps)."""
    from nicr_mt_scene_analysis.data.preprocessing.instance import InstanceTargetGenerator
    g = torch.Generator().manual_seed(29)
    B, H, W = 3, 72, 100
    is_thing = (False, True, False, True, True, False)          # with void
    out = {'is_thing': np.array(is_thing)}
    sems, inss = [], []
    res = {k: [] for k in ('instance_center', 'instance_offset', 'instance_foreground',
                           'instance_center_mask')}
    res_px = []
    enc, skip = [], []
    for b in range(B):
        ins = blocky(g, 1, H, W, 9, 13)[0].numpy().astype(np.uint16)
        ins[ins == 5] = 51234
        sem = np.zeros((H, W), np.uint8)
        for i in np.unique(ins):                     # mostly one class per instance, some noise
            cls = int(torch.randint(1, 6, (1,), generator=g))
            sem[ins == i] = cls
        noise = torch.rand(H, W, generator=g).numpy() < 0.2
        sem[noise] = torch.randint(0, 6, (int(noise.sum()),), generator=g).numpy().astype(np.uint8)
        ins[ins == 0] = 0
        # make the assert of the reference hold: stuff-majority instances are cleared
        for i in np.unique(ins):
            if i and not is_thing[int(np.bincount(sem[ins == i]).argmax())]:
                ins[ins == i] = 0
        sems.append(sem); inss.append(ins)
        for norm, store in ((True, res), (False, None)):
            gen = InstanceTargetGenerator(sigma=5, semantic_classes_is_thing=is_thing,
                                          normalized_offset=norm)
            s = gen({'semantic': sem.copy(), 'instance': ins.copy()})
            if norm:
                for k in res:
                    res[k].append(s[k])
            else:
                res_px.append(s['instance_offset'])
        enc.append(sorted(int(i) for i in np.unique(ins) if i))
    for k in res:
        out[k] = np.stack(res[k])
    out['instance_offset_px'] = np.stack(res_px)
    out['sem'] = np.stack(sems); out['ins'] = np.stack(inss).astype(np.int32)
    out['encoded'] = jdump(enc)
    np.savez_compressed(os.path.join(HERE, 'instance_targets.npz'), **out)
    print('instance targets', [len(e) for e in enc], out['instance_center'].max())


def run_pq():
    """compare_and_accumulate on random blocky panoptic maps (pq.py:60-179)."""
    g = torch.Generator().manual_seed(13)
    B, H, W = 6, 64, 80
    L, OFF, NC = 1 << 16, 256 ** 3, 7
    cat_t = blocky(g, B, H, W, NC, 16)
    inst_t = blocky(g, B, H, W, 3, 8)
    cat_p = torch.where(torch.rand(B, H, W, generator=g) < 0.15,
                        blocky(g, B, H, W, NC, 16), cat_t)
    inst_p = torch.roll(inst_t, 2, dims=-1)
    is_thing = torch.tensor([False, True, False, True, True, False, True])
    tgt = cat_t * L + torch.where(is_thing[cat_t], inst_t + 1, torch.zeros_like(inst_t))
    tgt = torch.where((cat_t == 0) & (inst_t == 2), torch.full_like(tgt, 5), tgt)  # ignored w/ id>0
    pred = cat_p * L + torch.where(is_thing[cat_p], inst_p + 1, torch.zeros_like(inst_p))
    res = {'iou': [], 'tp': [], 'fn': [], 'fp': [], 'matches': []}
    for b in range(B):
        iou, tp, fn, fp, m = compare_and_accumulate(pred[b], tgt[b], NC, 0, L, OFF, 0)
        res['iou'].append(iou.numpy().copy()); res['tp'].append(tp.numpy().copy())
        res['fn'].append(fn.numpy().copy()); res['fp'].append(fp.numpy().copy())
        res['matches'].append(sorted([list(x) for x in m]))
    state = [torch.zeros(NC, dtype=torch.float64) for _ in range(4)]
    for b in range(B):      # PanopticQuality.update accumulation order  pq.py:298-303
        for s, k in zip(state, ('iou', 'tp', 'fn', 'fp')):
            s += torch.from_numpy(res[k][b])
    np.savez_compressed(
        os.path.join(HERE, 'pq.npz'), pred=pred.numpy(), target=tgt.numpy(),
        is_thing=is_thing.numpy(), L=L, offset=OFF, num_categories=NC,
        iou=np.stack(res['iou']), tp=np.stack(res['tp']), fn=np.stack(res['fn']),
        fp=np.stack(res['fp']), matches=jdump(res['matches']),
        state=np.stack([s.numpy() for s in state]))
    print('pq tp', np.stack(res['tp']).sum(0), 'fp', np.stack(res['fp']).sum(0),
          'fn', np.stack(res['fn']).sum(0))


def run_miou():
    """MeanIntersectionOverUnion.update / compute (miou.py:44-94)."""
    g = torch.Generator().manual_seed(17)
    out = {}
    for n in (6, 41, 200):
        m0 = MeanIntersectionOverUnion(n_classes=n, ignore_first_class=False)
        m1 = MeanIntersectionOverUnion(n_classes=n, ignore_first_class=True)
        preds, tgts = [], []
        for _ in range(2):
            t = torch.randint(0, n - 1, (3, 32, 40), generator=g)     # last class never in GT
            p = torch.where(torch.rand(3, 32, 40, generator=g) < 0.7, t,
                            torch.randint(0, n, (3, 32, 40), generator=g))
            m0.update(p, t.to(torch.uint8) if n <= 255 else t)
            m1.update(p, t)
            preds.append(p); tgts.append(t)
        miou0, ious0 = m0.compute(return_ious=True)
        miou1, ious1 = m1.compute(return_ious=True)
        out[f'pred_{n}'] = torch.stack(preds).numpy()
        out[f'target_{n}'] = torch.stack(tgts).numpy()
        out[f'confmat_{n}'] = m0.confmat.numpy()
        out[f'miou0_{n}'] = miou0.numpy(); out[f'ious0_{n}'] = ious0.numpy()
        out[f'miou1_{n}'] = miou1.numpy(); out[f'ious1_{n}'] = ious1.numpy()
    np.savez_compressed(os.path.join(HERE, 'miou.npz'), **out)
    print('miou ok')


def run_orientation():
    """_get_instance_orientation stand-alone with a GT-style int32 instance map and
    foreground masks / None (instance.py:270-319)."""
    g = torch.Generator().manual_seed(19)
    B, H, W = 3, 40, 56
    seg = blocky(g, B, H, W, 6, 10).to(torch.int32)
    seg[seg == 5] = 300                                   # ids beyond uint8
    ang = torch.rand(B, H, W, generator=g) * 0.5 + seg.float()
    ori = torch.stack((torch.cos(ang), torch.sin(ang)), 1).contiguous()
    mask = blocky(g, B, H, W, 3, 7) > 0
    post = get_postprocessing_class('instance')()
    r_mask = post._get_instance_orientation(ori, seg, mask)
    r_none = post._get_instance_orientation(ori, seg, None)
    np.savez_compressed(os.path.join(HERE, 'orientation.npz'), ori=ori.numpy(), seg=seg.numpy(),
                        mask=mask.numpy(), with_mask=dicts_to_json(r_mask),
                        without_mask=dicts_to_json(r_none))
    print('orientation', [len(d) for d in r_mask])


def run_task_helpers():
    """Validation path of the reference's task helpers (task_helper/panoptic.py:87-212,
    task_helper/instance.py:289-436): two validation steps + epoch end on random blocky
    batches.  batch_idx != 0 skips the visualisation examples; the instance helper's loss
    computation (not on the evaluated path) is bypassed by overriding `_compute_losses`."""
    from types import SimpleNamespace
    from nicr_mt_scene_analysis.task_helper.instance import InstanceTaskHelper
    from nicr_mt_scene_analysis.task_helper.panoptic import PanopticTaskHelper

    g = torch.Generator().manual_seed(29)
    B, H, W, NC, L = 3, 64, 96, 7, 1 << 16
    is_thing = [False, True, False, True, True, False, True]     # with void
    it = torch.tensor(is_thing)
    labels = SimpleNamespace(colors=[(i, 2 * i, 3 * i) for i in range(NC)],
                             classes_is_thing=is_thing,
                             colors_array=np.zeros((NC, 3), np.uint8))

    class InstanceHelperNoLoss(InstanceTaskHelper):
        def _compute_losses(self, batch, batch_idx, predictions_post):
            self._with_orientation = 'orientations_present' in batch
            return {}

    pan_helper = PanopticTaskHelper(NC, is_thing, labels)
    ins_helper = InstanceHelperNoLoss(NC, is_thing)
    pan_helper.initialize(torch.device('cpu'))
    ins_helper.initialize(torch.device('cpu'))

    steps = []
    for step in range(2):
        cat_t = blocky(g, B, H, W, NC, 16)
        inst_t = blocky(g, B, H, W, 4, 8) + 1
        inst_gt = torch.where(it[cat_t], inst_t + 4 * (cat_t // 2), torch.zeros_like(inst_t))
        pan_t = cat_t * L + inst_gt
        # prediction of the panoptic branch: noisy classes, shifted instances
        cat_p = torch.where(torch.rand(B, H, W, generator=g) < 0.12, blocky(g, B, H, W, NC, 16), cat_t)
        inst_p = torch.where(it[cat_p], torch.roll(inst_t, 2, -1), torch.zeros_like(inst_t))
        pan_p = cat_p * L + inst_p
        # prediction of the instance branch inside the GT foreground (raw centre ids, uint8)
        inst_fg = torch.where(inst_gt > 0, torch.roll(inst_gt, 1, -2) % 7 + 1,
                              torch.zeros_like(inst_gt)).to(torch.uint8)
        tgt_ids, pred_ids, ori_t, ori_p, ori_fg, ori_full_gt = [], [], [], [], [], []
        for b in range(B):
            tgt_ids.append({int(v): int(v) % L for v in torch.unique(pan_t[b]) if int(v) % L})
            pred_ids.append({int(v): int(v) % L + 10 for v in torch.unique(pan_p[b]) if int(v) % L})
            # orientations for every second GT instance; predictions for most raw ids
            ori_t.append({i: float(0.37 * i + 0.1 * b + step) for i in set(tgt_ids[b].values())
                          if i % 2 == 0})
            ori_p.append({i: float(0.37 * (i - 10) + 0.25 - 0.05 * b) for i in set(pred_ids[b].values())
                          if i % 3})
            ori_fg.append({i: float(0.2 * i - 1.0 + step) for i in range(1, 8)})
            ori_full_gt.append({i: a + 0.125 for i, a in ori_t[b].items()})
        batch = {'panoptic_fullres': pan_t, 'semantic_fullres': cat_t.to(torch.uint8),
                 'instance_fullres': inst_gt.to(torch.int32),
                 'panoptic_ids_to_instance_dict': tgt_ids, 'orientations_present': ori_t}
        pan_post = {'panoptic_segmentation_deeplab_fullres': pan_p,
                    'panoptic_segmentation_deeplab_ids': pred_ids,
                    'orientations_panoptic_segmentation_deeplab_instance': ori_p}
        ins_post = {'instance_segmentation_gt_foreground_fullres': inst_fg,
                    'orientations_instance_segmentation_gt_orientation_foreground': ori_fg,
                    'orientations_gt_instance_gt_orientation_foreground': ori_full_gt}
        pan_helper.validation_step(batch, 1 + step, pan_post)
        ins_helper.validation_step(batch, 1 + step, ins_post)
        steps.append(dict(pan_t=pan_t.numpy(), sem_t=cat_t.to(torch.uint8).numpy(),
                          inst_gt=inst_gt.to(torch.int32).numpy(), pan_p=pan_p.numpy(),
                          inst_fg=inst_fg.numpy(), tgt_ids=tgt_ids, pred_ids=pred_ids, ori_t=ori_t,
                          ori_p=ori_p, ori_fg=ori_fg, ori_full_gt=ori_full_gt))

    def pack(result):
        artifacts, examples, logs = result
        assert examples == {}
        out = {}
        for kind, d in (('artifacts', artifacts), ('logs', logs)):
            for k, v in d.items():
                if k.endswith('_time'):
                    continue
                out[f'{kind}/{k}'] = torch.as_tensor(v).double().numpy()
        return out
    pan_out = pack(pan_helper.validation_epoch_end())
    ins_out = pack(ins_helper.validation_epoch_end())
    arrays = {'is_thing': np.array(is_thing), 'num_categories': NC, 'n_steps': len(steps)}
    for i, st in enumerate(steps):
        for k in ('pan_t', 'sem_t', 'inst_gt', 'pan_p', 'inst_fg'):
            arrays[f'step{i}/{k}'] = st[k]
        for k in ('tgt_ids', 'pred_ids', 'ori_t', 'ori_p', 'ori_fg', 'ori_full_gt'):
            arrays[f'step{i}/{k}'] = dicts_to_json(st[k])
    for k, v in pan_out.items():
        arrays[f'panoptic/{k}'] = v
    for k, v in ins_out.items():
        arrays[f'instance/{k}'] = v
    np.savez_compressed(os.path.join(HERE, 'task_helpers.npz'), **arrays)
    print('task helpers: panoptic all_pq', pan_out['logs/panoptic_all_deeplab_pq'],
          'mae', pan_out['logs/panoptic_mae_deeplab_rad'],
          '| instance all_pq', ins_out['logs/instance_all_deeplab_pq'],
          'mae_gt', ins_out['logs/orientation_mae_gt_rad'])


def run_semantic_helper():
    """Validation path of the reference's SemanticTaskHelper (task_helper/semantic.py:111-161):
    three validation steps + epoch end.  Predictions are network classes 0..C-1 (int64), the
    target is uint8 with 0 = void; the helper masks the void pixels and shifts the target.
    batch_idx != 0 skips the visualisation examples; the loss (not on the evaluated path) is
    bypassed by overriding `_compute_losses`."""
    from nicr_mt_scene_analysis.task_helper.semantic import SemanticTaskHelper

    class SemanticHelperNoLoss(SemanticTaskHelper):
        def _compute_losses(self, batch, batch_idx, predictions_post):
            return {}

    g = torch.Generator().manual_seed(31)
    B, H, W, C = 3, 75, 101, 9          # 75*101 is odd: the last group of the map is partial
    helper = SemanticHelperNoLoss(n_classes=C)
    helper.initialize(torch.device('cpu'))
    arrays = {'n_classes': C, 'n_steps': 3}
    for step in range(3):
        target = blocky(g, B, H, W, C + 1, 12)               # 0 = void
        target[:, :, -7:] = 0                                # a void border (invalid region)
        clean = (target - 1).clamp(min=0)
        noisy = torch.rand(B, H, W, generator=g) < 0.15
        preds = torch.where(noisy, torch.randint(0, C, (B, H, W), generator=g), clean)
        if step == 2:
            target = target[:1].clone(); preds = preds[:1].clone()
            target[0, :40] = 0                               # class C-1 may lose its ground truth
            target[target == C] = 1
        batch = {'semantic_fullres': target.to(torch.uint8)}
        post = {'semantic_segmentation_idx_fullres': preds}
        helper.validation_step(batch, 1 + step, post)
        arrays[f'step{step}/target'] = target.to(torch.uint8).numpy()
        arrays[f'step{step}/preds'] = preds.numpy()
    artifacts, examples, logs = helper.validation_epoch_end()
    assert examples == {}
    arrays['artifacts/semantic_cm'] = artifacts['semantic_cm'].numpy()
    arrays['artifacts/semantic_ious_per_class'] = artifacts['semantic_ious_per_class'].numpy()
    arrays['logs/semantic_miou'] = logs['semantic_miou'].numpy()
    np.savez_compressed(os.path.join(HERE, 'semantic_helper.npz'), **arrays)
    print('semantic helper: miou', float(logs['semantic_miou']),
          'pixels counted', int(artifacts['semantic_cm'].sum()))


if __name__ == '__main__':
    torch.set_num_threads(1)
    if len(sys.argv) > 1:       # regenerate selected fixtures only: make_golden.py task_helpers pq
        for name in sys.argv[1:]:
            if name == 'post_wrap':
                run_postprocess('wrap', B=3, C=6, H=72, W=96, K=5, seed=7, quantize='q10',
                                saturate=(1, 2), dist_thr=30)
            else:
                globals()[f'run_{name}']()
        sys.exit(0)
    run_postprocess('q10', B=3, C=8, H=96, W=128, K=5, seed=1, quantize='q10')
    run_postprocess('tie', B=3, C=6, H=96, W=128, K=6, seed=2, quantize='tie', top_k=3)
    run_postprocess('odd', B=2, C=5, H=75, W=91, K=4, seed=3, quantize='q10', ks=5,
                    apply_fg=True, dist_thr=20, top_k=8)
    run_postprocess('pixel_offsets', B=2, C=7, H=64, W=80, K=4, seed=4, quantize='q10',
                    normalized=False, with_orientation=False)
    run_postprocess('scores', B=2, C=6, H=64, W=96, K=4, seed=5, quantize='q10',
                    compute_scores=True)
    run_postprocess('nonfinite', B=2, C=6, H=48, W=64, K=4, seed=6, quantize='q10', poison=0.05)
    run_postprocess('wrap', B=3, C=6, H=72, W=96, K=5, seed=7, quantize='q10', saturate=(1, 2),
                    dist_thr=30)
    run_fullres()
    run_centers()
    run_merge()
    run_naive_merge()
    run_instance_targets()
    run_pq()
    run_miou()
    run_orientation()
    run_task_helpers()
    run_semantic_helper()
