"""Debug probe: the non-evaluating forward chain on a small frame, many times; which output goes wrong?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from nicr_mt_scene_analysis_b200 import testing
from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class

dev = torch.device('cuda:0')
from nicr_mt_scene_analysis_b200 import _lib
try:
    print('build info:', _lib.lib().npb_build_info().decode())
except Exception as e:
    print('no build info', e)
for (H, W) in ((64, 96), (75, 91)):
    B, C, K = 2, 7, 4
    d = testing.make_batch(B, C, H, W, K, seed=23, with_orientation=False, device=dev, quantize='q10')
    is_thing = testing.default_is_thing(C)
    post = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=get_postprocessing_class('instance')(),
        semantic_classes_is_thing=is_thing, semantic_class_has_orientation=(False,) * C)()
    first = None
    shown = 0
    nbad = 0
    for it in range(100):
        if it % 3 == 0:     # garbage into recycled buffers
            junk = torch.full((B * H * W * 16,), 0x97, dtype=torch.uint8, device=dev); del junk
        r = post.postprocess(((d['logits'], (d['heat'], d['offset'])), (None, None)),
                             testing.make_batch_dict(B, H, W), is_training=False)
        tabs = r['_panoptic_instance_tables']
        cur = dict(pan=r['panoptic_segmentation_deeplab'].cpu().numpy(),
                   inst=r['panoptic_segmentation_deeplab_instance_idx'].cpu().numpy(),
                   sem=r['_panoptic_segmentation_deeplab_semantic_idx_u8'].cpu().numpy(),
                   cls=tabs['inst_class'].copy(), pid=tabs['inst_pan_id'].copy(),
                   area=tabs['inst_area'].copy(), cyx=tabs['centers_yx'].copy(),
                   n=tabs['n_centers'].copy())
        if first is None:
            first = cur
            print('first: n', cur['n'], 'cls', cur['cls'][:, :6], 'pid', cur['pid'][:, :6], 'area', cur['area'][:, :6])
            continue
        diff = {k: int((cur[k] != first[k]).sum()) for k in cur}
        if any(diff.values()):
            nbad += 1
            if shown < 3:
                shown += 1
                print('iter', it, 'diffs', diff)
                w = np.argwhere(cur['pan'] != first['pan'])
                if len(w):
                    b, y, x = w[0]
                    print('  first pan diff at', (b, y, x), 'got', cur['pan'][b, y, x], 'want', first['pan'][b, y, x],
                          'inst', cur['inst'][b, y, x], first['inst'][b, y, x], 'sem', cur['sem'][b, y, x], first['sem'][b, y, x])
                print('  cls', cur['cls'][:, :6], 'pid', cur['pid'][:, :6], 'area', cur['area'][:, :6])
    print((H, W), 'lib', os.environ.get('NPB_LIB_PATH', 'new'), 'NPB_NO_PDL=' + os.environ.get('NPB_NO_PDL', ''), 'bad iterations', nbad)
