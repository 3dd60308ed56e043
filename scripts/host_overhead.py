"""Host-side (Python / allocator / launch) cost of one bench step, with cProfile."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nicr_mt_scene_analysis_b200 import testing
from nicr_mt_scene_analysis_b200.metric import MeanIntersectionOverUnion, PanopticEvaluation, PanopticQualityWithOrientationMAE
from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class

dev = torch.device('cuda:0')
w = bench.WORKLOAD
B, C, H, W, K = 8, w['C'], w['H'], w['W'], w['K']     # small batch: GPU time small, host time visible
is_thing = testing.default_is_thing(C)
has_ori = tuple(bool(t and c % 4 == 1) for c, t in enumerate(is_thing))
data = testing.make_batch(B, C, H, W, K, seed=1, device=dev, quantize=None)
batch = testing.make_batch_dict(B, H, W)
post = get_postprocessing_class('panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
    instance_postprocessing=get_postprocessing_class('instance')(), semantic_classes_is_thing=is_thing,
    semantic_class_has_orientation=has_ori, async_results=True)()
pq = PanopticQualityWithOrientationMAE(C + 1, 0, 1 << 16, 256 ** 3, (False,) + is_thing, device=dev)
miou = MeanIntersectionOverUnion(C + 1, True, device=dev)
ev = PanopticEvaluation(pq, miou)
raw = ((data['logits'], (data['heat'], data['offset'], data['orientation'])), (None, None))
r = post.postprocess(raw, batch, is_training=False)
tgt, tgt_sem = testing.make_eval_targets(r['panoptic_segmentation_deeplab'])

def step():
    r = post.postprocess(raw, batch, is_training=False)
    ev.update(r['panoptic_segmentation_deeplab'], tgt, tgt_sem)

for _ in range(5): step()
torch.cuda.synchronize()
N = 50
t0 = time.perf_counter()
for _ in range(N): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f'host issue time/step: {(t1-t0)/N*1e3:.3f} ms ; incl. drain: {(t2-t0)/N*1e3:.3f} ms')
pr = cProfile.Profile(); pr.enable()
for _ in range(N): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
