"""Where does the eager (non-graph) step lose time?  post only / eval only / both, device time
per step vs host issue time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nicr_mt_scene_analysis_b200 import testing
from nicr_mt_scene_analysis_b200.metric import MeanIntersectionOverUnion, PanopticEvaluation, PanopticQuality
from nicr_mt_scene_analysis_b200.model.postprocessing import get_postprocessing_class
dev = torch.device('cuda:0')
w = bench.WORKLOADS['sunrgbd']
B, C, H, W, K = w['B'], w['C'], w['H'], w['W'], w['K']
is_thing = testing.default_is_thing(C)
has_ori = tuple(bool(t and c % 4 == 1) for c, t in enumerate(is_thing))
frames = [testing.make_frame(C, H, W, K, seed=i, device=dev, quantize=None) for i in range(8)]
data = {k: torch.stack([frames[i % 8][k] for i in range(B)]).contiguous() for k in frames[0]}
batch = testing.make_batch_dict(B, H, W)
post = get_postprocessing_class('panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
    instance_postprocessing=get_postprocessing_class('instance')(), semantic_classes_is_thing=is_thing,
    semantic_class_has_orientation=has_ori, async_results=True)()
ev = PanopticEvaluation(PanopticQuality(C + 1, 0, 1 << 16, 256 ** 3, (False,) + is_thing, device=dev),
                        MeanIntersectionOverUnion(C + 1, True, device=dev))
raw = ((data['logits'], (data['heat'], data['offset'], data['orientation'])), (None, None))
r0 = post.postprocess(raw, batch, is_training=False)
pan = r0['panoptic_segmentation_deeplab']
tgt, tgt_sem = testing.make_eval_targets(pan)
def f_post(): return post.postprocess(raw, batch, is_training=False)
def f_eval(): ev.update(pan, tgt, tgt_sem)
def f_both():
    r = post.postprocess(raw, batch, is_training=False); ev.update(r['panoptic_segmentation_deeplab'], tgt, tgt_sem); return r
def f_kernels_only(): return post._forward_kernels(data['logits'], data['heat'], data['offset'], data['orientation'])
for name, fn in (('post', f_post), ('post kernels only', f_kernels_only), ('eval', f_eval), ('both', f_both)):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(200): fn()
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    print(f'{name:20s} device {e0.elapsed_time(e1)/200*1e3:8.1f} us/step   host issue {(t1-t0)/200*1e6:8.1f} us/step')

# split the combined step with events
N = 200
ev_list = []
for _ in range(20): f_both()
torch.cuda.synchronize()
for _ in range(N):
    a, b_, c_ = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    r = post.postprocess(raw, batch, is_training=False)
    b_.record()
    ev.update(r['panoptic_segmentation_deeplab'], tgt, tgt_sem)
    c_.record()
    ev_list.append((a, b_, c_))
torch.cuda.synchronize()
tp = sum(a.elapsed_time(b_) for a, b_, c_ in ev_list) / N * 1e3
te = sum(b_.elapsed_time(c_) for a, b_, c_ in ev_list) / N * 1e3
tg = sum(ev_list[i][2].elapsed_time(ev_list[i + 1][0]) for i in range(N - 1)) / (N - 1) * 1e3
print(f'combined: post {tp:.1f} us, eval {te:.1f} us, gap between steps {tg:.1f} us')
