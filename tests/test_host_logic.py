"""CPU-only tests: the C-ABI library loads and exports what the header declares, host-side
helpers behave like the reference's, the product package never touches the oracle, and the
cross-rank metric reduction works with world_size 2 (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'nicr-multitask-scene-analysis_b200')


@pytest.fixture(scope='session', autouse=True)
def built():
    sys.path.insert(0, ROOT)
    import __graft_entry__
    __graft_entry__.build()


def test_library_exports_every_declared_symbol():
    from nicr_mt_scene_analysis_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'nicr_panoptic_b200.h')).read()
    declared = set(re.findall(r'\b(npb_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found'
    handle = _lib.lib()
    for name in declared:
        assert hasattr(handle, name), f'{name} declared in the header but not exported'
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert handle.npb_abi_version() == 3
    assert handle.npb_error_string(-2).decode().startswith('more than 255')


def test_host_argument_validation_without_gpu():
    """bad arguments are rejected on the host before any launch (no GPU needed)"""
    from nicr_mt_scene_analysis_b200 import _lib
    L = _lib.lib()
    assert L.npb_semantic_argmax(None, 1, 4, 8, 8, None, None, None) == _lib.ERR_ARG
    assert L.npb_instance_centers(None, 1, 8, 8, 0.1, 3, 4, None, 0, None, None, None, None, None,
                                  None) == _lib.ERR_ARG
    assert L.npb_instance_centers_workspace_bytes(2, 480, 640, 3) >= 2 * 240 * 320 * 8
    assert L.npb_pq_update_workspace_bytes(4, 41) > 4 * 16384 * 12
    one = (ctypes.c_int64 * 1)()
    for fn in (L.npb_confmat_update, L.npb_confmat_update_nonvoid):
        # empty maps add nothing (their pointers may be null); bad sizes / dtypes / pointers
        assert fn(None, _lib.I64, None, _lib.U8, 0, 5, None, None, None) == _lib.OK
        assert fn(None, _lib.I64, None, _lib.U8, 16, 5, one, one, None) == _lib.ERR_ARG
        assert fn(one, _lib.I64, one, _lib.U8, 1, 0, one, one, None) == _lib.ERR_ARG
        assert fn(one, 9, one, _lib.U8, 1, 5, one, one, None) == _lib.ERR_ARG
        assert fn(one, _lib.I64, one, _lib.U8, -1, 5, one, one, None) == _lib.ERR_ARG


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(import|from)\s+oracle\b', text, re.M), f
                assert 'panoptic_oracle' not in text, f


def test_no_cpu_fallback():
    from nicr_mt_scene_analysis_b200.metric import MeanIntersectionOverUnion, PanopticQuality
    from nicr_mt_scene_analysis_b200.utils import deeplab_merge_batch
    m = MeanIntersectionOverUnion(4, device='cpu')
    with pytest.raises(RuntimeError):
        m.update(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.int64))
    pq = PanopticQuality(2, 0, 16, 256, [False, True], device='cpu')
    with pytest.raises(RuntimeError):
        pq.update(torch.zeros(1, 2, 2, dtype=torch.int64), torch.zeros(1, 2, 2, dtype=torch.int64))
    with pytest.raises(RuntimeError):
        deeplab_merge_batch(torch.zeros(1, 2, 2, dtype=torch.int64),
                            torch.zeros(1, 2, 2, dtype=torch.uint8),
                            torch.zeros(1, 2, 2, dtype=torch.bool), 16, [1], 0)


def test_factory_and_constructor_contract():
    from nicr_mt_scene_analysis_b200.model.postprocessing import (
        InstancePostprocessing, PanopticPostprocessing, get_postprocessing_class)
    cls = get_postprocessing_class('instance', top_k_instances=100, heatmap_threshold=0.2)
    post = cls()
    assert isinstance(post, InstancePostprocessing)
    assert post._top_k_instances == 100 and post._heatmap_threshold == 0.2
    assert get_postprocessing_class('instance', top_k_instances=100, heatmap_threshold=0.2) is cls
    with pytest.raises(ValueError):
        get_postprocessing_class('does-not-exist')
    with pytest.raises(AssertionError):
        InstancePostprocessing(top_k_instances=255)
    with pytest.raises(AssertionError):
        InstancePostprocessing(heatmap_nms_kernel_size=4)
    pan = get_postprocessing_class(
        'panoptic', semantic_postprocessing=get_postprocessing_class('semantic')(),
        instance_postprocessing=post, semantic_classes_is_thing=(False, True, True),
        semantic_class_has_orientation=(False, False, True))()
    assert isinstance(pan, PanopticPostprocessing)
    assert pan.max_instances_per_category == 65536
    assert pan._thing_ids_panoptic.tolist() == [2, 3] and pan._orientation_ids.tolist() == [3]
    # training mode only forwards the raw decoder outputs (panoptic.py:57-75)
    out = pan.postprocess((('S', 'I'), ('s', 'i')), {}, is_training=True)
    assert out == {'semantic_output': 'S', 'semantic_side_outputs': 's',
                   'instance_output': 'I', 'instance_side_outputs': 'i'}


def test_fullres_helpers():
    from nicr_mt_scene_analysis_b200.utils.fullres import (fullres_shape,
                                                           valid_region_and_fullres_shape)
    batch = {'rgb_fullres': torch.zeros(2, 3, 10, 12),
             '_applied_preprocessing': [[{'type': 'Other'}, {'type': 'Resize',
                                                             'valid_region_slice_y': slice(0, 5),
                                                             'valid_region_slice_x': slice(1, 6)}]]}
    assert fullres_shape(batch, 'semantic') == (10, 12)
    assert valid_region_and_fullres_shape(batch, 'instance') == ((slice(0, 5), slice(1, 6)), (10, 12))
    with pytest.raises(ValueError):
        fullres_shape({}, 'semantic')
    with pytest.raises(ValueError):
        valid_region_and_fullres_shape({'rgb_fullres': torch.zeros(1, 3, 4, 4)}, 'semantic')


def test_result_dict_defers_and_aliases():
    from nicr_mt_scene_analysis_b200._results import ResultDict
    calls = []
    r = ResultDict(a=1)
    r.defer('b', lambda: calls.append('b') or 2)
    r.alias('b_fullres', 'b')
    assert 'b' in r and set(r.keys()) == {'a', 'b', 'b_fullres'} and not calls
    assert r.is_deferred('b')
    assert r['b_fullres'] == 2 and r['b'] == 2 and calls == ['b']
    assert dict(r.items()) == {'a': 1, 'b': 2, 'b_fullres': 2}
    assert r.get('zzz', 5) == 5


def test_pq_compute_matches_oracle_formulae():
    """compute() on hand-set states (CPU): pq.py:304-361"""
    import numpy as np
    import oracle
    from nicr_mt_scene_analysis_b200.metric import PanopticQuality
    is_thing = [False, True, False, True, True]
    m = PanopticQuality(5, 0, 1 << 16, 256 ** 3, is_thing, device='cpu')
    m.iou_per_class = torch.tensor([0.3, 4.2, 0.0, 1.7, 0.0], dtype=torch.float64)
    m.tp_per_class = torch.tensor([1.0, 5.0, 0.0, 2.0, 0.0], dtype=torch.float64)
    m.fn_per_class = torch.tensor([0.0, 1.0, 0.0, 3.0, 0.0], dtype=torch.float64)
    m.fp_per_class = torch.tensor([2.0, 0.0, 0.0, 1.0, 4.0], dtype=torch.float64)
    got = m.compute(suffix='_x')
    ref = oracle.pq_results(m.iou_per_class.numpy(), m.tp_per_class.numpy(),
                            m.fn_per_class.numpy(), m.fp_per_class.numpy(), is_thing, 0, '_x')
    assert set(got) == set(ref)
    for k in ref:
        np.testing.assert_allclose(np.asarray(got[k], dtype=np.float64), ref[k], rtol=1e-12)
    assert int(got['all_x_num_categories']) == 3 and int(got['all_with_gt_x_num_categories']) == 2


def test_metric_states_allreduce_world_size_2(tmp_path):
    """the N>1 path: per-rank states are summed at compute() (dist_reduce_fx='sum' of the
    reference, miou.py:24 / pq.py:231-246) -- two gloo ranks on the CPU."""
    script = tmp_path / 'rank.py'
    script.write_text(f'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {ROOT!r})
from nicr_mt_scene_analysis_b200.metric import MeanIntersectionOverUnion, PanopticQuality
dist.init_process_group('gloo')
rank = dist.get_rank()
pq = PanopticQuality(3, 0, 16, 256, [False, True, True], device='cpu')
pq.tp_per_class += torch.tensor([0., 1., 2.], dtype=torch.float64) * (rank + 1)
pq.iou_per_class += torch.tensor([0., .75, 1.5], dtype=torch.float64) * (rank + 1)
pq.fn_per_class += torch.tensor([0., 1., 0.], dtype=torch.float64)
mi = MeanIntersectionOverUnion(3, ignore_first_class=True, device='cpu')
mi.confmat += torch.tensor([[1, 0, 0], [0, 2 + rank, 1], [0, 1, 3]])
r = pq.compute()
miou = mi.compute()
assert pq.tp_per_class.tolist() == [0., 1. * (rank + 1), 2. * (rank + 1)]   # local state untouched
if rank == 0:
    assert abs(float(r['all_sq']) - 0.75) < 1e-12, r
    assert abs(float(r['rq_per_class'][1]) - 3 / 4) < 1e-12, r
    want = ((5 / 9) + (6 / 10)) / 2
    assert abs(float(miou) - want) < 1e-6, (float(miou), want)
    print('RANK0 OK')
dist.destroy_process_group()
''')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
                          '--nproc-per-node=2', '--master-addr', '127.0.0.1', '--master-port',
                          '29541', str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert 'RANK0 OK' in out.stdout


def test_mean_absolute_angular_error_host_logic():
    """MeanAbsoluteAngularError / abs_angle_error_rad (mae.py:16-64): pure host arithmetic"""
    import math
    from nicr_mt_scene_analysis_b200.metric import MeanAbsoluteAngularError
    from nicr_mt_scene_analysis_b200.metric.mae import abs_angle_error_rad
    e = abs_angle_error_rad(torch.tensor(0.1), torch.tensor(2 * math.pi - 0.1))
    assert float(e) == pytest.approx(0.2, abs=1e-6)
    e = abs_angle_error_rad(torch.tensor(-3.0), torch.tensor(3.0))
    assert float(e) == pytest.approx(2 * math.pi - 6.0, abs=1e-6)
    m = MeanAbsoluteAngularError(device='cpu')
    m.update([{1: 0.5, 2: 1.0}, {7: -1.0}], [{1: 0.25, 2: 1.5, 3: 9.0}, {7: 1.0}])
    rad, deg = m.compute()
    assert int(m.n_elements) == 3
    assert float(rad) == pytest.approx((0.25 + 0.5 + 2.0) / 3, rel=1e-6)
    assert float(deg) == pytest.approx(math.degrees((0.25 + 0.5 + 2.0) / 3), rel=1e-6)
    m.reset()
    assert int(m.n_elements) == 0 and float(m.sum_angular_error) == 0.0


def test_batched_angular_errors_equal_the_per_pair_loop():
    """The MAAE update evaluates the errors of a batch with one float32 vector operation; every
    element must equal the reference's scalar form `abs_angle_error_rad(torch.tensor(p),
    torch.tensor(t))` (mae.py:157-160) bit for bit, and the float64 state the sequential sum."""
    import math
    import random
    from nicr_mt_scene_analysis_b200.metric import MeanAbsoluteAngularError
    from nicr_mt_scene_analysis_b200.metric.mae import abs_angle_error_rad
    rnd = random.Random(5)
    preds = [rnd.uniform(-10, 10) for _ in range(500)] + [0.0, math.pi, -math.pi, 2 * math.pi, 7.5]
    targets = [rnd.uniform(-10, 10) for _ in range(500)] + [2 * math.pi, -math.pi, math.pi, 0.0, 7.5]
    vec = abs_angle_error_rad(torch.tensor(preds), torch.tensor(targets))
    total = torch.tensor(0, dtype=torch.float64)
    for i, (p, t) in enumerate(zip(preds, targets)):
        e = abs_angle_error_rad(torch.tensor(p), torch.tensor(t))
        assert e.dtype == torch.float32 and float(e) == float(vec[i]), i
        total += e
    m = MeanAbsoluteAngularError(device='cpu')
    m.update([dict(enumerate(preds))], [dict(enumerate(targets))])
    assert int(m.n_elements) == len(preds)
    assert float(m.sum_angular_error) == float(total)


def test_instance_tables_layout_is_aligned():
    from nicr_mt_scene_analysis_b200._results import InstanceTables
    for B in (1, 3, 64):
        offsets, total = InstanceTables.layout(B)
        assert total % 16 == 0
        spans = sorted((off, off + nbytes) for off, nbytes, _, _ in offsets.values())
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))          # no overlap
        import numpy as np
        for off, _, dt, _ in offsets.values():
            assert off % np.dtype(dt).itemsize == 0


def test_task_helper_requires_cuda_device():
    import torch
    from nicr_mt_scene_analysis_b200.task_helper import PanopticTaskHelper
    helper = PanopticTaskHelper(3, [False, True, False])
    with pytest.raises(RuntimeError):
        helper.initialize(torch.device('cpu'))
    with pytest.raises(RuntimeError):
        helper.device


def test_eval_args_struct_layout_matches_the_header(tmp_path):
    """`_lib.EvalArgs` (ctypes) must mirror `npb_eval_args` of include/nicr_panoptic_b200.h
    field by field: compile a C probe against the header and compare sizes and offsets."""
    import ctypes
    import subprocess
    from nicr_mt_scene_analysis_b200 import _lib
    names = [f[0] for f in _lib.EvalArgs._fields_]
    src = tmp_path / 'probe.c'
    src.write_text(
        '#include <stddef.h>\n#include <stdio.h>\n#include "nicr_panoptic_b200.h"\n'
        'int main(void) {\n  printf("%zu\\n", sizeof(npb_eval_args));\n' +
        ''.join(f'  printf("%zu\\n", offsetof(npb_eval_args, {n}));\n' for n in names) +
        '  return 0;\n}\n')
    exe = tmp_path / 'probe'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    out = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True,
                                          text=True).stdout.split()]
    assert out[0] == ctypes.sizeof(_lib.EvalArgs)
    assert out[1:] == [getattr(_lib.EvalArgs, n).offset for n in names]


def test_table_constants_match_the_header():
    """Row lengths the Python layer allocates its tables with, and the error codes it decodes,
    are the header's."""
    import re
    from nicr_mt_scene_analysis_b200 import _lib
    text = open(os.path.join(ROOT, 'include', 'nicr_panoptic_b200.h')).read()
    macro = lambda name: int(re.search(r'#define\s+%s\s+\(?(-?\d+)\)?' % name, text).group(1))
    assert macro('NPB_MAX_INST') == _lib.MAX_INST
    assert macro('NPB_MAX_WIDE_CENTERS') == _lib.MAX_WIDE_CENTERS
    for name in ('ARG', 'TOO_MANY_CENTERS', 'ZERO_DIVISION', 'CATEGORY_RANGE', 'CAPACITY', 'CUDA'):
        assert macro('NPB_ERR_' + name) == getattr(_lib, 'ERR_' + name), name


def test_reference_citations_resolve():
    """Every `file.py:first-last` citation of the C header, the design / integration notes, the
    host package, the kernels and the oracle names a file of the reference that has that many
    lines (checked where /root/reference exists; the GPU box has no reference)."""
    import glob
    import re
    ref_root = '/root/reference'
    if not os.path.isdir(ref_root):
        pytest.skip('reference sources not present')
    files = glob.glob(os.path.join(ref_root, '**', '*.py'), recursive=True)
    n_lines = {f: sum(1 for _ in open(f, encoding='utf-8', errors='replace')) for f in files}
    pkg = os.path.join(ROOT, 'nicr-multitask-scene-analysis_b200')
    sources = [os.path.join(ROOT, p) for p in ('include/nicr_panoptic_b200.h', 'DESIGN.md',
                                               'INTEGRATION.md', 'oracle/panoptic_oracle.c',
                                               'oracle/__init__.py')]
    sources += glob.glob(os.path.join(pkg, '**', '*.py'), recursive=True)
    sources += glob.glob(os.path.join(pkg, 'csrc', '*.cu*'))
    total = 0
    for src in sources:
        for name, first, last in re.findall(r'([A-Za-z_/+]+\.py):(\d+)(?:-(\d+))?', open(src).read()):
            total += 1
            last = int(last or first)
            candidates = [f for f in files if f.endswith('/' + name)]
            where = f'{os.path.relpath(src, ROOT)}: {name}:{first}-{last}'
            assert candidates, where + ' (no such file in the reference)'
            assert any(n_lines[f] >= last for f in candidates), where + ' (beyond the end of the file)'
            assert int(first) <= last, where
    assert total >= 200


def _bench_reference_line(*extra):
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '0', *extra], capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['higher_is_better'] is True and d['scaling'] == 'weak'
    assert d['unit'] == 'frames/s' and d['value'] > 0 and d['ms_per_step'] > 0
    assert d['n_gpus'] == 1 and d['steps'] == 1 and d['vs_baseline'] is None
    assert d['data'] == 'synthetic' and d['dtype'] == 'f32'
    # the configuration the metric is quoted on (BASELINE.json "@480x640")
    assert d['config']['workload'] == 'nyuv2_480x640_c40_b8'
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0
    cb = d['cpu_baseline']
    assert cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    return d


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs without a GPU and prints ONE JSON line with the keys the
    driver reads.  Where the reference is installed (baseline/_ref) or present (/root/reference)
    it is the UNMODIFIED reference that is timed (`kind: "reference"`), on one batch of the
    headline workload per step."""
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import reference_arm
    d = _bench_reference_line()
    cb = d['cpu_baseline']
    if reference_arm.available():
        assert cb['kind'] == 'reference' and 'UNMODIFIED reference' in cb['sample']
        assert d['config']['sample_frames_per_step'] == 8
        assert 0.0 < d['quality']['all_pq'] <= 1.0 and 0.0 < d['quality']['miou'] <= 1.0
    else:
        assert cb['kind'] == 'port'


def test_bench_port_arm_prints_the_contract_line():
    """`--port`: the C oracle port as the CPU arm (second key of the GPU arm's line)."""
    d = _bench_reference_line('--port', '--frames', '4')
    assert d['cpu_baseline']['kind'] == 'port'
