# -*- coding: utf-8 -*-
"""Panoptic quality on the GPU (API of metric/pq.py:60-361).

`update` is one call of `npb_pq_update` (csrc/eval.cu): a streaming pixel pass building a
per-frame (gt segment, pred segment) contingency table, a per-frame matcher and an ordered
accumulation -- it replaces the reference's process pool + three `torch.unique` sorts per
frame.  The float64 states are bit-identical to the reference's (same summation orders).
`compute` mirrors pq.py:304-361 on the rank-summed states.
"""
from ctypes import c_int, c_int64
from typing import Dict, List, Optional, Set, Tuple, Union

import torch

from .. import _lib
from ._state import MetricState

_EPSILON = 1e-10
MATCH_CAP = 1024        # matched (gt, pred) pairs kept per frame


def _safe_divide(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """x / y, 0 where |y| < 1e-10 (pq.py:182-187)."""
    return torch.where(torch.abs(y) < _EPSILON, torch.zeros_like(x), x / y)


class _ScratchOwner:
    def __init__(self, scratch):
        self._scratch = scratch


class _PQKernel:
    """Workspace + launch helper shared by PanopticQuality and compare_and_accumulate."""

    _scratch = {}      # workspaces of the function-style entry points (compare_and_accumulate)

    @classmethod
    def _workspace(cls, dev, B, num_categories, scratch=None):
        """(device, stream, B, num_categories) -> reusable workspace.  `scratch`: the dict of the
        metric object that owns the update -- a workspace may hold the hand-over of an update
        whose matcher has not run yet (pipelined matching), so metric objects do not share one."""
        if scratch is not None:
            # one workspace per metric object, batch size and device -- NOT per stream: a pending
            # hand-over must be found again by a call that is issued on another (ordered) stream,
            # e.g. the capture stream of a CUDA graph behind eager warm-up calls, and a workspace
            # must never be created (zero-filled) inside a capture
            cls = _ScratchOwner(scratch)
            key = (dev, None, B, num_categories)
        else:
            key = (dev, torch.cuda.current_stream(dev).cuda_stream, B, num_categories)
        if key not in cls._scratch:
            # the size depends on the SM count of the device the kernels will run on
            with torch.cuda.device(dev):
                nbytes = _lib.lib().npb_pq_update_workspace_bytes(B, num_categories)
            # zeroed once: the hand-over tables between pixel pass and matcher are zero at rest
            # (the matcher cleans up behind itself; npb_panoptic_forward_eval_pipelined relies on it)
            cls._scratch[key] = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        return cls._scratch[key]

    @staticmethod
    def prepare(target: torch.Tensor, num_categories: int, ignored_label: int, offset: int,
                void_segment_id: int, iou, tp, fn, fp, sem_target: Optional[torch.Tensor] = None,
                confmat: Optional[torch.Tensor] = None, want_matches: bool = False,
                want_frame_stats: bool = False, status: Optional[torch.Tensor] = None,
                scratch: Optional[Dict] = None):
        """Everything of an update except the prediction: validated targets, scratch, optional
        outputs.  Returns (`_lib.EvalArgs`, dict of the tensors it points into)."""
        dev = iou.device
        if not dev.type == 'cuda':
            raise RuntimeError('PanopticQuality.update needs its states on a CUDA device')
        target = _lib.require_cuda(target.to(dev).to(torch.int64), 'targets', ndim=3)
        B = target.shape[0]
        ws = _PQKernel._workspace(dev, B, num_categories, scratch)
        if status is None or status.numel() < B:
            status = torch.zeros(B, dtype=torch.int32, device=dev)
        matches = n_matches = frame_stats = None
        if want_matches:
            matches = torch.empty((B, MATCH_CAP, 2), dtype=torch.int64, device=dev)
            n_matches = torch.zeros(B, dtype=torch.int32, device=dev)
        if want_frame_stats:
            # rows 0..B-1: per-frame results, row B: the states before this update (the journal
            # `PanopticQuality._replay` re-accumulates from)
            frame_stats = torch.empty((B + 1, 4, num_categories), dtype=torch.float64, device=dev)
        n_cm = 0
        if confmat is not None:
            sem_target = _lib.require_cuda(sem_target.to(dev), 'semantic target', torch.uint8, 3)
            assert sem_target.shape == target.shape
            n_cm = confmat.shape[0]
        else:
            sem_target = None
        p = lambda t: None if t is None else t.data_ptr()
        args = _lib.EvalArgs(
            target=p(target), sem_target=p(sem_target), num_categories=num_categories,
            confmat_n=n_cm, ignored_label=ignored_label, offset=offset,
            void_segment_id=void_segment_id, workspace=p(ws), iou=p(iou), tp=p(tp), fn=p(fn),
            fp=p(fp), confmat=p(confmat), frame_stats=p(frame_stats), matches=p(matches),
            match_cap=MATCH_CAP, n_matches=p(n_matches), status=p(status))
        keep = dict(target=target, sem_target=sem_target, workspace=ws, status=status,
                    matches=matches, n_matches=n_matches, frame_stats=frame_stats)
        return args, keep

    @staticmethod
    def run(pred: torch.Tensor, target: torch.Tensor, num_categories: int, ignored_label: int,
            max_instances_per_category: int, offset: int, void_segment_id: int,
            iou, tp, fn, fp, **kw):
        a, keep = _PQKernel.prepare(target, num_categories, ignored_label, offset,
                                    void_segment_id, iou, tp, fn, fp, **kw)
        dev = iou.device
        pred = _lib.require_cuda(pred.to(dev).to(torch.int64), 'preds', ndim=3)
        target = keep['target']
        assert target.shape == pred.shape
        B = pred.shape[0]
        P = pred.shape[1] * pred.shape[2]
        _lib.check(_lib.lib().npb_pq_update(
            _lib.ptr(pred), a.target, a.sem_target, c_int(B), c_int64(P),
            c_int(num_categories), c_int64(ignored_label), c_int64(max_instances_per_category),
            c_int64(offset), c_int64(void_segment_id), a.workspace, a.iou, a.tp, a.fn, a.fp,
            a.confmat, c_int(a.confmat_n), a.frame_stats, a.matches, c_int(MATCH_CAP),
            a.n_matches, a.status, _lib.stream_ptr(dev)), 'npb_pq_update')
        return keep['status'], keep['matches'], keep['n_matches'], keep['frame_stats']


def _evaluate_big_frame(pred: torch.Tensor, target: torch.Tensor, num_categories: int,
                        ignored_label: int, max_instances_per_category: int, offset: int,
                        void_segment_id: int, iou, tp, fn, fp, matches_row=None,
                        n_matches_row=None, where: str = 'PanopticQuality.update') -> None:
    """`npb_pq_update_big_frame`: one frame (H,W) that exceeded the capacities of the
    shared-memory matcher (it contributed nothing there), every table in global memory.
    Synchronous: the frame's own status word is read back before returning."""
    dev = iou.device
    pred = _lib.require_cuda(pred.to(dev).to(torch.int64), 'preds', ndim=2)
    target = _lib.require_cuda(target.to(dev).to(torch.int64), 'targets', ndim=2)
    P = pred.numel()
    L = _lib.lib()
    ws = torch.empty(L.npb_pq_update_big_frame_workspace_bytes(P, num_categories),
                     dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(L.npb_pq_update_big_frame(
        _lib.ptr(pred), _lib.ptr(target), c_int64(P), c_int(num_categories),
        c_int64(ignored_label), c_int64(max_instances_per_category), c_int64(offset),
        c_int64(void_segment_id), _lib.ptr(ws), _lib.ptr(iou), _lib.ptr(tp), _lib.ptr(fn),
        _lib.ptr(fp), _lib.ptr(matches_row), c_int(MATCH_CAP), _lib.ptr(n_matches_row),
        _lib.ptr(status), _lib.stream_ptr(dev)), 'npb_pq_update_big_frame')
    _lib.raise_for_status(status.cpu().tolist(), where + ' (large-frame path)')


def compare_and_accumulate(
    pred: torch.Tensor,
    target: torch.Tensor,
    num_categories: int,
    ignored_label: int,
    max_instances_per_category,
    offset: int,
    void_segment_id: int
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, Set[Tuple[int, int]]]:
    """One frame (H,W): returns (iou, tp, fn, fp) float64 [num_categories] (CPU) and the set
    of matched (gt_segment_id, pred_segment_id) pairs -- pq.py:60-179."""
    if not pred.is_cuda:
        raise RuntimeError('compare_and_accumulate: expected CUDA tensors (no CPU path)')
    dev = pred.device
    state = [torch.zeros(num_categories, dtype=torch.float64, device=dev) for _ in range(4)]
    status, matches, n_matches, _ = _PQKernel.run(
        pred[None], target[None], num_categories, ignored_label, max_instances_per_category,
        offset, void_segment_id, *state, want_matches=True)
    codes = status.cpu().tolist()
    if codes[0] == _lib.ERR_CAPACITY:       # beyond the shared-memory matcher: large-frame path
        _evaluate_big_frame(pred, target, num_categories, ignored_label,
                            max_instances_per_category, offset, void_segment_id, *state,
                            matches_row=matches[0], n_matches_row=n_matches[0:1],
                            where='compare_and_accumulate')
    else:
        _lib.raise_for_status(codes, 'compare_and_accumulate')
    n = int(n_matches[0].item())
    pairs = {(int(g), int(p)) for g, p in matches[0, :n].cpu().tolist()}
    iou, tp, fn, fp = (s.cpu() for s in state)
    return iou, tp, fn, fp, pairs


class PanopticQuality(MetricState):
    def __init__(
        self,
        num_categories: int,
        ignored_label: int,
        max_instances_per_category: int,
        offset: int,
        is_thing: Union[torch.Tensor, List[bool]],
        num_workers=None,          # accepted for API parity; there is no process pool
        device=None
    ) -> None:
        super().__init__(device)
        self.num_categories = num_categories
        self.ignored_label = ignored_label
        self.max_instances_per_category = max_instances_per_category
        self.offset = offset
        self.is_thing = torch.as_tensor(is_thing).to(torch.bool).cpu()
        self.is_stuff = torch.logical_not(self.is_thing)
        assert len(self.is_thing) == self.num_categories
        # one void segment with instance id 0 (pq.py:220-222)
        self.void_segment_id = self.ignored_label * self.max_instances_per_category
        for name in ('iou_per_class', 'tp_per_class', 'fn_per_class', 'fp_per_class'):
            self.add_state(name, torch.zeros(num_categories, dtype=torch.float64),
                           dist_reduce_fx='sum')
        # status words.  Eager updates get their own [B] buffer and stay in `_pending` until it has
        # been read: a frame that reports NPB_ERR_CAPACITY is evaluated again on the large-frame
        # path (its tensors are kept alive for that).  Updates recorded into a CUDA graph cannot
        # be followed up per replay; they share one buffer per batch size (errors merge by
        # atomicMin) and a capacity overflow stays an error.
        self._status: Dict[int, torch.Tensor] = {}
        self._pending: List[Dict] = []
        self._scratch: Dict = {}            # kernel workspaces of this metric object
        self._status_hosts: Dict[int, List[torch.Tensor]] = {}     # pinned status copies, recycled
        # pipelined matching (PanopticPostprocessing.fuse_evaluation(..., pipeline_matching=True)):
        # the update whose pixel pass has been issued but whose matcher has not run yet
        self._deferred: Optional[Dict] = None

    # ---- update --------------------------------------------------------------------------
    def _status_for(self, B: int) -> Tuple[torch.Tensor, bool]:
        """(status buffer, whether this update can be followed up)"""
        dev = self.iou_per_class.device
        if not torch.cuda.is_current_stream_capturing():
            return torch.zeros(B, dtype=torch.int32, device=dev), True
        status = self._status.get(B)
        if status is None or status.device != dev:
            raise RuntimeError('PanopticQuality: run one update outside of the graph capture first '
                               '(its status buffer cannot be allocated while capturing)')
        return status, False

    def _shared_status(self, B: int) -> None:
        dev = self.iou_per_class.device
        if torch.cuda.is_current_stream_capturing():
            return      # never allocated inside a capture (a memset node would clear it per replay)
        status = self._status.get(B)
        if status is None or status.device != dev:
            self._status[B] = torch.zeros(B, dtype=torch.int32, device=dev)

    FOLLOW_UP_DEPTH = 3     # eager updates whose status words have not been read yet

    def _follow_up(self, status, preds, targets, matches=None, n_matches=None,
                   frame_stats=None) -> None:
        """Remember an eager update until its status has been read; the oldest ones are resolved
        now (their kernels have long finished, so this does not stall the device)."""
        # the status words travel to pinned memory right behind the update's kernels; waiting for
        # THAT copy later does not wait for anything enqueued after it
        pool = self._status_hosts.setdefault(status.numel(), [])
        # (pinned allocations cost tens of microseconds: the buffers are recycled by _resolve)
        host = pool.pop() if pool else torch.empty(status.shape, dtype=status.dtype, pin_memory=True)
        host.copy_(status, non_blocking=True)
        landed = torch.cuda.Event()
        landed.record(torch.cuda.current_stream(status.device))
        self._pending.append(dict(status=status, host=host, landed=landed, preds=preds,
                                  targets=targets, matches=matches, n_matches=n_matches,
                                  frame_stats=frame_stats))
        # a few updates may be in flight: the host never waits for the batch it has just issued
        while len(self._pending) > self.FOLLOW_UP_DEPTH:
            self._resolve(self._pending.pop(0))

    def _codes(self, entry: Dict) -> List[int]:
        """Status words of a pending update (waits for their copy, not for the device)."""
        if entry.get('host') is None:
            return entry['status'].cpu().tolist()
        entry['landed'].synchronize()
        codes = entry['host'].tolist()
        pool = self._status_hosts.setdefault(entry['host'].numel(), [])
        if len(pool) < 2 * self.FOLLOW_UP_DEPTH:
            pool.append(entry['host'])
        entry['host'] = None
        return codes

    def _resolve(self, entry: Dict) -> None:
        codes = self._codes(entry)
        if _lib.ERR_CAPACITY not in codes:
            _lib.raise_for_status(codes, type(self).__name__ + '.update')
            return
        # a frame beyond the capacities of the batched matcher contributed nothing; it is
        # evaluated again on the large-frame path and put back IN ITS PLACE: the updates issued
        # since (all still pending) are re-accumulated with it
        later, self._pending = self._pending, []
        self._replay([(entry, codes)] + [(e, self._codes(e)) for e in later])

    def _states(self) -> Tuple[torch.Tensor, ...]:
        return self.iou_per_class, self.tp_per_class, self.fn_per_class, self.fp_per_class

    def _replay(self, entries: List[Tuple[Dict, List[int]]]) -> None:
        """Large-frame follow-up with the reference's float64 order (pq.py:298-303).

        `entries`: every update issued since (and including) the oldest one with an
        over-capacity frame, with their status words.  Each left a journal -- per-frame results
        + the states before it (`frame_stats`, see npb_pq_update) -- so the states are rebuilt
        from the oldest journal frame by frame, the large-frame results in the places of the
        frames that had contributed zeros.  Updates this object does not know about (replays of
        a CUDA graph on the same metric in between) are detected by rebuilding the CURRENT
        states first: when that does not reproduce them bit for bit the large-frame results are
        simply added (exact counts, float64 IoU sums out of frame order)."""
        where = type(self).__name__ + '.update'
        errors: List[int] = []
        redone = {}
        for k, (entry, codes) in enumerate(entries):
            errors += [c for c in codes if c != _lib.ERR_CAPACITY]
            for b, c in enumerate(codes):
                if c != _lib.ERR_CAPACITY:
                    continue
                fs = [torch.zeros_like(s) for s in self._states()]
                m, n = entry['matches'], entry['n_matches']
                _evaluate_big_frame(
                    entry['preds'][b], entry['targets'][b], self.num_categories,
                    self.ignored_label, self.max_instances_per_category, self.offset,
                    self.void_segment_id, *fs, matches_row=None if m is None else m[b],
                    n_matches_row=None if n is None else n[b:b + 1], where=where)
                redone[(k, b)] = torch.stack(fs)        # 0 + x = x: the frame's own result
        journals = [e.get('frame_stats') for e, _ in entries]
        exact = all(j is not None for j in journals)
        if exact:
            current = torch.stack(self._states())
            rebuilt = journals[0][-1].clone()
            for j in journals:
                for b in range(j.shape[0] - 1):
                    rebuilt += j[b]
            # (bit patterns: NaN-safe and -0.0 != 0.0)
            exact = bool(torch.equal(rebuilt.view(torch.int64), current.view(torch.int64)))
        if exact:
            state = journals[0][-1].clone()
            for k, j in enumerate(journals):
                for b in range(j.shape[0] - 1):
                    state += redone.get((k, b), j[b])
            for dst, src in zip(self._states(), state):
                dst.copy_(src)
        else:
            for fs in redone.values():
                for dst, src in zip(self._states(), fs):
                    dst += src
        _lib.raise_for_status(errors, where)

    # ---- pipelined matching ----------------------------------------------------------------
    def _set_deferred(self, args, keep: Dict, preds: Optional[torch.Tensor], B: int) -> None:
        """The pixel pass of an update has been issued, its matcher has not (it runs next to
        the kernels of the next pipelined call, or in `_flush_deferred`)."""
        self._deferred = dict(args=args, keep=keep, preds=preds, B=int(B))

    def _deferred_matcher_issued(self, captured: bool = False) -> None:
        """The matcher of the deferred update has just been enqueued: an eager update is followed
        up from here (its status words are final behind that matcher)."""
        d, self._deferred = self._deferred, None
        if d is None or not d['keep'].get('eager'):
            return
        k = d['keep']
        if captured:
            # handed over to a CUDA graph (the first replay runs its matcher and reports into the
            # shared status words): only its own words are looked at, at the next check
            self._pending.append(dict(status=k['status'], host=None, landed=None, preds=d['preds'],
                                      targets=k['target'], matches=None, n_matches=None))
        else:
            self._follow_up(k['status'], d['preds'], k['target'], k['matches'], k['n_matches'],
                            k['frame_stats'])

    def _flush_deferred(self) -> None:
        """Run the matcher of the deferred update now (current stream)."""
        d = self._deferred
        if d is None:
            return
        import ctypes
        dev = self.iou_per_class.device
        _lib.check(_lib.lib().npb_pq_match_pending(
            ctypes.byref(d['args']), c_int(d['B']), c_int64(self.max_instances_per_category),
            _lib.stream_ptr(dev)), 'npb_pq_match_pending')
        self._deferred_matcher_issued()

    def _launch(self, preds, targets, **kw):
        assert preds.ndim == 3
        assert targets.shape == preds.shape
        self._flush_deferred()      # the stand-alone update shares the hand-over workspace
        self._shared_status(preds.shape[0])
        status, eager = self._status_for(preds.shape[0])
        _, matches, n_matches, frame_stats = _PQKernel.run(
            preds, targets, self.num_categories, self.ignored_label,
            self.max_instances_per_category, self.offset, self.void_segment_id,
            self.iou_per_class, self.tp_per_class, self.fn_per_class, self.fp_per_class,
            status=status, scratch=self._scratch, want_frame_stats=eager, **kw)
        if eager:
            self._follow_up(status, preds, targets, matches, n_matches, frame_stats)
        return matches, n_matches, frame_stats

    def _eval_args(self, targets, **kw):
        """`npb_eval_args` of an update whose prediction is produced by a fused kernel
        (model/postprocessing/panoptic.py); same keyword arguments as `_launch`.  The caller
        hands the produced prediction to `_fused_issued` once the call has been issued."""
        assert targets.ndim == 3
        self._shared_status(targets.shape[0])
        status, eager = self._status_for(targets.shape[0])
        args, keep = _PQKernel.prepare(
            targets, self.num_categories, self.ignored_label, self.offset, self.void_segment_id,
            self.iou_per_class, self.tp_per_class, self.fn_per_class, self.fp_per_class,
            status=status, scratch=self._scratch, want_frame_stats=eager, **kw)
        keep['eager'] = eager
        return args, keep

    def _fused_issued(self, keep: Dict, preds: torch.Tensor) -> None:
        if keep.get('eager'):
            self._follow_up(keep['status'], preds, keep['target'], keep['matches'], keep['n_matches'],
                            keep['frame_stats'])

    def update(self, preds: torch.Tensor, targets: torch.Tensor) -> None:
        """preds, targets: (B,H,W) panoptic ids (class * max_instances + instance).
        Asynchronous; data-dependent errors surface at `compute()` / `check_status()`."""
        if preds.shape[0] == 0:          # an empty batch adds nothing (the loop of pq.py:276-303)
            assert targets.shape == preds.shape
            return
        self._launch(preds, targets)

    def check_status(self) -> None:
        self._flush_deferred()
        while self._pending:        # (a large-frame follow-up takes the later entries with it)
            self._resolve(self._pending.pop(0))
        for status in self._status.values():
            codes = status.cpu().tolist()
            status.zero_()
            _lib.raise_for_status(codes, type(self).__name__ + '.update')

    def reset(self) -> None:
        self._flush_deferred()      # its frames belong to the states that are being dropped
        super().reset()
        self._pending = []
        for status in self._status.values():
            status.zero_()

    # ---- compute (pq.py:254-361) -----------------------------------------------------------
    @staticmethod
    def _valid(counts: torch.Tensor, ignored_label: int) -> torch.Tensor:
        valid = counts != 0
        if 0 <= ignored_label < len(valid):
            valid[ignored_label] = False
        return valid

    def _host_states(self) -> Dict[str, torch.Tensor]:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # a data error on ONE rank must not leave the others blocked in the all-reduce of the
            # states: every rank learns whether any rank failed before anybody raises
            err = None
            try:
                self.check_status()
            except Exception as e:      # noqa: BLE001
                err = e
            flag = torch.tensor([0 if err is None else 1], dtype=torch.int32,
                                device=self.iou_per_class.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if err is not None:
                raise err
            if int(flag.item()):
                raise RuntimeError(type(self).__name__ + '.compute: another rank reported an '
                                   'error in its updates')
        else:
            self.check_status()
        return self.host_states()

    def result_per_category(self, states: Optional[Dict[str, torch.Tensor]] = None) -> Dict:
        s = states if states is not None else self._host_states()
        sq = _safe_divide(s['iou_per_class'], s['tp_per_class'])
        rq = _safe_divide(s['tp_per_class'], s['tp_per_class'] + 0.5 * s['fn_per_class'] +
                          0.5 * s['fp_per_class'])
        return {'sq_per_class': sq, 'rq_per_class': rq, 'pq_per_class': torch.multiply(sq, rq)}

    def compute(self, suffix: str = '') -> Dict:
        s = self._host_states()
        results = self.result_per_category(s)
        tp, fn, fp = s['tp_per_class'], s['fn_per_class'], s['fp_per_class']
        valid = self._valid(tp + fn + fp, self.ignored_label)        # panopticapi convention
        valid_gt = self._valid(tp + fn, self.ignored_label)          # categories with GT only
        subsets = {
            f'all{suffix}': valid,
            f'things{suffix}': valid & self.is_thing,
            f'stuff{suffix}': valid & self.is_stuff,
            f'all_with_gt{suffix}': valid_gt,
            f'things_with_gt{suffix}': valid_gt & self.is_thing,
            f'stuff_with_gt{suffix}': valid_gt & self.is_stuff,
        }
        for name, sel in subsets.items():
            if torch.any(sel):
                for q in ('pq', 'sq', 'rq'):
                    results[f'{name}_{q}'] = torch.mean(results[f'{q}_per_class'][sel])
                results[f'{name}_num_categories'] = torch.sum(sel.int())
            else:
                for q in ('pq', 'sq', 'rq', 'num_categories'):
                    results[f'{name}_{q}'] = torch.tensor(0)
        return results
