"""Minimal stand-in for `torchmetrics` (absent from this image).

TEST INFRASTRUCTURE ONLY: lets the *unmodified* reference under /root/reference be
imported in the authoring container so that golden vectors can be generated from it
(see tests/golden/make_golden.py).  Only the state book-keeping the reference uses
(`add_state`, `reset`) is provided; no arithmetic lives here.
"""
import torch


class Metric(torch.nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self._state_defaults = {}

    def add_state(self, name, default, dist_reduce_fx=None):
        self._state_defaults[name] = default.clone()
        setattr(self, name, default.clone())

    def reset(self):
        for name, default in self._state_defaults.items():
            setattr(self, name, default.clone())


class ConfusionMatrix(Metric):
    pass
