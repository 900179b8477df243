"""Shim: adaptive stepping is never taken on the reference path (adaptive=False)."""


def compute_error(*a, **k):
    raise NotImplementedError("shim: adaptive=False on the reference path")


def update_step_size(*a, **k):
    raise NotImplementedError("shim: adaptive=False on the reference path")
