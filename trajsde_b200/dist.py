"""Scene-sharded data parallelism for the SDE path: one process per GPU, rows of different scenes are
independent so the forward needs NO collective; training adds ONE all-reduce of the flat fp32 gradient bucket per step
(what torch DDP does for the reference when `--gpus > 1`, train.py:35,54 — here explicit and latency-sized: 88,003 floats
for SDE+GRU parameters)."""
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_scenes(n_scenes: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous scene range [start, end) of `rank`; sizes differ by at most one scene."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(n_scenes, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_row_offsets(n_scenes: int, agents_per_scene: int, world_size: int, rank: int, modes: int = 10) -> dict:
    """Global row ids of this rank's first encoder / decoder row, for the `row_offset` argument of the solver calls: the
    Philox Brownian streams are keyed by GLOBAL row id, so a sharded run draws the noise of the unsharded batch."""
    s0, _ = shard_scenes(n_scenes, world_size, rank)
    return {'scene_start': s0, 'enc_agent_row': s0 * agents_per_scene, 'dec_row': s0 * agents_per_scene * modes}


class FlatGradBucket:
    """All parameter gradients in ONE contiguous fp32 buffer, so a training step issues a single all-reduce per bucket.

    ``pack=False`` (DDP's gradient_as_bucket_view): every ``p.grad`` is a view into the buffer, zeroed by ``zero_()``; autograd then
    ACCUMULATES each gradient into its view — one small add launch per parameter tensor.
    ``pack=True``: ``zero_()`` sets the gradients to ``None`` (autograd hands over the tensors the fused backward calls produced, no
    launch), and the reduction first packs them into the buffer with one multi-tensor copy and re-points ``p.grad`` at the views.  With
    the fused operators every parameter gradient arrives exactly once, so nothing is lost — and a launch-bound step (128 scenes) sheds
    ~60 launches.  Parameters that received no gradient keep ``grad = None`` (the optimizer skips them, like the reference's ``pi``
    head under its L2 + DiffBCE losses); their slots of the buffer are zero."""

    def __init__(self, params: Iterable[torch.nn.Parameter], pack: bool = False):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.pack = bool(pack)
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self._side = None
        self._views = []
        off = 0
        for p in self.params:
            self._views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        if not self.pack:
            for p, v in zip(self.params, self._views):
                p.grad = v

    def zero_(self):
        if self.pack:
            for p in self.params:
                p.grad = None
            return
        self.flat.zero_()
        for p, view in zip(self.params, self._views):   # re-attach: optimizers / autograd may have replaced .grad
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                p.grad = view

    def pack_(self):
        """pack mode: copy the gradients autograd left on the parameters into the buffer (one multi-tensor launch) and make ``p.grad``
        the buffer views, so the reduced values are what the optimizer reads.  No-op for gradients that already are the views."""
        if not self.pack:
            return
        src, dst, missing = [], [], False
        for p, view in zip(self.params, self._views):
            if p.grad is None:
                missing = True
            elif p.grad.data_ptr() != view.data_ptr():
                src.append(p.grad)
                dst.append(view)
        if missing and not getattr(self, '_zeroed_missing', False):
            # slots of parameters that never receive a gradient: zero once (nobody writes them afterwards)
            for p, view in zip(self.params, self._views):
                if p.grad is None:
                    view.zero_()
            self._zeroed_missing = True
        if src:
            torch._foreach_copy_(dst, src)
            for p, view in zip(self.params, self._views):
                if p.grad is not None:
                    p.grad = view

    def all_reduce_mean(self, group: Optional[dist.ProcessGroup] = None, async_op: bool = False):
        """SUM all-reduce of the flat bucket followed by division by the world size (DDP's gradient averaging).  Also the once-per-step
        place where pending tensor-core backward status snapshots are looked at (non-blocking; see ops.poll_status)."""
        if self.flat.is_cuda and not torch.cuda.is_current_stream_capturing():
            from . import ops
            ops.poll_status(self.flat.device)
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world = dist.get_world_size(group)
        if world == 1:
            return None
        self.pack_()
        self.flat.div_(world)
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def all_reduce_mean_async(self, group: Optional[dist.ProcessGroup] = None) -> "PendingReduce":
        """The same reduction enqueued on a SIDE stream behind everything the current stream has produced so far (start
        the all-reduce as soon as the fused backward has written the bucket, so that it overlaps the backward work that follows on
        the main stream).  ``.wait()`` on the result makes the current stream wait for the reduced bucket; the host never blocks."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return PendingReduce(None, None)
        world = dist.get_world_size(group)
        self.pack_()
        if not self.flat.is_cuda:                                   # gloo / CPU tests: no streams
            self.flat.div_(world)
            return PendingReduce(dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=True), None)
        if self._side is None:
            self._side = torch.cuda.Stream(self.flat.device)
        self._side.wait_stream(torch.cuda.current_stream(self.flat.device))
        with torch.cuda.stream(self._side):
            self.flat.div_(world)
            work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        return PendingReduce(work, self._side)


class PendingReduce:
    def __init__(self, work, side_stream):
        self.work, self.side = work, side_stream

    def wait(self) -> None:
        if self.work is None:
            return
        if self.side is None:
            self.work.wait()
            return
        with torch.cuda.stream(self.side):
            self.work.wait()                                        # orders the side stream after NCCL's stream (no host block)
        torch.cuda.current_stream(self.side.device).wait_stream(self.side)


def bind_host_to_gpu(local_rank: int) -> Optional[int]:
    """Pin this process (and the pinned host buffers it allocates afterwards) to the CPU cores next to GPU `local_rank` — NVML's
    CPU affinity mask of the device, i.e. its NUMA node.  One rank per GPU feeds its GPU over its own PCIe link; without this a rank
    scheduled on the other socket pays the inter-socket hop on every host->device byte.  Returns the number of cores bound, or None
    when NVML / sched_setaffinity is unavailable (then nothing changes)."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:                                        # noqa: BLE001 — an optimisation only; never fatal
        return None
