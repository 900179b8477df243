"""The decoder stage's prologue and the two training losses around the SDE solve, each one fused fp32 launch with a
fused backward (csrc/stage_ops.cu):

    aggr_embed(module, local_embed, global_embed) -> hidden_0 [modes * N, 64]     dec_hivt_nusargo_sde.py:26-29, 82-85
    pi_head(module, local_embed, global_embed)    -> pi [N, modes]                 dec_hivt_nusargo_sde.py:63-67, 92-94 (fused forward)
    l2_loss(loc, target, reg_mask)               -> scalar                        losses/L2.py:10-27
    diff_bce_loss(diff_in, diff_out)             -> scalar                        losses/diff_BCE.py:11-16 (labels of enc…sep2.py:194-195)

CUDA only; the parameters stay in the caller's modules (reference checkpoints load unchanged)."""
import ctypes as C
from typing import Optional

import torch

from . import _lib
from .ops import LAUNCHES, _stream_ptr


def _ws(nbytes: int, dev) -> torch.Tensor:
    return torch.empty((max(int(nbytes), 1),), dtype=torch.uint8, device=dev)


def _cuda_only(t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError("trajsde_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")


# ---------------------------------------------------------------------------------------------------------------------
# aggr_embed
# ---------------------------------------------------------------------------------------------------------------------
def _aggr_args(local_embed, global_embed, w, b, g, beta, eps):
    modes, n = global_embed.shape[0], global_embed.shape[1]
    if global_embed.dim() != 3 or global_embed.shape[2] != 64 or tuple(local_embed.shape) != (n, 64):
        raise ValueError("global_embed [modes, N, 64] and local_embed [N, 64] expected")
    if tuple(w.shape) != (64, 128) or any(tuple(t.shape) != (64,) for t in (b, g, beta)):
        raise NotImplementedError("fused aggr_embed supports Linear(128, 64) + LayerNorm(64) only")
    keep = [t.detach().contiguous().float() for t in (global_embed, local_embed, w, b, g, beta)]
    a = _lib.AggrArgs()
    a.struct_bytes = C.sizeof(_lib.AggrArgs)
    a.n_modes, a.n_actors = modes, n
    a.global_embed, a.local_embed, a.w, a.b, a.ln_g, a.ln_b = (t.data_ptr() for t in keep)
    a.ln_eps = eps
    return a, keep, modes, n


class _AggrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, local_embed, global_embed, w, b, g, beta, eps):
        a, keep, modes, n = _aggr_args(local_embed, global_embed, w, b, g, beta, eps)
        out = torch.empty((modes * n, 64), dtype=torch.float32, device=global_embed.device)
        a.out = out.data_ptr()
        with torch.cuda.device(out.device):
            _lib.check(_lib.lib().trajsde_aggr_embed_fwd(C.byref(a), _stream_ptr(out.device)), "trajsde_aggr_embed_fwd")
        if modes * n > 0:
            LAUNCHES['n'] += 1
        ctx.save_for_backward(local_embed, global_embed, w, b, g, beta)
        ctx.eps = eps
        return out

    @staticmethod
    def backward(ctx, grad_out):
        local_embed, global_embed, w, b, g, beta = ctx.saved_tensors
        a, keep, modes, n = _aggr_args(local_embed, global_embed, w, b, g, beta, ctx.eps)
        dev = global_embed.device
        go = grad_out.contiguous()
        gg, gl = torch.empty_like(keep[0]), torch.empty_like(keep[1])
        gw, gb, gga, gbe = (torch.empty_like(t) for t in keep[2:6])
        L = _lib.lib()
        need = _lib.check(L.trajsde_aggr_embed_workspace_bytes(modes, n), "trajsde_aggr_embed_workspace_bytes")
        ws = _ws(need + 256, dev)
        base = (ws.data_ptr() + 255) & ~255
        a.grad_out, a.grad_global, a.grad_local = go.data_ptr(), gg.data_ptr(), gl.data_ptr()
        a.grad_w, a.grad_b, a.grad_ln_g, a.grad_ln_b = gw.data_ptr(), gb.data_ptr(), gga.data_ptr(), gbe.data_ptr()
        a.workspace, a.workspace_bytes = base, need
        with torch.cuda.device(dev):
            _lib.check(L.trajsde_aggr_embed_bwd(C.byref(a), _stream_ptr(dev)), "trajsde_aggr_embed_bwd")
        LAUNCHES['n'] += 5            # row flags + compaction + backward over the rows with a gradient + two reduces
        return gl, gg, gw, gb, gga, gbe, None


def aggr_embed(module: torch.nn.Module, local_embed: torch.Tensor, global_embed: torch.Tensor) -> torch.Tensor:
    """``module`` = the decoder's ``aggr_embed`` Sequential(Linear(128, 64), LayerNorm(64), ReLU) (dec…sde.py:26-29): returns
    ``hidden_0`` [modes * N, 64] = what :82-85 compute, without the [modes, N, 128] concatenation."""
    _cuda_only(global_embed)
    lin, ln = module[0], module[1]
    return _AggrFn.apply(local_embed, global_embed, lin.weight, lin.bias, ln.weight, ln.bias, float(ln.eps))


# ---------------------------------------------------------------------------------------------------------------------
# pi head
# ---------------------------------------------------------------------------------------------------------------------
def _pi_torch(local_embed, global_embed, w1, b1, g, beta, w2, b2, eps):
    """dec_hivt_nusargo_sde.py:92-94 in torch ops (the backward of ``pi_head`` differentiates this)."""
    x = torch.cat((local_embed.expand(global_embed.shape[0], *local_embed.shape), global_embed), dim=-1)
    z = torch.relu(torch.nn.functional.layer_norm(torch.nn.functional.linear(x, w1, b1), (64,), g, beta, eps))
    return torch.nn.functional.linear(z, w2.view(1, 64), b2).squeeze(-1).t()


class _PiFn(torch.autograd.Function):
    """Fused forward.  No loss of the reference configuration reads ``pi`` (yml:78-80: L2 + DiffBCE; pi feeds test-time metrics), so in
    training its gradient is None and backward returns at once; if a loss DOES use pi, backward recomputes the 4-layer head with
    torch ops under autograd (GPU, exact) — a cold path, not a fallback of the forward."""

    @staticmethod
    def forward(ctx, local_embed, global_embed, w1, b1, g, beta, w2, b2, eps):
        modes, n = global_embed.shape[0], global_embed.shape[1]
        if global_embed.dim() != 3 or global_embed.shape[2] != 64 or tuple(local_embed.shape) != (n, 64):
            raise ValueError("global_embed [modes, N, 64] and local_embed [N, 64] expected")
        keep = [t.detach().contiguous().float() for t in (global_embed, local_embed, w1, b1, g, beta, w2, b2)]
        a = _lib.PiArgs()
        a.struct_bytes = C.sizeof(_lib.PiArgs)
        a.n_modes, a.n_actors = modes, n
        a.global_embed, a.local_embed, a.w1, a.b1, a.ln_g, a.ln_b, a.w2, a.b2 = (t.data_ptr() for t in keep)
        a.ln_eps = eps
        out = torch.empty((n, modes), dtype=torch.float32, device=global_embed.device)
        a.out = out.data_ptr()
        with torch.cuda.device(out.device):
            _lib.check(_lib.lib().trajsde_pi_head_fwd(C.byref(a), _stream_ptr(out.device)), "trajsde_pi_head_fwd")
        if modes * n > 0:
            LAUNCHES['n'] += 1
        ctx.save_for_backward(local_embed, global_embed, w1, b1, g, beta, w2, b2)
        ctx.eps = eps
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, grad_pi):
        if grad_pi is None:
            return (None,) * 9
        saved = ctx.saved_tensors
        with torch.enable_grad():
            leaves = [t.detach().requires_grad_(need) for t, need in zip(saved, ctx.needs_input_grad[:8])]
            pi = _pi_torch(*leaves, ctx.eps)
            wanted = [t for t in leaves if t.requires_grad]
            grads = iter(torch.autograd.grad(pi, wanted, grad_pi, allow_unused=True)) if wanted else iter(())
        return tuple(next(grads) if t.requires_grad else None for t in leaves) + (None,)


def pi_head(module: torch.nn.Module, local_embed: torch.Tensor, global_embed: torch.Tensor) -> torch.Tensor:
    """``module`` = the decoder's ``pi`` Sequential(Linear(128, 64), LayerNorm(64), ReLU, Linear(64, 1)) (dec…sde.py:63-67): returns
    ``pi`` [N, modes] = what :92-94 compute, in one launch without the [modes, N, 128] concatenation."""
    _cuda_only(global_embed)
    lin, ln, proj = module[0], module[1], module[3]
    if tuple(lin.weight.shape) != (64, 128) or tuple(proj.weight.shape) != (1, 64) or tuple(ln.weight.shape) != (64,):
        raise NotImplementedError("fused pi head supports Linear(128, 64) + LayerNorm(64) + ReLU + Linear(64, 1) only")
    return _PiFn.apply(local_embed, global_embed, lin.weight, lin.bias, ln.weight, ln.bias, proj.weight.reshape(64), proj.bias, float(ln.eps))


# ---------------------------------------------------------------------------------------------------------------------
# L2
# ---------------------------------------------------------------------------------------------------------------------
class _L2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loc, target, reg_mask):
        if loc.dim() != 4 or loc.shape[3] not in (2, 4) or loc.dtype != torch.float32:
            raise ValueError("loc must be float32 [modes, N, T, 2 or 4]")
        modes, n, T, c = loc.shape
        if tuple(target.shape) != (n, T, 2) or tuple(reg_mask.shape) != (n, T) or reg_mask.dtype != torch.bool:
            raise ValueError("target [N, T, 2] float and reg_mask [N, T] bool expected")
        dev = loc.device
        locc, tg, rm = loc.detach().contiguous(), target.detach().contiguous().float(), reg_mask.contiguous().view(torch.uint8)
        out = torch.empty((2,), dtype=torch.float32, device=dev)              # loss, count
        best = torch.empty((n,), dtype=torch.int32, device=dev)
        a = _lib.L2Args()
        a.struct_bytes = C.sizeof(_lib.L2Args)
        a.n_modes, a.n_actors, a.n_t = modes, n, T
        a.loc, a.loc_stride, a.target, a.reg_mask = locc.data_ptr(), c, tg.data_ptr(), rm.data_ptr()
        a.loss, a.count, a.best_mode = out.data_ptr(), out.data_ptr() + 4, best.data_ptr()
        L = _lib.lib()
        need = _lib.check(L.trajsde_l2_loss_workspace_bytes(n), "trajsde_l2_loss_workspace_bytes")
        ws = _ws(need, dev)
        a.workspace, a.workspace_bytes = ws.data_ptr(), need
        with torch.cuda.device(dev):
            _lib.check(L.trajsde_l2_loss_fwd(C.byref(a), _stream_ptr(dev)), "trajsde_l2_loss_fwd")
        LAUNCHES['n'] += 2
        ctx.save_for_backward(locc, tg, rm, out, best)
        ctx.mark_non_differentiable(best)
        return out[0], best

    @staticmethod
    def backward(ctx, grad_loss, _grad_best):
        locc, tg, rm, out, best = ctx.saved_tensors
        modes, n, T, c = locc.shape
        dev = locc.device
        gl = grad_loss.detach().reshape(1).float().contiguous()
        grad_loc = torch.zeros_like(locc)
        a = _lib.L2Args()
        a.struct_bytes = C.sizeof(_lib.L2Args)
        a.n_modes, a.n_actors, a.n_t = modes, n, T
        a.loc, a.loc_stride, a.target, a.reg_mask = locc.data_ptr(), c, tg.data_ptr(), rm.data_ptr()
        a.count, a.best_mode = out.data_ptr() + 4, best.data_ptr()
        a.grad_loss, a.grad_loc, a.grad_loc_stride = gl.data_ptr(), grad_loc.data_ptr(), c
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().trajsde_l2_loss_bwd(C.byref(a), _stream_ptr(dev)), "trajsde_l2_loss_bwd")
        LAUNCHES['n'] += 1
        return grad_loc, None, None


def l2_loss(loc: torch.Tensor, target: torch.Tensor, reg_mask: torch.Tensor, return_best: bool = False):
    """``L2.forward`` (losses/L2.py:10-27) on ``output['loc']`` [modes, N, T, 4] (cat of loc and scale, :12) or a plain [modes, N, T, 2]
    tensor, ``data['y']`` [N, T, 2] and ``output['reg_mask']`` [N, T]: one launch instead of norm / clone / masked assign / mean /
    argmin / gather / masked mean.  ``return_best`` also returns the winning mode per actor (int32 [N])."""
    _cuda_only(loc)
    loss, best = _L2Fn.apply(loc, target, reg_mask)
    return (loss, best) if return_best else loss


# ---------------------------------------------------------------------------------------------------------------------
# DiffBCE
# ---------------------------------------------------------------------------------------------------------------------
class _BceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, diff_in, diff_out):
        dev = diff_in.device
        di, do = diff_in.detach().contiguous().float().reshape(-1), diff_out.detach().contiguous().float().reshape(-1)
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        gi, go = torch.empty_like(di), torch.empty_like(do)
        a = _lib.BceArgs()
        a.struct_bytes = C.sizeof(_lib.BceArgs)
        a.n_in, a.n_out = di.numel(), do.numel()
        a.diff_in, a.diff_out, a.loss, a.grad_in, a.grad_out = di.data_ptr(), do.data_ptr(), loss.data_ptr(), gi.data_ptr(), go.data_ptr()
        L = _lib.lib()
        need = _lib.check(L.trajsde_diff_bce_workspace_bytes(), "trajsde_diff_bce_workspace_bytes")
        ws = _ws(need, dev)
        a.workspace, a.workspace_bytes = ws.data_ptr(), need
        with torch.cuda.device(dev):
            _lib.check(L.trajsde_diff_bce(C.byref(a), _stream_ptr(dev)), "trajsde_diff_bce")
        LAUNCHES['n'] += 2
        ctx.save_for_backward(gi, go)
        ctx.shapes = (diff_in.shape, diff_out.shape)
        return loss[0]

    @staticmethod
    def backward(ctx, grad_loss):
        gi, go = ctx.saved_tensors
        return (gi * grad_loss).reshape(ctx.shapes[0]), (go * grad_loss).reshape(ctx.shapes[1])


def diff_bce_loss(diff_in: torch.Tensor, diff_out: torch.Tensor) -> torch.Tensor:
    """``DiffBCE.forward`` (losses/diff_BCE.py:11-16) with the encoder's labels (in -> 0, out -> 1): BCE(diff_in, 0) + BCE(diff_out, 1),
    mean reduction, in one launch that also produces the gradients.  Any shapes (the reference passes [B, 64] with 64 identical
    columns; the per-row diffusion [B] gives the same mean)."""
    _cuda_only(diff_in)
    return _BceFn.apply(diff_in, diff_out)
