"""Drop-in ``snr`` / ``sisnr`` criteria and ``apply_mask`` (brever/criterion.py).

Same registry names, signatures ``f(x, y, lengths) -> (B,)`` and quirks as the
reference; one fused kernel pass per call (``brv_snr_forward``) plus one
elementwise kernel for the gradient (``brv_masked_affine``).

Kept quirks: ``eps = finfo(float32).eps`` inside and outside the ratio
(criterion.py:60-61,99-100); no eps on ``||s||^2`` (:58); means divided by
``lengths`` (:48-49); a 2-D ``snr`` input reduces over ALL dims because
``mean(())`` does (:101, hit by DCCRN); PIT through the arg-max permutation.
Unlike the reference's ``sisnr`` (in-place ``/=`` on the ``amax`` output,
criterion.py:69-70) ours can be back-propagated.
"""
import inspect
import math
from itertools import permutations

import torch

from . import _lib
from .registry import Registry

eps = torch.finfo(torch.float32).eps

CriterionRegistry = Registry('criterion')

_K = 10.0 / math.log(10.0)
_workspaces = {}


def init_criterion(name, **kwargs):
    criterion = CriterionRegistry.get(name)
    if inspect.isclass(criterion):
        criterion = criterion(**kwargs)
    return criterion


def _lengths_on(lengths, device):
    if not isinstance(lengths, torch.Tensor):
        lengths = torch.as_tensor(lengths)
    return lengths.to(device=device, dtype=torch.int64).contiguous()


def _rows(t):
    """View (B, ..., L) as (B, R, L) with unit stride on L (copy only if needed)."""
    t3 = t.reshape(t.shape[0], -1, t.shape[-1])
    if t3.dtype != torch.float32:
        t3 = t3.float()
    if t3.stride(-1) != 1 and t3.shape[-1] > 1:
        t3 = t3.contiguous()
    return t3


def _workspace(nbytes, device):
    """Ticket buffer the kernel leaves zeroed; cached per (device, stream)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _moments(x3, y3, lengths, pairwise, sign=1.0):
    """Launch the fused reduction -> (sign * dB (pairs,), moments (pairs, 6) float64)."""
    batch, rows, length = x3.shape
    pairs = batch * rows * rows if pairwise else batch * rows
    db = torch.empty(pairs, dtype=torch.float32, device=x3.device)
    mom = torch.empty((pairs, 6), dtype=torch.float64, device=x3.device)
    if pairs:
        lib = _lib.lib()
        nbytes = lib.brv_snr_workspace_bytes(pairs, length)
        ws = _workspace(nbytes, x3.device)
        with _lib.on_device(x3.device):
            _lib.check(lib.brv_snr_forward(
                _lib.ptr(x3), _lib.ptr(y3), _lib.ptr(lengths), batch, rows,
                length, x3.stride(0), x3.stride(1), y3.stride(0), y3.stride(1),
                int(pairwise), float(eps), float(sign), _lib.ptr(db),
                _lib.ptr(mom), _lib.ptr(ws), ws.numel(),
                _lib.stream_ptr(x3.device)))
    return db, mom


def _masked_affine(x3, y3, lengths, ca, cb, c0, ymap=None):
    batch, rows, length = x3.shape
    gx = torch.empty((batch, rows, length), dtype=torch.float32, device=x3.device)
    if gx.numel():
        with _lib.on_device(x3.device):
            _lib.check(_lib.lib().brv_masked_affine(
                _lib.ptr(x3), _lib.ptr(y3), _lib.ptr(lengths), batch, rows,
                length, x3.stride(0), x3.stride(1), y3.stride(0), y3.stride(1),
                _lib.ptr(ca), _lib.ptr(cb), _lib.ptr(c0), _lib.ptr(ymap),
                _lib.ptr(gx), _lib.stream_ptr(x3.device)))
    return gx


class _SnrFunction(torch.autograd.Function):
    """-(mean over rows of 10 log10(sum y^2 / (sum (y-x)^2 + eps) + eps))."""

    @staticmethod
    def forward(ctx, x, y, lengths):
        x3, y3 = _rows(x), _rows(y)
        db, mom = _moments(x3, y3, lengths, False, -1.0)   # the kernel negates
        ctx.save_for_backward(x3, y3, lengths, mom)
        ctx.shape, ctx.dtype = x.shape, x.dtype
        db = db.view(x3.shape[0], x3.shape[1])
        if x.ndim == 2:       # torch's mean(()) reduces over every dim
            return db.mean()
        return db.view(-1) if x3.shape[1] == 1 else db.mean(1)

    @staticmethod
    def backward(ctx, grad):
        x3, y3, lengths, mom = ctx.saved_tensors
        batch, rows, _ = x3.shape
        p, d = mom[:, 4], mom[:, 5]
        r = p / (d + eps)
        # d(dB)/dx_n = K * 2P / ((r+eps)(D+eps)^2) * (y_n - x_n)
        coef = _K * 2 * p / ((r + eps) * (d + eps) ** 2)
        if len(ctx.shape) == 2:
            g = grad.reshape(1).expand(batch) / batch
        else:
            g = grad.reshape(batch) / rows
        cb = (-(g.double().repeat_interleave(rows)) * coef).float()
        gx = _masked_affine(x3, y3, lengths, (-cb).contiguous(), cb.contiguous(),
                            torch.zeros_like(cb))
        return gx.view(ctx.shape).to(ctx.dtype), None, None


class _SiSnrFunction(torch.autograd.Function):
    """Pairwise SI-SNR matrix + PIT (criterion.py:45-72)."""

    @staticmethod
    def forward(ctx, x, y, lengths):
        batch, n_src, _ = x.shape
        x3, y3 = _rows(x), _rows(y)
        if n_src == 1:                             # no permutation: the kernel negates
            loss, mom = _moments(x3, y3, lengths, True, -1.0)
            ctx.save_for_backward(x3, y3, lengths, mom)
            ctx.shape, ctx.dtype = x.shape, x.dtype
            return loss
        db, mom = _moments(x3, y3, lengths, True)
        si_snr = db.view(batch, n_src, n_src)      # [b, target i, estimate j]
        perms = torch.tensor(list(permutations(range(n_src))),
                             dtype=torch.int64, device=x.device)
        # snr_set[b, p] = sum_i si_snr[b, i, perms[p, i]]   (criterion.py:66-68)
        gathered = si_snr[:, torch.arange(n_src, device=x.device), perms]
        totals = gathered.sum(-1)
        best, which = totals.max(1)
        perm = perms[which]                        # target i <- estimate perm[b, i]
        best = best / n_src
        ctx.save_for_backward(x3, y3, lengths, mom, perm)
        ctx.shape, ctx.dtype = x.shape, x.dtype
        return -best

    @staticmethod
    def backward(ctx, grad):
        if len(ctx.saved_tensors) == 4:
            x3, y3, lengths, mom = ctx.saved_tensors
            perm = torch.zeros((x3.shape[0], 1), dtype=torch.int64, device=x3.device)
        else:
            x3, y3, lengths, mom, perm = ctx.saved_tensors
        batch, n_src, _ = x3.shape
        dev = x3.device
        # moments of the matched pairs: pair index = (b*S + i)*S + perm[b, i]
        tgt = torch.arange(n_src, device=dev).expand(batch, n_src)
        pair = (torch.arange(batch, device=dev)[:, None] * n_src + tgt) * n_src + perm
        m = mom[pair.reshape(-1)]                  # (B*S, 6) ordered by (b, target i)
        sx, sy, sxy, sxx, syy = m[:, 0], m[:, 1], m[:, 2], m[:, 3], m[:, 4]
        L = lengths.double().repeat_interleave(n_src)
        n = torch.minimum(L, torch.full_like(L, float(x3.shape[-1])))
        mx, my = sx / L, sy / L
        dot = sxy - mx * sy - my * sx + n * mx * my
        ea = sxx - 2 * mx * sx + n * mx * mx
        eb = syy - 2 * my * sy + n * my * my
        t = dot * dot / eb
        e = (ea - t).clamp_min(0)
        r = t / (e + eps)
        common = _K / ((r + eps) * (e + eps) ** 2)
        alpha = common * (e + eps + t) * (2 * dot / eb)   # multiplies b = y - my
        beta = -2 * t * common                            # multiplies a = x - mx
        g = (-grad.double() / n_src).repeat_interleave(n_src)
        ca_t, cb_t = g * beta, g * alpha
        c0_t = g * (-alpha * my - beta * mx)
        # scatter from (b, target i) order to estimate rows j = perm[b, i]
        row = (torch.arange(batch, device=dev)[:, None] * n_src + perm).reshape(-1)
        ca = torch.empty_like(ca_t).index_copy_(0, row, ca_t).float()
        cb = torch.empty_like(cb_t).index_copy_(0, row, cb_t).float()
        c0 = torch.empty_like(c0_t).index_copy_(0, row, c0_t).float()
        ymap = torch.empty(batch * n_src, dtype=torch.int32, device=dev)
        ymap.index_copy_(0, row, tgt.reshape(-1).to(torch.int32))
        gx = _masked_affine(x3, y3, lengths, ca, cb, c0, ymap)
        return gx.view(ctx.shape).to(ctx.dtype), None, None


@CriterionRegistry.register('sisnr')
def sisnr(x, y, lengths):
    """Scale-invariant SNR loss with PIT, ``(B, S, L) -> (B,)`` (criterion.py:21-72)."""
    assert x.shape == y.shape
    assert x.ndim == 3
    _lib.require_cuda(x, 'sisnr estimate')
    _lib.require_cuda(y, 'sisnr target')
    out = _SiSnrFunction.apply(x, y, _lengths_on(lengths, x.device))
    return out.to(x.dtype) if x.dtype == torch.float64 else out


@CriterionRegistry.register('snr')
def snr(x, y, lengths):
    """SNR loss without PIT, ``(B, ..., L) -> (B,)`` (criterion.py:75-101)."""
    assert x.shape == y.shape
    assert x.ndim >= 2
    _lib.require_cuda(x, 'snr estimate')
    _lib.require_cuda(y, 'snr target')
    out = _SnrFunction.apply(x, y, _lengths_on(lengths, x.device))
    return out.to(x.dtype) if x.dtype == torch.float64 else out


class _MaskFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, lengths):
        ctx.save_for_backward(lengths)
        return _mask_raw(t, lengths)

    @staticmethod
    def backward(ctx, grad):
        lengths, = ctx.saved_tensors
        return _mask_raw(grad, lengths), None


def _mask_raw(t, lengths):
    src = t.contiguous()
    if src.dtype != torch.float32:
        src = src.float()
    out = torch.empty_like(src)
    if src.numel():
        inner = src.numel() // (src.shape[0] * src.shape[-1])
        with _lib.on_device(src.device):
            _lib.check(_lib.lib().brv_apply_mask(
                _lib.ptr(src), _lib.ptr(lengths), src.shape[0], inner,
                src.shape[-1], _lib.ptr(out), _lib.stream_ptr(src.device)))
    return out if t.dtype in (torch.float32, torch.float16, torch.bfloat16) \
        else out.to(t.dtype)


def apply_mask(x, y, lengths):
    """Zero both tensors at and beyond ``lengths[i]`` (criterion.py:229-234)."""
    assert len(lengths) == x.shape[0]
    _lib.require_cuda(x, 'apply_mask input')
    _lib.require_cuda(y, 'apply_mask input')
    lengths = _lengths_on(lengths, x.device)
    return _MaskFunction.apply(x, lengths), _MaskFunction.apply(y, lengths)
