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
import functools
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
    """`lengths` as a contiguous int64 tensor on `device`.  A device tensor passes through; a
    host tensor / list costs one pageable host-to-device copy per call (it synchronises the
    host and cannot be captured in a CUDA graph): keep lengths on the device in hot loops."""
    if not isinstance(lengths, torch.Tensor):
        lengths = torch.as_tensor(lengths)
    if lengths.device == device and lengths.dtype == torch.int64 and lengths.is_contiguous():
        return lengths
    return lengths.to(device=device, dtype=torch.int64).contiguous()


@functools.lru_cache(maxsize=None)
def _perms_on(n_src, device):
    """All permutations of the sources, resident on `device` (built once per (S, device): a
    per-call torch.tensor(..., device=cuda) is a host-to-device copy inside the hot path)."""
    idx = torch.tensor(list(permutations(range(n_src))), dtype=torch.int64)
    return idx.to(device), torch.arange(n_src, device=device)


def _rows(t):
    """View (B, ..., L) as (B, R, L) with unit stride on L (copy only if needed)."""
    t3 = t.reshape(t.shape[0], -1, t.shape[-1])
    if t3.dtype != torch.float32:
        t3 = t3.float()
    if t3.stride(-1) != 1 and t3.shape[-1] > 1:
        t3 = t3.contiguous()
    return t3


def _workspace(nbytes, n_tickets, device):
    """Reduction workspace, cached per (device, stream): `n_tickets` uint32 counters first (the
    kernels need them zero on entry and leave them zero), partial sums after.  The partial sums
    stay dirty, so when a call puts its counters over bytes that no earlier call used as
    counters, that range is zeroed first (stream-ordered) -- a small-batch / long-rows call
    followed by a many-rows call would otherwise read stale partial sums as counters."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ticket_bytes = (4 * int(n_tickets) + 255) & ~255
    entry = _workspaces.get(key)
    if entry is None or entry[0].numel() < nbytes:
        ws = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        entry = [ws, ws.numel()]
        _workspaces[key] = entry
    ws, clean = entry                      # bytes [0, clean) are zero at rest
    if ticket_bytes > clean:
        ws[clean:ticket_bytes].zero_()
    entry[1] = ticket_bytes                # everything after this call's counters is dirty now
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
        ws = _workspace(nbytes, pairs, x3.device)
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


def _criterion_backward(x3, y3, lengths, mom, grad, gscale, pairwise, ymap=None):
    """One launch: coefficients from the saved moments + masked affine map."""
    batch, rows, length = x3.shape
    gx = torch.empty((batch, rows, length), dtype=torch.float32, device=x3.device)
    if gx.numel():
        g = grad.float().reshape(-1).contiguous()
        stride = 0 if g.numel() == 1 else 1
        with _lib.on_device(x3.device):
            _lib.check(_lib.lib().brv_criterion_backward(
                _lib.ptr(x3), _lib.ptr(y3), _lib.ptr(lengths), _lib.ptr(mom), _lib.ptr(g),
                stride, float(gscale), _lib.ptr(ymap), int(pairwise), batch, rows, length,
                x3.stride(0), x3.stride(1), y3.stride(0), y3.stride(1), float(eps),
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
        # loss = -mean dB: a 2-D input averages over the batch too (the mean(()) quirk)
        gscale = 1.0 / batch if len(ctx.shape) == 2 else 1.0 / rows
        gx = _criterion_backward(x3, y3, lengths, mom, grad, gscale, False)
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
        perms, src_idx = _perms_on(n_src, x.device)
        # snr_set[b, p] = sum_i si_snr[b, i, perms[p, i]]   (criterion.py:66-68)
        gathered = si_snr[:, src_idx, perms]
        totals = gathered.sum(-1)
        best, which = totals.max(1)
        perm = perms[which]                        # target i <- estimate perm[b, i]
        best = best / n_src
        ctx.save_for_backward(x3, y3, lengths, mom, perm)
        ctx.shape, ctx.dtype = x.shape, x.dtype
        return -best

    @staticmethod
    def backward(ctx, grad):
        if len(ctx.saved_tensors) == 4:          # one source: no permutation
            x3, y3, lengths, mom = ctx.saved_tensors
            gx = _criterion_backward(x3, y3, lengths, mom, grad, 1.0, True)
            return gx.view(ctx.shape).to(ctx.dtype), None, None
        x3, y3, lengths, mom, perm = ctx.saved_tensors
        batch, n_src, _ = x3.shape
        dev = x3.device
        # estimate row j = perm[b, i] is matched with target i: ymap[b, j] = i
        ymap = torch.empty((batch, n_src), dtype=torch.int64, device=dev)
        ymap.scatter_(1, perm, torch.arange(n_src, device=dev).expand(batch, n_src))
        gx = _criterion_backward(x3, y3, lengths, mom, grad, 1.0 / n_src, True,
                                 ymap.to(torch.int32).contiguous())
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


class _MseFunction(torch.autograd.Function):
    """sum_{n < length} (x - y)^2 per row: the `sum (y-x)^2` moment of the fused reduction."""

    @staticmethod
    def forward(ctx, x3, y3, lengths):
        _, mom = _moments(x3, y3, lengths, False)
        ctx.save_for_backward(x3, y3, lengths)
        return mom[:, 5].view(x3.shape[0], x3.shape[1])

    @staticmethod
    def backward(ctx, grad):
        x3, y3, lengths = ctx.saved_tensors
        c = (2 * grad).reshape(-1).float().contiguous()
        gx = _masked_affine(x3, y3, lengths, c, (-c).contiguous(), torch.zeros_like(c))
        return gx, None, None


@CriterionRegistry.register('mse')
def mse(x, y, lengths, weight=None):
    """Masked mean squared error, ``(B, ..., L) -> (B,)``, real or complex, optionally
    weighted per batch item (criterion.py:104-132).  One pass of the fused moment kernel;
    complex inputs are read as interleaved (re, im) rows of length ``2 L``."""
    assert x.shape == y.shape
    assert x.ndim >= 2
    _lib.require_cuda(x, 'mse estimate')
    _lib.require_cuda(y, 'mse target')
    lengths = _lengths_on(lengths, x.device)
    out_dtype = x.real.dtype if x.is_complex() else x.dtype
    mask_len = lengths
    if x.is_complex():
        if not y.is_complex():
            y = y.to(x.dtype)
        xr, yr = torch.view_as_real(x.resolve_conj()), torch.view_as_real(y.resolve_conj())
        x3 = _rows(xr.reshape(*xr.shape[:-2], -1))
        y3 = _rows(yr.reshape(*yr.shape[:-2], -1))
        mask_len = lengths * 2
    else:
        x3, y3 = _rows(x), _rows(y)
    loss = _MseFunction.apply(x3, y3, mask_len)                 # (B, R) float64
    loss = loss / lengths.view(-1, 1)
    if weight is not None:
        loss = loss * weight.to(loss.device).view(-1, 1)
    loss = loss.mean() if x.ndim == 2 else loss.mean(1)         # mean(()) quirk, as in snr
    return loss.to(out_dtype)


class _L1RowsFunction(torch.autograd.Function):
    """sum_{n < length} |s x - y| per row (time-domain term of MultiResYuLoss)."""

    @staticmethod
    def forward(ctx, x3, y3, lengths, scale):
        batch, rows, length = x3.shape
        out = torch.empty(batch * rows, dtype=torch.float32, device=x3.device)
        if out.numel():
            lib = _lib.lib()
            nbytes = lib.brv_l1_workspace_bytes(batch * rows, length)
            ws = _workspace(nbytes, batch * rows, x3.device)
            with _lib.on_device(x3.device):
                _lib.check(lib.brv_l1_forward(
                    _lib.ptr(x3), _lib.ptr(y3), _lib.ptr(lengths), _lib.ptr(scale),
                    batch, rows, length, x3.stride(0), x3.stride(1), y3.stride(0),
                    y3.stride(1), _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                    _lib.stream_ptr(x3.device)))
        ctx.save_for_backward(x3, y3, lengths, scale)
        return out.view(batch, rows)

    @staticmethod
    def backward(ctx, grad):
        x3, y3, lengths, scale = ctx.saved_tensors
        batch, rows, length = x3.shape
        coef = grad.reshape(-1).float()
        if scale is not None:
            coef = coef * scale
        coef = coef.contiguous()
        gx = torch.empty((batch, rows, length), dtype=torch.float32, device=x3.device)
        if gx.numel():
            with _lib.on_device(x3.device):
                _lib.check(_lib.lib().brv_l1_backward(
                    _lib.ptr(x3), _lib.ptr(y3), _lib.ptr(lengths), _lib.ptr(scale),
                    _lib.ptr(coef), batch, rows, length, x3.stride(0), x3.stride(1),
                    y3.stride(0), y3.stride(1), _lib.ptr(gx), _lib.stream_ptr(x3.device)))
        return gx, None, None, None


class _MagL1Function(torch.autograd.Function):
    """sum over bins and frames of | |X| - |Y| | per signal (spectral term).  X, Y are
    (n_sig, T*F) contiguous complex64: the frame-major memory STFT.forward returns."""

    @staticmethod
    def forward(ctx, X, Y):
        n_sig, n_elems = X.shape
        out = torch.empty(n_sig, dtype=torch.float32, device=X.device)
        if n_sig:
            lib = _lib.lib()
            nbytes = lib.brv_l1_workspace_bytes(n_sig, n_elems)
            ws = _workspace(nbytes, n_sig, X.device)
            with _lib.on_device(X.device):
                _lib.check(lib.brv_mag_l1_forward(
                    _lib.ptr(X), _lib.ptr(Y), n_sig, n_elems, _lib.ptr(out), _lib.ptr(ws),
                    ws.numel(), _lib.stream_ptr(X.device)))
        ctx.save_for_backward(X, Y)
        return out

    @staticmethod
    def backward(ctx, grad):
        X, Y = ctx.saved_tensors
        gX = torch.empty_like(X)
        if X.numel():
            coef = grad.float().contiguous()
            with _lib.on_device(X.device):
                _lib.check(_lib.lib().brv_mag_l1_backward(
                    _lib.ptr(X), _lib.ptr(Y), _lib.ptr(coef), X.shape[0], X.shape[1],
                    _lib.ptr(gX), _lib.stream_ptr(X.device)))
        return gX, None


@CriterionRegistry.register('multiresyu')
class MultiResYuLoss:
    """Multi-resolution STFT-magnitude + L1 time-domain loss (criterion.py:135-226), the
    default criterion of TF-GridNet (tfgridnet.py:58).

    Boxcar-window, unnormalised STFTs of the masked estimate and target on the tcgen05
    kernels; the time-domain and spectral L1 sums are fused reductions (``brv_l1_forward``,
    ``brv_mag_l1_forward``), so no magnitude or difference tensor is materialised.  The
    gradient flows back through ``brv_mag_l1_backward`` into the STFT gradient kernel.
    With ``scale_invariant=True`` the scaling factor comes from the fused moment kernel
    without gradients, and from a few broadcast ops kept in the autograd graph with them.
    """

    def __init__(self, frame_lengths=[512], hop_lengths=None, time_domain_weight=0.5,
                 spectral_weight=0.5, scale_invariant=False):
        from .modules.stft import STFT
        if hop_lengths is None:
            hop_lengths = [x // 2 for x in frame_lengths]
        self.stfts = [STFT(frame_length=fl, hop_length=hl, window=None, normalized=False)
                      for fl, hl in zip(frame_lengths, hop_lengths)]
        self.time_domain_weight = time_domain_weight
        self.spectral_weight = spectral_weight
        self.scale_invariant = scale_invariant

    def __call__(self, x, y, lengths):
        assert x.shape == y.shape
        _lib.require_cuda(x, 'multiresyu estimate')
        _lib.require_cuda(y, 'multiresyu target')
        lengths = _lengths_on(lengths, x.device)
        assert len(lengths) == x.shape[0]
        out_dtype = x.dtype
        x3, y3 = _rows(x), _rows(y)
        batch, rows, length = x3.shape
        scale = None
        if self.scale_invariant and torch.is_grad_enabled() and x.requires_grad:
            # the scaling factor depends on the estimate (criterion.py:207-211): keep it in the
            # autograd graph -- d(a x)/dx = a g + <g, x> (y - 2 a x) / (sum x^2 + eps) comes out
            # of these few broadcast ops on the masked rows; the L1 / STFT terms below see a x
            xm, ym = apply_mask(x3, y3, lengths)
            a = (xm * ym).sum(-1, keepdim=True) / (xm.pow(2).sum(-1, keepdim=True) + eps)
            x3 = a * xm
        elif self.scale_invariant:
            _, mom = _moments(x3, y3, lengths, True) if rows == 1 else (None, None)
            if mom is None:     # pairwise moments are S x S: take the matched pairs directly
                xm, ym = apply_mask(x3, y3, lengths)
                sxy, sxx = (xm * ym).sum(-1).double(), xm.pow(2).sum(-1).double()
            else:
                sxy, sxx = mom[:, 2].view(batch, rows), mom[:, 3].view(batch, rows)
            scale = (sxy / (sxx + eps)).float().reshape(-1).contiguous()
        total = self.time_domain_weight * _L1RowsFunction.apply(x3, y3, lengths, scale)
        xm, ym = apply_mask(x3, y3, lengths)
        if scale is not None:
            xm = xm * scale.view(batch, rows, 1)
        for stft in self.stfts:
            X, Y = stft(xm), stft(ym)                    # (B, R, F, T) over frame-major memory
            Xf = X.transpose(-1, -2).reshape(batch * rows, -1)
            Yf = Y.transpose(-1, -2).reshape(batch * rows, -1)
            spectral = _MagL1Function.apply(Xf, Yf).view(batch, rows)
            total = total + self.spectral_weight * spectral / len(self.stfts)
        total = total / lengths.view(-1, 1)
        total = total.mean() if x.ndim == 2 else total.mean(1)
        return total.to(out_dtype) if out_dtype == torch.float64 else total


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
