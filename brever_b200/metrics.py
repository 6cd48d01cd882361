"""Metric wrappers over the criterion path (mirror of ``brever/metrics.py:112-150``).

Only the two metrics that run on the hot path are here: ``snr`` and ``sisnr`` are the
negated criteria evaluated on ``(B, L)`` or ``(L,)`` waveforms, with the reference's
input checks, default lengths and ``.item()`` for unbatched input.  STOI / ESTOI / PESQ
(``brever/metrics.py:19-109,153-222``) call third-party CPU libraries and stay with the
reference (out of scope).
"""
import torch

from .criterion import CriterionRegistry
from .registry import Registry

MetricRegistry = Registry('metric')


def _check_input(x, y, lengths):
    """brever/metrics.py:126-150: same shapes, add the batch and source dimensions,
    default / validated lengths.  Same exceptions and messages."""
    if x.shape != y.shape:
        raise ValueError('inputs must have same shape, got '
                         f'{x.shape} and {y.shape}')
    # add batch dimension
    unbatched = x.ndim == 1
    if unbatched:
        x, y = x.unsqueeze(0), y.unsqueeze(0)
    # add source dimension
    if x.ndim == 2:
        x, y = x.unsqueeze(1), y.unsqueeze(1)
    else:
        raise ValueError(f'input must be 1 or 2 dimensional, got {x.ndim}')
    # check lengths items are smaller than input length
    if lengths is None:
        lengths = torch.full((x.shape[0],), x.shape[-1], device=x.device)
    else:
        if len(lengths) != x.shape[0]:
            raise ValueError('lengths must have same length as batch size, '
                             f'got {len(lengths)} and {x.shape[0]}')
        if any(length > x.shape[-1] for length in lengths):
            raise ValueError('lengths items must be smaller than input '
                             f'length, got lengths={lengths} and '
                             f'input.shape={x.shape}')
    return x, y, lengths, unbatched


@MetricRegistry.register('snr')
def snr(x, y, lengths=None):
    """brever/metrics.py:112-116."""
    x, y, lengths, unbatched = _check_input(x, y, lengths)
    output = - CriterionRegistry.get('snr')(x, y, lengths)
    return output.item() if unbatched else output


@MetricRegistry.register('sisnr')
def sisnr(x, y, lengths=None):
    """brever/metrics.py:119-123."""
    x, y, lengths, unbatched = _check_input(x, y, lengths)
    output = - CriterionRegistry.get('sisnr')(x, y, lengths)
    return output.item() if unbatched else output
