"""Batched on-device model ``transform`` after collation.

During validation the reference transforms a collated batch one utterance at a time on the
device and re-collates the outputs (``brever/training.py:336-338``:
``_collate_fn([model.transform(x[..., :l]) for x, l in zip(batch, lengths)])``,
``brever/data.py:408-491``).  The functions here take the zero-padded ``(B, 2, C, L)`` batch of
(mixture, foreground) pairs and ``lengths`` and return exactly what that loop returns -- the
transformed batch, zero beyond each item's own frame count, and the per-item frame counts --
with ONE launch per kernel for the whole batch instead of B Python iterations.

Why the batched result equals the per-item one: the samples past ``lengths[b]`` are zero in the
collated batch and in the right padding ``STFT.pad`` adds to the trimmed utterance alike, so
every frame ``t < n_frames(lengths[b])`` sees the same samples; the frames past it are masked to
the zeros ``_collate_fn`` pads with.
"""
import torch

from .ffnn import FFNNFrontEnd, _frame_mask


def _frames(stft, lengths):
    hop, fl = stft.hop_length, stft.frame_length
    frames0 = (torch.clamp(lengths - fl, min=0) + hop - 1) // hop + 1            # STFT.frame_count
    return 1 + ((frames0 - 1) * hop + fl + 2 * (stft.n_fft // 2) - stft.n_fft) // hop


def ffnn_transform_batched(front, batch, lengths):
    """``FFNN.transform`` (ffnn.py:77-91) -> ``(B, input_size + n_labels, T')``, ``(B,)`` frames."""
    return front.transform_batched(batch, lengths)


def sgmse_transform_batched(stft, batch, lengths, discard_nyquist=True):
    """``SGMSE.transform`` (models/sgmse/sgmse.py:148-158): channel mean, peak normalisation by
    the mixture, compressed STFT, optional Nyquist drop -> ``(B, 2, F, T)`` complex, ``(B,)`` frames."""
    if batch.ndim != 4 or batch.shape[1] != 2:
        raise ValueError(f'batch must be (B, 2, channels, samples), got {tuple(batch.shape)}')
    lengths = torch.as_tensor(lengths).to(device=batch.device, dtype=torch.int64)
    mono = batch.mean(dim=-2)                                     # (B, 2, L): make monaural
    mono = mono / mono[:, 0].abs().amax(dim=-1)[:, None, None]    # the padding is zero: same peak
    spec = stft(mono)                                             # (B, 2, F, T)
    if discard_nyquist:
        spec = spec[..., :-1, :]
    frames = _frames(stft, lengths)
    keep = torch.arange(spec.shape[-1], device=spec.device)[None, :] < frames[:, None]
    return spec * keep[:, None, None, :], frames


def transform_batched(model_kind, batch, lengths, **kw):
    """Dispatch on the reference's model registry names ('ffnn', 'sgmsep', 'sgmsepm', 'idmse')."""
    if model_kind == 'ffnn':
        front = kw.pop('front', None) or FFNNFrontEnd(**kw)
        return ffnn_transform_batched(front, batch, lengths)
    if model_kind in ('sgmsep', 'sgmsepm', 'sgmsepmheun', 'idmse'):
        return sgmse_transform_batched(kw.pop('stft'), batch, lengths, **kw)
    raise ValueError(f'no batched transform for model {model_kind!r}')


__all__ = ['transform_batched', 'ffnn_transform_batched', 'sgmse_transform_batched', '_frame_mask']
