"""CUDA-graph capture of a front-end step.

The hot path is a handful of short kernels (STFT ~50 us, iSTFT ~70 us, criterion
~10 us at 64 x 4 s): launched eagerly from Python the GPU waits on the host between
them.  ``capture`` records a callable that uses only ``brever_b200`` ops (and
stream-ordered torch ops) into one ``torch.cuda.CUDAGraph``; a replay re-issues the
whole chain with a single host call.

    step = brv.graphs.capture(lambda mix, tgt: brv.snr(stft.backward(net(stft(mix))), tgt, lengths),
                              mix_buf, tgt_buf)
    for batch in loader:
        mix_buf.copy_(batch.mix, non_blocking=True)      # refill the static inputs in place
        tgt_buf.copy_(batch.tgt, non_blocking=True)
        loss = step()                                    # replay; `loss` is a static output tensor

Everything the library allocates lazily (plans, DFT bases, envelope tables, criterion
workspaces) is created by the eager warm-up calls that precede the capture, on the
capture stream, so the captured region contains kernel launches only.  Shapes are
frozen at capture time: capture one graph per (batch, length) bucket, which is how
brever's bucket batch sampler (batching.py:219-276) feeds fixed shapes anyway.
"""
import torch

from . import _lib


class GraphedStep:
    """A captured callable.  ``inputs`` are the static tensors the graph reads
    (refill them in place), ``outputs`` whatever the callable returned."""

    def __init__(self, fn, inputs, warmup=3, stream=None):
        for t in inputs:
            _lib.require_cuda(t, 'graph input')
        device = inputs[0].device if inputs else torch.device('cuda', torch.cuda.current_device())
        self.inputs = tuple(inputs)
        self.stream = stream or torch.cuda.Stream(device)
        self.graph = torch.cuda.CUDAGraph()
        lib = _lib.lib()
        self.stream.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(self.stream):
            for _ in range(max(1, warmup)):      # lazy plans / workspaces are built here
                fn(*self.inputs)
            self.stream.synchronize()
            before = lib.brv_launch_count()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.outputs = fn(*self.inputs)
            self.launches = int(lib.brv_launch_count() - before)   # library kernels per replay
        torch.cuda.current_stream(device).wait_stream(self.stream)

    def __call__(self, *new_inputs):
        """Replay on the current stream.  Optional ``new_inputs`` are copied into
        the static input tensors first (device-to-device or pinned host-to-device)."""
        if new_inputs:
            if len(new_inputs) != len(self.inputs):
                raise ValueError(f'expected {len(self.inputs)} inputs, got {len(new_inputs)}')
            for dst, src in zip(self.inputs, new_inputs):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.outputs


def capture(fn, *inputs, warmup=3, stream=None):
    """Capture ``fn(*inputs)`` into a CUDA graph; see the module docstring."""
    return GraphedStep(fn, inputs, warmup=warmup, stream=stream)
