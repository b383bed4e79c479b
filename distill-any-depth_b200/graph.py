"""CUDA-graph replay of the hot path.  One forward (+ losses) is ~240 dependent kernel launches; capturing them once
and replaying the graph removes the per-launch host and front-end cost (bench.py: 48.6 -> 47.3 ms per ViT-L 518^2
batch-32 step).  Everything the library enqueues is capturable: ``dad_forward`` and the loss entry points allocate
nothing and never synchronise; TMA tensor maps travel as kernel parameters and are baked into the graph nodes, so
the static input / workspace / output buffers must keep their addresses - this wrapper owns them.

    step = capture(lambda x: model(x), example_batch)      # warm-up + capture
    depth, feat = step(batch)                              # copy into the static input, replay

After a weight update run the model once eagerly (that repacks the bf16 operands in place, same addresses); the
captured graph then reads the new weights.  A different (B, H, W) needs its own capture.
"""
import torch


class CapturedStep:
    def __init__(self, fn, example_inputs, warmup=2):
        self._static_in = [t.clone() if isinstance(t, torch.Tensor) else t for t in example_inputs]
        if not any(isinstance(t, torch.Tensor) and t.is_cuda for t in self._static_in):
            raise RuntimeError("capture() needs CUDA tensors: the B200 path has no CPU fallback")
        dev = next(t.device for t in self._static_in if isinstance(t, torch.Tensor) and t.is_cuda)
        stream = torch.cuda.Stream(dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream):
            for _ in range(max(1, warmup)):   # packs weights, builds position tables, sizes the workspaces
                fn(*self._static_in)
            stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=stream):
                self._static_out = fn(*self._static_in)
        torch.cuda.current_stream(dev).wait_stream(stream)

    @property
    def static_inputs(self):
        return self._static_in

    def __call__(self, *inputs):
        if len(inputs) != len(self._static_in):
            raise ValueError(f"expected {len(self._static_in)} inputs")
        for dst, src in zip(self._static_in, inputs):
            if isinstance(dst, torch.Tensor):
                if src.shape != dst.shape or src.dtype != dst.dtype:
                    raise ValueError(f"captured for {tuple(dst.shape)} {dst.dtype}, got {tuple(src.shape)} {src.dtype}")
                if src.data_ptr() != dst.data_ptr():
                    dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self._static_out


def capture(fn, *example_inputs, warmup=2):
    """Capture ``fn(*example_inputs)`` (any composition of this package's forwards and losses) in a CUDA graph.
    The returned callable copies its arguments into the captured input buffers, replays, and returns the captured
    output tensors (overwritten by the next call)."""
    return CapturedStep(fn, example_inputs, warmup)
