"""CUDA-graph capture of a whole training step (forward + backward [+ optimizer]).

The reference launches ~25 kernels per layer with CUDA_LAUNCH_BLOCKING=1 (models/mm_model.py:7);
here one step is ~300 launches of fused kernels and library GEMMs, which leaves the host as the
bottleneck at config-2 sizes.  Shapes are static per (batch, L), so the step is captured once and
replayed: one graph launch per step.

Dropout stays correct under replay: the seeds baked into the captured kernel arguments are combined
on the device with a counter (`mmdti_set_seed_offset`) that the graph itself increments at the start
of every replay, so each step draws fresh masks while forward and backward of one step agree."""
import torch

from . import _lib


class GraphedStep:
    """step_fn(*device_tensors) -> tensor or tuple of tensors (e.g. the loss).  step_fn must run the
    complete step on the current stream (forward, backward, optionally optimizer.step()) and must not
    synchronise with the host.

    params: every parameter whose gradient the step produces.  Their ``.grad`` is set to None after the
    warm-up and just before the capture, so that the captured backward WRITES each gradient (autograd's
    AccumulateGrad takes the incoming buffer when ``.grad`` is undefined) instead of recording
    ``grad += new`` against whatever the warm-up steps left behind -- with a defined ``.grad`` every replay
    would add to the sum of all previous steps.  The gradients then live in graph-pool buffers that are
    overwritten by each replay; do not set ``.grad`` to None (or zero it) after construction.

    example_inputs: tensors defining the static input signature; host (pinned) tensors are allowed,
    in which case every call copies them to the static device buffers inside the timed path.

    Data parallel: the step may contain in-stream NCCL collectives (dist.allreduce_grads(..., async_op=False));
    pass capture_error_mode="thread_local" so that NCCL's watchdog thread does not invalidate the capture,
    and run enough warm-up steps for the communicator to exist before the capture starts."""

    def __init__(self, step_fn, example_inputs, device=None, warmup=3, capture_error_mode="global", params=None):
        if not torch.cuda.is_available():
            raise _lib.MMDTIError("GraphedStep needs a CUDA device")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.static_in = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in example_inputs]
        for dst, src in zip(self.static_in, example_inputs):
            dst.copy_(src)
        self.staging, self.copy_stream, self._staged, self._consumed = None, None, None, None
        self.counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        _lib.call("mmdti_set_seed_offset", self.counter)
        _lib.launch_count -= 1                       # registration is not a kernel launch

        def body():
            self.counter.add_(1)
            return step_fn(*self.static_in)

        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                if params is not None:
                    for p in params:          # every warm-up step is a proper step: no accumulation onto the previous one's gradients
                        p.grad = None
                body()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        if params is not None:
            for p in params:
                p.grad = None
        self.params = list(params) if params is not None else None
        torch.cuda.empty_cache()          # the warm-up's cached blocks would sit next to the graph's private pool
        n0 = _lib.launch_count
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
            self.static_out = body()
        self.launches_per_replay = _lib.launch_count - n0      # own kernels captured in the graph

    def __call__(self, *inputs):
        """Replay the step on ``inputs`` (host or device tensors, copied into the static buffers).  With no arguments
        the batch handed to ``prefetch`` earlier is used."""
        if not inputs:
            if self._staged is None:
                raise _lib.MMDTIError("GraphedStep(): no inputs given and nothing prefetched")
            torch.cuda.current_stream(self.device).wait_event(self._staged)
            inputs, self._staged = self.staging, None
        staged = inputs is self.staging
        for dst, src in zip(self.static_in, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        if staged:
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream(self.device))
        self.graph.replay()
        _lib.launch_count += self.launches_per_replay
        return self.static_out

    def prefetch(self, *inputs):
        """Input pipeline: start copying the NEXT batch (pinned host tensors) into device staging buffers on a copy
        stream; the copy overlaps the replay that is in flight.  The next ``step()`` call without arguments waits for it
        and moves the staged batch into the static buffers with a device-to-device copy."""
        if self.staging is None:
            self.staging = [torch.empty_like(t) for t in self.static_in]
            self.copy_stream = torch.cuda.Stream(device=self.device)
        if self._consumed is not None:
            self.copy_stream.wait_event(self._consumed)      # the previous staged batch has been moved out
        with torch.cuda.stream(self.copy_stream):
            for dst, src in zip(self.staging, inputs):
                dst.copy_(src, non_blocking=True)
            self._staged = torch.cuda.Event()
            self._staged.record(self.copy_stream)

    def close(self):
        _lib.call("mmdti_set_seed_offset", None)
        _lib.launch_count -= 1
