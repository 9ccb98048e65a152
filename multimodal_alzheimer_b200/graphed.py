"""Whole-step CUDA-graph capture: the ~500 kernel launches of one training step (normalisation, forward, loss,
backward, gradient all-reduce, optimizer) are recorded once and replayed with a single host call.  The kernels and
their arguments (including the TMA descriptors, which embed tensor addresses) are identical to the eager path; the
graph's private memory pool keeps every address stable across replays.  B200-first: CUDA streams + graphs instead
of a tracing compiler."""
import torch
import torch.distributed as dist


class GraphedStep:
    """capture(step_fn) once, then replay().  `step_fn()` must read its inputs from fixed device tensors, perform no
    host synchronisation, and return a tensor (e.g. the loss) that stays valid after each replay."""

    def __init__(self, step_fn, warmup=3):
        """Call on a NON-default stream that has also run every earlier backward pass of the model (autograd pins
        each parameter's AccumulateGrad node to the stream of its first backward; a later sync with the legacy
        default stream would invalidate the capture)."""
        self.graph = None
        self.output = None
        self.error = None
        stream = torch.cuda.current_stream()
        if stream == torch.cuda.default_stream():
            raise RuntimeError("GraphedStep must be built under `with torch.cuda.stream(s)` / torch.cuda.set_stream(s)")
        for _ in range(warmup):
            step_fn()
        torch.cuda.synchronize()
        ok = 1
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                out = step_fn()
            self.graph, self.output = g, out
        except Exception as e:  # noqa: BLE001 - capture is an optimisation; the eager path is always available
            self.error = f"{type(e).__name__}: {e}"
            ok = 0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            flag = torch.tensor([ok], device="cuda", dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag) == 0 and self.graph is not None:
                self.graph, self.error = None, "capture failed on another rank"
        torch.cuda.synchronize()

    @property
    def captured(self):
        return self.graph is not None

    def replay(self):
        self.graph.replay()
        return self.output
