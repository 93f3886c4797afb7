"""Whole-step CUDA graph for the steady-state VQ training step (forward + loss + backward).

One VQ-IMG step is ~1200 kernel launches, several hundred of them 10-70 us long (the 16x16 and 32x32 levels), so the
Python/ctypes launch path starves the GPU there (~4 ms of a ~110 ms step at batch 32).  Every C-ABI entry enqueues on the
caller's stream without synchronising or allocating (include/mas_b200.h), which makes the whole step capturable; replaying
the captured graph removes the launch path from the critical path.  Host-side state that the reference keeps in Python
(Codebook.q_counter and the reservoir, modules.py:474-486) is advanced around the replay, not inside it.

Single process / single GPU only: with DistributedDataParallel the gradient all-reduce hooks are host-driven.
"""
import torch

from . import _lib as L


def _unwrap(model):
    return model.module if hasattr(model, "module") else model


class GraphedStep:
    """gs = GraphedStep(model, loss_fn, example_input); loss = gs(batch)

    loss_fn(model, x) -> 0-dim loss.  After a call, parameter .grad tensors hold this step's gradients (static buffers,
    overwritten by the next call) and the returned loss is a static 0-dim tensor.  Construction runs `warmup` real eager
    steps on example_input first (they advance BatchNorm running statistics and the codebook counters like any step).

    The optimizer runs OUTSIDE the graph, between calls.  `optimizer.zero_grad()` (set_to_none=True by default) detaches
    the static gradient buffers from the parameters; every call re-attaches them after the replay, so
    `gs(x); opt.step(); opt.zero_grad()` loops train exactly like eager steps.  The library's scratch buffer whose address
    is baked into the graph is held by this object until close(): a later, larger eager call allocates a new one instead
    of freeing the captured one under the graph.
    """

    def __init__(self, model, loss_fn, example_input, warmup=2):
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            raise RuntimeError("GraphedStep: single-GPU only (DDP's all-reduce hooks are host-driven)")
        self.model, self.loss_fn = model, loss_fn
        base = _unwrap(model)
        self.vq = getattr(base, "quantize", None)
        if self.vq is not None and base.training and self.vq.q_counter + 1 < self.vq.q_re_end:
            raise RuntimeError("GraphedStep: the codebook is still in its warm-up / re-initialisation schedule "
                               "(q_counter < q_re_end): that control flow is host-side, run those steps eagerly")
        self.static_in = example_input.detach().clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):   # sizes the workspace and the allocator, initialises lazily-built state
                model.zero_grad(set_to_none=True)
                loss_fn(model, self.static_in).backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        model.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
        if self.vq is not None:
            self.vq.defer_collect = True     # reservoir sampling changes shape over time: runs eagerly after each replay
        self.graph = torch.cuda.CUDAGraph()
        l0 = L.launch_count()
        try:
            with torch.cuda.graph(self.graph):
                self.loss = loss_fn(model, self.static_in)
                self.loss.backward()
        except Exception:
            if self.vq is not None:
                self.vq.defer_collect = False
            raise
        self.launches_per_step = L.launch_count() - l0
        self._grads = [(p, p.grad) for p in model.parameters() if p.grad is not None]   # static buffers of the capture
        self._ws = L.pin_workspaces()    # scratch addresses recorded in the graph stay allocated while it lives
        self._z = getattr(self.vq, "_deferred_z", None) if self.vq is not None else None
        if self.vq is not None and base.training:
            self.vq.q_counter -= 1           # the captured step was recorded, not executed

    def __call__(self, x):
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        for p, g in self._grads:         # zero_grad(set_to_none=True) / a manual `p.grad = None` between calls
            if p.grad is not g:
                p.grad = g
        if self.vq is not None and _unwrap(self.model).training:
            self.vq.q_counter += 1
            if self._z is not None:
                self.vq._collect(self._z)
        return self.loss

    def close(self):
        """Back to eager stepping; releases the graph and its memory pool (the static loss / gradient buffers with it)."""
        if self.vq is not None:
            self.vq.defer_collect = False
            self.vq._deferred_z = None
        self._z = None
        self.loss = None
        self._grads = []
        self._ws = None
        if self.graph is not None:
            self.model.zero_grad(set_to_none=True)
            self.graph.reset()
            self.graph = None
