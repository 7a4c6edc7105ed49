"""Flat-buffer Adam / AdamW for the train step of the hot path.

The reference steps ``torch.optim.Adam`` per mini-batch on the classification side
(graph_classification/graph_neural_networks/main.py:43, built at main.py:253) and
``torch.optim.AdamW(amsgrad=True)`` on the counting side (subgraph_isomorphism/train.py:831-838).  For the 50 k-parameter
GIN the stock capturable optimizer is ~75 tiny kernels per step (profiles/r1d); here parameters, gradients and moments
live in flat fp32 buffers (parameters / gradients become views, as in ``parallel.GradientBucket``) and one
``dn4gl_adam_f32`` launch does the whole update.  Same update rule as torch's single-tensor path
(tests/test_pipeline_gpu.py::test_flat_adam_matches_torch).  Hyper-parameters and the step counter are device
tensors, so the step can be captured in a CUDA graph and replayed (``capturable`` is always true); after changing
``param_groups[0]['lr']`` call ``sync_hyper()`` (done automatically by every eager ``step()``).
"""
import torch

from ._lib import lib, ptr
from .graph import _stream, require_cuda
from .parallel import GradientBucket


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False,
                 decoupled_weight_decay=False, bucket=None):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad,
                        decoupled_weight_decay=decoupled_weight_decay, capturable=True)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FlatAdam: one parameter group (the reference never uses more)")
        self.bucket = bucket if bucket is not None else GradientBucket(self.param_groups[0]["params"])
        self.flat_param = self.exp_avg = self.exp_avg_sq = self.max_exp_avg_sq = None
        self._hyper = self._hyper_host = self._step = self._counter = None

    # ------------------------------------------------------------------------------------------------
    def _build(self):
        """after the first backward: flat parameter buffer over the parameters that received a gradient (the same
        set, in the same order, as the gradient bucket); parameters become views into it."""
        b = self.bucket
        b._ensure()
        if b.flat is None:
            return False
        dev = b.flat.device
        require_cuda(b.flat, "gradients")
        self.flat_param = torch.zeros_like(b.flat)
        base = b.flat.data_ptr()
        for p, gview in zip(b.active, b.views):   # same (16-byte aligned) offsets as the gradient views
            off = (gview.data_ptr() - base) // 4
            view = self.flat_param[off: off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
        self.exp_avg = torch.zeros_like(b.flat)
        self.exp_avg_sq = torch.zeros_like(b.flat)
        if self.param_groups[0]["amsgrad"]:
            self.max_exp_avg_sq = torch.zeros_like(b.flat)
        self._hyper = torch.zeros(8, dtype=torch.float32, device=dev)
        self._step = torch.zeros(1, dtype=torch.float32, device=dev)
        self._counter = torch.zeros(1, dtype=torch.int32, device=dev)
        return True

    def sync_hyper(self):
        """push lr / betas / eps / weight_decay to the device (only when they changed)."""
        g = self.param_groups[0]
        h = (float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]))
        if h != self._hyper_host:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FlatAdam: hyper-parameters changed inside a CUDA-graph capture; call sync_hyper() before")
            self._hyper[:5].copy_(torch.tensor(h, dtype=torch.float32), non_blocking=False)
            self._hyper_host = h

    @property
    def num_steps(self):
        return 0 if self._step is None else int(self._step.item())

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self.flat_param is None:
            if not self._build():
                return loss
            self._apply_pending_state()      # a checkpoint loaded before the first backward
        if self._hyper_host is None or not torch.cuda.is_current_stream_capturing():
            self.sync_hyper()
        b = self.bucket
        b.gather()      # gradients that autograd left outside the flat buffer (stolen tensors) are copied in
        g = self.param_groups[0]
        lib().call("dn4gl_adam_f32", ptr(self.flat_param), ptr(b.flat), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                   ptr(self.max_exp_avg_sq), b.numel, ptr(self._hyper), ptr(self._step),
                   1 if g["decoupled_weight_decay"] else 0, ptr(self._counter), _stream())
        return loss

    def zero_grad(self, set_to_none=True):
        self.bucket.zero()

    # ---- checkpointing (train.py:1450-1470 / main.py:95-100 save and restore optimizer.state_dict()) ------------------
    def state_dict(self):
        """torch-format state: per parameter ``step`` / ``exp_avg`` / ``exp_avg_sq`` [/ ``max_exp_avg_sq``] sliced out of
        the flat buffers (parameters that never received a gradient have no entry, as with torch.optim.Adam), so a
        checkpoint resumes with the moments and the bias-correction step instead of restarting Adam."""
        if self.flat_param is not None:
            b = self.bucket
            base = self.flat_param.data_ptr()
            step = self._step.detach().clone().view(())
            for p in b.active:
                off = (p.data_ptr() - base) // 4
                sl = slice(off, off + p.numel())
                st = {"step": step.clone(), "exp_avg": self.exp_avg[sl].view_as(p).clone(),
                      "exp_avg_sq": self.exp_avg_sq[sl].view_as(p).clone()}
                if self.max_exp_avg_sq is not None:
                    st["max_exp_avg_sq"] = self.max_exp_avg_sq[sl].view_as(p).clone()
                self.state[p] = st
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """accepts what ``state_dict()`` (or a torch.optim.Adam / AdamW over the same parameters) produced.  The flat
        buffers are filled at the first step after the load (they need the gradient bucket's layout)."""
        super().load_state_dict(state_dict)
        self._pending_state = {p: dict(st) for p, st in self.state.items() if st}
        if self.flat_param is not None:
            self._apply_pending_state()

    def _apply_pending_state(self):
        pend = getattr(self, "_pending_state", None)
        if not pend:
            return
        b = self.bucket
        base = self.flat_param.data_ptr()
        steps = []
        for p in b.active:
            st = pend.get(p)
            if st is None:
                continue
            off = (p.data_ptr() - base) // 4
            sl = slice(off, off + p.numel())
            self.exp_avg[sl].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[sl].copy_(st["exp_avg_sq"].reshape(-1))
            if self.max_exp_avg_sq is not None and "max_exp_avg_sq" in st:
                self.max_exp_avg_sq[sl].copy_(st["max_exp_avg_sq"].reshape(-1))
            steps.append(float(st["step"]))
        if steps:
            self._step.fill_(max(steps))
        self._pending_state = None
