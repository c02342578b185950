"""CUDA-graph replay of one head step (forward + backward + fused update) -- the training_step glue of SURVEY.md
section 8 f3 for the part this package owns.

The reference's call site (model/FR_PartialFC.py:175-188) issues `loss = self.loss(feat, id_, self.opt)` and
`loss.backward()` every step: ~30 kernel launches plus Python, which on a B200 costs as much host time as the step
takes on the device.  `GraphedHeadStep` captures that exact call sequence once (static input / output buffers, the
collectives included -- NCCL or the peer-memory exchanges) and replays it:

    step = GraphedHeadStep(head, optimizer, b, d)          # after a few eager warm-up steps
    loss, dx = step(feat, labels)                          # device tensors: loss [] fp32, dx [b, d] fp32 (= W * dL/dfeat)

`feat` / `labels` may be device tensors or PINNED host tensors (copied with non_blocking=True on the current stream);
`loss` and `dx` are static device buffers that the next call overwrites.  Requirements (checked): the head runs with
conf.fused_optimizer (the update is part of the captured backward) and a constant batch size.  sample_rate < 1 (BASELINE
configs[2], [3]) is captured too: the sampler, the row normalisation through the index list and the in-place update
(conf.inplace_update) are ordinary kernels; what stays on the host is the random draw -- `torch.rand(num_local)` on the CPU
generator exactly like the reference (nets/PartialFC.py:110), uploaded from a pinned ring before each replay (or
conf.device_sampling: drawn on the device) -- and it needs num_sample >= global batch, so that the number of active classes
is not data dependent (:114-115 would otherwise need a host read).  Both heads are supported:
PartialFC (SGD) and PartialFCAdamW (the bias-correction step count lives in a device scalar
that every replay advances).  Hyper-parameters (lr, momentum / betas, weight decay) are kernel arguments and therefore
part of the graph: every call compares optimizer.param_groups[-1] and the storage of weight_activated with what was
captured and re-captures when the scheduler (utils/scheduler.py:87-88) or load_state_dict() changed them.  Capturing
needs a few real warm-up steps; the weights and the optimizer state of the head are saved before and restored after
them, so constructing / recapturing does not train.
"""
import torch

from . import kernels as K
from .hostrng import cpu_rand_


class GraphedHeadStep:
    def __init__(self, head, optimizer, batch, dim, device=None, warmup=2, autograd=False):
        """autograd=False (default) captures `head.fused_step` (no autograd between forward and backward: three small
        torch kernels fewer per step); autograd=True captures `head(x, labels, opt)` + `loss.backward()`.  Same
        results (tests/test_gpu_modes.py)."""
        if not head.fused_optimizer:
            raise RuntimeError("GraphedHeadStep needs conf.fused_optimizer = True (the update is part of the graph)")
        self._sampled = head.sample_rate < 1
        if self._sampled:
            if not head._indexed:
                raise RuntimeError("GraphedHeadStep with sample_rate < 1 needs conf.inplace_update (the default)")
            if head.num_sample < batch * head.world_size:
                raise RuntimeError("GraphedHeadStep with sample_rate < 1 needs num_sample >= global batch: with fewer, the "
                                   "number of active classes depends on the labels (nets/PartialFC.py:114-115)")
        self.head, self.optimizer = head, optimizer
        dev = device if device is not None else (head.weight if self._sampled else head.weight_activated).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedHeadStep needs a CUDA head")
        self.device = dev
        self._x = torch.zeros(batch, dim, device=dev).requires_grad_(True)
        self._labels = torch.zeros(batch, dtype=torch.int64, device=dev)
        self._loss = None
        self._dx = None
        self._autograd = bool(autograd)
        self._graph = None
        self._warmup = warmup
        self._captured = None
        self._perm = None                  # sample_rate < 1: the workspace's draw buffer [num_local] (static)
        self._perm_host, self._perm_ev, self._perm_k = None, None, 0
        if self._sampled:
            self._perm = head._ensure_workspace(batch, dev).perm
            self._perm_host = [torch.empty(head.num_local, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._perm_ev = [torch.cuda.Event() for _ in range(2)]
            self._draw()
        self.recapture()

    # ------------------------------------------------------------------
    def _draw(self, perm=None):
        """Next step's sampling scores into the static device buffer: the caller's, the CUDA generator's
        (conf.device_sampling) or -- the reference's semantics -- torch.rand on the CPU generator (nets/PartialFC.py:110)."""
        if perm is not None:
            self._perm.copy_(perm.reshape(-1), non_blocking=True)
        elif self.head.device_sampling:
            self._perm.uniform_()
        else:
            k = self._perm_k
            self._perm_ev[k].synchronize()             # the upload that last used this pinned buffer has finished
            cpu_rand_(self._perm_host[k])              # = torch.rand(num_local, out=...) on the CPU generator
            self._perm.copy_(self._perm_host[k], non_blocking=True)
            self._perm_ev[k].record()
            self._perm_k = 1 - k

    def _eager(self):
        if not self._autograd:
            loss, self._dx = self.head.fused_step(self._x.detach(), self._labels, self.optimizer, perm=self._perm)
            return loss
        self._x.grad = None
        loss = self.head(self._x, self._labels, self.optimizer, perm=self._perm)
        loss.backward()
        self._dx = self._x.grad
        return loss

    def _signature(self):
        """What the captured kernels have baked in: the head's hyper-parameters and the storage they update."""
        g = self.optimizer.param_groups[-1]
        hp = tuple((k, g[k]) for k in ("lr", "momentum", "weight_decay", "betas", "eps") if k in g)
        return hp, (self.head.weight if self._sampled else self.head.weight_activated).data_ptr()

    def _optimizer_state(self):
        head = self.head
        if self._sampled:
            return [getattr(head, "weight_" + nm) for nm in head._state_names]
        st = head._fused_state
        if st is None:
            return []
        return list(st) if isinstance(st, (tuple, list)) else [st]

    def recapture(self):
        """(Re)build the graph from the head's current state and the optimizer's current hyper-parameters."""
        head = self.head
        torch.cuda.synchronize(self.device)
        w = head.weight if self._sampled else head.weight_activated.data
        saved_w = w.clone()
        had_state = self._sampled or head._fused_state is not None
        saved_state = [t.clone() for t in self._optimizer_state()]
        saved_step = head.step
        head._graph_steps = head._optimizer_kind != "sgd"
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):                      # warm-up on a side stream, as graph capture requires
            for _ in range(max(2, self._warmup)):
                self._eager()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        ws = head._ws
        g = torch.cuda.CUDAGraph()
        self._x.grad = None
        l0 = K.launch_count()
        with torch.cuda.graph(g):
            self._loss = self._eager()
        self.launches_per_replay = K.launch_count() - l0      # libpfc_b200 kernels inside the captured step
        torch.cuda.synchronize(self.device)
        self._graph = g
        # undo the warm-up steps: weights, optimizer state, step count and the normalised bf16 shard the graph's first
        # replay will read
        w.copy_(saved_w)
        if had_state:
            for t, s in zip(self._optimizer_state(), saved_state):
                t.copy_(s)
        else:
            for t in self._optimizer_state():
                t.zero_()
        head.step = saved_step
        # Adam(W): the kernels use adam_step[0] + 1; a sampled shard is bias-corrected one step ahead (PartialFCAdamW)
        ws.adam_step.fill_(saved_step + (1 if self._sampled else 0))
        if not self._sampled:
            K.l2norm_rows(w, None, head._n, ws.wn, ws.inv_w)
            head._wn_valid = True
            if ws.wn_b is not ws.wn:       # AMP mode: the captured step expects the bf16 twin of the shard as well
                K.cast_f16_to_bf16(ws.wn, ws.wn_b, head._n * w.shape[1])
                head._wn_b_valid = True
        torch.cuda.synchronize(self.device)
        self._captured = self._signature()

    def __call__(self, feat, labels, perm=None):
        """perm (sample_rate < 1, optional): this step's sampling scores [num_local] (host or device); default: drawn as
        the reference draws them."""
        if self._signature() != self._captured:
            self.recapture()               # the scheduler changed lr, or load_state_dict() rebound the weights
        head = self.head
        if self._sampled:
            self._draw(perm)
        self._x.data.copy_(feat, non_blocking=True)
        self._labels.copy_(labels.reshape(-1), non_blocking=True)
        self._graph.replay()
        # host mirror of what the replayed step did on the device
        if head._optimizer_kind != "sgd":
            head.step += 1
        return self._loss, (self._x.grad if self._autograd else self._dx)
