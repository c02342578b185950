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
conf.fused_optimizer (the update is part of the captured backward), sample_rate == 1 (sampling draws host random
numbers and patches the optimizer every step) and a constant batch size -- the configuration of BASELINE configs[1].
Hyper-parameters (lr, momentum, weight decay) are baked in at capture time: call `recapture()` after the scheduler
changes them.  Capturing needs a few real warm-up steps; the weights and the optimizer state of the head are saved
before and restored after them, so constructing / recapturing does not train.
"""
import torch

from . import kernels as K


class GraphedHeadStep:
    def __init__(self, head, optimizer, batch, dim, device=None, warmup=2, autograd=True):
        """autograd=False captures `head.fused_step` (no autograd between forward and backward: three small torch
        kernels fewer per step) instead of `head(x, labels, opt)` + `loss.backward()`; same results."""
        if not head.fused_optimizer:
            raise RuntimeError("GraphedHeadStep needs conf.fused_optimizer = True (the update is part of the graph)")
        if head._optimizer_kind != "sgd":
            raise RuntimeError("GraphedHeadStep supports the SGD head (Adam's bias-correction step count is host state)")
        if head.sample_rate < 1:
            raise RuntimeError("GraphedHeadStep needs sample_rate == 1 (sampling patches the optimizer on the host)")
        self.head, self.optimizer = head, optimizer
        dev = device if device is not None else head.weight_activated.device
        if dev.type != "cuda":
            raise RuntimeError("GraphedHeadStep needs a CUDA head")
        self.device = dev
        # two slots: with conf.overlap_update the normalised shard ping-pongs between two buffers, so consecutive
        # steps are two different graphs that are replayed alternately; otherwise one graph serves every step
        self._x = [torch.zeros(batch, dim, device=dev).requires_grad_(True) for _ in range(2)]
        self._labels = [torch.zeros(batch, dtype=torch.int64, device=dev) for _ in range(2)]
        self._loss = [None, None]
        self._dx = [None, None]
        self._autograd = bool(autograd)
        self._graphs = None
        self._warmup = warmup
        self.recapture()

    # ------------------------------------------------------------------
    def _eager(self, k):
        if not self._autograd:
            loss, self._dx[k] = self.head.fused_step(self._x[k].detach(), self._labels[k], self.optimizer)
            return loss
        self._x[k].grad = None
        loss = self.head(self._x[k], self._labels[k], self.optimizer)
        loss.backward()
        self._dx[k] = self._x[k].grad
        return loss

    def _parity(self):
        return 0 if self.head._ws is None or self.head._ws.wn.data_ptr() == self._wn_ptr0 else 1

    def recapture(self):
        """(Re)build the graph(s) from the head's current state and the optimizer's current hyper-parameters."""
        head = self.head
        torch.cuda.synchronize(self.device)
        w = head.weight_activated.data
        saved_w = w.clone()
        saved_state = None if head._fused_state is None else head._fused_state.clone()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):                      # warm-up on a side stream, as graph capture requires
            for i in range(max(2, self._warmup)):
                self._eager(i % 2)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        ws = head._ws
        self._wn_ptr0 = ws.wn.data_ptr()
        graphs = {}
        for _ in range(2):
            k = self._parity()
            g = torch.cuda.CUDAGraph()
            self._x[k].grad = None
            with torch.cuda.graph(g):
                self._loss[k] = self._eager(k)              # capturing flips the ping-pong on the host side only
            graphs[k] = g
            if ws.wn_alt is None:
                graphs[1 - k] = g
                self._loss[1 - k] = self._loss[k]
                self._dx[1 - k] = self._dx[k]
                self._x[1 - k], self._labels[1 - k] = self._x[k], self._labels[k]
                break
        torch.cuda.synchronize(self.device)
        self._graphs = graphs
        # undo the warm-up steps: weights, momentum and the normalised bf16 shard the graph's first replay will read
        w.copy_(saved_w)
        if saved_state is not None:
            head._fused_state.copy_(saved_state)
        elif head._fused_state is not None:
            head._fused_state.zero_()
        K.l2norm_rows(w, None, head._n, ws.wn, ws.inv_w)
        head._wn_valid = True
        torch.cuda.synchronize(self.device)

    def __call__(self, feat, labels):
        k = self._parity()
        self._x[k].data.copy_(feat, non_blocking=True)
        self._labels[k].copy_(labels.reshape(-1), non_blocking=True)
        self._graphs[k].replay()
        ws = self.head._ws
        if ws.wn_alt is not None:
            ws.wn, ws.wn_alt = ws.wn_alt, ws.wn             # what the replayed step did on the device
        return self._loss[k], (self._x[k].grad if self._autograd else self._dx[k])
