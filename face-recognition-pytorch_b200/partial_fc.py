"""PartialFC / PartialFCAdamW with the reference's nn.Module interface (nets/PartialFC.py), backed by libpfc_b200.

Same constructor, same `forward(local_embeddings, local_labels, optimizer) -> loss`, same attributes
(rank, world_size, num_local, class_start, num_sample, weight, weight_mom, weight_activated, weight_activated_mom,
weight_index ...), same `{"weight": [num_local, d]}` per-rank state_dict, same optimizer patching when
sample_rate < 1 -- so `importlib.import_module(f"nets.{conf.loss}").PartialFC(conf=..., num_classes=...)`
(model/FR_PartialFC.py:102-109) keeps working when `nets` resolves to this package's shim.

What changes underneath (one rank = one GPU, all work enqueued on the current CUDA stream, no host syncs except
the data-dependent sampled size when num_sample < global batch):
  forward : l2norm(x)->bf16 | all-gather(bf16 x, labels) | localise labels | [sample + gather rows] | l2norm(W)->bf16
            | tcgen05 GEMM with margin/exp/row-sum epilogue spilling bf16 E' | row stats | ONE all-reduce [B,2] | loss
  backward: coefficients + target patch | tcgen05 dX GEMM (class-split) | scale [+ reduce-scatter] + normalise-bwd
            | tcgen05 dW GEMM | normalise-bwd (+ fused SGD / AdamW step and next step's bf16 shard)
The collectives are the naturally sharded ones issued through torch.distributed.

Optional keys read from `conf` beyond the reference's five (emd_size, sample_rate, mixed_precision, loss_s, loss_m):
  conf.fused_optimizer (bool, default False): run the SGD / AdamW update of the head inside the backward and leave
      weight_activated.grad = None so optimizer.step() skips it.  Hyper-parameters are re-read from
      optimizer.param_groups[-1] every step (the scheduler mutates lr, utils/scheduler.py:87-88).  A scaled loss
      (GradScaler flow, model/FR_PartialFC.py:178-184) is honoured: dX keeps the scale autograd expects and the fused
      update divides it out on the device; the scaler's inf-skipping cannot be (bf16 operands do not need a scaler).
  conf.inplace_update (bool, default True; only with fused_optimizer and sample_rate < 1): the fused step reads and
      updates the sampled rows IN PLACE in `weight` / `weight_mom` (or Adam's moments) through the index list, so the
      row gather before the step (nets/PartialFC.py:120-121) and the scatter after it (:142-143) -- 4 x n x d x 4 bytes
      of HBM traffic each -- disappear; `weight_activated` then stays the reference's initial (0, 0) placeholder and
      `weight`, `weight_mom`, `weight_index`, state_dict() are always current.  False: gather / scatter like the
      reference (what the un-fused path always does, because torch.optim steps on `weight_activated`).
  conf.device_sampling (bool, default False): draw the PartialFC sampling scores with the CUDA generator on the device
      instead of `torch.rand` on the CPU generator + H2D copy (nets/PartialFC.py:110).  Removes a host round trip per
      step; the sampled index set is then NOT the reference's for the same seed (same distribution, other stream).
  conf.dx_side_stream (True / False / "auto", default "auto" = on): after the dX
      partials exist the step forks -- the peer scatter + barrier + normalise-backward of dX (one GPU: just the
      normalise-backward) on a side stream, the rank-local dW GEMM / update on the main one -- and joins before
      backward returns.  The branches share no buffer; in a CUDA graph they become parallel branches.
  conf.dx_fork_gemm (True / False / "auto", default "auto" = off): move that fork in front of the dX GEMM, so the two
      gradient GEMMs run side by side (the clusters that run out of tiles in one persistent kernel's last wave pick up the
      other's); the fused update waits for the dX GEMM's last read of the shard.  Measured -2 us per step at the shard
      sizes of 2 / 4 / 8 GPUs, nothing on one (profiles/r02c_exp_shard_shapes.txt).
  conf.fuse_prepare (bool, default True; only `fused_step`, i.e. the no-autograd step, where d loss = 1 is known when the
      forward ends): the kernel that forms the loss also forms the backward coefficients (c_i, the scaled bf16 rows and
      the patched target column), so the step has one launch fewer; False: the two launches of the autograd path.
  conf.peer_collectives (True / False / "auto"): exchange the batch, the softmax statistics and dX through peer
      (NVLink) memory with the stores fused into the producing kernels instead of three NCCL collectives.
  conf.peer_timeout_ms (float, default 600 000 = NCCL's watchdog default; 0 = wait forever): how long a flag barrier of
      the peer exchange waits for a stalled rank before this rank's kernel traps (a CUDA error instead of a hang).
"""
import collections
import contextlib
from typing import Callable

import torch
from torch import distributed

from . import kernels as K
from .hostrng import cpu_rand
from .arcface import ArcFace


def shard_range(num_classes: int, rank: int, world_size: int):
    """nets/PartialFC.py:57-62."""
    num_local = num_classes // world_size + int(rank < num_classes % world_size)
    class_start = num_classes // world_size * rank + min(rank, num_classes % world_size)
    return num_local, class_start


class _Workspace:
    """Persistent device buffers for one (global batch, max active classes, d) shape -- allocated once so that the
    whole step is CUDA-graph capturable and nothing is allocated in the hot loop."""

    def __init__(self, dev, b, W, n_max, nl, d, sampled, op_dtype=torch.bfloat16):
        B = b * W
        f32, bf16, i32, i64 = torch.float32, torch.bfloat16, torch.int32, torch.int64
        self.op_dtype = op_dtype                  # Xn / Wn: bf16, or fp16 in the reference's AMP mode
        z = lambda *s, dt=f32: torch.zeros(*s, dtype=dt, device=dev)   # noqa: E731
        self.B, self.b, self.n_max = B, b, n_max
        self.n_pad_max = K.padded_classes(n_max)
        self.B_pad = K.padded_batch(B)
        self.xn_local = z(b, d, dt=op_dtype)
        self.inv_x = z(b)
        self.xn_all = z(B, d, dt=op_dtype) if W > 1 else self.xn_local
        self.labels_all = z(B, dt=i64)
        self.labels_local = z(B, dt=i32)
        self.labels_act = z(B, dt=i32) if sampled else self.labels_local
        self.wn = z(n_max, d, dt=op_dtype)
        # AMP mode: the dX contraction multiplies the bf16 spill with a bf16 copy of the fp16 shard
        self.wn_b = z(n_max, d, dt=bf16) if op_dtype != bf16 else self.wn
        self.inv_w = z(n_max)
        self.E = z(B * self.n_pad_max, dt=bf16)
        self.part_sum = z(K.num_class_tiles(n_max) * self.B_pad)
        self.tgt_raw, self.tgt_e, self.tgt_z = z(B), z(B), z(B)
        self.stats = z(B, 2)
        self.row_L = z(B)
        self.loss = z(1)
        self.ticket = z(1, dt=i32)
        self.coef = z(B)
        self.xs = z(B, d, dt=bf16)
        self.max_splits = max(1, K.dx_max_splits(B, d))
        self.dx_partial = z(self.max_splits * B * d)
        self.dxn_all = z(B, d) if W > 1 else None
        self.dxn_local = z(b, d) if W > 1 else None
        self.dwn = None                           # fp32 un-normalised dW (un-fused / AdamW), allocated on first use
        self.dwn_bf16 = None                      # bf16 spill of it (fused SGD), allocated on first use
        self.gscale = torch.ones(1, dtype=f32, device=dev)     # loss scale the gradient carries (fused update)
        self.adam_step = z(1, dt=i32)
        if sampled:
            self.perm = z(nl)
            self.index = z(n_max, dt=i64)
            self.n_out = z(1, dt=i32)
            self.sample_ws = z(K.sample_workspace_bytes(nl), dt=torch.uint8)

    def grad_buffer(self, bf16, d):
        if bf16:
            if self.dwn_bf16 is None:
                self.dwn_bf16 = torch.zeros(self.n_max, d, dtype=torch.bfloat16, device=self.wn.device)
            return self.dwn_bf16
        if self.dwn is None:
            self.dwn = torch.zeros(self.n_max, d, dtype=torch.float32, device=self.wn.device)
        return self.dwn


class _HeadFunction(torch.autograd.Function):
    """loss = head(local_embeddings, weight_activated); backward delivers W * dL/d local_embeddings
    (nets/PartialFC.py:521) and dL/d weight_activated (or None when the optimizer step is fused)."""

    @staticmethod
    def forward(ctx, local_embeddings, weight_activated, head):
        ctx.head = head
        ctx.x = local_embeddings
        ctx.step_id = head._step_id
        return head._forward_impl(local_embeddings, need_dx=ctx.needs_input_grad[0])

    @staticmethod
    def backward(ctx, grad_loss):
        head = ctx.head
        if ctx.step_id != head._step_id:
            raise RuntimeError("PartialFC backward called after a newer forward: the head keeps one step of state")
        dx, dw = head._backward_impl(ctx.x, grad_loss)
        return dx, dw, None


class _PartialFCBase(torch.nn.Module):
    _version = 1
    _optimizer_kind = "sgd"

    def __init__(self, conf, num_classes, margin_loss: Callable = ArcFace):
        super().__init__()
        assert distributed.is_initialized(), "must initialize distributed before create this"
        self.rank = distributed.get_rank()
        self.world_size = distributed.get_world_size()

        self.embedding_size = conf.emd_size
        self.sample_rate: float = conf.sample_rate
        # nets/PartialFC.py:198 runs the logits GEMM under autocast(fp16) when conf.mixed_precision: the normalised
        # operands Xn / Wn are then fp16 here as well (fp32 accumulation either way); False: bf16 operands
        self.fp16 = conf.mixed_precision
        self._op_dtype = torch.float16 if self.fp16 else torch.bfloat16
        self.fused_optimizer = bool(getattr(conf, "fused_optimizer", False))
        self.device_sampling = bool(getattr(conf, "device_sampling", False))
        # run the tail of the dX path (finalize / peer scatter + finalize) on a side stream next to the rank-local
        # dW GEMM + update, which it does not depend on; may be flipped between steps (before a graph capture)
        self.dx_side_stream = getattr(conf, "dx_side_stream", "auto")    # True / False / "auto"
        # fused update of a sampled shard in place through the index list (no gather / scatter of the active rows)
        self._indexed = (self.fused_optimizer and self.sample_rate < 1 and bool(getattr(conf, "inplace_update", True)))
        self.dw_first = getattr(conf, "dw_first", "auto")     # order of the two gradient GEMMs (see _backward_impl)
        # True / False / "auto": fork the side stream in FRONT of the dX GEMM (both gradient GEMMs side by side)
        self.dx_fork_gemm = getattr(conf, "dx_fork_gemm", "auto")
        # fused_step: form the backward coefficients in the kernel that forms the loss (one launch fewer)
        self.fuse_prepare = bool(getattr(conf, "fuse_prepare", True))
        self._prepared = False
        self._num_classes = int(num_classes)
        self.num_local, self.class_start = shard_range(num_classes, self.rank, self.world_size)
        self.num_sample: int = int(self.sample_rate * self.num_local)
        self.last_batch_size: int = 0
        self.is_updated: bool = True
        self.init_weight_update: bool = True
        self._state_names = self._optimizer_state_names()

        d = self.embedding_size
        if self.sample_rate < 1:
            self.register_buffer("weight", tensor=torch.normal(0, 0.01, (self.num_local, d)))
            for nm in self._state_names:
                self.register_buffer("weight_" + nm, tensor=torch.zeros(self.num_local, d))
                self.register_buffer("weight_activated_" + nm, tensor=torch.empty(0, 0))
            self.register_parameter("weight_activated", param=torch.nn.Parameter(torch.empty(0, 0)))
            self.register_buffer("weight_index", tensor=torch.empty(0, 0))
        else:
            self.weight_activated = torch.nn.Parameter(torch.normal(0, 0.01, (self.num_local, d)))

        if isinstance(margin_loss, Callable):
            self.margin_softmax = margin_loss(conf.loss_s, conf.loss_m)
        else:
            raise RuntimeError("margin_loss must be callable")
        if not hasattr(self.margin_softmax, "margin_spec"):
            raise TypeError("the fused head needs a margin from this package (ArcFace, CosFace, CombinedMarginLoss): "
                            "the margin is applied inside the GEMM epilogue, not by calling the module")
        self.step = 0
        self._ws = None
        self._step_id = 0
        self._wn_valid = False          # wn / inv_w in the workspace match weight_activated
        self._wn_b_valid = False        # AMP mode: the bf16 twin of the fp16 shard (ws.wn_b, read by the dX GEMM) matches wn
        self._act_store = None          # persistent storage behind weight_activated / its optimizer state (r < 1)
        self._fused_state = None        # optimizer state for the fused step when sample_rate == 1
        self._n = self.num_local        # active classes this step
        self._opt_args = None
        self._side_stream = None        # dX tail
        self._gscale_is_one = True
        self._graph_steps = False       # AdamW: take the bias-correction step count from ws.adam_step (graph replay)
        # True / False / "auto": exchange the batch, the softmax statistics and dX through peer (NVLink) memory with
        # the stores fused into the producing kernels (csrc/pfc_peer.cu) instead of three NCCL collectives
        self.peer_collectives = getattr(conf, "peer_collectives", "auto")
        self.peer_timeout_ms = getattr(conf, "peer_timeout_ms", None)
        self._peer = None

    # ------------------------------------------------------------------ reference-visible helpers
    def _optimizer_state_names(self):
        return ["mom"]

    def _patch_optimizer(self, optimizer):
        raise NotImplementedError

    @torch.no_grad()
    def sample(self, labels_local: torch.Tensor, index_positive, optimizer: torch.optim.Optimizer,
               perm: torch.Tensor = None):
        """nets/PartialFC.py:92-131.  labels_local: int32 shard-local ids with -1 for foreign rows (index_positive
        is implied by it and only kept for signature parity).  `perm` defaults to torch.rand on the CPU generator exactly like the reference
        (:110) -- one H2D copy per step; pass a device tensor to replay a recorded draw."""
        ws = self._ws
        if perm is ws.perm:
            pass                                     # the caller filled the workspace's draw buffer (GraphedHeadStep)
        elif perm is None and self.device_sampling and ws.perm.is_cuda:
            ws.perm.uniform_()                       # [0, 1) from the CUDA generator, no host round trip
        else:
            if perm is None:
                perm = cpu_rand(self.num_local)      # = torch.rand(size=[num_local]) on the CPU generator, drawn in bulk
            ws.perm.copy_(perm, non_blocking=True)
        K.sample(ws.perm, labels_local, self.num_local, self.num_sample, ws.index, ws.n_out, ws.labels_act,
                 ws.sample_ws)
        if self.num_sample >= ws.B:
            n = self.num_sample                      # positives (<= B distinct) always fit: no host sync needed
        else:
            n = int(ws.n_out.item())                 # data-dependent: more positives than num_sample (:114-115)
        self._n = n
        self.weight_index = ws.index[:n]
        self._wn_valid = False
        if self._indexed:
            return                 # rows are normalised (forward) and updated (backward) in place through weight_index
        names = self._state_names
        srcs = [self.weight] + [getattr(self, "weight_" + nm) for nm in names]
        dsts = [self._act_store[0][:n]] + [self._act_store[1 + i][:n] for i in range(len(names))]
        K.gather_rows(srcs, dsts, self.weight_index, n)                          # :120-121
        self.weight_activated = torch.nn.Parameter(dsts[0])
        for i, nm in enumerate(names):
            setattr(self, "weight_activated_" + nm, dsts[1 + i])
        self._wn_valid = False
        self._patch_optimizer(optimizer)                                         # :123-131

    @torch.no_grad()
    def update(self):
        """partial weight to global, nets/PartialFC.py:133-143."""
        if self.init_weight_update:
            self.init_weight_update = False
            return
        if self.sample_rate < 1 and not self._indexed:
            names = self._state_names
            n = self.weight_activated.shape[0]
            if n == 0:
                return
            srcs = [self.weight_activated.data] + [getattr(self, "weight_activated_" + nm) for nm in names]
            dsts = [self.weight] + [getattr(self, "weight_" + nm) for nm in names]
            K.scatter_rows(srcs, dsts, self.weight_index, n)

    # ------------------------------------------------------------------ forward / backward
    def _ensure_workspace(self, b, dev):
        d = self.embedding_size
        sampled = self.sample_rate < 1
        B = b * self.world_size
        n_max = self.num_local if not sampled else max(self.num_sample, min(B, self.num_local))
        if self._ws is None or self._ws.b != b or self._ws.xn_local.device != dev:
            self._ws = _Workspace(dev, b, self.world_size, n_max, self.num_local, d, sampled, self._op_dtype)
            self._wn_valid = False
            if sampled and not self._indexed:
                k = 1 + len(self._state_names)
                self._act_store = [torch.zeros(n_max, d, device=dev) for _ in range(k)]
            self._peer = None
            if self.world_size > 1 and dev.type == "cuda" and self.peer_collectives in (True, "auto"):
                try:
                    from .peer import PeerExchange
                    self._peer = PeerExchange(dev, self.rank, self.world_size, b, d, timeout_ms=self.peer_timeout_ms,
                                              operand_dtype=self._op_dtype)
                    self._ws.xn_all = self._peer.xn_all          # peers store straight into these
                    self._ws.labels_all = self._peer.labels_all
                except Exception as e:                            # no P2P / symmetric memory: keep the NCCL collectives
                    if self.peer_collectives is True:
                        raise
                    import warnings
                    warnings.warn(f"peer-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL collectives")
                    self._peer = None
        return self._ws

    def forward(self, local_embeddings: torch.Tensor, local_labels: torch.Tensor, optimizer: torch.optim.Optimizer,
                perm: torch.Tensor = None):
        local_labels.squeeze_()                                       # in place on the caller's tensor, :164
        local_labels = local_labels.long()
        self.update()

        batch_size = local_embeddings.size(0)
        if self.last_batch_size == 0:
            self.last_batch_size = batch_size
        assert self.last_batch_size == batch_size, (
            "last batch size do not equal current batch size: {} vs {}".format(self.last_batch_size, batch_size))
        if local_embeddings.dtype != torch.float32:
            local_embeddings = local_embeddings.float()
        ws = self._ensure_workspace(batch_size, local_embeddings.device)
        self._optimizer = optimizer
        self._step_id += 1
        self._prepare(local_embeddings, local_labels.contiguous(), optimizer, perm)
        # sampling may just have replaced weight_activated (:120): hand autograd the CURRENT parameter
        return _HeadFunction.apply(local_embeddings, self.weight_activated, self)

    def _read_optimizer(self, optimizer):
        raise NotImplementedError

    @torch.no_grad()
    def fused_step(self, local_embeddings: torch.Tensor, local_labels: torch.Tensor,
                   optimizer: torch.optim.Optimizer, perm: torch.Tensor = None):
        """forward + backward of one step WITHOUT autograd: the same kernel sequence as `forward(...)` followed by
        `loss.backward()` with d loss = 1 (nets/PartialFC.py:146-208 + :464-522), minus the three small torch kernels
        autograd puts between them (clone of the loss, ones_like for its gradient, gradient accumulation).  Returns
        (loss, dx): `loss` is a 0-dim VIEW of the head's static loss buffer (overwritten by the next step), `dx`
        [b, d] = world_size * dL/d local_embeddings.  With conf.fused_optimizer the update has been applied; otherwise
        dL/d weight_activated is left in `weight_activated.grad` for optimizer.step().  For callers that own the
        training loop (GraphedHeadStep(autograd=False)); GradScaler users keep the autograd path."""
        local_labels.squeeze_()
        local_labels = local_labels.long()
        self.update()
        batch_size = local_embeddings.size(0)
        if self.last_batch_size == 0:
            self.last_batch_size = batch_size
        assert self.last_batch_size == batch_size, (
            "last batch size do not equal current batch size: {} vs {}".format(self.last_batch_size, batch_size))
        if local_embeddings.dtype != torch.float32:
            local_embeddings = local_embeddings.float()
        self._ensure_workspace(batch_size, local_embeddings.device)
        self._optimizer = optimizer
        self._step_id += 1
        self._prepare(local_embeddings, local_labels.contiguous(), optimizer, perm)
        # d loss = 1 is known here, so the backward coefficients are formed by the kernel that forms the loss
        loss = self._forward_impl(local_embeddings, clone_loss=False, need_dx=True, fuse_prepare=self.fuse_prepare)
        dx, dw = self._backward_impl(local_embeddings, None, need_dx=True)
        if dw is not None:
            self.weight_activated.grad = dw
        return loss, dx

    @torch.no_grad()
    def _prepare(self, local_embeddings, labels_in, optimizer, perm):
        """Everything that precedes the differentiable part: normalise + gather the batch, localise the labels,
        sample the active classes (nets/PartialFC.py:175-196)."""
        ws, W = self._ws, self.world_size
        x = local_embeddings.detach().contiguous()
        self._x_local = x
        peer = self._peer
        if peer is not None:
            # normalise + all-gather in one kernel: every rank stores its bf16 rows and labels into every peer
            # (the barrier that publishes the rows is taken by the consumer of the labels, below)
            K.peer_l2norm_gather(x, labels_in, self.rank, W, peer.ptrs("xn_all"), peer.ptrs("labels_all"), ws.inv_x,
                                 fp16=self._op_dtype == torch.float16)
            labels_all = peer.labels_all
        elif W > 1:
            K.l2norm_rows(x, None, ws.b, ws.xn_local, ws.inv_x)
            if distributed.get_backend() == "nccl":
                # one NCCL launch for both gathers (ncclGroupStart/End); every collective of the step is latency-bound
                with distributed._coalescing_manager(device=x.device):
                    distributed.all_gather_into_tensor(ws.xn_all, ws.xn_local)    # :182 (bf16: half the bytes)
                    distributed.all_gather_into_tensor(ws.labels_all, labels_in)  # :183
            else:
                distributed.all_gather_into_tensor(ws.xn_all, ws.xn_local)
                distributed.all_gather_into_tensor(ws.labels_all, labels_in)
            labels_all = ws.labels_all
        else:
            # one rank: normalise + label localisation (:188-193) in one launch
            K.l2norm_rows_localize(x, ws.b, ws.xn_local, ws.inv_x, labels_in, self.class_start, self.num_local,
                                   ws.labels_local)
            labels_all = None
        if labels_all is None:
            pass
        elif peer is not None:
            K.peer_localize_labels(peer.ptrs("flags"), peer.counter, self.rank, W, labels_all, self.class_start,
                                   self.num_local, ws.labels_local)               # barrier + :188-193
        else:
            K.localize_labels(labels_all, self.class_start, self.num_local, ws.labels_local)   # :188-193
        if self.sample_rate < 1:
            self.sample(ws.labels_local, None, optimizer, perm)                   # :195-196
        else:
            self._n = self.num_local

    def _forward_impl(self, local_embeddings, clone_loss=True, need_dx=False, fuse_prepare=False):
        ws, W, d = self._ws, self.world_size, self.embedding_size
        self._prepared = False
        b, B = ws.b, ws.B
        n = self._n
        w = self.weight if self._indexed else self.weight_activated.data
        if self.fused_optimizer:
            self._opt_args = self._read_optimizer(self._optimizer)
        if not self._wn_valid:
            K.l2norm_rows(w, ws.index if self._indexed else None, n, ws.wn, ws.inv_w)   # :200 (+ :120 when indexed)
            self._wn_valid = True
            self._wn_b_valid = False
        kind, s, m2, m3, thr = self.margin_softmax.margin_spec()
        self._n_pad = K.padded_classes(n)
        K.forward(ws.xn_all, ws.wn, ws.labels_act, B, n, d, s, kind, m2, m3, thr, ws.E, self._n_pad, ws.part_sum,
                  ws.tgt_raw, ws.tgt_e, ws.tgt_z)                                 # :201-207
        peer = self._peer
        if peer is not None:
            # statistics straight into every peer's slot, then a rank-ordered local sum (identical bits on all ranks)
            K.peer_row_stats(ws.part_sum, K.num_class_tiles(n), B, ws.labels_act, ws.tgt_e, self.rank, W,
                             peer.ptrs("slots"))
            if fuse_prepare:                                                      # + :464-484 (d loss = 1)
                K.peer_loss_prepare(peer.ptrs("flags"), peer.counter, self.rank, peer.slots, W, B, ws.stats, ws.row_L,
                                    ws.loss, ws.ticket, None, s, d, ws.labels_act, ws.tgt_raw, kind, m2, ws.xn_all,
                                    ws.xs, ws.coef, ws.E, self._n_pad)
                self._prepared = True
            else:
                K.peer_loss(peer.ptrs("flags"), peer.counter, self.rank, peer.slots, W, B, ws.stats, ws.row_L,
                            ws.loss)                                              # barrier + :448, :453, :459, :461
        elif W == 1 and fuse_prepare:
            K.row_stats_loss_prepare(ws.part_sum, K.num_class_tiles(n), B, ws.labels_act, ws.tgt_e, ws.stats, ws.row_L,
                                     ws.loss, ws.ticket, None, s, d, ws.tgt_raw, kind, m2, ws.xn_all, ws.xs, ws.coef,
                                     ws.E, self._n_pad)                           # :446-461 + :464-484 in one launch
            self._prepared = True
        elif W == 1:
            K.row_stats_loss(ws.part_sum, K.num_class_tiles(n), B, ws.labels_act, ws.tgt_e, ws.stats, ws.row_L,
                             ws.loss, ws.ticket)                                  # :446-461 in one launch
        else:
            K.row_stats(ws.part_sum, K.num_class_tiles(n), B, ws.labels_act, ws.tgt_e, ws.stats)
            distributed.all_reduce(ws.stats, distributed.ReduceOp.SUM)            # replaces :448, :453, :459
            K.loss(ws.stats, B, ws.row_L, ws.loss)                                # :461
        return ws.loss[0].clone() if clone_loss else ws.loss[0]

    def _backward_impl(self, x_in, grad_loss, need_dx=None):
        """grad_loss None: d loss = 1 (no GradScaler); need_dx None: x_in.requires_grad."""
        ws, W, d = self._ws, self.world_size, self.embedding_size
        b, B, n, n_pad = ws.b, ws.B, self._n, self._n_pad
        kind, s, m2, m3, thr = self.margin_softmax.margin_spec()
        need_dx = x_in.requires_grad if need_dx is None else bool(need_dx)
        g = None if grad_loss is None else grad_loss.detach().to(torch.float32).reshape(1).contiguous()
        w = self.weight if self._indexed else self.weight_activated.data
        peer = self._peer
        if self.fused_optimizer:
            if g is not None:
                ws.gscale.copy_(g)                 # the fused update divides the loss scale out again
            elif not self._gscale_is_one:
                ws.gscale.fill_(1.0)
            self._gscale_is_one = g is None
        if not (self._prepared and g is None):
            K.backward_prepare(ws.stats, ws.row_L, g, s, B, d, ws.labels_act, ws.tgt_raw, kind, m2, ws.xn_all, ws.xs,
                               ws.coef, ws.E, n_pad)
        self._prepared = False
        spill_bf16 = self.fused_optimizer and self._optimizer_kind == "sgd" and d % 128 == 0
        dwn = ws.grad_buffer(spill_bf16, d)
        dw = None
        # Order.  N > 1: dX GEMM and its exchange first, THEN the rank-local dW GEMM and the update, so the
        # reduce-scatter / peer stores have the whole dW + update to complete.  N = 1: dW GEMM first (it walks the class
        # tiles from the end, where the forward's spill is still in L2), then dX, then the update.
        dw_first = (W == 1) if self.dw_first == "auto" else bool(self.dw_first)
        dx, rs_work = None, None
        # fork: the tail of the dX path runs on a side stream next to the dW GEMM / update (see conf.dx_side_stream);
        # conf.dx_fork_gemm moves the fork in front of the dX GEMM, so the two gradient GEMMs run side by side and the
        # clusters that run out of tiles in one persistent kernel's last wave pick up the other kernel's tiles
        want_fork = True if self.dx_side_stream == "auto" else bool(self.dx_side_stream)
        fork = want_fork and w.is_cuda and need_dx and (W == 1 or peer is not None)
        gemm_fork = fork and self._gemm_fork(n)
        if dw_first and not gemm_fork:
            K.backward_dw(ws.E, n_pad, ws.xs, B, n, d, dwn)
        tail = None
        if need_dx:
            splits = K.dx_splits(B, n, d)
            if fork and self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=w.device)
            if gemm_fork:
                tail = self._side_stream
                tail.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(tail) if tail is not None else contextlib.nullcontext():
                if ws.wn_b is not ws.wn and not self._wn_b_valid:
                    # AMP mode, first step / after an external optimizer / sampled rows: afterwards the fused update
                    # writes the bf16 twin next to the fp16 shard
                    K.cast_f16_to_bf16(ws.wn, ws.wn_b, n * d)
                    self._wn_b_valid = True
                K.backward_dx(ws.E, n_pad, ws.wn_b, B, n, d, ws.dx_partial, splits)
                if gemm_fork:        # the fused update rewrites the shard the dX GEMM is reading: it waits for this
                    dx_done = torch.cuda.Event()
                    dx_done.record()
            dx = torch.empty(b, d, dtype=torch.float32, device=x_in.device)
            if fork and not gemm_fork:
                tail = self._side_stream
                tail.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(tail) if tail is not None else contextlib.nullcontext():
                if W == 1:
                    K.dx_finalize(ws.dx_partial, splits, ws.coef, self._x_local, ws.inv_x, 1.0, B, B, d, dx)
                elif peer is not None:
                    # :505-522 -- every rank stores its scaled partial of row i into the owner's slot; the owner sums
                    # the W slots in rank order inside the normalise-backward kernel (x W, :521)
                    K.peer_dx_scatter(ws.dx_partial, splits, ws.coef, B, b, d, self.rank, W, peer.ptrs("dx_slots"))
                    if tail is not None:
                        # barrier + :521; also the fence that keeps a fast rank's NEXT gather out of xn_all while a
                        # slow rank still reads it (this rank signals after its last read of the gathered batch)
                        K.peer_dx_finalize(peer.ptrs("flags"), peer.counter, self.rank, W, peer.dx_slots,
                                           self._x_local, ws.inv_x, float(W), b, d, dx)
                else:
                    K.dx_finalize(ws.dx_partial, splits, ws.coef, None, None, 1.0, B, B, d, ws.dxn_all)
                    # :505-519 -- asynchronous: it overlaps the rank-local dW GEMM / update below
                    rs_work = distributed.reduce_scatter_tensor(ws.dxn_local, ws.dxn_all, distributed.ReduceOp.SUM,
                                                                async_op=True)
        if dw_first and gemm_fork:
            K.backward_dw(ws.E, n_pad, ws.xs, B, n, d, dwn)
        if not dw_first:
            K.backward_dw(ws.E, n_pad, ws.xs, B, n, d, dwn)
        if self.fused_optimizer:
            if gemm_fork and need_dx:
                torch.cuda.current_stream().wait_event(dx_done)
            self._fused_step(w, n, d, dwn, ws.wn)         # in place, after the last reader of this step's shard
        else:
            dw = torch.empty(n, d, dtype=torch.float32, device=w.device)
            K.dw_finalize(dwn, w, ws.inv_w, n, d, 1.0, dw)
            self._wn_valid = False    # an external optimizer is about to change the weights
        if rs_work is not None:
            rs_work.wait()
            K.dx_finalize(ws.dxn_local, 1, None, self._x_local, ws.inv_x, float(W), b, b, d, dx)        # :521
        if peer is not None and tail is None:
            # also the fence that keeps a fast rank's NEXT gather out of xn_all while a slow rank still reads it
            if dx is not None:
                K.peer_dx_finalize(peer.ptrs("flags"), peer.counter, self.rank, W, peer.dx_slots, self._x_local,
                                   ws.inv_x, float(W), b, d, dx)                  # barrier + :521
            else:
                K.peer_barrier(peer.ptrs("flags"), peer.counter, self.rank, W)
        if tail is not None:
            torch.cuda.current_stream().wait_stream(tail)                          # join
        return dx, dw

    def _gemm_fork(self, n):
        """conf.dx_fork_gemm; "auto": off (see DESIGN.md section 6 for the measurement)."""
        if self.dx_fork_gemm == "auto":
            return False
        return bool(self.dx_fork_gemm)

    def _fused_step(self, w, n, d, dwn, wn_out):
        raise NotImplementedError

    # ------------------------------------------------------------------ checkpoint layout (nets/PartialFC.py:210-232)
    def state_dict(self, destination=None, prefix="", keep_vars=False):
        if destination is None:
            destination = collections.OrderedDict()
            destination._metadata = collections.OrderedDict()
        for name, module in self._modules.items():
            if module is not None:
                module.state_dict(destination=destination, prefix=prefix + name + ".", keep_vars=keep_vars)
        if self.sample_rate < 1:
            destination["weight"] = self.weight.detach()
        else:
            destination["weight"] = self.weight_activated.data.detach()
        return destination

    def load_state_dict(self, state_dict, strict: bool = True):
        self._wn_valid = False
        if self.sample_rate < 1:
            self.weight = state_dict["weight"].to(self.weight.device)
            for nm in self._state_names:
                getattr(self, "weight_" + nm).zero_()
                getattr(self, "weight_activated_" + nm).zero_()
            self.weight_activated.data.zero_()
            if self._state_names == ["mom"]:
                self.weight_index.zero_()
        else:
            self.weight_activated.data = state_dict["weight"].to(self.weight_activated.data.device)
            self._fused_state = None


class PartialFC(_PartialFCBase):
    """Class-sharded margin-softmax head driven by torch.optim.SGD (nets/PartialFC.py:10-232)."""
    _optimizer_kind = "sgd"

    def _optimizer_state_names(self):
        return ["mom"]

    def _patch_optimizer(self, optimizer):
        if isinstance(optimizer, torch.optim.SGD):
            # the params of partial fc must be last in the params list (:124)
            optimizer.state.pop(optimizer.param_groups[-1]["params"][0], None)
            optimizer.param_groups[-1]["params"][0] = self.weight_activated
            optimizer.state[self.weight_activated]["momentum_buffer"] = self.weight_activated_mom
        else:
            raise RuntimeError("PartialFC needs torch.optim.SGD (nets/PartialFC.py:130-131)")

    def _read_optimizer(self, optimizer):
        if not isinstance(optimizer, torch.optim.SGD):
            raise RuntimeError("fused_optimizer: PartialFC needs torch.optim.SGD")
        g = optimizer.param_groups[-1]
        if g.get("dampening", 0) != 0 or g.get("nesterov", False) or g.get("maximize", False):
            raise RuntimeError("fused SGD supports dampening=0, nesterov=False, maximize=False")
        return dict(lr=float(g["lr"]), momentum=float(g["momentum"]), wd=float(g["weight_decay"]))

    def _momentum(self, w):
        if self.sample_rate < 1:
            return self.weight_activated_mom
        if self._fused_state is None:
            self._fused_state = torch.zeros_like(w)
            self.weight_activated_mom = self._fused_state
        return self._fused_state

    def _fused_step(self, w, n, d, dwn, wn_out):
        ws, o = self._ws, self._opt_args
        if self._indexed:
            # rows weight_index of the full shard / momentum, in place; the next step samples other rows, so no bf16 shard
            K.dw_sgd(dwn, self.weight, self.weight_mom, ws.inv_w, n, d, o["lr"], o["momentum"], o["wd"], ws.gscale, None,
                     None, index=ws.index)
            return
        K.dw_sgd(dwn, w, self._momentum(w), ws.inv_w, n, d, o["lr"], o["momentum"], o["wd"], ws.gscale, wn_out, ws.inv_w,
                 wn_next_b=ws.wn_b)
        self._wn_valid = True             # the update wrote next step's normalised bf16 rows and 1/norm in place
        self._wn_b_valid = True           # ... and, in AMP mode, their bf16 twin


class PartialFCAdamW(_PartialFCBase):
    """Same head with Adam / AdamW state (nets/PartialFC.py:235-432)."""
    _optimizer_kind = "adamw"

    def _optimizer_state_names(self):
        return ["exp_avg", "exp_avg_sq"]

    @torch.no_grad()
    def sample(self, labels_local, index_positive, optimizer, perm=None):
        self.step += 1                                                            # :306
        super().sample(labels_local, index_positive, optimizer, perm)

    def _patch_optimizer(self, optimizer):
        if isinstance(optimizer, (torch.optim.Adam, torch.optim.AdamW)):
            optimizer.state.pop(optimizer.param_groups[-1]["params"][0], None)
            optimizer.param_groups[-1]["params"][0] = self.weight_activated
            st = optimizer.state[self.weight_activated]
            st["exp_avg"] = self.weight_activated_exp_avg
            st["exp_avg_sq"] = self.weight_activated_exp_avg_sq
            # the reference stores a Python int (:327); current torch expects a tensor step for Adam/AdamW
            st["step"] = torch.tensor(float(self.step))
        else:
            raise RuntimeError("PartialFCAdamW needs torch.optim.Adam or AdamW (nets/PartialFC.py:328-329)")

    def _read_optimizer(self, optimizer):
        if not isinstance(optimizer, (torch.optim.Adam, torch.optim.AdamW)):
            raise RuntimeError("fused_optimizer: PartialFCAdamW needs torch.optim.Adam / AdamW")
        g = optimizer.param_groups[-1]
        if g.get("amsgrad", False) or g.get("maximize", False):
            raise RuntimeError("fused Adam supports amsgrad=False, maximize=False")
        decoupled = isinstance(optimizer, torch.optim.AdamW) or bool(g.get("decoupled_weight_decay", False))
        return dict(lr=float(g["lr"]), beta1=float(g["betas"][0]), beta2=float(g["betas"][1]), eps=float(g["eps"]),
                    wd=float(g["weight_decay"]), decoupled=decoupled)

    def _fused_step(self, w, n, d, dwn, wn_out):
        ws, o = self._ws, self._opt_args
        index = None
        if self.sample_rate < 1:
            if self._indexed:
                w, m, v, index, wn_out = self.weight, self.weight_exp_avg, self.weight_exp_avg_sq, ws.index, None
            else:
                m, v = self.weight_activated_exp_avg, self.weight_activated_exp_avg_sq
            # The reference hands torch.optim.AdamW state["step"] = self.step BEFORE optimizer.step(), which increments it
            # once more: forward call t is bias-corrected with t + 1 (nets/PartialFC.py:306, :327; pinned by
            # tests/golden/head_w*_adamw_sampled.npz).  The un-fused path inherits that from torch; mirror it here.
            step = self.step + 1
        else:
            if self._fused_state is None:
                self._fused_state = (torch.zeros_like(w), torch.zeros_like(w))
            m, v = self._fused_state
            self.step += 1
            step = self.step
        # under GraphedHeadStep the step count comes from a device scalar that every replay advances (the kernel uses
        # adam_step[0] + 1; for a sampled shard the scalar runs one ahead, see the quirk above)
        graphed = self._graph_steps
        K.dw_adam(dwn, w, m, v, ws.inv_w, n, d, o["lr"], o["beta1"], o["beta2"], o["eps"], o["wd"], step,
                  o["decoupled"], ws.gscale, wn_out, None if wn_out is None else ws.inv_w,
                  ws.adam_step if graphed else None, index=index, wn_next_b=None if wn_out is None else ws.wn_b)
        if graphed:
            ws.adam_step.add_(1)
        self._wn_valid = self._wn_b_valid = wn_out is not None
