"""Symmetric-memory buffers for the peer-memory exchanges (csrc/pfc_peer.cu).

One symmetric allocation per rank (torch.distributed._symmetric_memory), carved into
    flags      uint32 [W]            barrier flags, one per sender
    xn_all     bf16   [W*b, d]       all-gathered normalised batch       (written by every peer; fp16 in AMP mode)
    labels_all int64  [W*b]          all-gathered labels
    slots      fp32   [W, B, 2]      softmax statistics, one slot per sender
    dx_slots   fp32   [W, b, d]      scaled dXn rows owned by this rank, one slot per sender
After the rendezvous every rank holds the device pointers of all W buffers as mapped into its own address space;
they are handed to the kernels as small host arrays of pointers.  Only ranks of ONE node with NVLink / P2P access
qualify; anything else raises and the head keeps the NCCL collectives.
"""
import torch
import torch.distributed as dist

from . import _lib


def _align(v, a=256):
    return (v + a - 1) // a * a


class PeerExchange:
    def __init__(self, device, rank, world, b, d, timeout_ms=None, operand_dtype=torch.bfloat16):
        import torch.distributed._symmetric_memory as symm_mem
        if timeout_ms is not None:      # how long a flag barrier waits for a stalled peer before it traps (default 10 min)
            _lib.check(_lib.lib.pfc_peer_set_timeout_ms(float(timeout_ms)), "pfc_peer_set_timeout_ms")
        if world > _lib.lib.pfc_peer_max_ranks():
            raise RuntimeError("too many ranks for the peer-memory exchange")
        B = b * world
        self.rank, self.world, self.b, self.d, self.B = rank, world, b, d, B
        off = {}
        cur = 0
        for name, nbytes in (("flags", 4 * world), ("xn_all", B * d * 2), ("labels_all", B * 8),
                             ("slots", world * B * 2 * 4), ("dx_slots", world * b * d * 4)):
            off[name] = cur
            cur = _align(cur + nbytes)
        self.nbytes = cur
        self.buf = symm_mem.empty(cur, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, dist.group.WORLD.group_name)
        torch.cuda.synchronize(device)
        dist.barrier()                      # every rank has zeroed its flags before anybody signals
        bases = [int(p) for p in self.handle.buffer_ptrs]
        if len(bases) != world:
            raise RuntimeError("symmetric memory rendezvous returned an unexpected number of peers")
        self._ptrs = {name: _lib.ptr_array([base + o for base in bases]) for name, o in off.items()}
        # local typed views
        view = lambda name, dtype, shape: self.buf[off[name]: off[name] + _nbytes(dtype, shape)].view(dtype).view(shape)  # noqa: E731
        self.xn_all = view("xn_all", operand_dtype, (B, d))       # bf16, or fp16 (conf.mixed_precision)
        self.labels_all = view("labels_all", torch.int64, (B,))
        self.slots = view("slots", torch.float32, (world, B, 2))
        self.dx_slots = view("dx_slots", torch.float32, (world, b, d))
        self.counter = torch.zeros(2, dtype=torch.int32, device=device)   # this rank's barrier state {epoch, ticket}

    def ptrs(self, name):
        return self._ptrs[name]


def _nbytes(dtype, shape):
    n = 1
    for s in shape:
        n *= s
    return n * torch.empty((), dtype=dtype).element_size()
