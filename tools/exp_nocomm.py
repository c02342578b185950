"""EXPERIMENT ONLY (results are numerically wrong): run bench.py with the head's collectives replaced by no-ops to
measure how much of the multi-GPU step is collective latency.  Usage: torchrun ... tools/exp_nocomm.py --gpus N ..."""
import contextlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch.distributed as dist   # noqa: E402


class _Done:
    def wait(self):
        return True


def _noop(*a, **k):
    return _Done() if k.get("async_op") else None


_real_ag, _real_ar, _real_rs, _real_cm = (dist.all_gather_into_tensor, dist.all_reduce, dist.reduce_scatter_tensor,
                                          dist._coalescing_manager)
import face_recognition_pytorch_b200.partial_fc as pf   # noqa: E402


class _FakeDist:
    """stands in for torch.distributed inside partial_fc only (bench.py keeps the real module for its barriers)"""
    ReduceOp = dist.ReduceOp

    def __getattr__(self, name):
        return getattr(dist, name)

    all_gather_into_tensor = staticmethod(_noop)
    all_reduce = staticmethod(_noop)
    reduce_scatter_tensor = staticmethod(_noop)

    @staticmethod
    @contextlib.contextmanager
    def _coalescing_manager(*a, **k):
        yield


pf.distributed = _FakeDist()
import bench   # noqa: E402

if __name__ == "__main__":
    bench.main()
