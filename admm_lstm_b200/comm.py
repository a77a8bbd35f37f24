"""Sample-sharding plumbing: one process per GPU, torch.distributed for the few
small all-reduces of a step (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


class Comm:
    """Thin view of the default process group; a no-op when the job has a single rank."""

    def __init__(self, group=None):
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.group = group
        self.world_size = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0
        self.events = None            # set by enable_timing()

    def _reduce(self, t: torch.Tensor, op) -> None:
        if self.events is not None and t.is_cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_reduce(t, op=op, group=self.group)
            e1.record()
            self.events.append((t.numel() * t.element_size(), e0, e1))
        else:
            dist.all_reduce(t, op=op, group=self.group)

    def allreduce_sum_(self, *tensors: torch.Tensor) -> None:
        """In-place sum over ranks.  Identical reduced values on every rank keep the replicated
        theta decisions (and therefore the weights) bit-identical without a broadcast."""
        if not self.active:
            return
        for t in tensors:
            self._reduce(t, dist.ReduceOp.SUM)

    def allreduce_max_(self, *tensors: torch.Tensor) -> None:
        """In-place maximum over ranks (the global torch.max scalars of ADMM-LSTM-L, admm_lstm.py:168,179,225)."""
        if not self.active:
            return
        for t in tensors:
            self._reduce(t, dist.ReduceOp.MAX)

    def allreduce_max_and_sum_(self, max_t: torch.Tensor, sum_t: torch.Tensor) -> None:
        """One collective for a MAX-reduced and a SUM-reduced scalar (ADMM-LSTM-L needs both between the two kernels of every
        timestep, admm_lstm.py:225,230): all-gather the pair, reduce locally in rank order -- identical on every rank."""
        if not self.active:
            return
        pair = torch.stack([max_t.reshape(()).double(), sum_t.reshape(()).double()])
        out = torch.empty((self.world_size, 2), dtype=torch.float64, device=pair.device)
        if self.events is not None and pair.is_cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_gather(list(out.unbind(0)), pair, group=self.group)
            e1.record()
            self.events.append((16, e0, e1))
        else:
            dist.all_gather(list(out.unbind(0)), pair, group=self.group)
        max_t.copy_(out[:, 0].max().to(max_t.dtype).reshape(max_t.shape))
        sum_t.copy_(out[:, 1].sum().to(sum_t.dtype).reshape(sum_t.shape))

    def enable_timing(self, enabled: bool = True) -> None:
        """Bracket every collective with CUDA events on the launching stream: the elapsed time of a collective is its
        transfer plus the wait for the slowest rank to arrive -- the evidence for what limits multi-GPU scaling."""
        self.events = [] if enabled else None

    def timing_summary(self):
        """{'calls', 'ms', 'bytes', 'ms_small' (payload <= 64 KB), 'ms_large'} of the recorded collectives; synchronises."""
        torch.cuda.synchronize()
        out = {"calls": 0, "ms": 0.0, "bytes": 0, "calls_small": 0, "ms_small": 0.0, "ms_large": 0.0}
        for nbytes, e0, e1 in self.events or []:
            ms = e0.elapsed_time(e1)
            out["calls"] += 1
            out["ms"] += ms
            out["bytes"] += nbytes
            if nbytes <= 65536:
                out["calls_small"] += 1
                out["ms_small"] += ms
            else:
                out["ms_large"] += ms
        return out

    def shard_range(self, n_total: int) -> Tuple[int, int]:
        """Contiguous, balanced slice [lo, hi) of n_total samples owned by this rank."""
        base, rem = divmod(n_total, self.world_size)
        lo = self.rank * base + min(self.rank, rem)
        return lo, lo + base + (1 if self.rank < rem else 0)

    def sum_int(self, value: int, device) -> int:
        if not self.active:
            return int(value)
        t = torch.tensor([value], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return int(t.item())
