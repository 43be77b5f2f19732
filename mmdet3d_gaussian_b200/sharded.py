"""Row sharding of the GD loss across the GPUs of one box.

Rows are independent; the only cross-row step is the final sum.  Rank ``g`` owns
a contiguous block of rows (block starts are multiples of 4 rows so every shard
of a contiguous ``[N,7]`` fp32 tensor stays 16-byte aligned for the bulk-copy
kernel), computes ``scale_global * sum_i w_i loss_i`` over its block with the
fused kernel -- gradients stay sharded, matching the reference's one process per
GPU (``tools/dist_train.sh:8``) -- and ONE ``all_reduce(SUM)`` of the scalar
(NCCL over NVLink on GPUs, gloo in the CPU tests) produces the global loss.
The reference itself issues no collective on this path.
"""
import torch
import torch.distributed as dist

ROW_ALIGN = 4


def shard_bounds(n, rank, world_size, align=ROW_ALIGN):
    """``[lo, hi)`` of rank's rows: contiguous, balanced, ``lo`` a multiple of
    ``align``; the union over ranks is exactly ``[0, n)``."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f'bad rank {rank} / world_size {world_size}')
    per = -(-n // world_size)                 # ceil
    per = -(-per // align) * align            # round up to the alignment
    lo = min(rank * per, n)
    hi = min(lo + per, n)
    return lo, hi


class ShardedGDLoss(torch.nn.Module):
    """Wraps a ``GDLoss``-like module: every rank passes ITS rows; the returned
    scalar is the loss over ALL ranks' rows.

    ``reduction='mean'`` without ``avg_factor`` divides by the global row count
    (all-reduced together with the loss: one collective, two floats);
    ``avg_factor`` is taken as already global (mmdet passes the all-reduced
    ``num_total_samples``).  The gradient of the returned scalar w.r.t. the local
    ``pred`` is the local shard's gradient, as with DDP.
    """

    def __init__(self, loss_module, group=None):
        super().__init__()
        self.loss_module = loss_module
        self.group = group

    def forward(self, pred, target, weight=None, avg_factor=None,
                reduction_override=None, **kwargs):
        reduction = reduction_override or self.loss_module.reduction
        if reduction == 'none':
            return self.loss_module(pred, target, weight, avg_factor=avg_factor,
                                    reduction_override='none', **kwargs)
        n_local = pred.numel() // 7
        if reduction == 'mean' and avg_factor is None:
            local = self.loss_module(pred, target, weight, reduction_override='sum',
                                     **kwargs)
            packed = torch.stack([local.detach(),
                                  torch.tensor(float(n_local), device=local.device,
                                               dtype=local.dtype)])
            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.group)
            n_global = packed[1]
            total = packed[0] / n_global
            # value = global mean; gradient = d(local sum)/d pred / n_global
            return total + (local - local.detach()) / n_global
        local = self.loss_module(pred, target, weight, avg_factor=avg_factor,
                                 reduction_override=reduction, **kwargs)
        total = local.detach().clone()
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
        return total + (local - local.detach())
