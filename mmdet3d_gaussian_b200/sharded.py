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


class PeerSumContext:
    """Exchange buffers for the in-kernel cross-GPU sum (``gd_peer_sum``,
    ``include/gd_loss_b200.h``): one small symmetric-memory allocation per rank
    (``torch.distributed._symmetric_memory``: every rank's buffer is mapped into every other
    rank's address space over NVLink), zero-filled, rendezvoused once.  The fused kernel's last
    CTA then writes its partial straight into the peers' buffers -- no NCCL launch on the path.

    ``PeerSumContext.create(group)`` returns ``None`` when symmetric memory is unavailable
    (single process, gloo, no P2P): callers then keep the NCCL all-reduce."""

    def __init__(self, handle, buf, shim_obj, world, rank):
        self.handle, self.buf, self.shim_obj = handle, buf, shim_obj     # keep the mapping alive
        self.world, self.rank = world, rank

    @classmethod
    def create(cls, group=None, device=None):
        if not (dist.is_available() and dist.is_initialized()):
            return None
        group = group if group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world < 2 or dist.get_backend(group) != 'nccl':
            return None
        from . import _lib
        sh = _lib.shim()
        ok = torch.ones(1, device=device or torch.device('cuda', torch.cuda.current_device()))
        ctx = None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            nbytes = int(sh.peer_sum_buffer_bytes())
            buf = symm_mem.empty(max(nbytes, 1024), dtype=torch.uint8, device=ok.device)
            buf.zero_()
            handle = symm_mem.rendezvous(buf, group.group_name)
            ptrs = [int(p) for p in handle.buffer_ptrs]
            ctx = cls(handle, buf, sh.PeerSum(world, rank, ptrs), world, rank)
            torch.cuda.synchronize()
        except Exception:                              # noqa: BLE001 -- capability probe
            ok.zero_()
        # all ranks take the same path: fused only if EVERY rank could map the buffers
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        return ctx if float(ok.item()) > 0 else None


class ShardedGDLoss(torch.nn.Module):
    """Wraps a ``GDLoss``-like module: every rank passes ITS rows; the returned
    scalar is the loss over ALL ranks' rows.

    ``reduction='mean'`` without ``avg_factor`` divides by the global row count
    (all-reduced together with the loss: one collective, two floats);
    ``avg_factor`` is taken as already global (mmdet passes the all-reduced
    ``num_total_samples``).  The gradient of the returned scalar w.r.t. the local
    ``pred`` is the local shard's gradient, as with DDP.
    """

    def __init__(self, loss_module, group=None, fused=False):
        """``fused=True``: sum the loss over the GPUs INSIDE the fused launch through peer
        memory (``PeerSumContext``) instead of a separate NCCL all-reduce; needs
        ``GDLoss(host_sync=False)`` (or ``weight=None`` calls) and falls back to NCCL when
        symmetric memory is unavailable (``self.fused`` tells which)."""
        super().__init__()
        self.loss_module = loss_module
        self.group = group
        self.peer = PeerSumContext.create(group) if fused else None
        self.fused = self.peer is not None

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
        if self.fused:
            # the kernel's last CTA exchanges the scaled partials over NVLink: the value IS the
            # global sum (identical bits on every rank), the gradient is the local shard's
            self.loss_module._peer_sum = self.peer.shim_obj
            try:
                return self.loss_module(pred, target, weight, avg_factor=avg_factor,
                                        reduction_override=reduction, **kwargs)
            finally:
                self.loss_module._peer_sum = None
        local = self.loss_module(pred, target, weight, avg_factor=avg_factor,
                                 reduction_override=reduction, **kwargs)
        total = local.detach().clone()
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
        return total + (local - local.detach())


# ---------------------------------------------------------------------------
# pairwise path ("Pairwise"): anchors (rows) sharded, the M GT
# boxes replicated.  Row minima are local; the per-GT minima over ALL anchors need one
# all-reduce(MIN) of M packed (value, global anchor index) 64-bit keys.
# ---------------------------------------------------------------------------
_NO_INDEX = 0x7fffffff          # "no anchor" (empty shard); sorts after every real index


def pack_min_keys(values, index):
    """float32 ``values`` [M] + int64 ``index`` [M] (< 2^31 - 1, negative = none) ->
    int64 keys whose signed order is (value, index) lexicographic with NaN FIRST and
    ties -> lowest index, the order the kernel's own column reduction uses
    (``csrc/gd_pairwise.cuh``) and ``torch.min`` follows."""
    bits = values.detach().to(torch.float32).contiguous().view(torch.int32).to(torch.int64)
    # order-preserving float -> signed int map: negatives have their low 31 bits flipped
    ordered = torch.where(bits < 0, bits ^ 0x7fffffff, bits)
    ordered = torch.where(torch.isnan(values), torch.full_like(ordered, -(1 << 31)), ordered)
    idx = index.to(torch.int64)
    idx = torch.where(idx < 0, torch.full_like(idx, _NO_INDEX), idx)
    return (ordered << 32) | idx


def unpack_min_keys(keys):
    """Inverse of ``pack_min_keys``: ``(values float32 [M], index int64 [M])``; index -1
    where no rank had an anchor; a NaN value comes back as the canonical NaN."""
    ordered = keys >> 32                                     # arithmetic shift keeps the sign
    idx = keys & 0xffffffff
    nan = ordered == -(1 << 31)
    bits = torch.where(ordered < 0, ordered ^ 0x7fffffff, ordered)
    bits = torch.where(nan, torch.full_like(bits, 0x7fc00000), bits)
    values = bits.to(torch.int32).view(torch.float32)
    idx = torch.where(idx == _NO_INDEX, torch.full_like(idx, -1), idx)
    return values, idx


def merge_column_minima(col_min, col_argmin, row_offset, group=None):
    """Per-GT ``(min, argmin)`` over all ranks' anchor shards from each rank's local
    ``(col_min [M], col_argmin [M])`` (local row indices, -1 = empty shard): local
    indices are shifted by ``row_offset`` and ONE ``all_reduce(MIN)`` of M int64 keys
    merges them.  Returns ``(global_min float32 [M], global_argmin int64 [M])``, the same
    on every rank and identical to the reduction of the unsharded matrix."""
    idx = col_argmin.to(torch.int64)
    idx = torch.where(idx >= 0, idx + int(row_offset), idx)
    keys = pack_min_keys(col_min, idx)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
    return unpack_min_keys(keys)


class ShardedGDMaxSimAssigner(torch.nn.Module):
    """``GDMaxSimAssigner`` over row-sharded anchors: every rank passes ITS anchors and
    the (replicated) GT boxes; labels come back for the local anchors only and equal
    the unsharded assignment.  ``assigner`` needs the ``GDMaxSimAssigner`` attributes
    (thresholds, ``cfg``); ``pairwise_fn`` / ``assign_fn`` default to the fused CUDA
    operators and exist so the CPU tests can inject the oracle."""

    def __init__(self, assigner, group=None, pairwise_fn=None, assign_fn=None):
        super().__init__()
        self.assigner = assigner
        self.group = group
        self._pairwise = pairwise_fn
        self._assign = assign_fn

    @torch.no_grad()
    def assign(self, bboxes, gt_bboxes, row_offset):
        a = self.assigner
        if self._pairwise is None:
            from . import ops
            row_min, row_arg, col_min, col_arg, _ = ops.pairwise_assign(
                bboxes[..., :7], gt_bboxes[..., :7], a.cfg)
        else:
            row_min, row_arg, col_min, col_arg = self._pairwise(bboxes, gt_bboxes)
        n = bboxes.shape[0]
        gmin, garg = merge_column_minima(col_min, col_arg, row_offset, self.group)
        local = garg - int(row_offset)
        local = torch.where((garg >= 0) & (local >= 0) & (local < n), local,
                            torch.full_like(local, -1))
        if self._assign is None:
            from . import ops
            assigned, max_ov = ops.assign_from_minima(
                row_min, row_arg, gmin, local, a.pos_iou_thr, a.neg_lo, a.neg_hi,
                a.min_pos_iou, a.match_low_quality)
        else:
            assigned, max_ov = self._assign(row_min, row_arg, gmin, local)
        if gt_bboxes.shape[0] == 0:
            assigned.zero_()
            max_ov.zero_()
        return dict(assigned_gt_inds=assigned, max_overlaps=max_ov,
                    gt_max_overlaps=1.0 - gmin, gt_argmax_overlaps=garg)

    forward = assign
