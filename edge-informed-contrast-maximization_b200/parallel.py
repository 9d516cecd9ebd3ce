"""Multi-GPU partitioning of the EINCM objective (one process per GPU, ``torch.distributed``).

Two ways the path shards:

* **Window / sequence sharding** - no data-path collective.  Windows of one sequence are chained through the handover prior
  (reference src/eincm/solver.py:254-256, 302-347), so with handover on the unit of distribution is the *sequence*
  (``shard_sequences_lpt``); with handover off (or first-sample semantics) windows are independent (``shard_windows``).
* **Event split of one huge window** - the splat is additive in events, everything after it needs the complete image:
  each rank votes its share of the events into private images, one all-reduce, every rank runs the (cheap) image pass
  redundantly, the backward runs on the local events and the flow gradient is all-reduced (``EventSplitObjective``).

The collectives are ``torch.distributed`` (NCCL over NVLink on the GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


# ---------------------------------------------------------------------------------------------------------------
# window / sequence sharding (no collective)
# ---------------------------------------------------------------------------------------------------------------
def shard_windows(n_windows: int, world: int, rank: int) -> List[int]:
    """Round-robin assignment of independent windows."""
    return list(range(rank, n_windows, world))


def shard_sequences_lpt(window_counts: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time bin packing of whole sequences (windows inside a sequence are chained by the handover).
    Returns, per rank, the list of sequence indices.  E.g. the 7 DSEC test sequences (286/541/91/376/376/56/361 windows,
    reference docs/assets/dsec_extended_evals/*.csv)."""
    order = sorted(range(len(window_counts)), key=lambda i: (-window_counts[i], i))
    loads = [0] * world
    bins: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        bins[r].append(i)
        loads[r] += window_counts[i]
    return bins


# ---------------------------------------------------------------------------------------------------------------
# event split
# ---------------------------------------------------------------------------------------------------------------
def split_events(xs, ys, ts, world: int, rank: int):
    """Contiguous slab ``rank`` of ``world`` of the (time-sorted) event stream; every event lands on exactly one rank."""
    n = len(xs)
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    return xs[lo:hi], ys[lo:hi], ts[lo:hi]


class EventSplitObjective:
    """Objective + gradient of ONE window whose events are split over the ranks of ``group``.

    ``plan`` is an ``eincm_b200.plan.Plan`` created with ``FLAG_EVENT_SPLIT`` (or any object with the same split-phase
    methods - the CPU tests drive this class with an oracle-backed stand-in over gloo).  Every rank passes ITS events to
    ``set_datasample``; ``value_and_grad`` returns the same ``(loss, grad)`` on every rank.

    ``fixed_point=True``: the per-evaluation all-reduce runs on the int64 fixed-point images (same bytes, order-independent sums, the
    fused image pass follows directly).

    ``p2p=True`` (GPUs of one NVLink domain, NCCL group): the ranks exchange the CUDA IPC handles of their fixed-point image
    buffers once, and from then on every splat adds its votes to the images of ALL ranks over NVLink - the all-reduce of the
    ``R*H*W`` images is fused into the splat kernel (``include/eincm.h``); the only collectives left per evaluation are two
    one-element barriers and the all-reduce of the small flow gradient."""

    def __init__(self, plan, make_hparams: Callable, group=None, rank: Optional[int] = None, world: Optional[int] = None,
                 p2p: bool = False, fixed_point: bool = False):
        import torch.distributed as dist
        self.dist = dist
        self.plan = plan
        self.group = group
        self.make_hparams = make_hparams
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        plan.set_event_split(self.rank, self.world)
        self.n_collectives = 0
        self.collective_bytes = 0
        self.p2p = bool(p2p)
        # all-reduce the int64 fixed-point images instead of the float64 ones: bit-identical sums on every rank whatever the reduction
        # order, and the fused image pass runs on them directly (include/eincm.h: eincm_plan_set_split_fixed_point)
        self.fixed_point = bool(fixed_point) and not self.p2p
        if self.fixed_point:
            plan.set_split_fixed_point(True)
        self._token = None
        if self.p2p:
            handles = [None] * self.world
            dist.all_gather_object(handles, plan.ipc_handle(), group=group)
            plan.set_peers(handles)

    def _allreduce(self, t, op=None):
        op = self.dist.ReduceOp.SUM if op is None else op
        if self.world > 1:
            self.dist.all_reduce(t, op=op, group=self.group)
        self.n_collectives += 1
        self.collective_bytes += t.numel() * t.element_size()

    def _barrier(self):
        """Stream-ordered cross-rank barrier: a one-element all-reduce on the current stream."""
        import torch
        if self._token is None:
            self._token = torch.zeros(1, dtype=torch.float32, device=f'cuda:{self.plan.device}')
        self._allreduce(self._token)

    def set_datasample(self, xs_local, ys_local, ts_local, edges, edge_ts, need_mask: bool = True):
        if self.p2p:
            self.plan.split_prepare()
            self._barrier()                                              # every rank's image buffer is clean
            self.plan.set_window(xs_local, ys_local, ts_local, edges, edge_ts)   # votes of the zero-warp image go to all ranks
            self._barrier()                                              # all votes have landed
            self.plan.split_window_images()
        else:
            self.plan.set_window(xs_local, ys_local, ts_local, edges, edge_ts)
            self._allreduce(self.plan.zero_iwe())                        # sum of the partial un-warped images
        if need_mask:
            self._allreduce(self.plan.event_mask(), self.dist.ReduceOp.MAX)   # union of the per-rank event masks (TV only)
        self.plan.window_finalize()

    def value_and_grad(self, theta, cur_pyr_lvl: int, loss_out=None, grad_out=None):
        """theta: (h, w, 2) float64 tensor on the plan's device (identical on every rank)."""
        import torch
        hp = self.make_hparams(cur_pyr_lvl)
        if loss_out is None:
            loss_out = torch.zeros(1, dtype=torch.float64, device=theta.device)
        if grad_out is None:
            grad_out = torch.zeros_like(theta)
        if self.p2p:
            self.plan.split_prepare()
            self._barrier()                                              # nobody still reads / clears the image buffers
            self.plan.forward_events(theta, hp)                          # splat fused with the all-reduce (NVLink reductions)
            self._barrier()                                              # all votes have landed: images complete everywhere
        else:
            self.plan.forward_events(theta, hp)
            self._allreduce(self.plan.iwe_fix() if self.fixed_point else self.plan.iwe())    # C1: partial images of warped events
        self.plan.backward(hp, loss_out, grad_out)
        self._allreduce(grad_out)                                        # C2: partial flow-parameter gradients
        return loss_out, grad_out
