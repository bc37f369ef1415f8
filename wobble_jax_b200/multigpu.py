"""One process per GPU: chunk sharding (no collective) and epoch sharding (one all-reduce).

SURVEY.md section 8e.  The reference's only parallelism is a ``multiprocessing.Pool`` over echelle
orders (notebooks/02-51peg.py:66-78): chunks/orders are independent fits, so they are dealt
round-robin to the ranks and nothing is exchanged.  When ONE chunk has more epochs than a GPU
should hold, its rows are split across ranks on ``which_key`` boundaries (every epoch key lives on
one rank); loss and control-point gradients are sums over rows, so ONE
``ncclAllReduce`` of ``[loss, template gradients]`` enqueued by the library behind the kernels
(``jb_group``, include/jabble_b200.h) completes them; per-epoch entries belong to one rank.
"""

import numpy as np

from . import engine
from .dataset import DataBlock


def shard_chunks(n_chunks, rank, world_size):
    """Chunk c -> rank c mod world_size."""
    return [c for c in range(n_chunks) if c % world_size == rank]


def shard_rows_by_key(keys, world_size):
    """Split rows into ``world_size`` parts of nearly equal row count such that all rows sharing a
    key are in the same part and parts are contiguous in key order.

    ``keys``: (E,) integer key of every row (the metablock column the epoch-specific models use).
    Returns a list of row-index arrays (ascending), one per rank.
    """
    keys = np.asarray(keys).astype(np.int64).reshape(-1)
    if keys.size == 0:
        return [np.zeros(0, dtype=np.int64) for _ in range(world_size)]
    counts = np.bincount(keys)
    cum = np.cumsum(counts)
    total = cum[-1]
    # key k goes to the part whose row quota contains the midpoint of k's row span
    mid = cum - counts / 2.0
    part_of_key = np.minimum((mid * world_size / total).astype(np.int64), world_size - 1)
    part_of_row = part_of_key[keys]
    return [np.nonzero(part_of_row == r)[0] for r in range(world_size)]


def epoch_keys(model, metablock):
    """The key column shared by every epoch-specific leaf of the tree (rows are sharded on it)."""
    spec = engine.TreeSpec(model)
    names = {getattr(lf.node, "which_key") for lf in spec.leaves if hasattr(lf.node, "which_key")}
    if not names:
        return np.arange(len(metablock), dtype=np.int64)
    if len(names) > 1:
        cols = [np.asarray(metablock[n]) for n in names]
        if not all(np.array_equal(cols[0], c) for c in cols[1:]):
            raise ValueError(f"epoch sharding needs one key column; the model uses {sorted(names)} with different values")
    return np.asarray(metablock[next(iter(names))]).astype(np.int64)


def take_rows(datablock, metablock, rows):
    """Row subset of a (datablock, metablock) pair; keys keep their GLOBAL values."""
    db = DataBlock(**{k: np.ascontiguousarray(np.asarray(datablock[k])[rows]) for k in ("xs", "ys", "yivar", "mask")})
    mb = DataBlock(**{k: np.asarray(metablock[k])[rows] for k in metablock.keys()})
    return db, mb


def all_reduce_sum(vec, group=None):
    """Sum a [loss, gradient...] vector (torch tensor, CPU or CUDA) over the ranks, in place."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec


class EpochShardedObjective:
    """(func, fprime) for ONE chunk whose rows are split over the ranks of ``group``.

    Every rank holds its rows in a plan of the WHOLE model; an evaluation is ``jb_group_eval_device``
    (the local fused kernels, then one NCCL all-reduce of [loss, template gradients], 8 (1 + 2 M)
    bytes, enqueued behind them by the library), then -- because scipy on every rank needs the whole
    gradient -- one all-reduce of the per-epoch entries (each key has rows on one rank only, the other
    ranks contribute zeros), then a single D2H copy.  All ranks must call ``value`` / ``gradient``
    with the same parameter vector.  A rank whose trial point leaves a knot grid contributes a NaN
    loss, so every rank takes the same branch; knot-window overflows (the shifts drifted since the
    rows were sorted) are settled inside the library, rank by rank, before the exchange.
    """

    def __init__(self, loss, model, datablock, metablock, rank, world_size, device=0, group=None):
        import torch
        from . import loss as L
        if not isinstance(loss, L.ChiSquare):
            raise NotImplementedError("only ChiSquare runs on the CUDA path")
        self.torch = torch
        self.group = group
        parts = shard_rows_by_key(epoch_keys(model, metablock), world_size)
        empty = [r for r, rows in enumerate(parts) if rows.size == 0]
        if empty:      # the same on every rank: nobody enters a collective
            raise ValueError(f"epoch sharding over {world_size} ranks leaves rank(s) {empty} without rows "
                             f"({len(np.unique(epoch_keys(model, metablock)))} epoch keys): use fewer ranks")
        rows = parts[rank]
        self.rows = rows
        db, mb = take_rows(datablock, metablock, rows)
        self.plan = engine.build_plan(model, db, mb, coefficient=loss.coefficient, device=device)
        self.n_fit = self.plan.n_fit
        self.dev = torch.device("cuda", device)
        self.d_p = torch.empty(max(self.n_fit, 1), dtype=torch.float64, device=self.dev)
        self.d_out = torch.empty(self.n_fit + 1, dtype=torch.float64, device=self.dev)
        # which outputs are per-epoch (not template control points): they cross the ranks separately
        spec = engine.TreeSpec(model)
        local, o = [], 1
        for lf in spec.leaves:
            if lf.node._fit:
                if hasattr(lf.node, "which_key") and not hasattr(lf.node, "xs"):
                    local.append((o, o + lf.size))
                o += lf.size
        self.local_slices = local
        # without an initialised torch.distributed the shard stands alone (its sums are the caller's to add up)
        import torch.distributed as dist
        comm_world = world_size if (dist.is_available() and dist.is_initialized()) else 1
        nccl_id = None
        if comm_world > 1:
            torch.cuda.set_device(self.dev)
            nccl_id = engine.Group.broadcast_unique_id(rank, comm_world, group)
        self.jgroup = engine.Group([self.plan], rank if comm_world > 1 else 0, comm_world, nccl_id)
        self.jgroup.bind([self.d_p.data_ptr()], [self.d_out.data_ptr()])
        # a real stream (the C ABI reads NULL as "the plan's own stream"), made current for the copies
        # and the second all-reduce so that everything is ordered on it
        self.stream = torch.cuda.Stream(self.dev)
        self.world_size = comm_world
        self._p = None
        self._f = None
        self._g = None
        self._f_inside = None

    def _eval(self, p):
        torch = self.torch
        p = np.asarray(p, dtype=np.float64).reshape(-1)
        if self._p is None or not np.array_equal(p, self._p):
            with torch.cuda.stream(self.stream):
                self.d_p[:self.n_fit].copy_(torch.from_numpy(p), non_blocking=True)
                self.jgroup.eval_device_checked(self.stream.cuda_stream)
                if self.world_size > 1:
                    for a, b in self.local_slices:
                        all_reduce_sum(self.d_out[a:b], self.group)
                out = self.d_out.cpu().numpy()
            if not np.isfinite(out[0]):
                # some rank's pixels left a knot grid (its k_finish wrote NaN; the all-reduce spread it):
                # the first evaluation raises, a later one (a line-search excursion) gets the penalty
                # engine.Objective gives it -- on every rank alike
                if self._f_inside is None:
                    raise ValueError("an unmasked warped pixel lies outside the knot grid of a spline component (on some rank)")
                out = np.concatenate([[1e8 * max(abs(self._f_inside), 1.0)], np.zeros(self.n_fit)])
            else:
                self._f_inside = float(out[0])
            self._f, self._g, self._p = float(out[0]), out[1:].copy(), p.copy()
        return self._f, self._g

    def value(self, p):
        return self._eval(p)[0]

    def gradient(self, p):
        return self._eval(p)[1].copy()

    def close(self):
        self.jgroup.close()
        self.plan.close()
