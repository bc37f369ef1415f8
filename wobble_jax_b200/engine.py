"""Model tree -> flat plan -> C ABI calls.

This is the host half of the drop-in boundary: it recognises the tree
shapes the reference's recipes build (jabble/quickplay.py:131-236)

    top      := component | AdditiveModel[component, ...]
    component:= spline | CompositeModel[ShiftingModel?, spline, StretchingModel?]
    spline   := CardinalSplineMixture | NormalizationModel(CardinalSplineMixture)

lays every leaf's parameters out in one vector in DFS order (the order of
``ContainerModel.get_parameters``, jabble/model.py:478-495, when everything is in fit mode) and
hands data + descriptors to ``jb_plan_create``.  Anything outside the grammar raises
``ValueError``: there is no CPU fallback.
"""

import ctypes as C

import numpy as np

from . import _lib
from . import model as M


class Leaf:
    __slots__ = ("node", "offset", "size")

    def __init__(self, node, offset, size):
        self.node, self.offset, self.size = node, offset, size


class TreeSpec:
    """Flat description of a model tree."""

    def __init__(self, root):
        self.root = root
        self.leaves = []        # DFS order
        self.n_params = 0
        self.components = []    # dicts: shift/spline/stretch leaves
        self._walk(root)
        tops = root.models if isinstance(root, M.AdditiveModel) else [root]
        for top in tops:
            self.components.append(self._component(top))
        if len(self.components) > 3:
            raise ValueError(f"{len(self.components)} additive components; the CUDA plan supports up to 3")

    def _walk(self, node):
        if isinstance(node, M.ContainerModel):
            for m in node.models:
                self._walk(m)
            return
        if isinstance(node, (M.EpochSpecificModel, M.CardinalSplineMixture, M.NormalizationModel)):
            size = int(np.asarray(node.p).size)
            self.leaves.append(Leaf(node, self.n_params, size))
            self.n_params += size
            return
        raise ValueError(f"unsupported model node {type(node).__name__}")

    def leaf_of(self, node):
        for lf in self.leaves:
            if lf.node is node:
                return lf
        raise KeyError(node)

    def _component(self, top):
        def flatten(n):
            if isinstance(n, M.CompositeModel):
                out = []
                for m in n.models:
                    out += flatten(m)
                return out
            if isinstance(n, M.AdditiveModel):
                raise ValueError("an AdditiveModel nested below the top level is outside the supported grammar")
            return [n]

        chain = flatten(top)
        comp = {"shift": None, "spline": None, "stretch": None}
        for n in chain:
            if isinstance(n, M.ShiftingModel):
                if comp["spline"] is not None or comp["shift"] is not None:
                    raise ValueError("component grammar is [ShiftingModel?, spline, StretchingModel?]")
                comp["shift"] = n
            elif isinstance(n, (M.CardinalSplineMixture, M.NormalizationModel)):
                if comp["spline"] is not None:
                    raise ValueError("two spline models in series are outside the supported grammar")
                comp["spline"] = n
            elif isinstance(n, M.StretchingModel):
                if comp["spline"] is None or comp["stretch"] is not None:
                    raise ValueError("component grammar is [ShiftingModel?, spline, StretchingModel?]")
                comp["stretch"] = n
            else:
                raise ValueError(f"unsupported model node {type(n).__name__}")
        if comp["spline"] is None:
            raise ValueError("every additive component needs a CardinalSplineMixture or NormalizationModel")
        sp = comp["spline"]
        if isinstance(sp, M.NormalizationModel) and not isinstance(sp.model, M.CardinalSplineMixture):
            raise ValueError("NormalizationModel over anything but a CardinalSplineMixture is unsupported")
        return comp

    # ----------------------------------------------------------------------------------------
    def params_all(self):
        out = np.empty(self.n_params, dtype=np.float64)
        for lf in self.leaves:
            out[lf.offset:lf.offset + lf.size] = np.asarray(lf.node.p, dtype=np.float64).reshape(-1)
        return out

    def fit_index(self):
        idx = [np.arange(lf.offset, lf.offset + lf.size, dtype=np.int64) for lf in self.leaves if lf.node._fit]
        return np.concatenate(idx) if idx else np.zeros(0, dtype=np.int64)

    def signature(self):
        sig = []
        for comp in self.components:
            sp = comp["spline"]
            inner = sp.model if isinstance(sp, M.NormalizationModel) else sp
            grid = np.asarray(inner.xs, dtype=np.float64).reshape(-1)
            sig.append((type(sp).__name__, inner.p_val, float(grid[0]), float(grid[1] - grid[0]) if grid.size > 1 else 0.0, grid.size,
                        None if comp["shift"] is None else (comp["shift"].which_key, comp["shift"].p.size),
                        None if comp["stretch"] is None else (comp["stretch"].which_key, comp["stretch"].p.size),
                        getattr(sp, "which_key", None), getattr(sp, "size", None)))
        return tuple(sig)


def _key_array(metablock, key, n_rows, n_keys, what):
    try:
        col = metablock[key]
    except KeyError:
        raise ValueError(f"metablock has no column {key!r} needed by {what}")
    arr = np.ascontiguousarray(np.asarray(col).reshape(-1), dtype=np.int32)
    if arr.size != n_rows:
        raise ValueError(f"metablock[{key!r}] has {arr.size} entries for {n_rows} rows")
    if arr.size and (arr.min() < 0 or arr.max() >= n_keys):
        raise ValueError(f"metablock[{key!r}] indexes outside the {n_keys} parameters of {what}")
    return arr


class Plan:
    """Owner of one ``jb_plan`` (data block resident in HBM + compiled model)."""

    def __init__(self, spec, xs, ys, yivar, mask, metablock, coefficient=1.0, device=0, rows_per_group=0,
                 general_kernel=False, float32=False, no_strip=False):
        lib = _lib.load()
        self.lib = lib
        self.spec = spec
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        if xs.ndim != 2:
            raise ValueError("xs must be (rows, pixels)")
        E, P = xs.shape
        self.shape = (E, P)
        keep = [xs]
        f64p = lambda a: a.ctypes.data_as(_lib.c_f64p)
        desc = _lib.PlanDesc()
        desc.abi_version = _lib.JB_ABI_VERSION
        desc.device = int(device)
        desc.n_rows, desc.n_pix = E, P
        desc.xs = f64p(xs)
        if ys is not None:
            ys = np.ascontiguousarray(ys, dtype=np.float64)
            yivar = np.ascontiguousarray(yivar, dtype=np.float64)
            if ys.shape != xs.shape or yivar.shape != xs.shape:
                raise ValueError("ys / yivar must have the shape of xs")
            keep += [ys, yivar]
            desc.ys, desc.yivar = f64p(ys), f64p(yivar)
        if mask is not None:
            mask = np.ascontiguousarray(np.asarray(mask).astype(bool), dtype=np.uint8)
            if mask.shape != xs.shape:
                raise ValueError("mask must have the shape of xs")
            keep.append(mask)
            desc.mask = mask.ctypes.data_as(_lib.c_u8p)
        comps = (_lib.ComponentDesc * len(spec.components))()
        i32p = lambda a: a.ctypes.data_as(_lib.c_i32p)
        for cd, comp in zip(comps, spec.components):
            sp = comp["spline"]
            is_norm = isinstance(sp, M.NormalizationModel)
            inner = sp.model if is_norm else sp
            grid = np.asarray(inner.xs, dtype=np.float64).reshape(-1)
            if grid.size < 2:
                raise ValueError("a spline needs at least two knots")
            lf = spec.leaf_of(sp)
            cd.kind = _lib.JB_COMP_NORM_SPLINE if is_norm else _lib.JB_COMP_SPLINE
            cd.p_val = int(inner.p_val)
            cd.xp0 = float(grid[0])
            cd.dx = float(grid[1] - grid[0])            # model.py:974
            cd.n_knots = grid.size
            cd.ctrl_offset = lf.offset
            if is_norm:
                if lf.size != sp.size * grid.size:
                    raise ValueError("NormalizationModel.p must hold size * len(model.p) values")
                cd.n_ctrl_keys = sp.size
                k = _key_array(metablock, sp.which_key, E, sp.size, "NormalizationModel")
                keep.append(k)
                cd.ctrl_key = i32p(k)
            elif lf.size != grid.size:
                raise ValueError("CardinalSplineMixture.p must have the shape of its knot grid")
            cd.shift_offset = cd.stretch_offset = -1
            for name in ("shift", "stretch"):
                node = comp[name]
                if node is None:
                    continue
                nlf = spec.leaf_of(node)
                k = _key_array(metablock, node.which_key, E, nlf.size, type(node).__name__)
                keep.append(k)
                setattr(cd, name + "_offset", nlf.offset)
                setattr(cd, "n_" + name, nlf.size)
                setattr(cd, name + "_key", i32p(k))
        desc.n_components = len(spec.components)
        desc.components = comps
        desc.n_params = spec.n_params
        desc.coefficient = float(coefficient)
        desc.rows_per_group = int(rows_per_group)
        desc.flags = ((_lib.JB_PLAN_GENERAL_KERNEL if general_kernel else 0) | (_lib.JB_PLAN_FLOAT32 if float32 else 0) |
                      (_lib.JB_PLAN_NO_STRIP if no_strip else 0))
        handle = C.c_void_p()
        _lib.check(lib.jb_plan_create(C.byref(desc), C.byref(handle)))
        self.handle = handle
        self.n_fit = 0
        self._last_params = None      # what jb_plan_set_params last uploaded ...
        self._params_clean = False    # ... and whether the device copy still equals it (an evaluation scatters its fit vector)
        self._last_fit = None
        self._owner = None            # the Objective whose fit flags / fixed parameters the plan currently holds
        del keep

    def close(self):
        if getattr(self, "handle", None):
            self.lib.jb_plan_destroy(self.handle)
            self.handle = None

    __del__ = close

    # ----------------------------------------------------------------------------------------
    def set_params(self, params_all):
        p = np.ascontiguousarray(params_all, dtype=np.float64)
        if self._params_clean and self._last_params is not None and np.array_equal(p, self._last_params):
            return                     # nothing changed since the last upload: no copy, no re-sort
        _lib.check(self.lib.jb_plan_set_params(self.handle, p.ctypes.data_as(_lib.c_f64p), p.size))
        self._last_params, self._params_clean = p.copy(), True

    def set_fit(self, fit_index):
        idx = np.ascontiguousarray(fit_index, dtype=np.int64)
        if self._last_fit is not None and np.array_equal(idx, self._last_fit):
            return
        _lib.check(self.lib.jb_plan_set_fit(self.handle, idx.ctypes.data_as(_lib.c_i64p), idx.size))
        self.n_fit = idx.size
        self._last_fit = idx.copy()

    def sync_from_model(self):
        """Upload the tree's current parameter values and fit flags."""
        self.set_params(self.spec.params_all())
        self.set_fit(self.spec.fit_index())

    def eval(self, p_fit, want_grad=True):
        p = np.ascontiguousarray(p_fit, dtype=np.float64).reshape(-1)
        if p.size:
            self._params_clean = False
        loss = C.c_double()
        grad = np.empty(p.size, dtype=np.float64) if want_grad else None
        _lib.check(self.lib.jb_plan_eval(
            self.handle, p.ctypes.data_as(_lib.c_f64p), p.size, C.byref(loss),
            grad.ctypes.data_as(_lib.c_f64p) if want_grad else None))
        return loss.value, grad

    def forward(self, p_fit=()):
        p = np.ascontiguousarray(p_fit, dtype=np.float64).reshape(-1)
        if p.size:
            self._params_clean = False
        out = np.empty(self.shape, dtype=np.float64)
        _lib.check(self.lib.jb_plan_forward(self.handle, p.ctypes.data_as(_lib.c_f64p), p.size,
                                            out.ctypes.data_as(_lib.c_f64p)))
        return out

    def fisher(self, component):
        n = self.spec.leaf_of(self.spec.components[component]["shift"]).size
        out = np.empty(n, dtype=np.float64)
        _lib.check(self.lib.jb_plan_fisher(self.handle, component, out.ctypes.data_as(_lib.c_f64p), n))
        return out

    def curvature(self, component):
        """d^2 loss / d shift_k^2 for every shift key of ``component`` (jb_plan_curvature)."""
        n = self.spec.leaf_of(self.spec.components[component]["shift"]).size
        out = np.empty(n, dtype=np.float64)
        _lib.check(self.lib.jb_plan_curvature(self.handle, component, out.ctypes.data_as(_lib.c_f64p), n))
        return out

    def grid_search(self, component, grid):
        """(n_keys, M) losses of the rows of every shift key of ``component`` at M candidate shifts."""
        g = np.ascontiguousarray(grid, dtype=np.float64)
        n = self.spec.leaf_of(self.spec.components[component]["shift"]).size
        if g.ndim != 2 or g.shape[0] != n:
            raise ValueError(f"grid must be ({n}, M), got {g.shape}")
        out = np.empty(g.shape, dtype=np.float64)
        _lib.check(self.lib.jb_plan_grid_search(self.handle, component, g.ctypes.data_as(_lib.c_f64p), n, g.shape[1],
                                                out.ctypes.data_as(_lib.c_f64p)))
        return out

    def grad_all(self):
        out = np.empty(self.spec.n_params, dtype=np.float64)
        _lib.check(self.lib.jb_plan_grad_all(self.handle, out.ctypes.data_as(_lib.c_f64p), out.size))
        return out

    def info(self):
        info = _lib.PlanInfo()
        _lib.check(self.lib.jb_plan_info(self.handle, C.byref(info)))
        return {n: getattr(info, n) for n, _ in _lib.PlanInfo._fields_}

    def eval_device(self, d_p_fit_ptr, d_out_ptr, stream_ptr=0):
        """Enqueue one evaluation with device-resident in/out (see jb_plan_eval_device)."""
        self._params_clean = False
        _lib.check(self.lib.jb_plan_eval_device(self.handle, C.c_void_p(d_p_fit_ptr), C.c_void_p(d_out_ptr),
                                                C.c_void_p(stream_ptr)))

    def check(self):
        _lib.check(self.lib.jb_plan_check(self.handle))

    def stream(self):
        """The plan's cudaStream_t as an integer (for torch.cuda.ExternalStream / event ordering)."""
        out = C.c_void_p()
        _lib.check(self.lib.jb_plan_stream(self.handle, C.byref(out)))
        return out.value or 0

    def profile(self, enable=True):
        _lib.check(self.lib.jb_plan_profile(self.handle, int(bool(enable))))

    def profile_read(self):
        """(summed fused-kernel milliseconds, launches) since the last read."""
        ms, n = C.c_double(), C.c_int64()
        _lib.check(self.lib.jb_plan_profile_read(self.handle, C.byref(ms), C.byref(n)))
        return ms.value, n.value


class Batch:
    """Independent plans (chunks) evaluated with one launch per kernel (``jb_batch``)."""

    def __init__(self, plans):
        self.lib = _lib.load()
        self.plans = list(plans)
        for p in self.plans:
            p._params_clean = False       # batched evaluations scatter fit vectors into the plans' parameters
        arr = (C.c_void_p * len(self.plans))(*[p.handle for p in self.plans])
        handle = C.c_void_p()
        _lib.check(self.lib.jb_batch_create(arr, len(self.plans), C.byref(handle)))
        self.handle = handle

    def bind(self, d_p_fit_ptrs, d_out_ptrs):
        n = len(self.plans)
        pf = (C.c_void_p * n)(*d_p_fit_ptrs) if d_p_fit_ptrs is not None else None
        po = (C.c_void_p * n)(*d_out_ptrs)
        _lib.check(self.lib.jb_batch_bind(self.handle, pf, po))

    def eval_device(self, stream_ptr=0):
        _lib.check(self.lib.jb_batch_eval_device(self.handle, C.c_void_p(stream_ptr)))

    def eval_host(self, p_list, out_list=None):
        """All chunks with HOST buffers: ``p_list[i]`` the fit vector of plan i; returns (or fills)
        ``out_list[i]`` = [loss, gradient] arrays.  One H2D copy, one launch per kernel, one D2H copy."""
        n = len(self.plans)
        ps = [np.ascontiguousarray(p, dtype=np.float64).reshape(-1) for p in p_list]
        if out_list is None:
            out_list = [np.empty(p.size + 1, dtype=np.float64) for p in ps]
        pf = (C.c_void_p * n)(*[p.ctypes.data for p in ps])
        po = (C.c_void_p * n)(*[o.ctypes.data for o in out_list])
        _lib.check(self.lib.jb_batch_eval_host(self.handle, pf, po))
        return out_list

    def host_buffers(self):
        """The batch's page-locked host buffers as numpy views: ``(p_views, out_views)`` with
        ``p_views[i]`` the fit vector of plan i (write it in place) and ``out_views[i]`` its
        [loss, gradient] after ``eval_pinned`` (read in place: the next evaluation overwrites it).
        Valid until a plan's fit vector changes size; then ask again."""
        n = len(self.plans)
        hp, ho = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
        off_p, off_o = (C.c_int64 * (n + 1))(), (C.c_int64 * (n + 1))()
        _lib.check(self.lib.jb_batch_host_buffers(self.handle, C.byref(hp), C.byref(ho), off_p, off_o))
        flat_p = np.ctypeslib.as_array(hp, shape=(max(off_p[n], 1),))
        flat_o = np.ctypeslib.as_array(ho, shape=(off_o[n],))
        self._pinned = (flat_p, flat_o)
        return ([flat_p[off_p[i]:off_p[i + 1]] for i in range(n)],
                [flat_o[off_o[i]:off_o[i + 1]] for i in range(n)])

    def eval_pinned(self):
        """Evaluate every plan from / into the ``host_buffers`` views: one H2D copy, one launch per
        kernel, one D2H copy, one synchronisation -- and no staging copy on the host."""
        _lib.check(self.lib.jb_batch_eval_pinned(self.handle))

    def eval_pinned_begin(self):
        """Copy in, enqueue the evaluation and the copy back on the batch's stream; returns at once."""
        _lib.check(self.lib.jb_batch_eval_pinned_begin(self.handle))

    def eval_pinned_end(self):
        """Wait for ``eval_pinned_begin``; afterwards the ``host_buffers`` views hold the results."""
        _lib.check(self.lib.jb_batch_eval_pinned_end(self.handle))

    def soft_range(self, enable=True):
        """Out-of-range plans give a NaN loss in ``eval_host`` instead of failing the whole call."""
        _lib.check(self.lib.jb_batch_soft_range(self.handle, int(bool(enable))))

    def eval_device_checked(self, stream_ptr=0):
        """eval_device, then wait and settle knot-window overflows the way jb_plan_eval does."""
        _lib.check(self.lib.jb_batch_eval_device_checked(self.handle, C.c_void_p(stream_ptr)))

    def eval_host_begin(self, p_list):
        """Copy the fit vectors and enqueue the evaluation on the batch's own stream; returns at once."""
        n = len(self.plans)
        self._ps = [np.ascontiguousarray(p, dtype=np.float64).reshape(-1) for p in p_list]
        pf = (C.c_void_p * n)(*[p.ctypes.data for p in self._ps])
        _lib.check(self.lib.jb_batch_eval_host_begin(self.handle, pf))

    def eval_host_end(self, out_list=None):
        """Wait for ``eval_host_begin`` and return (or fill) the [loss, gradient] arrays."""
        n = len(self.plans)
        if out_list is None:
            out_list = [np.empty(p.size + 1, dtype=np.float64) for p in self._ps]
        po = (C.c_void_p * n)(*[o.ctypes.data for o in out_list])
        _lib.check(self.lib.jb_batch_eval_host_end(self.handle, po))
        return out_list

    def check(self):
        _lib.check(self.lib.jb_batch_check(self.handle))

    def launches(self):
        return int(self.lib.jb_batch_launches(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.jb_batch_destroy(self.handle)
            self.handle = None

    __del__ = close


class Group:
    """One chunk whose epochs are sharded over the GPUs of a box, one process per GPU (``jb_group``):
    the rank's plans evaluated with one launch per kernel, then ONE NCCL all-reduce of
    [loss, gradient of the template control points] enqueued behind them on the same stream.

    ``rank`` / ``world_size``: this process's place; ``nccl_id``: the 128 bytes of
    ``Group.unique_id()`` made on rank 0 and handed to every rank (``broadcast_unique_id`` does it
    through torch.distributed).  ``world_size`` 1 needs neither NCCL nor an id."""

    def __init__(self, plans, rank=0, world_size=1, nccl_id=None):
        self.lib = _lib.load()
        self.plans = list(plans)
        arr = (C.c_void_p * len(self.plans))(*[p.handle for p in self.plans])
        handle = C.c_void_p()
        idbuf = C.create_string_buffer(bytes(nccl_id), 128) if nccl_id is not None else None
        _lib.check(self.lib.jb_group_create(arr, len(self.plans), int(rank), int(world_size), idbuf, C.byref(handle)))
        self.handle = handle
        self.rank, self.world_size = int(rank), int(world_size)
        for p in self.plans:
            p._params_clean = False

    @staticmethod
    def unique_id():
        lib = _lib.load()
        buf = C.create_string_buffer(128)
        _lib.check(lib.jb_nccl_unique_id(buf))
        return buf.raw

    @staticmethod
    def broadcast_unique_id(rank, world_size, group=None):
        """The NCCL id of rank 0 on every rank, through torch.distributed (any backend)."""
        if world_size <= 1:
            return None
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(Group.unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0, group=group)
        return bytes(t.cpu().numpy().tobytes())

    def bind(self, d_p_fit_ptrs, d_out_ptrs):
        n = len(self.plans)
        pf = (C.c_void_p * n)(*d_p_fit_ptrs) if d_p_fit_ptrs is not None else None
        po = (C.c_void_p * n)(*d_out_ptrs)
        _lib.check(self.lib.jb_group_bind(self.handle, pf, po))

    def eval_device(self, stream_ptr=0):
        _lib.check(self.lib.jb_group_eval_device(self.handle, C.c_void_p(stream_ptr)))

    def eval_device_checked(self, stream_ptr=0):
        _lib.check(self.lib.jb_group_eval_device_checked(self.handle, C.c_void_p(stream_ptr)))

    def info(self):
        ns, nb, w = C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(self.lib.jb_group_info(self.handle, C.byref(ns), C.byref(nb), C.byref(w)))
        return {"n_shared": ns.value, "collective_bytes": nb.value, "world": w.value}

    def launches(self):
        return int(self.lib.jb_group_launches(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.jb_group_destroy(self.handle)
            self.handle = None

    __del__ = close


def build_plan(model, datablock, metablock, coefficient=1.0, device=0, rows_per_group=0, general_kernel=False,
               float32=False, no_strip=False):
    spec = TreeSpec(model)
    plan = Plan(spec, datablock["xs"], datablock["ys"], datablock["yivar"], datablock["mask"], metablock,
                coefficient=coefficient, device=device, rows_per_group=rows_per_group, general_kernel=general_kernel,
                float32=float32, no_strip=no_strip)
    plan.sync_from_model()
    return plan


# ------------------------------------------------------------------------------------------
# One resident copy of a data block serves every call made on it: Model.optimize stage after stage,
# grid_search, f_info, the curvature, get_loss_array, train_norm's forward (the reference re-bakes the
# data into a new XLA program per optimize() call, loss.py:47).  Between calls only parameter values
# and fit flags change (jb_plan_set_params / jb_plan_set_fit).
_PLAN_CACHE = []          # [(key, datablock, metablock, plan)], most recently used last
PLAN_CACHE_SIZE = 2


def clear_plan_cache():
    """Forget every cached plan; its HBM is freed as soon as nothing else refers to it."""
    del _PLAN_CACHE[:]


def get_plan(model, datablock, metablock, coefficient=1.0, device=0, unit_weights=None, float32=None):
    """The cached Plan of (data block, tree structure, data term), synchronised with the model's
    current parameter values and fit flags.  ``unit_weights``: a constant that replaces ``yivar``
    (L2Loss).  ``float32``: the float32 tier; None = what ``jabble.config`` says (``jax_enable_x64``
    False -> float32).  The plan belongs to the cache: do not close it."""
    if float32 is None:
        from . import config
        float32 = not config.jax_enable_x64
    spec = TreeSpec(model)
    key = (id(datablock), id(metablock), spec.signature(), float(coefficient), unit_weights, int(device), bool(float32))
    for i, (k, db, mb, plan) in enumerate(_PLAN_CACHE):
        if k == key and db is datablock and mb is metablock and plan.handle:
            _PLAN_CACHE.append(_PLAN_CACHE.pop(i))
            plan.spec = spec
            plan.sync_from_model()
            plan._owner = None
            return plan
    yivar = datablock["yivar"]
    if unit_weights is not None and datablock["ys"] is not None:
        yivar = np.full_like(np.asarray(datablock["ys"], dtype=np.float64), float(unit_weights))
    plan = Plan(spec, datablock["xs"], datablock["ys"], yivar, datablock["mask"], metablock,
                coefficient=coefficient, device=device, float32=bool(float32))
    plan.sync_from_model()
    _PLAN_CACHE.append((key, datablock, metablock, plan))
    while len(_PLAN_CACHE) > PLAN_CACHE_SIZE:
        _PLAN_CACHE.pop(0)          # an objective still using it keeps it alive; Plan.__del__ frees the HBM
    return plan


def _loss_key(loss):
    """The values a loss tree is made of (not its repr, which rounds the coefficient and omits constants)."""
    from . import loss as L
    terms = loss.loss_funcs if isinstance(loss, L.LossSequential) else [loss]
    out = []
    for t in terms:
        inds = getattr(t, "submodel_inds", None)
        out.append((type(t).__name__, float(getattr(t, "coefficient", 1.0)), float(getattr(t, "constant", 0.0)),
                    None if inds is None else tuple(np.asarray(inds).reshape(-1).tolist())))
    return tuple(out)


class Objective:
    """The (func, fprime) pair scipy receives in Model.optimize (jabble/model.py:226-232), backed by
    one fused evaluation per distinct parameter vector."""

    _cache = {}

    def __init__(self, loss, model, datablock, metablock, device=0):
        from . import loss as L
        terms = loss.loss_funcs if isinstance(loss, L.LossSequential) else [loss]
        data_terms = [t for t in terms if isinstance(t, (L.ChiSquare, L.L2Loss))]
        self.regs = [t for t in terms if isinstance(t, L.L2Reg)]          # L1Reg is an L2Reg subclass
        other = [t for t in terms if t not in data_terms and t not in self.regs]
        if other:
            raise NotImplementedError(f"{type(other[0]).__name__} is not a loss this package evaluates")
        if len(data_terms) != 1:
            raise NotImplementedError("exactly one data term (ChiSquare or L2Loss) per objective is supported")
        term = data_terms[0]
        # (ys - m)**2 == 0.5 * 2 * (ys - m)**2 : an L2Loss is a ChiSquare with unit weights x 2
        self.plan = get_plan(model, datablock, metablock, coefficient=term.coefficient, device=device,
                             unit_weights=2.0 if isinstance(term, L.L2Loss) else None)
        self._params_all = self.plan.spec.params_all()        # what this objective's evaluations start from
        self._fit_index = self.plan.spec.fit_index()
        self.plan._owner = self
        self.n_rows = len(datablock)
        for r in self.regs:
            if r.indices is None:
                r.ready_indices(model)
        self.n_evals = 0
        self.n_out_of_range = 0
        self._p = None
        self._f = None
        self._g = None
        self._f_inside = None                 # loss of the last evaluation that stayed inside the knot grids
        self._data_eval = self._plan_eval     # (loss, gradient) of the data term; LockstepEvaluator reroutes it

    def _plan_eval(self, p):
        if self.plan._owner is not self:      # another objective used the shared plan in between
            self.plan.set_params(self._params_all)
            self.plan.set_fit(self._fit_index)
            self.plan._owner = self
        return self.plan.eval(p)

    @classmethod
    def cached(cls, loss, model, datablock, metablock):
        """One objective per (data block, model, loss VALUES, fit state, stored parameters) between
        loss_all / grad_all calls: scipy's func and fprime then share one fused evaluation."""
        spec = TreeSpec(model)
        fit = tuple(bool(lf.node._fit) for lf in spec.leaves)
        from . import config
        key = (id(datablock), id(metablock), id(model), _loss_key(loss), spec.signature(), fit, config.jax_enable_x64)
        obj = cls._cache.get(key)
        if obj is not None and obj._db is datablock and np.array_equal(obj._params_all[~obj._fit_mask], spec.params_all()[~obj._fit_mask]):
            return obj
        cls._cache.clear()              # one at a time
        obj = cls(loss, model, datablock, metablock)
        obj._db = datablock
        obj._fit_mask = np.zeros(spec.n_params, dtype=bool)
        obj._fit_mask[spec.fit_index()] = True
        cls._cache[key] = obj
        return obj

    def _eval(self, p):
        p = np.asarray(p, dtype=np.float64).reshape(-1)
        if self._p is None or p.shape != self._p.shape or not np.array_equal(p, self._p):
            try:
                self._f, self._g = self._data_eval(p)
                self._f_inside = self._f
            except ValueError as exc:
                # A trial point of a line search that carries an unmasked pixel off a knot grid (a first
                # L-BFGS step has unit length: metres per second for an RV).  The reference's gather
                # clamps there (model.py:985) and scipy sees a large finite loss and shortens the step;
                # here the kernel refuses the point, so the line search is told the same thing: a loss
                # far above the last valid one, no slope.  The first evaluation of a fit still raises.
                if "outside the knot grid" not in str(exc) or self._f_inside is None:
                    raise
                self.n_out_of_range += 1
                self._f, self._g = 1e8 * max(abs(self._f_inside), 1.0), np.zeros_like(p)
            for r in self.regs:
                # the reference's loss_all evaluates a regulariser inside the vmap over rows
                # (loss.py:61-74): its value and gradient are counted once per row of the block
                rv, rg = r.value_and_grad(p)
                self._f += self.n_rows * rv
                self._g = self._g + self.n_rows * rg
            self._p = p.copy()
            self.n_evals += 1
        return self._f, self._g

    def value(self, p):
        return self._eval(p)[0]

    def value_and_gradient(self, p):
        """(f, g) in one call: what scipy's ``fmin_l_bfgs_b(func, x0)`` takes when ``fprime`` is None --
        one Python callback per evaluation instead of two (same numbers, same iterates)."""
        f, g = self._eval(p)
        return f, g.copy()

    def gradient(self, p):
        return self._eval(p)[1].copy()


_LBFGS_STATUS = {4: "CONVERGENCE", 5: "STOP", 7: "ERROR", 8: "ABNORMAL"}      # scipy/optimize/_lbfgsb_py.py status_messages
_LBFGS_TASK = {0: "", 401: "NORM OF PROJECTED GRADIENT <= PGTOL", 402: "RELATIVE REDUCTION OF F <= FACTR*EPSMCH",
               502: "TOTAL NO. OF F,G EVALUATIONS EXCEEDS LIMIT", 504: "TOTAL NO. OF ITERATIONS REACHED LIMIT",
               799: "STARTING POINT OUTSIDE THE KNOT GRID"}
_LBFGS_DEFAULTS = {"m": 10, "factr": 1e7, "pgtol": 1e-5, "maxfun": 15000, "maxiter": 15000, "maxls": 20}


def lbfgs_options_supported(options):
    """Can the device optimiser take these ``fmin_l_bfgs_b`` keyword options?  (No bounds -- the
    reference never passes any, model.py:226-232 --, no callback, analytic gradient.)"""
    for k, v in options.items():
        if k in _LBFGS_DEFAULTS:
            continue
        if k in ("bounds", "callback") and v is None:
            continue
        if k in ("approx_grad",) and not v:
            continue
        if k in ("args",) and tuple(v) == ():
            continue
        if k in ("epsilon",):
            continue
        return False
    return 1 <= int(options.get("m", 10)) <= 32 and int(options.get("maxls", 20)) > 0


def lbfgs_run(plans, x0s, options=None, regs=None):
    """``jb_lbfgs`` on plans whose fit state and fixed parameters are set: L-BFGS-B from ``x0s[i]`` for
    every plan at once.  ``regs[i]``: list of (kind, weight, constant, indices) parameter-only terms
    (kind 0: weight * 0.5 * sum (x[indices] - constant)^2, 1: weight * sum |x[indices] - constant|).
    Returns ([(x, f, d)], info): scipy's ``fmin_l_bfgs_b`` triple per plan, and {rounds, results} with
    the number of batched evaluations and the raw ``jb_lbfgs_result`` fields."""
    options = dict(options or {})
    if not lbfgs_options_supported(options):
        raise NotImplementedError(f"options {sorted(options)} need scipy's optimiser")
    lib = _lib.load()
    opt = _lib.LbfgsOptions()
    for k, v in _LBFGS_DEFAULTS.items():
        val = options.get(k, v)
        setattr(opt, k, float(val) if k in ("factr", "pgtol") else int(val))
    n = len(plans)
    handle = C.c_void_p()
    arr = (C.c_void_p * n)(*[p.handle for p in plans])
    import time
    t_create = time.perf_counter()
    _lib.check(lib.jb_lbfgs_create(arr, n, C.byref(opt), C.byref(handle)))
    t_create = time.perf_counter() - t_create
    t_run = 0.0
    try:
        for i, terms in enumerate(regs or []):
            for kind, weight, constant, indices in terms:
                idx = np.ascontiguousarray(np.asarray(indices, dtype=np.int64).reshape(-1))
                _lib.check(lib.jb_lbfgs_add_reg(handle, i, int(kind), float(weight), float(constant),
                                                idx.ctypes.data_as(_lib.c_i64p), idx.size))
        x0 = [np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1)) for x in x0s]
        for p, x in zip(plans, x0):
            if x.size != p.n_fit:
                raise ValueError(f"x0 has {x.size} entries, the model has {p.n_fit} parameters in fit mode")
        xs = [np.empty_like(x) for x in x0]
        gs = [np.empty_like(x) for x in x0]
        res = (_lib.LbfgsResult * n)()
        as_ptrs = lambda arrs: (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        t_run = time.perf_counter()
        _lib.check(lib.jb_lbfgs_run(handle, as_ptrs(x0), as_ptrs(xs), as_ptrs(gs), res))
        t_run = time.perf_counter() - t_run
        rounds = int(lib.jb_lbfgs_rounds(handle))
    finally:
        t_destroy = time.perf_counter()
        lib.jb_lbfgs_destroy(handle)
        t_destroy = time.perf_counter() - t_destroy
    out, raw = [], []
    for x, g, r in zip(xs, gs, res):
        d = {"grad": g, "task": _LBFGS_STATUS.get(int(r.task), "?") + ": " + _LBFGS_TASK.get(int(r.task_msg), ""),
             "funcalls": int(r.funcalls), "nit": int(r.nit), "warnflag": int(r.warnflag)}
        out.append((x, float(r.f), d))
        raw.append({k: getattr(r, k) for k, _ in _lib.LbfgsResult._fields_})
    return out, {"rounds": rounds, "results": raw, "create_s": t_create, "run_s": t_run, "destroy_s": t_destroy}


def minimize_lbfgs(objectives, x0s, options=None):
    """``scipy.optimize.fmin_l_bfgs_b(func=objective.value_and_gradient, x0=x0, **options)`` for every
    objective at once, on the device (``jb_lbfgs``, include/jabble_b200.h): one batched evaluation per
    round for all chunks still iterating, optimiser state and vectors in HBM.  Returns scipy's
    ``(x, f, d)`` per objective, ``d`` = {grad, task, funcalls, nit, warnflag}."""
    from . import loss as L
    for o in objectives:          # the plan carries this objective's fixed parameters and fit state
        if o.plan._owner is not o:
            o.plan.set_params(o._params_all)
            o.plan.set_fit(o._fit_index)
            o.plan._owner = o
    # the reference's loss_all evaluates a regulariser inside the vmap over rows (loss.py:61-74): its
    # value and gradient count once per row of the block
    regs = [[(1 if isinstance(r, L.L1Reg) else 0, float(o.n_rows) * float(r.coefficient), float(r.constant), r.indices)
             for r in o.regs] for o in objectives]
    out, info = lbfgs_run([o.plan for o in objectives], x0s, options, regs)
    for o, r in zip(objectives, info["results"]):
        o.n_evals += int(r["funcalls"])
        o.n_out_of_range += int(r["n_out_of_range"])
        o.lbfgs_rounds = info["rounds"]
        if r["task"] == 7:
            raise ValueError("an unmasked warped pixel lies outside the knot grid of a spline component (at the starting point)")
    return out


class LockstepEvaluator:
    """Evaluations of several independent objectives (chunks / orders, each driven by its own scipy
    L-BFGS-B in its own thread) sent to the GPU together: a thread that asks for (loss, gradient)
    waits until every thread still optimising has asked, then ONE batched launch per kernel serves
    them all (``jb_batch_eval_host``).  Every optimiser sees exactly the numbers ``jb_plan_eval``
    would have given it (the batched kernels are bit-identical to the per-plan ones), so the
    iterates are those of ``Model.optimize`` run chunk by chunk.
    """

    def __init__(self, objectives):
        import threading
        self.objs = list(objectives)
        self.cv = threading.Condition()
        self.pending, self.results = {}, {}
        self.active = set(range(len(self.objs)))
        self.batches = {}
        self.n_flushes = 0
        for i, o in enumerate(self.objs):
            o._data_eval = (lambda p, i=i: self.request(i, p))

    def request(self, i, p):
        with self.cv:
            self.pending[i] = np.array(p, dtype=np.float64).reshape(-1)
            if set(self.pending) >= self.active:
                self._flush()
            else:
                self.cv.wait_for(lambda: i in self.results)
            res = self.results.pop(i)
        if isinstance(res, BaseException):
            raise res
        return res

    def done(self, i):
        """Thread i has stopped optimising: the others no longer wait for it."""
        with self.cv:
            self.active.discard(i)
            if self.pending and set(self.pending) >= self.active:
                self._flush()

    def _flush(self):          # called with the lock held, by the thread whose request completed the set
        idx = tuple(sorted(self.pending))
        try:
            entry = self.batches.get(idx)
            if entry is None:
                batch = Batch([self.objs[i].plan for i in idx])
                batch.soft_range(True)
                entry = self.batches[idx] = (batch,) + batch.host_buffers()
            batch, p_views, outs = entry
            for v, i in zip(p_views, idx):
                v[:] = self.pending[i]
            batch.eval_pinned()
            for i, o in zip(idx, outs):
                if np.isfinite(o[0]):
                    self.results[i] = (float(o[0]), o[1:].copy())
                else:      # this chunk's trial left a knot grid: its optimiser hears what jb_plan_eval would say
                    self.results[i] = ValueError("an unmasked warped pixel lies outside the knot grid of a spline component")
        except BaseException as exc:       # every waiting optimiser gets the error, nobody is left waiting
            for i in idx:
                self.results[i] = exc
        self.pending.clear()
        self.n_flushes += 1
        self.cv.notify_all()

    def close(self):
        for entry in self.batches.values():
            entry[0].close()
        self.batches.clear()
        for o in self.objs:
            o._data_eval = o._plan_eval


def _meta_value(meta, key):
    try:
        v = meta[key]
    except (KeyError, TypeError):
        raise ValueError(f"meta has no entry {key!r}")
    return int(np.asarray(v).reshape(-1)[0])


def evaluate_row(model, p, x, meta):
    """``model(p, x, meta)`` for one row (jabble/model.py:138-160, 456-462, 730-736)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    p = np.asarray(p, dtype=np.float64).reshape(-1)
    if isinstance(model, M.EpochSpecificModel):
        # a bare warp has no spline to evaluate: one subtraction / multiplication per pixel on the
        # host (jabble/model.py:927, 946); not a hot-path operation
        val = (p if len(p) else model.p)[_meta_value(meta, model.which_key)]
        return x - val if isinstance(model, M.ShiftingModel) else val * x
    spec = TreeSpec(model)
    keys = {}
    for lf in spec.leaves:
        k = getattr(lf.node, "which_key", None)
        if k is not None:
            keys[k] = np.array([_meta_value(meta, k)], dtype=np.int32)
    plan = Plan(spec, x[None, :], None, None, None, keys)
    try:
        plan.sync_from_model()
        if len(p) not in (0, plan.n_fit):
            raise ValueError(f"p has {len(p)} entries, the model has {plan.n_fit} parameters in fit mode")
        return plan.forward(p)[0]
    finally:
        plan.close()


def shift_curvature(model, shift_model, datablock, metablock, coefficient=1.0, device=0):
    """Hessian diagonal of the loss w.r.t. a ShiftingModel of the tree (quickplay.get_RV_error_2d)."""
    spec = TreeSpec(model)
    comp = None
    for i, c in enumerate(spec.components):
        if c["shift"] is shift_model:
            comp = i
    if comp is None:
        raise NotImplementedError("the curvature runs on the CUDA path for the ShiftingModel of an additive component only")
    return get_plan(model, datablock, metablock, coefficient=coefficient, device=device, float32=False).curvature(comp)


def grid_search(model, shift_model, datablock, metablock, grid, coefficient=1.0, device=0):
    """EpochSpecificModel.grid_search (jabble/model.py:738-786) for a ShiftingModel of the tree."""
    spec = TreeSpec(model)
    comp = None
    for i, c in enumerate(spec.components):
        if c["shift"] is shift_model:
            comp = i
    if comp is None:
        raise NotImplementedError("grid_search runs on the CUDA path for the ShiftingModel of an additive component only")
    return get_plan(model, datablock, metablock, coefficient=coefficient, device=device, float32=False).grid_search(comp, grid)


def fisher_information(model, rv_model, datablock, metablock, device=0):
    """EpochSpecificModel.f_info (jabble/model.py:857-901) for a ShiftingModel of the tree."""
    spec = TreeSpec(model)
    comp = None
    for i, c in enumerate(spec.components):
        if c["shift"] is rv_model:
            comp = i
    if comp is None:
        raise ValueError("f_info: the model is not the ShiftingModel of any additive component of the tree")
    return get_plan(model, datablock, metablock, device=device).fisher(comp)
