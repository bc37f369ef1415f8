#!/usr/bin/env python
"""Benchmark of the fused loss+gradient+Fisher hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port on host cores

Workload (BASELINE.json configs[3]): synthetic HARPS-like chunks of
2,000 epochs x 4,096 pixels, WobbleModel tree (stellar + telluric cubic splines), shifts, both
templates and airmass in fit mode; 64 chunks resident per GPU (13.1 GB of float64 inputs, far larger
than the 126 MB L2, so every evaluation streams from HBM).  One "step" = one fused
loss+gradient+Fisher evaluation of every chunk.  Chunks are independent fits: ranks share nothing
and there is no data-path collective (weak scaling: 64 chunks per GPU at every N).

Printed JSON (one line, rank 0): see the bench contract in the task statement; `value` is G
epoch-pixels/s with parameters and outputs device-resident, `e2e` the same metric through the
host-buffer C-ABI call (jb_plan_eval: p_fit H2D, loss+gradient D2H, every chunk, every step).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "loss+grad evals/s (G epoch-pixels/s) and % HBM peak at 1/2/4/8 B200"
UNIT = "G epoch-pixels/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chunks", type=int, default=64, help="chunks per GPU")
    ap.add_argument("--float32", action="store_true",
                    help="the float32 tier (JB_PLAN_FLOAT32: float records, k_fused32; 13 B/pixel) instead of float64")
    ap.add_argument("--epochs", type=int, default=None, help="default: 2000 (chunks workload), 50000 (epochs workload)")
    ap.add_argument("--pixels", type=int, default=None, help="default: 4096 (chunks workload), 8192 (epochs workload)")
    ap.add_argument("--p-val", type=int, default=3)
    ap.add_argument("--rows-per-group", type=int, default=0)
    ap.add_argument("--e2e", default="pinned", choices=["pinned", "batch", "threads"],
                    help="host-buffer path of the end-to-end leg (chunks workload)")
    ap.add_argument("--e2e-parts", type=int, default=2, help="sub-batches in flight in the pinned end-to-end leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true",
                    help="skip the strong-scaling leg (64 chunks in total) and the epoch-sharded leg (configs[4])")
    ap.add_argument("--es-epochs", type=int, default=50000, help="epochs of the epoch-sharded leg's single chunk")
    ap.add_argument("--es-pixels", type=int, default=8192)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline time budget")
    ap.add_argument("--workload", default="chunks", choices=["chunks", "epochs"],
                    help="chunks: BASELINE configs[3] (default, the metric's config); epochs: configs[4], one chunk "
                         "of --epochs x --pixels sharded by epoch over the ranks with an NCCL all-reduce of the "
                         "loss and control-point gradients")
    args = ap.parse_args()
    big = args.workload == "epochs"          # BASELINE.json configs[4]: one chunk of 50000 epochs x 8192 px
    if args.epochs is None:
        args.epochs = 50000 if big else 2000
    if args.pixels is None:
        args.pixels = 8192 if big else 4096
    return args


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ reference arm
def cpu_baseline(args, budget_s, nthreads=0):
    """The oracle port (oracle/jb_oracle.c: plain-C restatement of the reference's arithmetic,
    OpenMP over rows) on a bounded sample of the same workload: whole rows of chunk 0."""
    from oracle import c_oracle
    from wobble_jax_b200 import synth
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which is not what the
    # CPU arm is about)
    cores = len(os.sched_getaffinity(0)) if nthreads <= 0 else nthreads
    rows = min(args.epochs, max(8 * cores, 64))
    ch = synth.make_chunk(0, n_epochs=rows, n_pix=args.pixels, p_val=args.p_val, as_block=True)
    model = synth.fit_all(ch.model)
    p = np.asarray(model.get_parameters())
    t0 = time.perf_counter()
    c_oracle.evaluate(p, ch.datablock, ch.metablock, model, nthreads=cores)
    t1 = time.perf_counter() - t0
    # grow the sample towards the time budget (bounded by one full chunk)
    want = int(min(args.epochs, max(rows, rows * (budget_s / 4.0) / max(t1, 1e-6))))
    if want > rows * 1.5:
        rows = want
        ch = synth.make_chunk(0, n_epochs=rows, n_pix=args.pixels, p_val=args.p_val, as_block=True)
        model = synth.fit_all(ch.model)
        p = np.asarray(model.get_parameters())
        c_oracle.evaluate(p, ch.datablock, ch.metablock, model, nthreads=cores)   # warm-up
    return ch, model, p, rows, cores


def run_reference(args, rank):
    if rank != 0:
        return
    from oracle import c_oracle
    ch, model, p, rows, cores = cpu_baseline(args, args.cpu_seconds)
    for _ in range(max(args.warmup, 0)):
        c_oracle.evaluate(p, ch.datablock, ch.metablock, model, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.evaluate(p, ch.datablock, ch.metablock, model, nthreads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    px = rows * args.pixels
    value = px / dt / 1e9
    sample = (f"{rows} epochs x {args.pixels} px of chunk 0 per step (1/{args.epochs * args.chunks / rows:.0f} of the "
              f"{args.chunks}-chunk workload), fused loss+grad+Fisher, oracle/jb_oracle.c, OpenMP {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.float32 else "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args):
    return {"workload": f"synthetic {args.epochs} epochs x {args.pixels} px x {args.chunks} chunks per GPU "
                        f"(BASELINE.json configs[3]), WobbleModel p_val={args.p_val}, shifts+templates+airmass in fit mode",
            "epochs": args.epochs, "pixels": args.pixels, "chunks_per_gpu": args.chunks,
            "parallelism": "chunk-sharded, no collective",
            "cache": "inputs 13.1 GB per GPU >> 126 MB L2; parameters perturbed every step"}


# ------------------------------------------------------------------------------------ own arm
def run_b200(args, rank, world, local_rank):
    import torch
    import wobble_jax_b200  # noqa: F401
    from wobble_jax_b200 import engine, synth

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- build: generate chunks on host threads, upload each into its plan ------------------
    t_setup = time.perf_counter()

    # rows per task: a LONE plan sizes its tasks to one wave of the GPU; in a batched launch of many
    # chunks there are dozens of waves anyway and long tasks amortise their set-up best (the hint)
    rows_per_group = args.rows_per_group or (32 if args.chunks > 1 else 0)

    def make(c):
        # generated, canonicalised and uploaded on a host thread of its own (ctypes drops the GIL)
        torch.cuda.set_device(local_rank)
        if gomp is not None:
            gomp.omp_set_num_threads(omp_inner)
        t0 = time.perf_counter()
        ch = synth.make_chunk(rank * args.chunks + c, n_epochs=args.epochs, n_pix=args.pixels,
                              p_val=args.p_val, as_block=True)
        model = synth.fit_all(ch.model)
        t1 = time.perf_counter()
        plan = engine.Plan(engine.TreeSpec(model), ch.datablock["xs"], ch.datablock["ys"], ch.datablock["yivar"],
                           ch.datablock["mask"], ch.metablock, device=local_rank,
                           rows_per_group=rows_per_group, float32=args.float32)
        plan.sync_from_model()
        setup_split[0] += t1 - t0                     # synthetic data (numpy, host): the bench's own cost
        setup_split[1] += time.perf_counter() - t1    # jb_plan_create: canonical block, upload, sort, strip layout
        return plan, np.asarray(model.get_parameters()).copy(), model

    plans, p_host, models = [], [], []
    setup_split = [0.0, 0.0]
    # one host thread per core of this rank, each generating and uploading its chunks; the library's own
    # OpenMP loop (row canonicalisation in jb_plan_create) is kept to one thread meanwhile, or the
    # worker threads x OpenMP threads oversubscribe the cores
    cores_here = max(1, len(os.sched_getaffinity(0)) // max(world, 1))
    nworkers = max(1, min(cores_here, args.chunks))
    try:
        import ctypes
        gomp = ctypes.CDLL("libgomp.so.1")          # (the ICV is per calling thread: every worker sets its own)
    except OSError:
        gomp = None
    omp_inner = max(1, cores_here // nworkers)
    with ThreadPoolExecutor(nworkers) as pool:
        for plan, p, model in pool.map(make, range(args.chunks)):
            plans.append(plan)
            p_host.append(p)
            models.append(model)
    info = plans[0].info()
    n_fit = info["n_fit"]
    setup_s = time.perf_counter() - t_setup

    # device-resident parameters and outputs: one flat buffer each, a view per chunk
    offs = np.concatenate([[0], np.cumsum([len(p) for p in p_host])])
    P_all = torch.from_numpy(np.concatenate(p_host)).cuda()
    O_all = torch.empty(int(offs[-1]) + len(plans), dtype=torch.float64, device="cuda")
    d_p = [P_all[int(offs[i]):int(offs[i + 1])] for i in range(len(plans))]
    d_out = [O_all[int(offs[i]) + i:int(offs[i + 1]) + i + 1] for i in range(len(plans))]
    # a real (non-default) stream: the C ABI reads a NULL stream as "the batch's own stream"
    main = torch.cuda.Stream()
    torch.cuda.set_stream(main)
    batch = engine.Batch(plans)            # every kernel of the evaluation launched once for all chunks
    batch.bind([t.data_ptr() for t in d_p], [t.data_ptr() for t in d_out])

    def step_device():
        P_all.add_(1e-13)                  # new parameter vectors every step
        batch.eval_device(main.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    batch.check()
    launches0 = batch.launches()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the library brackets the fused kernel of every batched launch with CUDA events on the launch
    # stream (jb_plan_profile on the batch's first plan): its duration inside the timed steps
    plans[0].profile(True)
    plans[0].profile_read()
    barrier()
    e0.record(main)
    for _ in range(args.steps):
        step_device()
    e1.record(main)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    fused_batch_ms, fused_batch_n = plans[0].profile_read()
    plans[0].profile(False)
    fused_batch_avg_ms = fused_batch_ms / max(fused_batch_n, 1)
    info = plans[0].info()                 # now with the kernel variant the evaluations used
    clocks = sampler.stop()
    launches = batch.launches() - launches0
    # outside the timed region: the same launch timed one at a time (spread of the kernel's duration)
    plans[0].profile(True)
    plans[0].profile_read()
    each = []
    for _ in range(12):
        step_device()
        ms, nl = plans[0].profile_read()
        each.append(ms / max(nl, 1))
    plans[0].profile(False)
    batch.check()
    losses = [float(o[0]) for o in d_out]
    assert all(np.isfinite(losses)), "non-finite loss"

    # ---- the fused kernel alone, ONE chunk per launch: what Model.optimize runs for a single order.
    # A plan made for that (no batching hint: it sizes its tasks to fill the GPU by itself), chunk 0.
    lone, lone_p = None, None
    try:
        ch0 = synth.make_chunk(rank * args.chunks, n_epochs=args.epochs, n_pix=args.pixels, p_val=args.p_val, as_block=True)
        m0 = synth.fit_all(ch0.model)
        lone = engine.Plan(engine.TreeSpec(m0), ch0.datablock["xs"], ch0.datablock["ys"], ch0.datablock["yivar"],
                           ch0.datablock["mask"], ch0.metablock, device=local_rank, rows_per_group=args.rows_per_group,
                           float32=args.float32)
        lone.sync_from_model()
        lone_p = torch.from_numpy(np.asarray(m0.get_parameters())).cuda()
        lone_out = torch.empty(lone_p.numel() + 1, dtype=torch.float64, device="cuda")
        del ch0
        for _ in range(3):
            lone.eval_device(lone_p.data_ptr(), lone_out.data_ptr())
        torch.cuda.synchronize()
        lone.profile(True); lone.profile_read()
        t_l0 = time.perf_counter()
        n_lone = max(5, min(args.steps, 20))
        for _ in range(n_lone):
            lone_p.add_(1e-13)
            lone.eval_device(lone_p.data_ptr(), lone_out.data_ptr())
        torch.cuda.synchronize()
        lone_eval_ms = (time.perf_counter() - t_l0) * 1e3 / n_lone
        fused_ms, fused_n = lone.profile_read()
        lone.profile(False)
        lone.check()
        lone_info = lone.info()
        fused_avg_ms = fused_ms / max(fused_n, 1)
    finally:
        if lone is not None:
            lone.close()

    # ---- end to end through the host-buffer C-ABI call ---------------------------------------
    # Host fit vectors in, host [loss, gradient] out, the copies inside the calls and so inside the
    # timed region.  With host cores to spare every chunk is driven from its own thread
    # (jb_plan_eval: what a Pool of scipy optimisers does); when the ranks of a node share few cores
    # one thread hands all chunks over in one call (jb_batch_eval_host).
    grads = [np.empty(n_fit) for _ in plans]
    cores_per_rank = max(1, len(os.sched_getaffinity(0)) // max(world, 1))
    e2e_path = args.e2e

    def eval_host(i):
        p_host[i] += 1e-13
        f, g = plans[i].eval(p_host[i])
        grads[i] = g
        return f

    outs = [np.empty(n_fit + 1) for _ in plans]

    def eval_host_batch():
        for p in p_host:
            p += 1e-13
        batch.eval_host(p_host, outs)
        return [float(o[0]) for o in outs]

    if e2e_path == "pinned":
        # the batch's own page-locked host buffers: the fit vectors are written there (as the optimiser
        # threads of model.optimize_many do), one H2D copy, the evaluation, one D2H copy, read in place.
        # The chunks go as two half-batches on two streams (begin, begin, end, end): the copies of one
        # half overlap the kernels of the other.
        nparts = max(1, min(args.e2e_parts, len(plans)))
        cuts = [len(plans) * i // nparts for i in range(nparts + 1)]
        halves = [engine.Batch(plans[cuts[i]:cuts[i + 1]]) for i in range(nparts)]
        pins = [hb.host_buffers() for hb in halves]
        pin_out = [o for _, outs_ in pins for o in outs_]
        for v, p in zip([v for vs, _ in pins for v in vs], p_host):
            v[:] = p

    def eval_host_pinned():
        for hb in halves:
            hb._pinned[0][::61] += 1e-13      # every step's inputs differ (a stride through all the fit vectors,
                                              # which lie back to back in the pinned buffer); all of it is copied
        for hb in halves:
            hb.eval_pinned_begin()
        for hb in halves:
            hb.eval_pinned_end()
        return [float(o[0]) for o in pin_out]

    with ThreadPoolExecutor(min(16, args.chunks)) as pool:
        def e2e_step():
            if e2e_path == "threads":
                return list(pool.map(eval_host, range(len(plans))))
            return eval_host_pinned() if e2e_path == "pinned" else eval_host_batch()
        for _ in range(min(args.warmup, 2)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            host_losses = e2e_step()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    assert all(np.isfinite(host_losses))

    # ---- strong scaling of the same workload: 64 chunks IN TOTAL (BASELINE.json configs[3]) -------
    strong = None
    total_chunks = 64
    if not args.no_extra_legs and total_chunks % world == 0 and total_chunks // world <= len(plans):
        per_gpu = total_chunks // world
        if per_gpu == len(plans):
            strong_ms = dev_ms / args.steps
        else:
            sub = engine.Batch(plans[:per_gpu])
            sub.bind([t.data_ptr() for t in d_p[:per_gpu]], [t.data_ptr() for t in d_out[:per_gpu]])

            def step_sub():
                P_all.add_(1e-13)
                sub.eval_device(main.cuda_stream)
            for _ in range(args.warmup):
                step_sub()
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(main)
            for _ in range(args.steps):
                step_sub()
            s1.record(main)
            barrier()
            strong_ms = s0.elapsed_time(s1) / args.steps
            sub.check()
            sub.close()
        tt = torch.tensor([strong_ms], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        strong_ms = float(tt[0])
        strong = {"workload": f"synthetic {args.epochs} epochs x {args.pixels} px x {total_chunks} chunks IN TOTAL (BASELINE.json configs[3]), "
                              f"{per_gpu} per GPU, chunk-sharded, no collective",
                  "value": args.epochs * args.pixels * total_chunks / (strong_ms * 1e-3) / 1e9, "unit": UNIT,
                  "ms_per_step": strong_ms, "chunks_per_gpu": per_gpu, "scaling": "strong",
                  "hbm_frac_per_gpu": info["algorithmic_bytes"] * per_gpu / (strong_ms * 1e-3) / 1e9 / measured_peaks()[0]}

    # ---- reduce over ranks (max time) ---------------------------------------------------------
    times = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(times[0]), float(times[1])
    px_step = args.epochs * args.pixels * args.chunks * world
    value = px_step * args.steps / (dev_ms * 1e-3) / 1e9
    e2e_value = px_step * args.steps / (e2e_ms * 1e-3) / 1e9
    peak, peak_src = measured_peaks()
    bytes_per_launch = info["algorithmic_bytes"] * args.chunks       # one launch serves every chunk of the GPU
    achieved = bytes_per_launch / (fused_batch_avg_ms * 1e-3) / 1e9
    traffic = None     # DRAM bytes per launch of the same kernel from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tpath) and (args.epochs, args.pixels, args.p_val) == (2000, 4096, 3) and not args.float32:
        with open(tpath) as f:
            traffic = json.load(f)["traffic_bytes_per_launch"]

    # the other ceiling of this path: the FP64 pipe, measured on this device
    import ctypes
    from wobble_jax_b200 import _lib
    tf = ctypes.c_double()
    _lib.check(_lib.load().jb_double_peak(local_rank, ctypes.byref(tf)))
    # FP64 instructions per pixel of the fused kernel (SASS of the shipped library, profiles/r2_sass_k_strip.txt):
    # pixel code + its share of the transposed row sums
    fp64_instr_px = 50.5 if info["fused_variant"] >= 128 else 68.0
    px_per_s_kernel = args.epochs * args.pixels * args.chunks / (fused_batch_avg_ms * 1e-3)
    fp64_info = {"peak_tflops_measured": tf.value,
                 "peak_warp_instr_per_s": tf.value * 1e12 / 64.0,            # one DFMA warp instruction = 32 lanes x 2 flop
                 "kernel_fp64_instr_per_pixel": fp64_instr_px,
                 "kernel_fp64_pipe_frac": fp64_instr_px * px_per_s_kernel / 32.0 / (tf.value * 1e12 / 64.0),
                 "note": "fraction of the measured FP64 issue rate the fused kernel's FP64 instructions take (any FP64 "
                         "instruction holds the pipe like a DFMA); beside roofline.frac = the HBM side"}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.float32 else "f64", "data": "synthetic",
        "config": workload_config(args),
        "evals_per_s": args.chunks * world * args.steps / (dev_ms * 1e-3),
        "hbm_frac_whole_step": (info["algorithmic_bytes"] * args.chunks * args.steps / (dev_ms * 1e-3) / 1e9) / peak,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(8 * n_fit * args.chunks * world),
                "d2h_bytes_per_step": int(8 * (n_fit + 1) * args.chunks * world),
                "ms_per_step": e2e_ms / args.steps,
                "api": {"threads": "jb_plan_eval (host p_fit in, host loss+gradient out), 16 host threads over the chunks",
                        "batch": "jb_batch_eval_host (host fit vectors in, host loss+gradient out, all chunks of the rank in one call, one host thread)",
                        "pinned": "jb_batch_host_buffers + jb_batch_eval_pinned_begin/_end (fit vectors written into / [loss, gradient] "
                                  "read from the batch's page-locked host buffers; the chunks as two half-batches on two streams, "
                                  "each one H2D copy, 6 launches, one D2H copy, one sync per step)"}[e2e_path],
                "host_cores_per_rank": cores_per_rank},
        "device_api": "jb_batch_eval_device: 6 launches per step for all chunks of the GPU",
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic * args.chunks if traffic else None,
                     "traffic_source": "profiles/r2_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per chunk of the committed "
                                       "ncu --set full capture of this kernel, times the chunks of the launch (not measured in this run)" if traffic else None,
                     "kernel": (f"k_strip<2,{args.p_val},{info['fused_variant'] & 15}>" if info["fused_variant"] >= 128 else
                                f"k_fused32<2,{args.p_val},{info['fused_variant'] & 15}>" if info["fused_variant"] >= 64 else
                                f"k_fused2<2,{args.p_val},{info['fused_variant']}>" if info["fused_variant"] >= 0 else f"k_fused<2,{args.p_val}>"),
                     "kernel_ms": fused_batch_avg_ms, "launches_timed": int(fused_batch_n),
                     "kernel_ms_one_by_one_after": {"min": min(each), "median": float(np.median(each)), "max": max(each)},
                     "rows_per_task": int(info["rows_per_group"]), "strip_rows_per_block": int(info.get("strip_rows_per_block", 0)),
                     "algorithmic_bytes_per_launch": int(bytes_per_launch), "peak_source": peak_src,
                     "timed": "CUDA events around the kernel on the launch stream inside the timed steps (one launch = all chunks of the GPU)",
                     "single_chunk_launch": {"kernel_ms": fused_avg_ms,
                                             "achieved": info["algorithmic_bytes"] / (fused_avg_ms * 1e-3) / 1e9,
                                             "frac": info["algorithmic_bytes"] / (fused_avg_ms * 1e-3) / 1e9 / peak,
                                             "eval_ms": lone_eval_ms, "rows_per_task": int(lone_info["rows_per_group"]),
                                             "n_tasks": int(lone_info["n_tasks"]),
                                             "note": "one chunk alone, as Model.optimize launches it (jb_plan_eval_device, device-resident; "
                                                     "eval_ms = all six launches, back to back): the plan sizes its tasks to one wave of the GPU"}},
        "fp64": fp64_info,
        "plan": {k: info[k] for k in ("n_tasks", "rows_per_group", "tile_pixels", "device_bytes", "scratch_bytes", "n_fit")},
        "setup_s": setup_s,
        "setup_split": {"host_threads": nworkers, "generate_synthetic_thread_s": setup_split[0], "plan_create_thread_s": setup_split[1],
                        "plan_create_ms_per_chunk": 1e3 * setup_split[1] / max(len(plans), 1)},
        "loss_chunk0": losses[0],
    }
    out["strong_scaling"] = strong
    # ---- the optimiser on the resident chunks: jb_lbfgs, all chunks in one optimiser -----------------
    out["lbfgs"] = None
    if not args.no_extra_legs and not args.float32:
        # every rank fits its own chunks (no collective: chunks are independent fits); whole-job numbers
        # are total evaluations over the slowest rank's wall time
        barrier()
        mine = lbfgs_leg(args, plans, models, main)
        if dist is not None:
            every = [None] * world
            dist.all_gather_object(every, mine)
            best = [min(r["runs"], key=lambda q: q["wall_s"]) for r in every]
            wall = max(b["wall_s"] for b in best)
            calls = sum(b["funcalls_total"] for b in best)
            mine = dict(every[0], per_rank=[{"wall_s": b["wall_s"], "rounds": b["rounds"], "funcalls_total": b["funcalls_total"]} for b in best],
                        whole_job={"wall_s": wall, "funcalls_total": calls, "value": args.epochs * args.pixels * calls / wall / 1e9})
        out["lbfgs"] = mine
    # ---- BASELINE.json configs[4]: one chunk, epochs sharded over the ranks, NCCL all-reduce ---------
    out["epoch_sharded"] = None
    if not args.no_extra_legs and not args.float32:
        if e2e_path == "pinned":
            for hb in halves:
                hb.close()
        batch.close()
        for pl in plans:
            pl.close()
        del P_all, O_all, d_p, d_out
        torch.cuda.empty_cache()
        out["epoch_sharded"] = epoch_sharded_leg(args, rank, world, local_rank, dist, args.es_epochs, args.es_pixels,
                                                 args.steps, args.warmup)
    out["api"] = None
    if rank == 0 and world == 1 and not args.no_extra_legs and not args.float32:
        out["api"] = api_leg(args, local_rank)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle
        ch, model, p, rows, cores = cpu_baseline(args, args.cpu_seconds)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            c_oracle.evaluate(p, ch.datablock, ch.metablock, model, nthreads=cores)
        dt = (time.perf_counter() - t0) / reps
        out["cpu_baseline"] = {
            "value": rows * args.pixels / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{rows} epochs x {args.pixels} px of chunk 0 ({reps} evals, {dt:.2f} s each), fused "
                      f"loss+grad+Fisher in oracle/jb_oracle.c (plain-C restatement of the reference, OpenMP)"}
    elif rank == 0:
        out["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def lbfgs_leg(args, plans, models, stream, maxiter=10):
    """Model.optimize for every resident chunk at once (jabble/model.py:195-238; the Pool over orders of
    notebooks/02-51peg.py:66-78), through the library's device optimiser (jb_lbfgs): templates in fit
    mode, a fixed number of L-BFGS-B iterations.  One round = one batched evaluation of all chunks + one
    launch of the optimiser's kernel (a block per chunk) + one 4-byte-per-chunk readback; reported next to
    the batched evaluation alone, timed with the same fit state."""
    import torch
    from wobble_jax_b200 import engine

    x0s = []
    for plan, model in zip(plans, models):
        model.fix(); model.fit(0, 1); model.fit(1, 1)
        plan.set_fit(engine.TreeSpec(model).fit_index())
        x0s.append(np.asarray(model.get_parameters(), dtype=np.float64).copy())
    n = len(plans)
    # the batched evaluation alone, in this fit state
    d_p = [torch.from_numpy(x).cuda() for x in x0s]
    d_out = [torch.empty(len(x) + 1, dtype=torch.float64, device="cuda") for x in x0s]
    batch = engine.Batch(plans)
    batch.bind([t.data_ptr() for t in d_p], [t.data_ptr() for t in d_out])
    for _ in range(3):
        batch.eval_device(stream.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        batch.eval_device(stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    eval_ms = e0.elapsed_time(e1) / reps
    batch.close()
    runs = []
    for it in range(2):
        t0 = time.perf_counter()
        res, info = engine.lbfgs_run(plans, x0s, {"maxiter": maxiter})
        dt = time.perf_counter() - t0
        calls = sum(d["funcalls"] for _, _, d in res)
        runs.append({"wall_s": dt, "create_s": info["create_s"], "run_s": info["run_s"], "destroy_s": info["destroy_s"],
                     "rounds": info["rounds"], "funcalls_total": calls, "ms_per_round": dt * 1e3 / info["rounds"],
                     "value": args.epochs * args.pixels * calls / dt / 1e9})
    best = min(runs, key=lambda r: r["wall_s"])
    f0 = [float(o[0]) for o in d_out]
    return {"workload": f"{n} chunks of {args.epochs} x {args.pixels} in ONE optimiser, templates in fit mode "
                        f"({len(x0s[0])} parameters per chunk), maxiter {maxiter}, scipy's default m / factr / pgtol",
            "unit": UNIT + " of evaluations performed", "runs": runs, "batched_eval_ms": eval_ms,
            "round_over_eval": best["ms_per_round"] / eval_ms,
            "loss_drop_chunk0": [f0[0], float(res[0][1])], "nit": [int(d["nit"]) for _, _, d in res][:4]}


def api_leg(args, local_rank):
    """The reference-facing Python API, timed: Model.optimize on one chunk (what a notebook calls per
    order, jabble/model.py:195-238) and model.optimize_many on several (the Pool over orders,
    notebooks/02-51peg.py:66-78), templates in fit mode, a fixed number of L-BFGS-B iterations.  Everything
    the call does is inside the clock: blockify, the upload into a plan (first call only: the second call
    finds the resident block, engine.get_plan), scipy, every evaluation's H2D / D2H copies."""
    import copy
    import wobble_jax_b200 as jabble
    from wobble_jax_b200 import engine, synth
    from wobble_jax_b200.dataset import Data

    def chunk(c):
        ch = synth.make_chunk(c, n_epochs=args.epochs, n_pix=args.pixels, p_val=args.p_val, as_block=True)
        db = ch.datablock
        data = Data.from_lists(list(db["xs"]), list(db["ys"]), list(db["yivar"]), list(db["mask"]))
        m = ch.model
        m.fix(); m.fit(0, 1); m.fit(1, 1)
        return data, m

    loss = jabble.loss.ChiSquare()
    px = args.epochs * args.pixels
    out = {}
    engine.clear_plan_cache()
    data, model = chunk(0)
    res = []
    for it in range(2):           # first call uploads the block, the second finds it resident
        m = copy.deepcopy(model)
        t0 = time.perf_counter()
        x, f, d = m.optimize(loss, data, options={"maxiter": 20})
        dt = time.perf_counter() - t0
        res.append({"wall_s": dt, "funcalls": int(d["funcalls"]), "nit": int(d["nit"]),
                    "ms_per_eval": dt * 1e3 / d["funcalls"], "ms_per_iteration": dt * 1e3 / max(d["nit"], 1),
                    "value": px * d["funcalls"] / dt / 1e9})
    out["optimize"] = {"workload": f"1 chunk of {args.epochs} x {args.pixels}, templates in fit mode ({len(x)} parameters), maxiter 20",
                       "optimizer": "jb_lbfgs (device)",
                       "unit": UNIT + " of evaluations performed", "first_call": res[0], "second_call_block_resident": res[1]}
    # the same call with scipy's L-BFGS-B on the host (one callback, one H2D and one D2H per evaluation)
    jabble.model.OPTIMIZER = "scipy"
    try:
        m = copy.deepcopy(model)
        t0 = time.perf_counter()
        x, f, d = m.optimize(loss, data, options={"maxiter": 20})
        dt = time.perf_counter() - t0
    finally:
        jabble.model.OPTIMIZER = "device"
    out["optimize"]["scipy_host_optimizer_block_resident"] = {
        "wall_s": dt, "funcalls": int(d["funcalls"]), "nit": int(d["nit"]), "ms_per_eval": dt * 1e3 / d["funcalls"],
        "value": px * d["funcalls"] / dt / 1e9}
    engine.clear_plan_cache()
    n = 8
    pairs = [chunk(c) for c in range(n)]
    engine.PLAN_CACHE_SIZE = max(engine.PLAN_CACHE_SIZE, n)
    many = []
    for it in range(2):           # first call uploads the blocks, the second finds them resident
        models = [copy.deepcopy(m) for _, m in pairs]
        t0 = time.perf_counter()
        outs = jabble.model.optimize_many(models, loss, [dt_ for dt_, _ in pairs], options={"maxiter": 10})
        dt = time.perf_counter() - t0
        calls = sum(int(o[2]["funcalls"]) for o in outs)
        rounds = max(int(o[2]["funcalls"]) for o in outs)
        many.append({"wall_s": dt, "funcalls_total": calls, "lockstep_rounds": rounds, "ms_per_round": dt * 1e3 / rounds,
                     "value": px * calls / dt / 1e9})
    out["optimize_many"] = {"workload": f"{n} chunks of {args.epochs} x {args.pixels} at once, templates in fit mode, maxiter 10",
                            "optimizer": "jb_lbfgs (device)",
                            "unit": UNIT + " of evaluations performed", "first_call": many[0], "second_call_blocks_resident": many[1]}
    jabble.model.OPTIMIZER = "scipy"
    try:
        models = [copy.deepcopy(m) for _, m in pairs]
        t0 = time.perf_counter()
        outs = jabble.model.optimize_many(models, loss, [dt_ for dt_, _ in pairs], options={"maxiter": 10})
        dt = time.perf_counter() - t0
    finally:
        jabble.model.OPTIMIZER = "device"
    calls = sum(int(o[2]["funcalls"]) for o in outs)
    rounds = max(int(o[2]["funcalls"]) for o in outs)
    out["optimize_many"]["scipy_host_optimizers_blocks_resident"] = {
        "wall_s": dt, "funcalls_total": calls, "lockstep_rounds": rounds, "ms_per_round": dt * 1e3 / rounds,
        "value": px * calls / dt / 1e9}
    engine.clear_plan_cache()
    return out


def epoch_sharded_leg(args, rank, world, local_rank, dist, epochs, pixels, steps, warmup):
    """BASELINE configs[4]: ONE chunk of `epochs` x `pixels`, its epochs sharded contiguously over the
    ranks.  A step is jb_group_eval_device: the rank's slabs in one launch per kernel, then the sum of
    [loss, dS, dT] over the slabs, ONE ncclAllReduce of it (8 (1 + 2 M) bytes, NVLink) and the scatter
    back, all enqueued by the library on one stream; per-epoch gradients and Fisher sums stay local."""
    import torch
    from wobble_jax_b200 import engine, synth

    per = (epochs + world - 1) // world
    e0 = rank * per
    E = max(0, min(per, epochs - e0))
    slab = 2048             # rows per plan: generated, canonicalised and uploaded slab by slab on host threads
    starts = list(range(0, E, slab))
    t_setup = time.perf_counter()

    def make(s0):
        torch.cuda.set_device(local_rank)
        n = min(slab, E - s0)
        ch = synth.make_chunk(0, n_epochs=n, n_pix=pixels, p_val=args.p_val, as_block=True,
                              epoch_offset=e0 + s0 + 1, total_epochs=epochs)
        model = synth.fit_all(ch.model)
        plan = engine.Plan(engine.TreeSpec(model), ch.datablock["xs"], ch.datablock["ys"], ch.datablock["yivar"],
                           ch.datablock["mask"], ch.metablock, device=local_rank, rows_per_group=32)
        plan.sync_from_model()
        return plan, np.asarray(model.get_parameters()).copy(), len(ch.grid)

    plans, p_host = [], []
    M = 0
    nworkers = max(1, min(8, len(os.sched_getaffinity(0)) // max(world, 1), len(starts)))
    with ThreadPoolExecutor(nworkers) as pool:
        for plan, p, m in pool.map(make, starts):
            plans.append(plan); p_host.append(p); M = m
    setup_s = time.perf_counter() - t_setup
    offs = np.concatenate([[0], np.cumsum([len(p) for p in p_host])])
    P_all = torch.from_numpy(np.concatenate(p_host)).cuda()
    O_all = torch.empty(int(offs[-1]) + len(plans), dtype=torch.float64, device="cuda")
    d_p = [P_all[int(offs[i]):int(offs[i + 1])] for i in range(len(plans))]
    d_out = [O_all[int(offs[i]) + i:int(offs[i + 1]) + i + 1] for i in range(len(plans))]
    nccl_id = engine.Group.broadcast_unique_id(rank, world) if world > 1 else None
    group = engine.Group(plans, rank, world, nccl_id)
    group.bind([t.data_ptr() for t in d_p], [t.data_ptr() for t in d_out])
    stream = torch.cuda.current_stream()

    def step():
        P_all.add_(1e-13)
        group.eval_device(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    group.eval_device_checked(stream.cuda_stream)        # settles window sizes before the timed steps
    for _ in range(warmup):
        step()
    barrier()
    launches0 = group.launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    launches1 = group.launches()
    # the fused kernel's share, from a few more steps with the library's event pair around it (those steps
    # are enqueued launch by launch; the timed ones above replay the step's recorded CUDA graph)
    plans[0].profile(True); plans[0].profile_read()
    for _ in range(max(3, min(steps, 10))):
        step()
    barrier()
    fused_ms, fused_n = plans[0].profile_read()
    plans[0].profile(False)
    for pl in plans:
        pl.check()
    ms = float(t[0]) / steps
    loss = float(d_out[0][0])
    info = plans[0].info()
    ginfo = group.info()
    launches = (launches1 - launches0) // max(steps, 1)
    peak, _ = measured_peaks()
    px = epochs * pixels
    bytes_total = px * 25 + 4 * M * 8 + epochs * 56
    out = {
        "workload": f"synthetic {epochs} epochs x {pixels} px single chunk (BASELINE.json configs[4]), epoch-sharded x{world}",
        "value": px / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms, "scaling": "strong",
        "epochs_per_gpu": per, "plans_per_gpu": len(plans),
        "hbm_frac_per_gpu": bytes_total / world / (ms * 1e-3) / 1e9 / peak,
        "fused_kernel_ms": fused_ms / max(fused_n, 1),
        "fused_kernel_hbm_frac": (bytes_total / world) / (fused_ms / max(fused_n, 1) * 1e-3) / 1e9 / peak if fused_n else None,
        "collective": "ncclAllReduce(sum, f64) of [loss, dS, dT]" if world > 1 else "none (one rank)",
        "collective_bytes": int(ginfo["collective_bytes"]), "launches_per_step": int(launches),
        "api": "jb_group_eval_device (include/jabble_b200.h): kernels + gather + all-reduce + scatter on one stream, "
               "replayed as one recorded CUDA graph per step",
        "kernel": "k_strip" if info["fused_variant"] >= 128 else "k_fused2" if info["fused_variant"] >= 0 else "k_fused",
        "loss": loss, "setup_s": setup_s,
    }
    group.close()
    for pl in plans:
        pl.close()
    del P_all, O_all
    torch.cuda.empty_cache()
    return out


def run_epochs(args, rank, world, local_rank):
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    main_stream = torch.cuda.Stream()
    torch.cuda.set_stream(main_stream)
    res = epoch_sharded_leg(args, rank, world, local_rank, dist, args.epochs, args.pixels, args.steps, args.warmup)
    if rank == 0:
        peak, src = measured_peaks()
        print(json.dumps({
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": res["workload"], "epochs": args.epochs, "pixels": args.pixels,
                       "parallelism": f"epoch-sharded x{world}"},
            "epoch_sharded": res}))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != max(args.gpus, 1) and world != 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    if args.workload == "epochs":
        run_epochs(args, rank, world, local_rank)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
