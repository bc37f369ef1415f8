"""Epoch-sharded evaluation (BASELINE.json configs[4]) on two REAL ranks: one process per GPU, NCCL,
through jb_group (the library enqueues the all-reduce of [loss, template gradients] behind the fused
kernels).  Every rank compares what it gets with the C oracle on the whole chunk (1e-10: loss, every
gradient block, the RV Fisher information) and with the unsharded plan (1e-12).  Needs two GPUs:
``gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu``; skipped on one."""

import os
import socket
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_epochs, n_pix, expect_strip):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import wobble_jax_b200 as jabble
    from wobble_jax_b200 import engine, multigpu, synth
    from oracle import c_oracle
    from helpers import rel_err
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ch = synth.make_chunk(chunk=2, n_epochs=n_epochs, n_pix=n_pix, p_val=3, n_lines=10)      # seeded: the same on every rank
        model = synth.fit_all(ch.model)
        db, mb = ch.data.blockify()
        dbd, mbd = {k: db[k] for k in ("xs", "ys", "yivar", "mask")}, {k: mb[k] for k in mb.keys()}
        p = np.asarray(model.get_parameters())
        E = db["xs"].shape[0]
        oval, ograd, ofisher, _ = c_oracle.evaluate(p, dbd, mbd, model)
        loss = jabble.loss.ChiSquare()
        obj = multigpu.EpochShardedObjective(loss, model, db, mb, rank, world, device=rank)
        assert obj.world_size == world
        val, grad = obj.value(p), obj.gradient(p)
        assert obj.jgroup.info()["collective_bytes"] == 8 * (1 + 2 * len(ch.grid))      # [loss, dS, dT], not the whole gradient
        if expect_strip:
            assert obj.plan.info()["fused_variant"] & 128
        spec = engine.TreeSpec(model)
        blocks, o = [], 0
        for lf in spec.leaves:
            if lf.node._fit:
                blocks.append((o, o + lf.size)); o += lf.size
        assert abs(val - oval) <= 1e-10 * abs(oval)
        for a, b in blocks:
            assert rel_err(grad[a:b], ograd[a:b]) < 1e-10
        # the unsharded plan on this rank's GPU
        whole = engine.build_plan(model, db, mb, device=rank)
        wval, wgrad = whole.eval(p)
        assert abs(val - wval) <= 1e-12 * abs(wval)
        for a, b in blocks:
            assert rel_err(grad[a:b], wgrad[a:b]) < 1e-12
        # Fisher information: every key's sum lives on the rank that owns its rows
        F = torch.from_numpy(np.asarray(obj.plan.fisher(0))).cuda()
        dist.all_reduce(F)
        assert rel_err(F.cpu().numpy(), ofisher[:E]) < 1e-10
        assert rel_err(F.cpu().numpy(), whole.fisher(0)) < 1e-12
        # a second evaluation at a moved point: same through the sharded and the unsharded path
        p2 = p.copy(); p2[E:E + 50] += 1e-3
        v2, g2 = obj.value(p2), obj.gradient(p2)
        w2, wg2 = whole.eval(p2)
        assert abs(v2 - w2) <= 1e-12 * abs(w2) and rel_err(g2, wg2) < 1e-12
        obj.close(); whole.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_epochs,n_pix,expect_strip", [(200, 2048, False), (3200, 330, True)])
def test_epoch_sharded_objective_on_two_gpus(n_epochs, n_pix, expect_strip):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), n_epochs, n_pix, expect_strip), nprocs=2, join=True)
