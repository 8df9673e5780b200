"""world_size-2/3 gloo tests of the z-slab driver (arterynetwork_b200/distributed.py) on CPU.

The per-slab compute is the NumPy slab engine (test-only); what is under test is the host-side
multi-rank logic: slab bounds, halo exchange, statistics all-reduce, level union, exit handling.
The result must equal the whole-volume oracle bit for bit, whatever the number of slabs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case(name):
    from arterynetwork_b200.phantom import make_phantom
    if name == "forest":
        data, vm, _ = make_phantom((36, 40, 44), seed=3, cell=(36, 40, 44), margin=3, depth=3, root_r2=9,
                                   min_len=8, max_len=16)
        return data, vm, 10 ** 12
    if name == "excl":  # label 4 + seeds straddling the slab boundary of a 2-rank split
        data, vm, _ = make_phantom((24, 28, 32), seed=1, cell=(24, 28, 32), margin=3, depth=3, root_r2=9,
                                   min_len=6, max_len=10, exclude_below_k=40)
        return data, vm, 10 ** 12
    if name == "maxseg":
        data, vm, _ = make_phantom((24, 28, 32), seed=2, cell=(24, 28, 32), margin=3, depth=3, root_r2=9,
                                   min_len=6, max_len=10)
        return data, vm, 60
    if name == "manylevels":  # more than 4095 distinct levels per slab: the level gather grows to its large form
        data, vm, _ = make_phantom((20, 20, 24), seed=2, cell=(20, 20, 24), margin=3, depth=2, root_r2=4, min_len=5,
                                   max_len=8)
        data = data + np.random.default_rng(9).normal(0, 1e-3, data.shape)  # 9600 distinct values
        return data, vm, 10 ** 12
    raise KeyError(name)


def _worker(rank, world, port, name, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from arterynetwork_b200.distributed import DistributedVRG, slab_bounds
    from cpu_slab_engine import NumpySlabEngine
    data, vm, max_seg = _case(name)
    b = slab_bounds(data.shape[0], world)
    eng = NumpySlabEngine(data.shape, b[rank], b[rank + 1], data, vm, max_segment_size=max_seg)
    drv = DistributedVRG(eng, rank, world, check_every=3)
    drv.prepare_levels()
    drv.init()
    res = drv.run()
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), labels=eng.labels(), trace=drv.trace(),
             iterations=res["iterations"], exit=res["exit_reason"], n_in=res["n_in"],
             quirks=[res["q_add_to_inside"], res["q_remove_to_outside"], res["q_cancel_repromoted"], res["q_cancelled"]])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("forest", 2), ("excl", 2), ("forest", 3), ("maxseg", 2), ("manylevels", 2)])
def test_slab_driver_equals_whole_volume_oracle(name, world, tmp_path):
    from oracle.vrg_oracle import vrg_oracle
    data, vm, max_seg = _case(name)
    ref = vrg_oracle(data, vm, max_segment_size=max_seg)
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    labels = np.concatenate([p["labels"] for p in parts])
    assert np.array_equal(labels, ref["labels"])
    for p in parts:  # every rank saw the same global trajectory
        assert int(p["iterations"]) == ref["iterations"] and int(p["exit"]) == ref["exit"]
        assert np.array_equal(p["trace"], ref["trace"])
        # the order-dependence counters are summed over the slabs (one more all-reduce after the exit)
        q = ref["quirk_potential"]
        assert p["quirks"].tolist() == [q["add_to_inside"], q["remove_to_outside"], q["cancel_repromoted"], q["cancelled"]]


def test_slab_bounds():
    from arterynetwork_b200.distributed import slab_bounds
    assert slab_bounds(640, 8) == [0, 80, 160, 240, 320, 400, 480, 560, 640]
    assert slab_bounds(10, 3) == [0, 4, 7, 10]
    with pytest.raises(ValueError):
        slab_bounds(5, 4)
