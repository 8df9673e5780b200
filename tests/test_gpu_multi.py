"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): the z-slab run (peer-memory and NCCL transports) must equal the
whole-volume C oracle and the single-GPU run bit for bit -- labels, iteration count and trace."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case(name):
    from arterynetwork_b200.phantom import make_phantom
    if name == "forest":
        return make_phantom((64, 96, 160), seed=4, cell=(64, 96, 80), margin=4, depth=3, root_r2=9, min_len=10, max_len=22)[:2] + (10 ** 12,)
    if name == "excl":
        return make_phantom((40, 60, 100), seed=1, cell=(40, 60, 100), margin=4, depth=3, root_r2=9, min_len=8, max_len=16,
                            exclude_below_k=40)[:2] + (10 ** 12,)
    raise KeyError(name)


def _worker(rank, world, port, name, mode, transport, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from arterynetwork_b200.distributed import DistributedVRG, GpuSlabEngine, slab_bounds
    from arterynetwork_b200.engine import VRGEngine
    data, vm, max_seg = _case(name)
    b = slab_bounds(data.shape[0], world)
    eng = VRGEngine(data.shape, max_segment_size=max_seg, intensity=mode, device=rank, z_begin=b[rank], z_end=b[rank + 1])
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.upload(data, vm)
    drv = DistributedVRG(GpuSlabEngine(eng, rank), rank, world, check_every=3, transport=transport)
    drv.prepare_levels()
    drv.init()
    res = drv.run()
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), labels=eng.labels(), trace=drv.trace(),
             iterations=res["iterations"], exit=res["exit_reason"])
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["p2p", "collective"])
@pytest.mark.parametrize("name,mode", [("forest", "f64_dense"), ("forest", "index"), ("excl", "f64_band")])
def test_slabs_equal_oracle(name, mode, transport, tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    from oracle.c_oracle import vrg_oracle_c
    data, vm, max_seg = _case(name)
    ref = vrg_oracle_c(data, vm, max_segment_size=max_seg)
    mp.spawn(_worker, args=(world, _free_port(), name, mode, transport, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    assert np.array_equal(np.concatenate([p["labels"] for p in parts]), ref["labels"])
    for p in parts:
        assert int(p["iterations"]) == ref["iterations"] and int(p["exit"]) == ref["exit"]
        assert np.array_equal(p["trace"], ref["trace"])
