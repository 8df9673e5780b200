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
    if name == "shift":  # the decision table changes twice in the middle of the run: the pipelined run repeats two sweeps
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from random_cases import table_shift_case
        return table_shift_case() + (10 ** 12,)
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
             iterations=res["iterations"], exit=res["exit_reason"], hash=np.uint64(eng.labels_hash()),
             quirks=[res["q_add_to_inside"], res["q_remove_to_outside"], res["q_cancel_repromoted"], res["q_cancelled"]])
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["p2p", "collective"])
@pytest.mark.parametrize("name,mode", [("forest", "f64_dense"), ("forest", "index"), ("excl", "f64_band"), ("shift", "f64_dense"),
                                       ("shift", "index")])
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
        q = ref["quirk_potential"]  # summed over the slabs by the exchange after the exit
        assert p["quirks"].tolist() == [q["add_to_inside"], q["remove_to_outside"], q["cancel_repromoted"], q["cancelled"]]
    from oracle.c_oracle import hash_labels
    assert sum(int(p["hash"]) for p in parts) % 2 ** 64 == hash_labels(ref["labels"])  # slab hashes add up


def _noisy():
    from arterynetwork_b200.phantom import make_phantom
    data, vm, _ = make_phantom((48, 30, 34), seed=1, cell=(48, 30, 34), margin=3, depth=3, root_r2=9, min_len=6, max_len=12,
                               quantum=16, sigma_k=5)
    vm[10:40, 8:22, 8:26] = 0  # a big seed across every slab boundary: removals, cancelled and re-promoted additions
    return data, vm


@pytest.mark.parametrize("mode", ["index", "f64_dense"])
def test_dropin_function_runs_on_every_gpu_from_one_process(mode):
    """variationalRegionGrowing() itself on z-slabs over all GPUs of the box -- no torchrun, one process, one host thread per
    GPU, peer access instead of CUDA IPC (module switch DEVICES): same labels, stdout and counters as the whole-volume oracle."""
    import contextlib
    import io
    import torch
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    from arterynetwork_b200 import variationalRegionGrowing as mod
    from oracle.c_oracle import vrg_oracle_c
    keep = (mod.DEVICES, mod.INTENSITY, mod.MAX_SECONDS)
    try:
        mod.DEVICES, mod.INTENSITY = "all", mode
        for case in ("forest", "excl", "noisy"):
            data, vm = _noisy() if case == "noisy" else _case(case)[:2]
            ref = vrg_oracle_c(data, vm, max_segment_size=10 ** 12)
            vm64 = vm.astype(np.int64)
            with contextlib.redirect_stdout(io.StringIO()) as buf, __import__("warnings").catch_warnings():
                __import__("warnings").simplefilter("ignore")
                segmented, seg_map, out = mod.variationalRegionGrowing(data, vm64, maxSegmentSize=10 ** 12)
            assert out is vm64 and np.array_equal(vm64, ref["labels"]), case
            assert np.array_equal(seg_map == 1, ref["seg"]) and np.array_equal(segmented, np.argwhere(ref["seg"]))
            assert buf.getvalue().startswith("Finished at iteration %d\n" % ref["iterations"])
            q = ref["quirk_potential"]
            r = mod.LAST_RUN
            assert (r["q_add_to_inside"], r["q_remove_to_outside"], r["q_cancel_repromoted"], r["q_cancelled"]) == (
                q["add_to_inside"], q["remove_to_outside"], q["cancel_repromoted"], q["cancelled"]), case
        # the wall-clock exit is collective on slabs: every rank leaves at the same update (and nobody hangs)
        mod.MAX_SECONDS = 1e-9
        data, vm = _case("forest")[:2]
        vm64 = vm.astype(np.int64)
        with contextlib.redirect_stdout(io.StringIO()) as buf:
            mod.variationalRegionGrowing(data, vm64, maxSegmentSize=10 ** 12)
        assert "(Max time reached)" in buf.getvalue()
    finally:
        mod.DEVICES, mod.INTENSITY, mod.MAX_SECONDS = keep
