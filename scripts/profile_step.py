#!/usr/bin/env python
"""One bounded VRG run for ncu (never a bench number): python scripts/profile_step.py --workload c3 --iters 12"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--intensity", default="f64_dense")
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--shape", default=None, help="ZxYxX instead of a named workload")
    a = ap.parse_args()
    import torch
    from arterynetwork_b200.engine import VRGEngine
    shape = tuple(int(t) for t in a.shape.split("x")) if a.shape else bench.WORKLOADS[a.workload]
    d, v = bench.device_phantom(shape, a.seed, 0, shape[0], 0)
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.Stream())
    with VRGEngine(shape, max_segment_size=10 ** 15, intensity=a.intensity, iter_max=a.iters) as eng:
        eng.set_stream(torch.cuda.current_stream().cuda_stream)
        eng.attach_device(d.data_ptr(), v.data_ptr())
        eng.init()
        print(eng.run())


if __name__ == "__main__":
    main()
