#!/usr/bin/env python
"""Strict (list-order) mode, SURVEY.md section 8(f) N4: wall time, wavefront rounds and parity with the list-order oracle on
config C1 (128^3 forest, the shape the unmodified reference needs 596 s for) and on a noisy 96^3 volume with a fat seed
(thousands of order-dependent flips).  The oracle comparison runs outside the timed region."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def noisy(shape, seed, noise, q, cube):
    rng = np.random.default_rng(seed)
    d = np.zeros(shape)
    c = [s // 2 for s in shape]
    d[c[0] - 4:c[0] + 4, c[1] - 4:c[1] + 4, 4:shape[2] - 4] = 1.0
    d[4:shape[0] - 4, c[1] - 3:c[1] + 3, c[2] - 3:c[2] + 3] = 1.0
    k = np.round((d + rng.normal(0, noise, shape)) * q).astype(np.int64)
    vm = np.full(shape, 3, dtype=np.uint8)
    vm[c[0] - cube // 2:c[0] + cube // 2, c[1] - cube // 2:c[1] + cube // 2, c[2] - cube // 2:c[2] + cube // 2] = 0
    return k / q, vm


def run(name, data, vm, check):
    import torch
    from arterynetwork_b200.strict import StrictEngine
    ms = data.size + 1
    times = []
    with StrictEngine(data.shape, max_segment_size=ms) as eng:
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.init(data, vm)
            res = eng.run()
            times.append(time.perf_counter() - t0)
        labels = eng.value_map()
        seg_rows = eng.segmented()
        trace = eng.trace()
    out = {"workload": name, "shape": list(data.shape), "seconds": times, "iterations": res["iterations"], "segmented": res["n_in"],
           "flips": int(trace[1:, 0].sum()), "wavefront_rounds": res["rounds"] // 3, "skipped": res["skipped"] // 3,
           "dropped": res["dropped"] // 3, "kernel_launches": res["kernel_launches"] // 3}
    if check:
        from oracle.strict_oracle import vrg_strict_oracle
        t0 = time.perf_counter()
        o = vrg_strict_oracle(data, vm, max_segment_size=ms)
        out["oracle_seconds_1_thread"] = time.perf_counter() - t0
        out["parity"] = {"value_map": bool(np.array_equal(labels, o["value_map"])), "segmented_rows": bool(np.array_equal(seg_rows, o["segmented"])),
                         "trace": bool(np.array_equal(trace, o["trace"])), "iterations": res["iterations"] == o["iterations"]}
    print(json.dumps(out))


def main():
    import torch
    from arterynetwork_b200.phantom import make_phantom
    torch.cuda.set_device(0)
    kw = dict(cell=(128, 128, 128), margin=8, depth=4, root_r2=16, min_len=16, max_len=34)
    data, vm, info = make_phantom((128, 128, 128), seed=0, **kw)
    run("C1 128^3 forest (reference: 596 s)", data, vm, True)
    d2, v2 = noisy((96, 96, 96), 11, 0.4, 8, 40)
    run("96^3 noisy bar + 40^3 seed", d2, v2, True)
    kw2 = dict(cell=(128, 128, 128), margin=8, depth=4, root_r2=16, min_len=16, max_len=34)
    d3, v3, _ = make_phantom((256, 256, 256), seed=0, **kw2)
    run("256^3 forest, 8 trees", d3, v3, True)


if __name__ == "__main__":
    main()
