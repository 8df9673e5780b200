#!/usr/bin/env python
"""bench.py -- VRG Gvoxel-updates/s on the BASELINE.json configs (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--intensity f64_dense]
    python bench.py --impl reference ...      # CPU arm: the oracle port on the host cores

One "step" is one complete variational region growing run (init branch + every
iteration to convergence) on the synthetic vessel-forest phantom of the named
shape.  ``value`` = N_voxels * sweeps / time with the inputs resident in HBM;
``e2e`` is the same run through the C-ABI with HOST (pinned) buffers, the
host->device copies of intensities + labels and the device->host copy of the
result labels inside the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {  # (Z, Y, X): BASELINE.json configs; the 640 / 170 / 1024 axis is the slab axis
    "c1": (128, 128, 128),
    "c2": (170, 512, 512),
    "c3": (640, 880, 880),
    "c4": (1024, 1024, 1024),
    "c5": (1024, 2048, 2048),
}
ALGO_BYTES_PER_UPDATE = 10.0  # 8 B fp64 intensity + 1 B label read + 1 B label write
METRIC = "VRG Gvoxel-updates/s at 880x880x640, 1/2/4/8 B200; % of HBM roofline"
CPU_SAMPLE_PLANES = 128


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md): NVML polled every 10 ms from a thread
    (first sample taken synchronously in start(), so that even a 50 ms timed region is covered); nvidia-smi -lms as the
    fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.nvml, self.handle, self.stop_flag, self.thread = None, None, threading.Event(), None
        self.sm, self.mx, self.reasons = [], [], set()

    def _nvml_sample(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        for bit, name in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def _nvml_loop(self):
        while not self.stop_flag.wait(0.01):
            try:
                self._nvml_sample()
            except Exception:
                return

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml_sample()
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "samples": len(self.sm), "reasons": sorted(self.reasons), "source": "nvml, 10 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 200"}


def forest_for(shape, seed):
    from arterynetwork_b200.phantom import forest_segments
    return forest_segments(shape, seed=seed)


def device_phantom(shape, seed, z0, nz, device):
    """fp64 intensities + uint8 seeds of planes [z0, z0+nz) generated on the device."""
    import ctypes
    import torch
    from arterynetwork_b200 import _native as nat
    segs, roots = forest_for(shape, seed)
    d = torch.empty((nz,) + tuple(shape[1:]), dtype=torch.float64, device="cuda:%d" % device)
    v = torch.empty((nz,) + tuple(shape[1:]), dtype=torch.uint8, device="cuda:%d" % device)
    shp = (ctypes.c_int64 * 3)(*shape)
    segs = np.ascontiguousarray(segs); roots = np.ascontiguousarray(roots)
    nat.check(nat.load().vrg_phantom_device(device, ctypes.addressof(shp), z0, nz, segs.ctypes.data, len(segs),
                                            roots.ctypes.data, len(roots), seed, 256, 31, 0, 0,
                                            d.data_ptr(), v.data_ptr()))
    return d, v


def cpu_sample(shape, seed):
    """Bounded CPU sample of the workload: a window of its planes (about 1e8 voxels at most) that holds seeds, as its
    own volume.  Returns (data, value_map, description, z0)."""
    from arterynetwork_b200.phantom import forest_segments, make_phantom
    nz = min(CPU_SAMPLE_PLANES, shape[0], max(8, int(1.0e8 // (shape[1] * shape[2]))))
    _, roots = forest_segments(shape, seed=seed)
    rz = int(roots[:, 0].min())
    z0 = 0 if rz + 2 <= nz else max(0, min(shape[0] - nz, rz - nz // 2))  # planes [0, nz) unless they hold no seed
    data, vm, _ = make_phantom(shape, seed=seed, z0=z0, nz=nz)
    return data, vm, "planes [%d,%d) of the %dx%dx%d phantom as a %dx%dx%d volume, run to convergence" % (
        z0, z0 + nz, shape[2], shape[1], shape[0], shape[2], shape[1], nz), z0


def reference_python_c1():
    """The unmodified reference's own time on config C1 (128^3), measured in the build container when the golden
    fixture was made (tests/golden/make_golden.py c1_128): the reference is pure Python and cannot travel to the GPU
    box, so this is a recorded number, not one measured in this run."""
    try:
        f = np.load(os.path.join(ROOT, "tests", "golden", "c1_128.npz"))
        sec, it = float(f["reference_wall_s"]), int(f["iterations"])
        return {"seconds": sec, "iterations": it, "Gvoxel_updates_per_s": 128 ** 3 * it / sec / 1e9, "cores": 1,
                "where": "build container, recorded in tests/golden/c1_128.npz (unmodified reference, NumPy %s)" % str(f["numpy_version"])}
    except Exception:
        return None


def time_cpu_port(data, vm, threads):
    from oracle.c_oracle import vrg_oracle_c
    t0 = time.perf_counter()
    o = vrg_oracle_c(data, vm, max_segment_size=10 ** 15, nthreads=threads)
    dt = time.perf_counter() - t0
    return data.size * o["iterations"] / dt / 1e9, dt, o["iterations"]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    data, vm, sample, _ = cpu_sample(shape, args.seed)
    for _ in range(args.warmup):
        time_cpu_port(data, vm, threads)
    t0 = time.perf_counter()
    iters = 0
    for _ in range(args.steps):
        _, _, it = time_cpu_port(data, vm, threads)
        iters += it
    dt = time.perf_counter() - t0
    value = data.size * iters / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gvoxel-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s %dx%dx%d vessel-forest phantom, seed %d" % (args.workload, shape[2], shape[1], shape[0], args.seed)},
        "cpu_baseline": {"value": value, "unit": "Gvoxel-updates/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Gvoxel-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is pure Python and exists only in the build container; this arm times oracle/vrg_oracle.c, "
                "the C restatement pinned to the reference's golden outputs, on all host threads",
    }
    print(json.dumps(line))


def bench_mode(eng, torch, mode_name, d_data, d_vm, steps, warmup, nvox, profile_in_timed_region=True):
    """Resident-input timing of one intensity mode: step = attach (zero-copy) + level scan + init + run.

    ``profile_in_timed_region``: CUDA events around every sweep launch inside the timed steps (the primary mode; this
    keeps vrg_run on plain stream launches).  Otherwise the timed steps run unprofiled -- vrg_run then replays CUDA
    graphs -- and the per-launch sweep time comes from one extra profiled step right after them.
    """
    def step():
        eng.attach_device(d_data.data_ptr(), d_vm.data_ptr())
        eng.init()
        return eng.run()

    for _ in range(warmup):
        res = step()
    eng.profile(profile_in_timed_region)
    l0 = eng.poll()["kernel_launches"]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    sweeps = 0
    for _ in range(steps):
        res = step()
        sweeps += res["sweeps"]
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end)
    launches = eng.poll()["kernel_launches"] - l0
    if not profile_in_timed_region:
        eng.profile(True)
        step()
    prof = eng.get_profile()
    eng.profile(False)
    return {"ms": ms, "sweeps": sweeps, "res": res, "prof": prof, "launches": launches, "mode": mode_name,
            "profiled_in_timed_region": profile_in_timed_region, "value": nvox * sweeps / (ms * 1e-3) / 1e9}


def roofline_of(r, nvox, peak, peak_kind, traffic=None):
    prof = r["prof"]
    per_launch_ms = prof["decide_ms"] / max(1, prof["decide_launches"])
    achieved = ALGO_BYTES_PER_UPDATE * nvox / (per_launch_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "k_sweep_dense" if r.get("mode") == "f64_dense" else "k_sweep", "achieved": achieved, "peak": peak, "peak_kind": peak_kind,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "algorithmic_bytes_per_launch": ALGO_BYTES_PER_UPDATE * nvox, "ms_per_launch": per_launch_ms,
            "launches_timed": prof["decide_launches"],
            "share_of_step": prof["decide_ms"] / r["ms"] if r.get("profiled_in_timed_region", True) else None,
            "cancel_ms_per_launch": prof["cancel_ms"] / max(1, prof["cancel_launches"])}


def workload_shape(args, n_gpus):
    """(Z, Y, X) of the run: the named config, or under --scaling weak its planes / 8 per GPU."""
    Z, Y, X = WORKLOADS[args.workload]
    if args.scaling == "weak":
        Z = (Z // 8) * n_gpus
    return (Z, Y, X)


def run_single(args):
    import torch
    from arterynetwork_b200.engine import VRGEngine
    dev = 0
    torch.cuda.set_device(dev)
    shape = workload_shape(args, 1)
    nvox = shape[0] * shape[1] * shape[2]
    peak, peak_kind = measured_peak()
    d_data, d_vm = device_phantom(shape, args.seed, 0, shape[0], dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.Stream())  # a capturable (non-default) stream: vrg_run replays CUDA graphs on it
    stream = torch.cuda.current_stream().cuda_stream
    modes = [args.intensity] + [m for m in ("f64_dense", "f64_band", "index") if m != args.intensity]
    results = {}
    clocks = None
    for i, mode in enumerate(modes):
        with VRGEngine(shape, max_segment_size=10 ** 15, intensity=mode, device=dev) as eng:
            eng.set_stream(stream)
            if i == 0:
                sampler = ClockSampler(dev)
                sampler.start()
                results[mode] = bench_mode(eng, torch, mode, d_data, d_vm, args.steps, args.warmup, nvox)
                clocks = sampler.stop()
                # end to end through the C-ABI with host buffers (pinned), same mode
                h_data = torch.empty(d_data.shape, dtype=torch.float64, pin_memory=True)
                h_vm = torch.empty(d_vm.shape, dtype=torch.uint8, pin_memory=True)
                h_out = torch.empty(d_vm.shape, dtype=torch.uint8, pin_memory=True)
                h_data.copy_(d_data); h_vm.copy_(d_vm)
                torch.cuda.synchronize()

                def e2e_step():
                    eng.upload(h_data.numpy(), h_vm.numpy())
                    eng.init()
                    r = eng.run()
                    from arterynetwork_b200 import _native as nat
                    nat.check(eng.lib.vrg_download_labels(eng._h, h_out.data_ptr()))
                    return r
                for _ in range(max(1, min(args.warmup, 2))):
                    e2e_step()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                s.record()
                sw = 0
                for _ in range(args.steps):
                    sw += e2e_step()["sweeps"]
                e.record()
                torch.cuda.synchronize()
                e2e_ms = s.elapsed_time(e)
                e2e = {"value": nvox * sw / (e2e_ms * 1e-3) / 1e9, "unit": "Gvoxel-updates/s",
                       "h2d_bytes_per_step": int(h_data.numel() * 8 + h_vm.numel()),
                       "d2h_bytes_per_step": int(h_out.numel()), "ms_per_step": e2e_ms / args.steps,
                       "intensity_mode": mode}
                labels_primary = h_out.numpy().copy()
                del h_data, h_vm
            else:
                results[mode] = bench_mode(eng, torch, mode, d_data, d_vm, max(1, min(args.steps, 3)), 2, nvox,
                                           profile_in_timed_region=False)
                out = torch.empty(d_vm.shape, dtype=torch.uint8, device="cuda")
                eng.labels_device(out.data_ptr())
                torch.cuda.synchronize()
                same = bool(np.array_equal(out.cpu().numpy(), labels_primary))
                results[mode]["labels_equal_primary"] = same
                del out
    prim = results[args.intensity]
    # CPU baseline (rank 0, N=1): the oracle port on a bounded sample of the same phantom
    data_s, vm_s, sample, z0_s = cpu_sample(shape, args.seed)
    threads = os.cpu_count() or 1
    cpu_val, cpu_dt, cpu_it = time_cpu_port(data_s, vm_s, threads)
    # the device generator and the NumPy generator must agree bit for bit on the sample
    gen_equal = bool(np.array_equal(d_data[z0_s: z0_s + data_s.shape[0]].cpu().numpy(), data_s))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_%s_%s.json" % (args.workload, args.intensity))
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": prim["value"], "unit": "Gvoxel-updates/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": prim["ms"] / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s %dx%dx%d vessel-forest phantom, seed %d" % (args.workload, shape[2], shape[1], shape[0], args.seed),
                   "intensity_mode": args.intensity, "sweeps_per_step": prim["sweeps"] // args.steps,
                   "segmented_voxels": prim["res"]["n_in"], "levels": prim["res"]["n_levels"],
                   "l2": "inputs (%.1f GB) larger than L2; no flush" % (nvox * 9 / 1e9),
                   "step": "attach resident inputs (zero-copy) + level scan + init + all iterations"},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": prim["launches"],
        "roofline": roofline_of(prim, nvox, peak, peak_kind, traffic),
        "cpu_baseline": {"value": cpu_val, "unit": "Gvoxel-updates/s", "cores": threads, "kind": "port",
                         "sample": sample, "seconds": cpu_dt, "iterations": cpu_it,
                         "device_phantom_equals_numpy_phantom": gen_equal, "reference_python_c1": reference_python_c1()},
        "modes": {m: {"value": r["value"], "ms_per_step": r["ms"] / max(1, (args.steps if m == args.intensity else min(args.steps, 3))),
                      "roofline_frac_10B": roofline_of(r, nvox, peak, peak_kind)["frac"],
                      "decide_ms_per_launch": r["prof"]["decide_ms"] / max(1, r["prof"]["decide_launches"]),
                      "labels_equal_primary": r.get("labels_equal_primary", True)} for m, r in results.items()},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--intensity", default="f64_dense", choices=["f64_dense", "f64_band", "index"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the named volume over N slabs (the BASELINE metric); weak: every GPU gets 1/8 of the "
                         "named volume's planes, i.e. the volume grows with N (c5: 2048x2048x128 per GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 or args.gpus > 1:
        from arterynetwork_b200.distributed import run_bench_distributed
        return run_bench_distributed(args, WORKLOADS, METRIC, ALGO_BYTES_PER_UPDATE, measured_peak())
    return run_single(args)


if __name__ == "__main__":
    main()
