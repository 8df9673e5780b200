#!/usr/bin/env python
"""profiles/traffic_<workload>_<mode>.json from a full ncu capture of the dense sweep: DRAM bytes per launch, stamped with the hash
of the kernel sources the capture was taken on (bench.py uses the number for `roofline.traffic` only if the hash still matches).

    python scripts/make_traffic_json.py gpurun_out/r2_final_sweep_dense.ncu-rep c3 f64_dense "<how it was captured>"
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rep, workload, mode, how = sys.argv[1:5]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]

    def col(name):
        i = hdr.index(name)
        return [float(r[i].replace(",", "")) for r in rows[2:]]
    units = dict(zip(hdr, rows[1]))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = [v * scale[units["dram__bytes_read.sum"]] for v in col("dram__bytes_read.sum")]
    wr = [v * scale[units["dram__bytes_write.sum"]] for v in col("dram__bytes_write.sum")]
    t = col("gpu__time_duration.sum")
    tu = units["gpu__time_duration.sum"]
    t_us = [v / 1e3 if tu in ("ns", "nsecond") else (v if tu in ("us", "usecond") else v * 1e3) for v in t]
    pct = col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") if "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed" in hdr else []
    d = {"kernel": rows[2][hdr.index("Kernel Name")], "workload": workload + ", " + mode,
         "dram_bytes_per_launch": int(sum(a + b for a, b in zip(rd, wr)) / len(rd)), "dram_bytes_read": rd, "dram_bytes_write": wr,
         "gpu_time_us": t_us, "dram_throughput_pct_of_peak": pct, "kernel_source_sha": bench.kernel_source_sha(), "source": how}
    path = os.path.join(ROOT, "profiles", "traffic_%s_%s.json" % (workload, mode))
    json.dump(d, open(path, "w"), indent=1)
    print(open(path).read())


if __name__ == "__main__":
    main()
