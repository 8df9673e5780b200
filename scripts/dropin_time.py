import sys, time, io, contextlib, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from arterynetwork_b200 import variationalRegionGrowing as mod
shape = bench.WORKLOADS['c3']
data, vm8 = bench.host_phantom_via_device(shape, 0, 0)
vm = vm8.astype(np.int64)
mod.MAX_SECONDS = None
for mode in ('index','index','f64_dense'):
    mod.INTENSITY = mode
    v = vm.copy()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        mod.variationalRegionGrowing(data, v, maxSegmentSize=10**15)
    print(mode, round(time.perf_counter()-t0,3), json.dumps({k: round(x,3) for k,x in mod.LAST_RUN['host_seconds'].items()}))
