import sys, time, os
sys.path.insert(0, os.getcwd())
import bench
from meep_nl_b200 import _cabi
class A: pass
a = A(); a.workload="vac3d"; a.n=256; a.dtype="f64"
b = _cabi.Backend()
gv, f = bench.build_problem(a, b)
t0=time.perf_counter(); f.init(); print("init %.3f s"%(time.perf_counter()-t0))
f.step(3)
st=f.stats(); print("pinned arrays", st.reserved[1], "bytes", st.reserved[2], "device bytes", st.device_bytes)
for k in range(3):
    t0=time.perf_counter(); f.sync_to_host(); t1=time.perf_counter(); f.sync_from_host(); t2=time.perf_counter()
    print("to_host %.1f ms  from_host %.1f ms"%((t1-t0)*1e3,(t2-t1)*1e3))
f.close()
