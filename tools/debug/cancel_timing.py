#!/usr/bin/env python
"""Debug aid: how long a cancelled / an uncancelled 8192-spp Cornell 1080p render takes, sliced and unsliced."""
import ctypes as C, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from racer_tracer_b200 import harness

cfg = harness.load_config(os.path.join(ROOT, "tests", "golden", "config.yml"))
w, h = 1920, 1080
job = harness.prepare_job(os.path.join(ROOT, "tests", "golden", "scenes", "cornell_box.yml"), cfg, w, h)
r = harness.CudaRenderer([0])
r.upload(job)
out = np.empty((h, w, 3))
r.render(harness.make_params(w, h, 8, 20, seed=1, specialize=2), out=out)
for slices in (None, "1", "2"):
    if slices is None: os.environ.pop("RC_SLICES", None)
    else: os.environ["RC_SLICES"] = slices
    for spp in (1024, 8192):
        p = harness.make_params(w, h, spp, 20, seed=1, specialize=2)
        t0 = time.perf_counter(); r.render(p, out=out); t1 = time.perf_counter()
        flag = C.c_int32(0)
        t2 = time.perf_counter(); r.render(p, cancel=flag, out=out); t3 = time.perf_counter()
        flag = C.c_int32(0)
        tm = threading.Timer(0.03, lambda: setattr(flag, "value", 1))
        t4 = time.perf_counter(); tm.start(); r.render(p, cancel=flag, out=out); t5 = time.perf_counter(); tm.join()
        st = r.stats()
        print(f"RC_SLICES={slices} spp={spp}: plain {t1 - t0:.3f} s, flag never raised {t3 - t2:.3f} s, raised at 30 ms {t5 - t4:.3f} s, launches {st.kernel_launches}", flush=True)
