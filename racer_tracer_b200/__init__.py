"""racer_tracer_b200 — B200 (sm_100a) path-tracing backend for racer-tracer.

csrc/      hand-written CUDA kernels + the C ABI (libracer_cuda.so, include/racer_cuda.h)
host/      C++ mirror of the reference's host-side Renderer interface (scene/config
           loading, flattening, BVH build, PNG output) over the C ABI
capi.py    ctypes mirror of the C ABI
harness.py Python stand-in for the Rust host, used by tests and bench.py
"""
__version__ = "0.1.0"
