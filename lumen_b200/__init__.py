"""lumen_b200: B200-native (sm_100a) replacement for Lumen's unidirectional path tracer with NEE + MIS.

Layout (only what the hot path needs):
  csrc/   CUDA kernels (LBVH build, wavefront traversal / shading / film) + the C-ABI of include/lumen_b200.h
  host/   C++ host shell: scene ingest (JSON / Mitsuba XML), camera, EXR, Integrator/PathB200 shim, headless main
  host.py, integrator.py   ctypes mirrors used by tests/ and bench.py
"""
from . import host  # noqa: F401
from . import integrator  # noqa: F401
