"""dcs_b200 — thin ctypes view of the C-ABI (include/dcs_b200.h) and of the C++ host code
(host/g2o_util.h, host/synth.h) for the Python test / bench harness.

The product is the CUDA library and the C++ host; this module adds no arithmetic.  It never
imports anything from oracle/.  Loading fails loudly if libdcs_b200.so has not been built
(`python -c "import __graft_entry__ as g; g.build()"`), and every compute call fails with the
library's CUDA error when no B200 is usable: there is no CPU fallback.
"""
from .capi import (DcsError, Graph, Options, Solver, Summary, device_count, launch_count, lib_path,
                   host_lib_path, nccl_unique_id, partition, pinned_empty, version, load_library, load_host_library,
                   DECLARED_SYMBOLS, solve_batch)

__all__ = ["DcsError", "Graph", "Options", "Solver", "Summary", "device_count", "launch_count", "lib_path",
           "host_lib_path", "nccl_unique_id", "partition", "pinned_empty", "version", "load_library", "load_host_library", "DECLARED_SYMBOLS", "solve_batch"]
