"""spike_petsc_b200 -- B200-native (sm_100a) SPIKE banded factor/solve behind the PETSc-shaped
PC/KSP surface of spikegpu/spike-petsc.

Only the hot path lives here: `csrc/` (hand-written CUDA kernels + the C ABI of include/spike_b200.h),
`host/` (C mirror of PCBANDED / KSPREORDER that binds the C ABI into PETSc-style ops tables) and
this thin ctypes layer used by tests and bench.py.  There is no CPU fallback: importing works
anywhere, but every compute call raises unless libspike_b200.so is built and a B200 is present.
"""
from .capi import (Spike, SpikeError, lib, library_path, exported_symbols, GMRES, BCGS,  # noqa: F401
                   LAYOUT_ROWS, LAYOUT_DIAGS, MEM_HOST, MEM_DEVICE)
from .sharded import ShardedSpike, shard_rows  # noqa: F401

__all__ = ["Spike", "SpikeError", "lib", "library_path", "exported_symbols", "GMRES", "BCGS",
           "LAYOUT_ROWS", "LAYOUT_DIAGS", "MEM_HOST", "MEM_DEVICE", "ShardedSpike", "shard_rows"]
