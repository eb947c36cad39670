"""amplisolve_b200 -- B200-native (sm_100a) implementation of AmpliSolve's hot path.

The product is libamplisolve_b200.so (CUDA kernels behind a C ABI, include/amplisolve_b200.h) and the
two drop-in programs under amplisolve_b200/bin; this package is the thin Python mirror used by the
tests and the benchmark.
"""
from .api import (ABSENT, CALL_DTYPE, AmpliSolveError, Context, calls_from_device, fisher_test, hash_iteration_order, lib,  # noqa: F401
                  pack_counts, shard_bounds, to_wire16, to_wire_packed, twin_links)
