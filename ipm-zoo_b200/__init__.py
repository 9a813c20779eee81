"""ipm-zoo_b200 -- B200-native (sm_100a) numerical interior-point hot path of albfre/ipm-zoo.

The product is the C-ABI shared library ``libipmz_b200.so`` (include/ipmz.h) built from
``csrc/*.cu`` by ``build.sh``.  This Python package is only the ctypes binding used by
tests/, bench.py and __graft_entry__.py; the reference-facing host layer is C++
(``host/``).  There is no CPU fallback: importing works anywhere, but every compute call
fails loudly without a CUDA device or without the built library.
"""
from .capi import (  # noqa: F401
    AUGMENTED, NORMAL, FULL, DUAL_NORMAL, EQ_OFF, EQ_SLACKED_SLACKS, EQ_NONE, EQ_REGULARIZATION, EQ_PENALTY, NONE, LOWER, UPPER, BOTH,
    IpmzError, Problem, Options, Result, Solver, BatchSolver, Factor,
    ldlt_decomposition, overwriting_solve_ldlt, symmetric_indefinite_factorization, overwriting_solve_bunch_kaufman, bk_factor_time, schedule_check, assembly_schedule_check, full_layout, solve_group, lib, lib_path, build, device_count, shard_range, pinned_empty, launch_count, fp64_peak_tflops,
    EXPORTED_SYMBOLS,
)
