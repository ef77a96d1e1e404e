"""redtime_b200 -- B200-native Time-RG hot path of michaelbuehlmann/redTime.

The product is the C-ABI shared library ``libredtime_b200.so`` (CUDA kernels for sm_100a +
C++ host code, built in-tree by ``redtime_b200/csrc/Makefile``) and the drop-in executable
``redTime_b200``.  This package is a thin ctypes binding used by the tests and ``bench.py``;
it contains no numerics and no CPU fallback: every compute call fails loudly when the
library or a CUDA device is missing.
"""
from .binding import (RedTimeB200, RtrgError, Config, load_library, library_path,
                      read_run_dir, grid_info, table_T, table_G, table_windows, table_extrap,
                      assembly_terms, print_result, dfma_peak_tflops, dmma_peak_tflops, kshard_nccl_id,
                      LoopbackGroup, pack_cosmologies, Pipeline)

__all__ = ["RedTimeB200", "RtrgError", "Config", "load_library", "library_path",
           "read_run_dir", "grid_info", "table_T", "table_G", "table_windows", "table_extrap",
           "assembly_terms", "print_result", "dfma_peak_tflops", "dmma_peak_tflops", "kshard_nccl_id", "LoopbackGroup", "pack_cosmologies", "Pipeline"]
