"""CPU oracle binding — TEST INFRASTRUCTURE ONLY (see oracle/hr_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from .binding import OracleCalc, RefCalc, kernels, lib_path, num_threads, ref_available, set_num_threads, ref_lib_path  # noqa: F401
