"""smartdenovo_b200 -- B200-native (sm_100a) kernels behind SMARTdenovo's `wtzmo` overlapper.

The product is the C host `smartdenovo_b200/bin/wtzmo` plus `smartdenovo_b200/lib/libzmo_b200.so`
(C ABI in include/zmo_b200.h).  This Python package is a thin ctypes mirror of that ABI used by the
tests and bench.py; it contains no compute and no CPU fallback: importing works anywhere, creating a
context without a Blackwell GPU raises ZmoError.
"""
from .api import (Zmo, ZmoError, ZmoParams, lib_path, load_lib, pack_reads, default_params,
                  PKG_DIR, REPO_DIR, wtzmo_path)

__all__ = ["Zmo", "ZmoError", "ZmoParams", "lib_path", "load_lib", "pack_reads", "default_params",
           "PKG_DIR", "REPO_DIR", "wtzmo_path"]
