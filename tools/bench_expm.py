#!/usr/bin/env python
"""Batched matrix-exponential throughput (the second half of BASELINE.json's metric): device-resident timing of the
standalone Pade-13 kernel `k_expm` through `qocb_expm_batched_time`, anti-Hermitian inputs -i H dt, "soft" (one-norm 1.5,
s = 0) and "stiff" (one-norm 12, s = 2).  GFLOP/s by the algorithmic count 8 n^3 (6 + 4/3 + s).
Usage (GPU box): python tools/bench_expm.py > gpurun_out/expm_batched.jsonl"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qoc_b200 import _lib  # noqa: E402

lib = _lib.load()
PEAK = 37.1
for n, batch in ((2, 1 << 20), (4, 1 << 18), (8, 1 << 16), (16, 1 << 14), (32, 1 << 12), (64, 1 << 12)):
    for label, norm, s in (("soft", 1.5, 0), ("stiff", 12.0, 2)):
        ms = ctypes.c_double()
        rc = lib.qocb_expm_batched_time(n, batch, ctypes.c_double(norm), 10, ctypes.byref(ms), 0)
        if rc != 0:
            print(json.dumps({"n": n, "error": lib.qocb_last_error(None).decode()}))
            continue
        flops = 8.0 * n ** 3 * (6 + 4.0 / 3 + s) * batch
        gbytes = 32.0 * n * n * batch / 1e9
        print(json.dumps({"probe": "expm_batched", "n": n, "batch": batch, "case": label, "squarings": s, "ms": ms.value,
                          "gflops": flops / ms.value / 1e6, "frac_fp64_peak": flops / ms.value / 1e9 / PEAK,
                          "hbm_gbs_algorithmic": gbytes / (ms.value * 1e-3), "matrices_per_s": batch / (ms.value * 1e-3)}))
