#!/usr/bin/env python
"""Measured read ceilings of this GPU with libbb25's own streaming microkernel
(bb25_measure_read_bandwidth): L2-resident buffers (L2 -> SM) and buffers far beyond L2 (HBM).
Writes one JSON document to stdout (committed as profiles/r02/peaks_l2.json)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

if not os.path.exists(entry.SO):
    entry.build()
import torch  # noqa: E402

from bayesian_bm25_b200 import _lib  # noqa: E402

assert torch.cuda.is_available()
rows = []
for nbytes, iters in ((16 << 20, 100), (32 << 20, 60), (48 << 20, 40), (64 << 20, 30), (96 << 20, 20), (256 << 20, 8),
                      (1 << 30, 3), (4 << 30, 1), (8 << 30, 1)):
    g, ms = C.c_double(), C.c_double()
    _lib.check(_lib.lib().bb25_measure_read_bandwidth(0, nbytes, iters, 7, C.byref(g), C.byref(ms)))
    rows.append({"buffer_mb": nbytes / 2**20, "reads_per_launch": iters, "best_ms": ms.value, "gbs": g.value})
peaks = {}
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peaks = json.load(open(p))
print(json.dumps({
    "gpu": torch.cuda.get_device_name(0),
    "kernel": "bb25::stream_read_kernel: ld.global.nc.L1::no_allocate.v4, 4 independent 128-bit loads per thread, 148 x 8 CTAs x 256 threads",
    "l2_read_gbs": max(r["gbs"] for r in rows if r["buffer_mb"] <= 64),
    "hbm_read_gbs": max(r["gbs"] for r in rows if r["buffer_mb"] >= 4096),
    "driver_hbm_copy_gbs": peaks.get("hbm_gbs"),
    "rows": rows,
}, indent=1))
