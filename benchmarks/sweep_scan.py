#!/usr/bin/env python
"""Tuning sweep of the dense scan kernel variants (PQV_SCAN_VARIANT / PQV_SCAN_CTAS_PER_SM) on the C2 workload.
Prints scan-kernel ms and GB/s per variant and checks every variant returns the default variant's result."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
variants = [(0, 0), (0, 0), (1, 0), (4, 0), (5, 0), (9, 0), (10, 0), (11, 0), (12, 0), (4, 0), (0, 0)]
names = {0: "rb8 cbv1 minb2", 1: "rb16 cbv1 minb2", 2: "rb4 cbv1 minb3", 3: "rb4 cbv2 minb2", 4: "rb8 cbv2 minb2",
         5: "rb4 cbv1 minb4", 6: "rb2 cbv2 minb3", 7: "rb8 cbv1 minb1", 8: "rb16 cbv1 minb1",
         9: "rb16 cbv2 minb2", 10: "rb8 cbv3 minb2", 11: "rb4 cbv3 minb2", 12: "rb16 cbv2 minb1"}
ref = None
out = []
for v, occ in variants:
    os.environ["PQV_SCAN_VARIANT"] = str(v)
    if occ:
        os.environ["PQV_SCAN_CTAS_PER_SM"] = str(occ)
    else:
        os.environ.pop("PQV_SCAN_CTAS_PER_SM", None)
    ctx = P.Context([0])
    ds = ctx.dataset(dim, rows)
    ds.fill_synthetic(rows, 1234)
    q = ctx.dataset(dim, 1)
    q.fill_synthetic(1, 7)
    qv = q.read(0, 1)[0]
    r, d = ds.l2_topk(qv, 100, P.PQV_SQRT)
    if ref is None:
        ref = (r, d)
    ok = r.tolist() == ref[0].tolist() and d.view(np.uint32).tolist() == ref[1].view(np.uint32).tolist()
    ds.bench_scan(qv, 100, P.PQV_SQRT, 5)
    ms = ds.bench_scan(qv, 100, P.PQV_SQRT, 30)
    t = ctx.last_timing()
    rec = {"variant": v, "name": names[v], "ctas_per_sm_override": occ, "grid": t["grid"], "scan_ms": ms,
           "gbs": rows * dim * 4 / ms / 1e6, "post_ms": t["post_ms"], "same_result": ok}
    out.append(rec)
    print(json.dumps(rec), flush=True)
    ds.drop()
    ctx.close()
