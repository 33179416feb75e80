#!/usr/bin/env python
"""Table load throughput: pqv_dataset_append from ordinary (pageable) host memory, the way a Rust caller hands over the
values buffer of each Arrow record batch."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pq_vector_b200 as P  # noqa: E402

dim = 768
ctx = P.Context([0])
out = {}
for batch_rows in (2048, 65536, 1_000_000):
    total = 2_000_000
    host = np.random.default_rng(0).random((batch_rows, dim), dtype=np.float32)
    ds = ctx.dataset(dim, total)
    ds.append(host[:16])
    t0 = time.perf_counter()
    done = 16
    while done + batch_rows <= total:
        ds.append(host)
        done += batch_rows
    dt = time.perf_counter() - t0
    out[f"batch_{batch_rows}"] = {"gbs": (done - 16) * dim * 4 / dt / 1e9, "seconds": dt, "rows": done - 16}
    # the bytes arrived intact (head, an interior stretch across staging-chunk borders, tail)
    assert np.array_equal(ds.read(16, 8), host[:8])
    if batch_rows >= 65536:
        assert np.array_equal(ds.read(16 + 2000, 3000), host[2000:5000])
        assert np.array_equal(ds.read(16 + batch_rows - 5, 5), host[-5:])
    ds.drop()
# the same large append through the driver's own pageable path
os.environ["PQV_APPEND_DIRECT"] = "1"
host = np.random.default_rng(0).random((1_000_000, dim), dtype=np.float32)
ds = ctx.dataset(dim, 1_000_016)
ds.append(host[:16])
t0 = time.perf_counter()
ds.append(host)
dt = time.perf_counter() - t0
out["batch_1000000_direct_pageable"] = {"gbs": host.nbytes / dt / 1e9, "seconds": dt}
ds.drop()
# cold streaming top-k (pqv_topk_stream_*, the VectorTopKExec feed) from PAGEABLE batches: staged lanes vs the driver's copy
del os.environ["PQV_APPEND_DIRECT"]
m, batch = 2_000_000, 65536
rows = np.random.default_rng(1).random((m, dim), dtype=np.float32)       # ordinary numpy memory
q = np.random.default_rng(2).random(dim, dtype=np.float32)
res = {}
for name, env in (("staged", None), ("direct_pageable", "1")):
    if env:
        os.environ["PQV_APPEND_DIRECT"] = env
    for rep in range(2):
        t0 = time.perf_counter()
        st = ctx.topk_stream(q, 10, P.PQV_SUM_SEQ)
        for s0 in range(0, m, batch):
            st.push(rows[s0:s0 + batch])
        r, d = st.finish()
        dt = time.perf_counter() - t0
    res[name] = (r.tolist(), d.view(np.uint32).tolist())
    out[f"stream_push_{name}"] = {"gbs": rows.nbytes / dt / 1e9, "seconds": dt, "batch_rows": batch}
out["stream_push_results_identical"] = res["staged"] == res["direct_pageable"]
print(json.dumps(out))
