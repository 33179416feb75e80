"""k-means++ init (1023 sweep + pick steps over a 50 000-row init set): the general sweep kernel ("off"), the default and a
range of L2 keep fractions, one line per configuration.  Centroid checksums must agree across all of them (same arithmetic).

    python benchmarks/probe_kpp_init.py [rows] [dim] [clusters]
"""
import hashlib, json, os, re, sys, tempfile, time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PQV_TRACE"] = "1"
import numpy as np
import pq_vector_b200 as P

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
clusters = int(sys.argv[3]) if len(sys.argv) > 3 else 1024

ctx = P.Context([0])
ds = ctx.dataset(dim, rows)
ds.fill_synthetic(rows, 1234)


def run(keep, reps=2):
    if keep is None:
        os.environ.pop("PQV_SWEEP_KEEP", None)
    else:
        os.environ["PQV_SWEEP_KEEP"] = str(keep)
    best, digest, phases = None, None, None
    for _ in range(reps):
        with tempfile.TemporaryFile(mode="w+b") as tf:
            saved = os.dup(2)
            os.dup2(tf.fileno(), 2)
            try:
                t0 = time.perf_counter()
                cent = ctx.kmeans_train(ds, clusters, max_iters=1, seed=42, sum_workers=16)
                wall = (time.perf_counter() - t0) * 1e3
            finally:
                os.dup2(saved, 2)
                os.close(saved)
            tf.seek(0)
            txt = tf.read().decode(errors="replace")
        m = re.search(r"k-means\+\+ on the device .*?: ([0-9.]+) ms", txt)
        ms = float(m.group(1)) if m else None
        ph = re.search(r"pick kernel, mean cycles: (.*)", txt)
        if best is None or (ms is not None and ms < best):
            best, phases = ms, ph.group(1) if ph else None
        digest = hashlib.sha1(np.ascontiguousarray(cent).tobytes()).hexdigest()[:12]
    print(json.dumps({"keep_pct": "default (0.55 x L2)" if keep is None else keep, "init_ms": best, "wall_ms_last": round(wall, 1),
                      "centroids": digest, "pick_phases": phases}), flush=True)


run("off")
run(None)
for keep in (0, 30, 40, 45, 50, 60, 100):
    run(keep)
