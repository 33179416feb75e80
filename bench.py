#!/usr/bin/env python
"""bench.py -- headline benchmark of the pq-vector B200 hot path (BASELINE.json config[1]):
brute-force squared-L2 top-100 of ONE query over 10M x 768 synthetic f32 rows resident in HBM,
results bit-identical to the reference loop (src/ivf/search.rs:112-141).

  python bench.py --gpus 1 --steps K --warmup W               (our arm, one JSON line on stdout)
  python -m torch.distributed.run ... bench.py --gpus N ...   (one rank per GPU; weak scaling: every
                                                               rank holds its own 10M x 768 slice)
  python bench.py --impl reference ...                        (the reference's CPU loop, oracle port)

A step = one query scanned over every resident row (all ranks), per-rank heap-entrant candidates
exchanged with one all-gather, reference heap replayed.  `value` is measured with the inputs already in
HBM (device-side loop, CUDA events on the launch stream); `e2e` goes through the public call with HOST
query/result buffers every step."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS_PER_GPU = 10_000_000
DIM = 768
K = 100
DATA_SEED = 1234   # benches/bench_util.rs:29
QUERY_SEED = 7     # benches/bench_util.rs:61-64 random_query(dim, seed 7)
METRIC = "queries/sec, brute-force L2 top-100 over 10M x 768 f32 per GPU (HBM GB/s + % roofline alongside)"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help="rows per GPU (debug; default = BASELINE config)")
    ap.add_argument("--dim", type=int, default=DIM)
    ap.add_argument("--k", type=int, default=K)
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cold", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: candidate exchange over NVLink peer memory (default; falls back to nccl if the IPC "
                         "set-up fails) or one NCCL all-gather per query")
    ap.add_argument("--cold-rows", type=int, default=2_000_000)
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(scan_bytes):
    """dram bytes per launch of the scan kernel from the committed ncu --set full capture -- only when this run's launch
    is the shape that was captured (the capture is of the BASELINE config); null for any other --rows / --dim."""
    p = os.path.join(ROOT, "profiles", "scan_kernel_ncu.json")
    if os.path.exists(p):
        try:
            cap = json.load(open(p))
            return cap.get("dram_bytes_per_launch") if cap.get("algorithmic_bytes_per_launch", ROWS_PER_GPU * DIM * 4) == scan_bytes else None
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 50 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference loop) -- bounded sample, extrapolated linearly in rows
# ------------------------------------------------------------------------------------------------------
def cpu_baseline(sample_rows: int, dim: int, k: int, full_rows: int, reps: int, warm: int = 1):
    import oracle as O
    host = O.synth(sample_rows, dim, DATA_SEED)
    queries = O.synth(reps + warm, dim, QUERY_SEED)
    cores = os.cpu_count() or 1
    scale = full_rows / sample_rows

    def timed(workers):
        ts = []
        for i in range(reps + warm):
            t0 = time.perf_counter()
            O.scan_topk_mt(host, queries[i], k, 0, workers)
            dt = time.perf_counter() - t0
            if i >= warm:
                ts.append(dt)
        return float(np.median(ts)), ts

    t1, ts1 = timed(1)
    tall, _ = timed(cores)
    base = {
        "value": 1.0 / (t1 * scale), "unit": UNIT, "cores": 1, "kind": "port",
        "sample": (f"{sample_rows} x {dim} rows of the same synthetic stream in RAM, median of {reps} single-query "
                   f"scans after {warm} warm-up, 1 thread (the reference re-rank loop src/ivf/search.rs:115-127 is "
                   f"serial); QPS extrapolated x{scale:g} in rows to {full_rows}"),
        "sample_seconds_per_scan": t1,
        "gbs": sample_rows * dim * 4 / t1 / 1e9,
        "all_cores": {"value": 1.0 / (tall * scale), "cores": cores,
                      "note": "charitable: same loop split over all host threads (not what the reference does)",
                      "gbs": sample_rows * dim * 4 / tall / 1e9},
    }
    return base, ts1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    sample = args.cpu_sample_rows
    host = O.synth(sample, args.dim, DATA_SEED)
    queries = O.synth(args.steps + args.warmup, args.dim, QUERY_SEED)
    scale = args.rows / sample
    for i in range(args.warmup):
        O.scan_topk_mt(host, queries[i], args.k, 0, 1)
    t0 = time.perf_counter()
    for i in range(args.steps):
        O.scan_topk_mt(host, queries[args.warmup + i], args.k, 0, 1)
    dt = (time.perf_counter() - t0) / args.steps
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    O.scan_topk_mt(host, queries[0], args.k, 0, cores)
    tall = time.perf_counter() - t0
    qps = 1.0 / (dt * scale)
    sample_txt = (f"each step = the reference loop (oracle port, 1 thread: src/ivf/search.rs:115-127 is serial) over "
                  f"{sample} x {args.dim} rows in RAM; seconds scaled x{scale:g} in rows to {args.rows}")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * scale * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"brute-force L2 top-{args.k}, 1 query, {args.rows} x {args.dim} f32 (BASELINE configs[1])",
                   "rows": args.rows, "dim": args.dim, "k": args.k,
                   "value_definition": "as the GPU arm: global queries/s x n_gpus, row-normalised (a serial scan of n_gpus x rows "
                                       "rows takes n_gpus x as long, so the figure is queries/s over `rows` rows at every N)"},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample_txt,
                         "gbs": sample * args.dim * 4 / dt / 1e9,
                         "all_cores": {"value": 1.0 / (tall * scale), "cores": cores,
                                       "note": "charitable split over all host threads, 1 run"}},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import pq_vector_b200 as P
    from pq_vector_b200.sharded import ShardedTopk

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n, dim, k = args.rows, args.dim, args.k
    pos_base = rank * n
    ctx = P.Context([local])
    ds = ctx.dataset(dim, n)
    ds.fill_synthetic(n, DATA_SEED, stream_first_row=pos_base)   # rank r holds rows [r*n, (r+1)*n) of one stream
    flags = P.PQV_SQRT                                            # TopkBuilder semantics (search.rs:129-140)

    nq = args.steps + args.warmup
    # query stream: same generator, query seed 7 (row i of that stream = query i)
    qds = ctx.dataset(dim, nq)
    qds.fill_synthetic(nq, QUERY_SEED)
    queries = qds.read(0, nq)
    qds.drop()

    sharded = ShardedTopk(lambda q, k_, f_, pb: ds.l2_topk_candidates(q, k_, f_, pb), pos_base, dev)
    exchange = "none" if world == 1 else "nccl"
    if world > 1 and args.exchange == "p2p":
        try:
            sharded.enable_p2p(ctx, ds)
            ok = 1.0
        except Exception as e:  # noqa: BLE001  (IPC not available in this container, ...)
            print(f"[bench] rank {rank}: peer exchange set-up failed ({e}); using the NCCL all-gather", file=sys.stderr)
            ok = 0.0
        t_ok = torch.tensor([ok], dtype=torch.float64, device=dev)
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)     # all ranks or none
        if float(t_ok.item()) < 1.0:
            sharded._p2p = None
        else:
            exchange = "p2p"

    sampler = ClockSampler(local) if rank == 0 else None

    # ---- (1) resident, device-side: K scans back to back, CUDA events on the launch stream --------------
    ds.bench_scan(queries[0], k, flags, max(args.warmup, 3))
    barrier()
    if sampler:
        sampler.start()
    t0 = time.perf_counter()
    ds.bench_scan(queries[args.warmup], k, flags, args.steps)
    tm = ctx.last_timing()
    torch.cuda.synchronize()
    barrier()
    wall_dev = (time.perf_counter() - t0) / args.steps
    step_ms = max_over_ranks(tm["total_ms"])      # device time per step (scan + merge + filter), max over ranks
    scan_ms = max_over_ranks(tm["scan_ms"])
    post_ms = tm["post_ms"]

    # ---- (2) end to end through the public call, host buffers every step -----------------------------------
    # N=1: the plain public call (pqv_l2_topk); N>1: per-rank candidates + one all-gather + replay
    search = (lambda q: ds.l2_topk(q, k, flags)) if world == 1 else (lambda q: sharded.search(q, k, flags))
    for i in range(args.warmup):
        search(queries[i])
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    res = None
    for i in range(args.steps):
        res = search(queries[args.warmup + i])
        t = ctx.last_timing()
        if exchange == "p2p":   # query in; the exchanged block of all ranks (world x (1 + cap) keys) out
            h2d += dim * 4
            d2h += world * (sharded.cap + 1) * 8
        else:
            h2d += dim * 4 + (sharded.cap + 1) * 8 * (world > 1)
            d2h += 8 * (1 + max(t["entrants"], 8192)) + (sharded.last_gather_bytes if world > 1 else 0)
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    clocks = sampler.stop() if sampler else None

    # ---- (3) cold path: rows streamed from pinned host memory through pqv_topk_stream_* (PCIe-bound) ------
    cold = None
    if rank == 0 and world == 1 and not args.no_cold:
        try:
            m = min(n, args.cold_rows)
            host_rows = torch.empty((m, dim), dtype=torch.float32, pin_memory=True)
            hr = host_rows.numpy()
            step = 1 << 17
            for s0 in range(0, m, step):
                hr[s0:s0 + step] = ds.read(s0, min(step, m - s0))
            batch = 1 << 16
            for rep in range(2):
                t0 = time.perf_counter()
                st = ctx.topk_stream(queries[0], k, P.PQV_SUM_SEQ)
                for s0 in range(0, m, batch):
                    st.push(hr[s0:s0 + batch])
                cr, cd = st.finish()
                dt = time.perf_counter() - t0
            cold = {"rows": m, "batch_rows": batch, "seconds": dt, "gbs": m * dim * 4 / dt / 1e9,
                    "qps_extrapolated_to_workload": 1.0 / (dt * n / m),
                    "note": "VectorTopKExec-style: every batch copied host->device inside the timed region "
                            "(pinned source, copy overlapped with the previous batch's scan); PCIe-bound"}
            del host_rows
        except Exception as e:  # pinned allocation can fail on small hosts; the headline does not depend on it
            cold = {"error": str(e)[:200]}

    # ---- parity spot check of the last e2e result against the oracle (rank 0, outside the timed region) ----
    parity = None
    if rank == 0:
        import oracle as O
        rows, dd = res
        ok = True
        for r_, d_ in zip(rows[:5].tolist(), dd[:5]):
            v = O.synth(1, dim, DATA_SEED, first_row=int(r_))[0]
            ok &= bool(np.sqrt(O.squared_l2_unroll4(queries[nq - 1], v)).view(np.uint32) == d_.view(np.uint32))
        parity = {"checked": "top-5 distances of the last step recomputed by the oracle from regenerated rows",
                  "bit_exact": ok, "results": int(rows.size), "ascending": bool(np.all(np.diff(dd) >= 0))}

    if rank == 0:
        peak, peak_src = measured_peak()
        scan_bytes = n * dim * 4
        achieved = scan_bytes / (scan_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC,
            "value": world / (step_ms * 1e-3),
            "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"brute-force L2 top-{k}, 1 query per step, {n} x {dim} f32 rows per GPU resident in HBM "
                            f"(BASELINE configs[1]); uniform[0,1) 24-bit grid, data seed {DATA_SEED}, query seed {QUERY_SEED}",
                "rows_per_gpu": n, "global_rows": n * world, "dim": dim, "k": k,
                "value_definition": "global queries/s x n_gpus (every query scans all n_gpus x rows_per_gpu rows; "
                                    "row-normalised so that N=1 is plain queries/s on 10M x 768)",
                "global_qps": 1.0 / (step_ms * 1e-3),
                "l2_policy": f"inputs ({scan_bytes / 1e9:.2f} GB per GPU per step) are larger than L2 (126 MB); no flush needed",
                "tie_order": "reference BinaryHeap replay (bit-exact row order)",
                "sharding": ("single GPU" if world == 1 else
                             "contiguous row ranges; per-rank heap-entrant candidates exchanged " +
                             ("by peer writes over NVLink from the scan's tail kernel (pqv_peer.cuh)" if exchange == "p2p"
                              else "with one NCCL all-gather")),
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(scan_bytes), "peak_source": peak_src, "kernel": "l2_scan_topk_kernel<ORDER=0,VEC4,dense,WARPS=8,RB=8,CBV=2,MINB=2>",
                         "algorithmic_bytes_per_launch": scan_bytes, "kernel_ms": scan_ms, "post_kernels_ms": post_ms},
            "e2e": {"value": world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps,
                    "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": e2e_s * 1e3,
                    "exchange": exchange,
                    "path": ("pqv_l2_topk: host query in, host (row_idx, distance) out" if world == 1 else
                             ("pqv_l2_topk_candidates_p2p (host query in, union of all ranks' candidate keys out) + "
                              "pqv_replay_candidates" if exchange == "p2p" else
                              "pqv_l2_topk_candidates (host query in, host candidate keys out) + one all-gather + "
                              "pqv_replay_candidates"))},
            "gpu_launches": 4 * args.steps,
            "clocks": clocks,
            "aggregate_gbs": world * scan_bytes / (step_ms * 1e-3) / 1e9,
            "host_wall_ms_per_step_device_loop": wall_dev * 1e3,
            "parity": parity,
            "e2e_cold_stream": cold,
        }
        if world == 1 and not args.no_cpu_baseline:
            base, _ = cpu_baseline(min(args.cpu_sample_rows, n), dim, k, n, reps=5)
            line["cpu_baseline"] = base
        emit(line)
    ds.drop()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout; everything else that writes to fd 1 while the bench
    runs (NCCL's "NCCL version ..." banner, library chatter) is diverted to stderr so the line stays alone."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
