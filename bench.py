#!/usr/bin/env python
"""bench.py -- driver benchmark of the pq-vector B200 hot path.

Headline (BASELINE.json configs[1], "C2"): brute-force squared-L2 top-100 of ONE query over 10M x 768 synthetic f32
rows resident in HBM, results bit-identical to the reference loop (src/ivf/search.rs:112-141).  The other BASELINE
configs ride in the same JSON line under "configs", each with its own roofline / e2e / cpu_baseline, the way the
reference's own bench covers build + indexed + un-indexed arms in one run (benches/query.rs:76-193):

  c3        10M x 768, IVF n_clusters=1024 build (k-means assign kernel) + nprobe=32 search        (N = 1)
  c4_shape  12.5M x 1536 per GPU, 1 query, top-100 (configs[3] = 100M x 1536 at N = 8)             (every N)
  c5        6.25M x 768 per GPU, 1024 queries per batch, top-10 (configs[4] = 50M x 768 at N = 8)  (every N)
  strong    C2's own 10M x 768 table cut N ways (strong scaling of the headline)                   (N > 1)

  python bench.py --gpus 1 --steps K --warmup W               (our arm, one JSON line on stdout)
  python -m torch.distributed.run ... bench.py --gpus N ...   (one rank per GPU; weak scaling: every rank holds its
                                                               own 10M x 768 slice of one synthetic stream)
  python bench.py --impl reference ...                        (the reference's CPU loop, oracle port, FULL 10M x 768)

A step = one query scanned over every resident row (all ranks), per-rank heap-entrant candidates exchanged, the
reference heap replayed.  `value` is measured with the inputs already in HBM (device-side loop, CUDA events on the
launch stream); `e2e` goes through the public call with HOST query/result buffers every step."""
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
UNIT = "queries/s"
ALL_LEGS = ("c3", "c5", "c4", "strong", "cold")


def metric_name(rows, dim, k):
    return (f"queries/sec, brute-force L2 top-{k} over {fmt_rows(rows)} x {dim} f32 per GPU "
            f"(HBM GB/s + % roofline alongside)")


def fmt_rows(n):
    if n % 1_000_000 == 0:
        return f"{n // 1_000_000}M"
    if n % 1000 == 0 and n >= 10_000:
        return f"{n / 1e6:g}M"
    return str(n)


def baseline_tag(rows, dim, k):
    return " (BASELINE configs[1])" if (rows, dim, k) == (ROWS_PER_GPU, DIM, K) else " (NOT a BASELINE config: --rows/--dim/--k override)"


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
    ap.add_argument("--legs", default="all",
                    help="comma list of the extra legs to run after the headline: c3,c5,c4,strong,cold | all | none")
    ap.add_argument("--no-cold", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: candidate exchange over NVLink peer memory (default; falls back to nccl if the IPC "
                         "set-up fails) or one NCCL all-gather per query")
    ap.add_argument("--cold-rows", type=int, default=2_000_000)
    ap.add_argument("--leg-scale", type=float, default=1.0, help="scale the rows of the extra legs (debug on small boxes)")
    ap.add_argument("--ref-rows", type=int, default=0, help="--impl reference: rows scanned per step (0 = all of --rows)")
    a = ap.parse_args()
    legs = set(ALL_LEGS) if a.legs == "all" else (set() if a.legs == "none" else set(a.legs.split(",")))
    if a.no_cold:
        legs.discard("cold")
    a.legs = legs
    return a


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return {"hbm_gbs": float(j["hbm_gbs"]), "tf_burst": float(j["bf16_tflops"]),
                    "tf_sustained": float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                    "source": "measured (MEASURED_PEAKS.json)"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(scan_bytes):
    """dram bytes per launch of the scan kernel from the committed ncu --set full capture (profiles/scan_kernel_ncu.json:
    a constant recorded under ncu, NOT measured by this run) -- only when this run's launch is the captured shape."""
    p = os.path.join(ROOT, "profiles", "scan_kernel_ncu.json")
    if os.path.exists(p):
        try:
            cap = json.load(open(p))
            if cap.get("algorithmic_bytes_per_launch", ROWS_PER_GPU * DIM * 4) == scan_bytes:
                return cap.get("dram_bytes_per_launch"), cap.get("source", "profiles/scan_kernel_ncu.json")
        except Exception:
            pass
    return None, None


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
# CPU side (oracle port of the reference loops) -- the ONLY users of oracle/ in this file
# ------------------------------------------------------------------------------------------------------
def synth_host(n_rows, dim, seed, first_row=0, threads=None):
    """n_rows x dim of the synthetic stream in host RAM, generated on all host threads (the generator is counter-based,
    so pieces are independent; ctypes releases the GIL)."""
    import ctypes as C
    import oracle as O
    out = np.empty((n_rows, dim), dtype=np.float32)
    threads = threads or (os.cpu_count() or 1)
    piece = max(1, (n_rows + threads * 4 - 1) // (threads * 4))
    fill = O.lib().pqo_synth_fill

    def work(t):
        for s in range(t * piece, n_rows, threads * piece):
            e = min(n_rows, s + piece)
            fill(out[s:e].ctypes.data_as(C.POINTER(C.c_float)), (first_row + s) * dim, (e - s) * dim, seed)

    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return out


def cpu_scan_baseline(sample_rows, dim, k, full_rows, reps, order=0, warm=1):
    """the reference re-rank loop (src/ivf/search.rs:115-127; order 1: src/df_vector/exec.rs:467-482) on a bounded sample"""
    import oracle as O
    host = synth_host(sample_rows, dim, DATA_SEED)
    queries = O.synth(reps + warm, dim, QUERY_SEED)
    cores = os.cpu_count() or 1
    scale = full_rows / sample_rows

    def timed(workers):
        ts = []
        for i in range(reps + warm):
            t0 = time.perf_counter()
            O.scan_topk_mt(host, queries[i], k, order, workers)
            dt = time.perf_counter() - t0
            if i >= warm:
                ts.append(dt)
        return float(np.median(ts))

    t1 = timed(1)
    tall = timed(cores)
    return {
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


def run_reference(args):
    """--impl reference: the reference's own CPU loop (oracle port, faithful = ONE thread: src/ivf/search.rs:115-127 is
    a serial loop) over the FULL rows x dim table held in host RAM, one scan per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    n = args.ref_rows or args.rows
    need = n * args.dim * 4
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = None
    shrunk = False
    if avail is not None and need > 0.8 * avail:   # slice only where the table exceeds host RAM
        n = int(0.6 * avail) // (args.dim * 4)
        shrunk = True
    scale = args.rows / n
    t_gen = time.perf_counter()
    host = synth_host(n, args.dim, DATA_SEED)
    t_gen = time.perf_counter() - t_gen
    queries = O.synth(args.steps + args.warmup, args.dim, QUERY_SEED)
    for i in range(args.warmup):
        O.scan_topk_mt(host, queries[i], args.k, 0, 1)
    t0 = time.perf_counter()
    for i in range(args.steps):
        O.scan_topk_mt(host, queries[args.warmup + i], args.k, 0, 1)
    dt = (time.perf_counter() - t0) / args.steps
    cores = os.cpu_count() or 1
    talls = []
    for i in range(3):
        t0 = time.perf_counter()
        O.scan_topk_mt(host, queries[i % len(queries)], args.k, 0, cores)
        talls.append(time.perf_counter() - t0)
    tall = float(np.median(talls))
    qps = 1.0 / (dt * scale)
    sample_txt = (f"each step = the reference loop (oracle port, 1 thread: src/ivf/search.rs:115-127 is serial) over "
                  f"{'ALL ' if not shrunk and n == args.rows else ''}{n} x {args.dim} rows in host RAM"
                  + ("" if scale == 1.0 else f"; seconds scaled x{scale:g} in rows to {args.rows} "
                     f"({'table exceeds host RAM' if shrunk else '--ref-rows'})"))
    line = {
        "impl": "reference", "metric": metric_name(args.rows, args.dim, args.k), "value": qps, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * scale * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"brute-force L2 top-{args.k}, 1 query per step, {args.rows} x {args.dim} f32"
                               f"{baseline_tag(args.rows, args.dim, args.k)}",
                   "rows": args.rows, "rows_scanned_per_step": n, "dim": args.dim, "k": args.k,
                   "host_table_gb": n * args.dim * 4 / 1e9, "host_table_generation_s": t_gen,
                   "value_definition": "as the GPU arm: global queries/s x n_gpus, row-normalised (a serial scan of n_gpus x rows "
                                       "rows takes n_gpus x as long, so the figure is queries/s over `rows` rows at every N)"},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample_txt,
                         "gbs": n * args.dim * 4 / dt / 1e9,
                         "all_cores": {"value": 1.0 / (tall * scale), "cores": cores,
                                       "note": "charitable split of the same loop over all host threads, median of 3"}},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
class Env:
    """rank / world / device plumbing shared by the legs"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(self, ok: bool) -> bool:
        if self.world == 1:
            return ok
        t = self.torch.tensor([1.0 if ok else 0.0], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return float(t.item()) >= 1.0


def synth_queries(ctx, dim, nq, seed=QUERY_SEED):
    qds = ctx.dataset(dim, nq)
    qds.fill_synthetic(nq, seed)
    q = qds.read(0, nq)
    qds.drop()
    return q


def scan_leg(env, P, ctx, n, dim, k, steps, warmup, exchange_pref, peaks, sampler=None, parity=True):
    """One brute-force single-query leg (C2 headline, the strong-scaling cut, the C4 shape): rank r holds rows
    [r*n, (r+1)*n) of the synthetic stream.  Returns the leg record (rank 0) -- device loop + e2e through the public call."""
    from pq_vector_b200.sharded import ShardedTopk
    world, rank, dev = env.world, env.rank, env.dev
    pos_base = rank * n
    ds = ctx.dataset(dim, n)
    ds.fill_synthetic(n, DATA_SEED, stream_first_row=pos_base)
    flags = P.PQV_SQRT                                            # TopkBuilder semantics (search.rs:129-140)
    nq = steps + warmup
    queries = synth_queries(ctx, dim, nq)
    sharded = ShardedTopk(lambda q, k_, f_, pb: ds.l2_topk_candidates(q, k_, f_, pb), pos_base, dev)
    exchange = "none" if world == 1 else "nccl"
    if world > 1 and exchange_pref == "p2p":
        try:
            sharded.enable_p2p(ctx, ds)
            ok = True
        except Exception as e:  # noqa: BLE001  (IPC not available in this container, ...)
            print(f"[bench] rank {rank}: peer exchange set-up failed ({e}); using the NCCL all-gather", file=sys.stderr)
            ok = False
        if env.all_ok(ok):     # all ranks or none
            exchange = "p2p"
        else:
            sharded._p2p = None

    # ---- (1) resident, device-side: `steps` scans back to back, CUDA events on the launch stream ----------
    ds.bench_scan(queries[0], k, flags, max(warmup, 3))
    env.barrier()
    if sampler:
        sampler.start()
    t0 = time.perf_counter()
    ds.bench_scan(queries[warmup], k, flags, steps)
    tm = ctx.last_timing()
    env.torch.cuda.synchronize()
    env.barrier()
    wall_dev = (time.perf_counter() - t0) / steps
    step_ms = env.max_over_ranks(tm["total_ms"])      # device time per step (scan + merge + filter), max over ranks
    scan_ms = env.max_over_ranks(tm["scan_ms"])
    post_ms = tm["post_ms"]
    dev_launches = int(tm["launches"]) * steps        # the library's own count for one scan (scan + 2 merges + filter)

    # ---- (2) end to end through the public call, host buffers every step -----------------------------------
    # N=1: the plain public call (pqv_l2_topk); N>1: per-rank candidates + exchange + replay
    search = (lambda q: ds.l2_topk(q, k, flags)) if world == 1 else (lambda q: sharded.search(q, k, flags))
    for i in range(warmup):
        search(queries[i])
    env.barrier()
    t0 = time.perf_counter()
    h2d = d2h = e2e_launches = 0
    res = None
    for i in range(steps):
        res = search(queries[warmup + i])
        t = ctx.last_timing()
        e2e_launches += int(t["launches"])
        if exchange == "p2p":   # query in; header + the live candidate keys of all ranks, written to pinned host memory by the pack kernel
            h2d += dim * 4
            d2h += (2 + world + int(t["entrants"])) * 8
        else:
            h2d += dim * 4 + (sharded.cap + 1) * 8 * (world > 1)
            d2h += 8 * (1 + max(t["entrants"], 8192)) + (sharded.last_gather_bytes if world > 1 else 0)
    env.barrier()
    e2e_s = env.max_over_ranks((time.perf_counter() - t0) / steps)
    clocks = sampler.stop() if sampler else None

    # ---- parity spot check of the last e2e result against the oracle (rank 0, outside the timed region) ----
    par = None
    if rank == 0 and parity:
        import oracle as O
        rows, dd = res
        ok = True
        for r_, d_ in zip(rows[:5].tolist(), dd[:5]):
            v = O.synth(1, dim, DATA_SEED, first_row=int(r_))[0]
            ok &= bool(np.sqrt(O.squared_l2_unroll4(queries[nq - 1], v)).view(np.uint32) == d_.view(np.uint32))
        par = {"checked": "top-5 distances of the last step recomputed by the oracle from regenerated rows (the full top-k "
                          "set at this size is checked against the oracle by tests/test_gpu_parity.py::test_full_scale_*)",
               "bit_exact": ok, "results": int(rows.size), "ascending": bool(np.all(np.diff(dd) >= 0))}
    rec = None
    if rank == 0:
        scan_bytes = n * dim * 4
        achieved = scan_bytes / (scan_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(scan_bytes)
        rec = {
            "value": world / (step_ms * 1e-3), "unit": UNIT, "ms_per_step": step_ms,
            "global_qps": 1.0 / (step_ms * 1e-3),
            "rows_per_gpu": n, "global_rows": n * world, "dim": dim, "k": k,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                         "traffic_source": (f"constant from the committed ncu --set full capture of this kernel instance and shape "
                                            f"({traffic_src}); not measured by this run" if traffic else None),
                         "peak_source": peaks["source"] + " hbm_gbs",
                         "kernel": "l2_scan_topk_kernel<ORDER=0,VEC4,dense,WARPS=8,RB=8,CBV=2,MINB=2>",
                         "algorithmic_bytes_per_launch": scan_bytes, "kernel_ms": scan_ms, "post_kernels_ms": post_ms},
            "e2e": {"value": world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d // steps,
                    "d2h_bytes_per_step": d2h // steps, "ms_per_step": e2e_s * 1e3, "global_qps": 1.0 / e2e_s,
                    "gpu_launches": e2e_launches, "exchange": exchange,
                    "path": ("pqv_l2_topk: host query in, host (row_idx, distance) out" if world == 1 else
                             ("pqv_l2_topk_p2p: host query in, host (row_idx, distance) out; scan + NVLink peer-write "
                              "candidate exchange + heap replay in one native call" if exchange == "p2p" else
                              "pqv_l2_topk_candidates (host query in, host candidate keys out) + one all-gather + "
                              "pqv_replay_candidates"))},
            "gpu_launches": dev_launches,
            "aggregate_gbs": world * scan_bytes / (step_ms * 1e-3) / 1e9,
            "host_wall_ms_per_step_device_loop": wall_dev * 1e3,
            "parity": par, "clocks": clocks,
        }
    return rec, ds, queries


def leg_cold(env, P, ctx, ds, queries, n, dim, k, cold_rows):
    """cold path: rows streamed from pinned host memory through pqv_topk_stream_* (VectorTopKExec feed), PCIe-bound;
    the pinned host->device peak is measured in the same run and is the roofline denominator"""
    torch = env.torch
    try:
        m = min(n, cold_rows)
        host_rows = torch.empty((m, dim), dtype=torch.float32, pin_memory=True)
        hr = host_rows.numpy()
        step = 1 << 17
        for s0 in range(0, m, step):
            hr[s0:s0 + step] = ds.read(s0, min(step, m - s0))
        # pinned H2D peak: the same buffer copied whole, CUDA events, best of 4 after one warm-up
        dst = torch.empty((m, dim), dtype=torch.float32, device=env.dev)
        peaks = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(host_rows, non_blocking=True)
            e1.record()
            e1.synchronize()
            if rep:
                peaks.append(m * dim * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        del dst
        h2d_peak = max(peaks)
        batch = 1 << 16
        ts = []
        for rep in range(6):   # 1 warm-up (staging allocations) + 5 timed
            t0 = time.perf_counter()
            st = ctx.topk_stream(queries[0], k, P.PQV_SUM_SEQ)
            for s0 in range(0, m, batch):
                st.push(hr[s0:s0 + batch])
            cr, cd = st.finish()
            if rep:
                ts.append(time.perf_counter() - t0)
        dt = float(np.median(ts))
        gbs = m * dim * 4 / dt / 1e9
        rr, rd = ds_prefix_topk(ds, queries[0], k, P.PQV_SUM_SEQ, m)
        same = bool(np.array_equal(cr, rr) and np.array_equal(cd.view(np.uint32), rd.view(np.uint32))) if rr is not None else None
        del host_rows
        return {"rows": m, "batch_rows": batch, "seconds": dt, "seconds_all": ts, "gbs": gbs,
                "roofline": {"bound": "pcie", "achieved": gbs, "peak": h2d_peak, "unit": "GB/s", "frac": gbs / h2d_peak,
                             "peak_source": "pinned host->device copy of the same buffer measured in this run (best of 4)",
                             "h2d_peak_samples": peaks},
                "qps_extrapolated_to_workload": 1.0 / (dt * n / m),
                "same_result_as_resident_scan_of_those_rows": same,
                "note": "VectorTopKExec-style: every batch copied host->device inside the timed region (pinned source, copy "
                        "overlapped with the previous batch's scan); median of 5 passes after one warm-up"}
    except Exception as e:  # pinned allocation can fail on small hosts; the headline does not depend on it
        return {"error": str(e)[:300]}


def ds_prefix_topk(ds, q, k, flags, m):
    """top-k of the first m resident rows (the stream leg's expected answer): via a gathered search over ids [0, m)"""
    try:
        ids = np.arange(m, dtype=np.uint32)
        return ds.l2_topk_gather(q, ids, k, flags)
    except Exception:
        return None, None


def leg_c3(env, P, ctx, ds, n, dim, peaks, cpu=True):
    """BASELINE configs[2]: IVF n_clusters=1024 build (k-means assign kernel) + nprobe=32 search over the resident table"""
    C, nprobe, k = 1024, 32, 100
    out = {"workload": f"{n} x {dim} f32, IVF n_clusters={C} build (max_iters 20, seed 42) + nprobe={nprobe} top-{k} search, 1 GPU"
                       + (" (BASELINE configs[2])" if (n, dim) == (ROWS_PER_GPU, DIM) else " (scaled-down shape)")}
    builds = []
    ix = None
    for _ in range(3):
        if ix is not None:
            ix.drop()
        t0 = time.perf_counter()
        ix = ctx.ivf_build(ds, n_clusters=C, max_iters=20, seed=42)
        builds.append({"seconds": time.perf_counter() - t0, **ix.build_stats()})
    best = min(builds, key=lambda b: b["seconds"])
    out["build"] = {"seconds": best["seconds"], "seconds_all": [b["seconds"] for b in builds],
                    "breakdown_ms": {k_: v for k_, v in best.items() if k_ != "seconds"},
                    "note": "pqv_ivf_build end to end (sample gather, k-means++ over 50k rows, <= 20 Lloyd sweeps over the 100k "
                            "sample, final assignment of all rows, inverted lists); first call pays scratch allocations"}
    cent = ix.centroids()
    # the assignment sweep (src/ivf/index.rs:189-206) as a device-resident loop: the tensor-core roofline record
    ctx.bench_assign(ds, cent, iters=1)
    ta = ctx.bench_assign(ds, cent, iters=3)
    flops = 2.0 * n * C * dim
    filt_tf = flops / (ta["filter_ms"] * 1e-3) / 1e12
    out["assign"] = {
        "value": n / (ta["total_ms"] * 1e-3), "unit": "rows/s", "ms_per_sweep": ta["total_ms"], "timing": ta,
        "roofline": {"bound": "tensor", "achieved": filt_tf, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                     "frac": filt_tf / peaks["tf_burst"], "traffic": None,
                     "peak_source": peaks["source"] + " bf16_tflops (burst; the filter kernel is timed alone)",
                     "kernel": "tc_rows_x_table_pair_kernel<AssignEpi> (tcgen05 filter of the sweep)",
                     "algorithmic_flops_per_launch": flops, "kernel_ms": ta["filter_ms"],
                     "whole_sweep": {"ms": ta["total_ms"], "algorithmic_bytes": n * dim * 4 + C * dim * 4 + n * 4,
                                     "gbs": (n * dim * 4 + C * dim * 4 + n * 4) / (ta["total_ms"] * 1e-3) / 1e9,
                                     "frac_of_hbm_peak": (n * dim * 4) / (ta["total_ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                     "tflops": flops / (ta["total_ms"] * 1e-3) / 1e12}}}
    # single-query search: TopkBuilder::search over the resident table + index
    nq = 100
    queries = synth_queries(ctx, dim, nq)
    for q in queries[:5]:
        ix.search(ds, q, k, nprobe)
    lat, cands, kern = [], [], []
    t_all = time.perf_counter()
    for q in queries:
        t0 = time.perf_counter()
        r, d = ix.search(ds, q, k, nprobe)
        lat.append(time.perf_counter() - t0)
        t = ctx.last_timing()
        cands.append(t["scan_bytes"] // (dim * 4))
        kern.append(t["scan_ms"])
    t_all = time.perf_counter() - t_all
    recall = []
    for q in queries[:10]:
        r, _ = ix.search(ds, q, k, nprobe)
        br, _ = ds.l2_topk(q, k)
        recall.append(len(set(r.tolist()) & set(br.tolist())) / max(len(br), 1))
    mc, mk = float(np.mean(cands)), float(np.mean(kern))
    gbytes = mc * (dim * 4 + 4)
    out["search"] = {
        "value": nq / t_all, "unit": UNIT, "ms_per_step": t_all / nq * 1e3, "mean_candidates": mc,
        "recall_at_k_vs_bruteforce": float(np.mean(recall)),
        "recall_note": "uniform i.i.d. data has no cluster structure: IVF pruning cannot help recall on this distribution; the "
                       "figure is stated so that the QPS is not read as a quality claim",
        "roofline": {"bound": "hbm", "achieved": gbytes / (mk * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": gbytes / (mk * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": None,
                     "peak_source": peaks["source"] + " hbm_gbs", "kernel": "l2_scan_topk_kernel<ORDER=0,VEC4,gather>",
                     "algorithmic_bytes_per_launch": gbytes, "kernel_ms": mk},
        "e2e": {"value": 1.0 / float(np.mean(lat)), "unit": UNIT, "ms_per_step": float(np.mean(lat)) * 1e3,
                "h2d_bytes_per_step": dim * 4, "d2h_bytes_per_step": 8 * 8193 + 4 * 8192 + 32,
                "path": "pqv_ivf_search: host query in, host (row_idx, distance) out, one round trip"}}
    # batched search: 1024 independent searches in one masked tensor-core pass
    bq = synth_queries(ctx, dim, 1024)
    sb = {}
    for kk in (10, k):
        ix.search_batch(ds, bq, kk, nprobe)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            br, bd, bc = ix.search_batch(ds, bq, kk, nprobe)
            ts.append(time.perf_counter() - t0)
        bt = ctx.last_batch_timing()
        same = True
        for i in range(0, 1024, 128):
            r1, d1 = ix.search(ds, bq[i], kk, nprobe)
            same &= bool(bc[i] == r1.size and np.array_equal(br[i, :bc[i]], r1) and
                         np.array_equal(bd[i, :bc[i]].view(np.uint32), d1.view(np.uint32)))
        tb = float(np.median(ts))
        sb[f"k{kk}"] = {"value": 1024 / tb, "unit": UNIT, "ms_per_batch": tb * 1e3, "timing": bt,
                        "identical_to_single_searches_on": 8 if same else -1}
    out["search_batch"] = sb
    if cpu:
        import oracle as O
        cores = os.cpu_count() or 1
        samp = min(20_000, n)
        host = synth_host(samp, dim, DATA_SEED)
        t0 = time.perf_counter()
        O.assign(host, cent, workers=cores)              # index.rs:193-201 runs on available_parallelism() threads
        dt_assign = time.perf_counter() - t0
        samp_c = min(200_000, n)
        host_c = synth_host(samp_c, dim, DATA_SEED)
        t0 = time.perf_counter()
        O.topk_rerank(queries[0], host_c, None, k, 0, True)   # search.rs:112-141 is serial
        dt_rerank = time.perf_counter() - t0
        out["assign"]["cpu_baseline"] = {
            "value": samp / dt_assign, "unit": "rows/s", "cores": cores, "kind": "port",
            "sample": f"{samp} rows x {C} centroids x {dim} on {cores} threads (index.rs:193-201 uses available_parallelism()): "
                      f"{dt_assign:.3f} s; a full sweep of {n} rows extrapolates to {dt_assign * n / samp:.1f} s"}
        out["search"]["cpu_baseline"] = {
            "value": 1.0 / (dt_rerank * mc / samp_c), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{samp_c} candidate rows x {dim}, 1 thread: {dt_rerank:.3f} s, scaled to {mc:.0f} candidates (distance loop "
                      "only; the reference also re-reads index and rows from Parquet per query)"}
    ix.drop()
    return out


def leg_c5(env, P, ctx, n, dim, peaks, cpu=True, exchange_pref="p2p"):
    """BASELINE configs[4] shape: rows sharded over the ranks (6.25M x 768 per GPU = 50M at N = 8), 1024 queries per batch,
    k = 10, VectorTopKExec arithmetic (PQV_SUM_SEQ).  Every rank answers the batch over its slice in one tensor-core pass,
    ONE all-gather, host merge, tie queries through the candidate exchange."""
    from pq_vector_b200.sharded import ShardedBatchTopk
    world, rank, dev = env.world, env.rank, env.dev
    nq, k, flags = 1024, 10, P.PQV_SUM_SEQ
    pos_base = rank * n
    ds = ctx.dataset(dim, n)
    ds.fill_synthetic(n, DATA_SEED, stream_first_row=pos_base)
    queries = synth_queries(ctx, dim, nq)
    sb = ShardedBatchTopk(lambda q, k_, f_, pb: ds.l2_topk_batch_keys(q, k_, f_, pb),
                          lambda q, k_, f_, pb: ds.l2_topk_candidates(q, k_, f_, pb), pos_base, dev,
                          tie_fn=lambda qi, q: ds.l2_topk_batch_tie_candidates(qi, q))
    exchange = "none" if world == 1 else "nccl"
    if world > 1 and exchange_pref == "p2p":
        try:
            sb.enable_p2p(ctx, ds, nq, k)
            ok = True
        except Exception as e:  # noqa: BLE001
            print(f"[bench] rank {rank}: peer exchange set-up failed ({e}); using the NCCL all-gather", file=sys.stderr)
            ok = False
        if env.all_ok(ok):
            exchange = "p2p"
        else:
            sb._p2p = None
            sb.single._p2p = None
    search = (lambda: ds.l2_topk(queries, k, flags)) if world == 1 else (lambda: sb.search(queries, k, flags))
    for _ in range(3):
        search()   # warm-up: allocations, row statistics / shadow, NCCL channels
    ts, dev_ms = [], []
    for _ in range(5):
        env.barrier()
        t0 = time.perf_counter()
        rows, dd, cnt = search()
        env.barrier()
        ts.append(env.max_over_ranks(time.perf_counter() - t0))
        dev_ms.append(env.max_over_ranks(ctx.last_batch_timing()["total_ms"]))
    tm = ctx.last_batch_timing()
    e2e = float(np.median(ts))
    dms = float(np.median(dev_ms))
    ok = True
    single = sb.single if world > 1 else None
    for q in range(0, nq, 256):
        r, d = (single.search(queries[q], k, flags) if single else ds.l2_topk(queries[q], k, flags))
        ok &= rows[q, :cnt[q]].tolist() == r.tolist() and dd[q, :cnt[q]].view(np.uint32).tolist() == d.view(np.uint32).tolist()
    rec = None
    if rank == 0:
        n_glob = n * world
        flops = 2.0 * n * nq * dim           # per GPU
        filt_tf = flops / (tm["filter_ms"] * 1e-3) / 1e12
        rec = {
            "workload": f"{n_glob} x {dim} f32 over {world} GPU(s) ({n} rows each), {nq} queries per batch, top-{k}, sequential-order "
                        f"sums (VectorTopKExec arithmetic, src/df_vector/exec.rs:529-533)"
                        + (" (BASELINE configs[4])" if (n_glob, dim) == (50_000_000, 768) else " (BASELINE configs[4] shape per GPU)"),
            "value": nq / (dms * 1e-3), "unit": UNIT, "ms_per_step": dms,
            "value_definition": "queries/s over ALL global rows, device time of one batch (prep + sample + filter + re-rank kernels, "
                                "max over ranks); e2e adds host buffers, the all-gather and the host merge",
            "roofline": {"bound": "tensor", "achieved": filt_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": filt_tf / peaks["tf_sustained"], "traffic": None,
                         "peak_source": peaks["source"] + " bf16_tflops_sustained (the filter runs inside a long step)",
                         "kernel": "tc_rows_x_table_pair_kernel<BatchEpi<FILTER>> (per GPU)",
                         "algorithmic_flops_per_launch": flops, "kernel_ms": tm["filter_ms"],
                         "hbm_gbs_of_the_pass": n * dim * 4 / (tm["filter_ms"] * 1e-3) / 1e9},
            "e2e": {"value": nq / e2e, "unit": UNIT, "ms_per_step": e2e * 1e3,
                    "h2d_bytes_per_step": nq * dim * 4,
                    "d2h_bytes_per_step": (nq * (k + 1) * 8 + nq * 8) * (world if exchange == "p2p" else 1) + (sb.last_gather_bytes if world > 1 else 0),
                    "exchange": exchange,
                    "path": ("pqv_l2_topk(n_queries = 1024): host queries in, host results out" if world == 1 else
                             ("pqv_l2_topk_batch_p2p: host queries in, host results out; per-rank tensor-core pass + NVLink peer-write "
                              "exchange of the key lists + host merge + tie replays in one native call" if exchange == "p2p" else
                              "pqv_l2_topk_batch_keys per rank + one all-gather + pqv_merge_batch_keys (+ candidate exchange for tie queries)"))},
            "rank0_batch_timing": tm, "replayed_queries": sb.last_replayed if world > 1 else tm["tie_queries"],
            "tflops_aggregate_e2e": 2.0 * n_glob * nq * dim / e2e / 1e12,
            "identical_to_single_query_search_on": 4 if ok else -1,
            "rank0_e2e_phases_ms": (getattr(sb, "last_phase_ms", None) if world > 1 else None)}
        if cpu and world == 1:
            import oracle as O
            samp, sq = min(250_000, n), 8
            host = synth_host(samp, dim, DATA_SEED)
            t0 = time.perf_counter()
            for i in range(sq):
                O.scan_topk_mt(host, queries[i], k, 1, 1)
            dt = (time.perf_counter() - t0) / sq
            cores = os.cpu_count() or 1
            t0 = time.perf_counter()
            for i in range(sq):
                O.scan_topk_mt(host, queries[i], k, 1, cores)
            dta = (time.perf_counter() - t0) / sq
            rec["cpu_baseline"] = {
                "value": 1.0 / (dt * n / samp), "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"{sq} of the queries x {samp} rows x {dim}, 1 thread, sequential-order loop (exec.rs:467-482 is serial and "
                          f"single-query: a batch is {nq} such scans); scaled x{n / samp:g} in rows",
                "all_cores": {"value": 1.0 / (dta * n / samp), "cores": cores, "note": "charitable split over all host threads"}}
    ds.drop()
    return rec


def run_ours(args):
    import pq_vector_b200 as P

    env = Env()
    rank, world, local = env.rank, env.world, env.local
    peaks = measured_peaks()
    n, dim, k = args.rows, args.dim, args.k
    ctx = P.Context([local])
    sampler = ClockSampler(local) if rank == 0 else None
    cpu = (world == 1 and not args.no_cpu_baseline)
    scale = args.leg_scale

    head, ds, queries = scan_leg(env, P, ctx, n, dim, k, args.steps, args.warmup, args.exchange, peaks, sampler)
    configs = {}
    if "cold" in args.legs and world == 1:
        configs["cold_stream"] = leg_cold(env, P, ctx, ds, queries, n, dim, k, args.cold_rows)
    if "c3" in args.legs and world == 1:
        configs["c3"] = leg_c3(env, P, ctx, ds, n, dim, peaks, cpu)
    ds.drop()
    if "strong" in args.legs and world > 1:
        ns = n // world
        rec, d2, _ = scan_leg(env, P, ctx, ns, dim, k, args.steps, args.warmup, args.exchange, peaks, None)
        d2.drop()
        if rank == 0:
            rec["workload"] = (f"strong scaling of the headline: the SAME {ns * world} x {dim} table cut into {world} slices of "
                               f"{ns} rows, 1 query per step, top-{k}")
            rec["value"] = rec["global_qps"]           # plain queries/s over the one table
            rec["e2e"]["value"] = rec["e2e"]["global_qps"]
            rec["value_definition"] = "queries/s over the whole table (compare with the N=1 headline: same table, one GPU)"
            configs["strong"] = rec
    if "c5" in args.legs:
        rec = leg_c5(env, P, ctx, int(6_250_000 * scale), dim, peaks, cpu, args.exchange)
        if rank == 0:
            configs["c5"] = rec
    if "c4" in args.legs:
        n4, d4 = int(12_500_000 * scale), 1536
        rec, d4s, _ = scan_leg(env, P, ctx, n4, d4, 100, max(args.steps // 2, 5), 3, args.exchange, peaks, None)
        d4s.drop()
        if rank == 0:
            rec["workload"] = (f"brute-force L2 top-100, 1 query per step, {n4 * world} x {d4} f32 over {world} GPU(s) ({n4} rows each)"
                               + (" (BASELINE configs[3])" if (n4 * world, d4) == (100_000_000, 1536) else
                                  " (BASELINE configs[3] shape per GPU; the 614 GB table needs >= 4 GPUs)"))
            rec["value"] = rec["global_qps"]
            rec["e2e"]["value"] = rec["e2e"]["global_qps"]
            rec["value_definition"] = "queries/s over all global rows (weak scaling: the table grows with N)"
            if cpu:
                rec["cpu_baseline"] = cpu_scan_baseline(min(500_000, n4), d4, 100, n4, reps=3)
            configs["c4_shape"] = rec

    if rank == 0:
        line = {
            "metric": metric_name(n, dim, k),
            "value": head["value"], "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"brute-force L2 top-{k}, 1 query per step, {n} x {dim} f32 rows per GPU resident in HBM"
                            f"{baseline_tag(n, dim, k)}; uniform[0,1) 24-bit grid, data seed {DATA_SEED}, query seed {QUERY_SEED}",
                "rows_per_gpu": n, "global_rows": n * world, "dim": dim, "k": k,
                "value_definition": "global queries/s x n_gpus (every query scans all n_gpus x rows_per_gpu rows; "
                                    f"row-normalised so that N=1 is plain queries/s on {fmt_rows(n)} x {dim})",
                "global_qps": head["global_qps"],
                "l2_policy": f"inputs ({n * dim * 4 / 1e9:.2f} GB per GPU per step) are larger than L2 (126 MB); no flush needed",
                "tie_order": "reference BinaryHeap replay (bit-exact row order)",
                "sharding": ("single GPU" if world == 1 else
                             "contiguous row ranges; per-rank heap-entrant candidates exchanged " +
                             ("by peer writes over NVLink from the scan's tail kernel (pqv_peer.cuh)" if head["e2e"]["exchange"] == "p2p"
                              else "with one NCCL all-gather")),
            },
            "roofline": head["roofline"],
            "e2e": head["e2e"],
            "gpu_launches": head["gpu_launches"],
            "clocks": head["clocks"],
            "aggregate_gbs": head["aggregate_gbs"],
            "host_wall_ms_per_step_device_loop": head["host_wall_ms_per_step_device_loop"],
            "parity": head["parity"],
            "configs": configs,
        }
        if cpu:
            line["cpu_baseline"] = cpu_scan_baseline(min(args.cpu_sample_rows, n), dim, k, n, reps=5)
        emit(line)
    ctx.close()
    if world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()


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
