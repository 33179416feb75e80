"""CPU-only: the C-ABI library loads, exports every symbol include/pqv.h declares, and refuses to
run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "pqv.h")).read()
    return sorted(set(re.findall(r"PQV_API[^;(]*?\b(pqv_\w+)\s*\(", txt)))


def test_header_declares_expected_surface():
    syms = header_symbols()
    for must in ["pqv_init", "pqv_destroy", "pqv_last_error", "pqv_dataset_create", "pqv_dataset_append",
                 "pqv_l2_topk", "pqv_l2_topk_gather", "pqv_topk_stream_begin", "pqv_topk_stream_push",
                 "pqv_topk_stream_finish", "pqv_kmeans_assign", "pqv_min_dist_update", "pqv_centroid_rank"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from pq_vector_b200 import _native as N
    assert os.path.exists(N.LIB_PATH)
    lib = C.CDLL(N.LIB_PATH)
    for s in header_symbols():
        assert getattr(lib, s) is not None, s
    assert sorted(N.SIGNATURES) == header_symbols()          # the ctypes table covers the whole header


def test_every_entry_point_cites_the_reference():
    txt = open(os.path.join(ROOT, "include", "pqv.h")).read()
    for cite in ["src/ivf/search.rs:112-141", "src/df_vector/exec.rs:257-277", "src/ivf/index.rs:193-201",
                 "src/ivf/index.rs:344-370", "src/ivf/index.rs:130-149", "src/ivf/index.rs:461-480"]:
        assert cite in txt


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pq_vector_b200 as P
    with pytest.raises(P.PqvError) as ei:
        P.Context()
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pq_vector_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "pqv_oracle" not in src, f


def test_hot_kernels_do_not_spill():
    """The HBM-bound scan kernels run at 2 CTAs/SM x 256 threads = 128 registers per thread; a change that pushes
    them into local-memory spills costs ~3 % of the scan (seen once).  Checked from the built cubin, no GPU needed."""
    import shutil
    import subprocess
    from pq_vector_b200 import _native as N
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--dump-resource-usage", N.LIB_PATH], capture_output=True, text=True).stdout
    usage = dict(re.findall(r"Function (\S+):\n\s*(REG:.*)", out))
    scans = {k: v for k, v in usage.items() if "l2_scan_topk_kernel" in k}
    assert scans, "scan kernels not found in libpqv.so"
    for name, u in scans.items():
        regs, stack, local = (int(re.search(p + r":(\d+)", u).group(1)) for p in ("REG", "STACK", "LOCAL"))
        if "ILi0ELb1ELb0ELi8ELi8ELi2ELi2E" in name or "ILi0ELb1ELb1ELi8ELi4ELi2ELi2E" in name or "ILi1ELb1ELb" in name:
            assert stack == 0 and local == 0 and regs <= 128, (name, u)   # the shipped vector variants
