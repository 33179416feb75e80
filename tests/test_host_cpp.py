"""The compiled host side (pq_vector_b200/host/pq_vector.hpp: IndexBuilder / TopkBuilder / has_pq_vector_index /
vector_topk in C++ over Arrow C++ and libpqv.so) through its command-line driver.

CPU part: the embedded-index file format (parquet.rs:105-208, 542-611) written by the C++ side is read by the Python mirror
and by pyarrow, and the other way round; the column validation of read_parquet_with_embeddings (parquet.rs:210-303) and the
builders' argument errors carry the crate's messages; without a GPU the compute entry points fail loudly.
GPU part: builds and searches through the C++ side are identical -- blob bytes, row ids, distance bits -- to the Python
mirror (which tests/test_gpu_builders.py holds against the oracle)."""
import os
import shutil
import subprocess

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "pq_vector_b200", "host")
CLI = os.path.join(HOST, "pqv_host_cli")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(CLI):
        r = subprocess.run(["make", "-C", HOST], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]


def cli(*args, ok=True):
    r = subprocess.run([CLI, *map(str, args)], capture_output=True, text=True)
    if ok:
        assert r.returncode == 0, r.stderr
    return r


def write_table(path, emb, value_type=pa.float32(), extra_kv=None, compression="snappy"):
    emb = np.asarray(emb)
    ids = pa.array(np.arange(len(emb), dtype=np.int64))
    col = pa.array([row.tolist() for row in emb], type=pa.list_(value_type))
    t = pa.table({"id": ids, "embedding": col, "name": pa.array([f"r{i}" for i in range(len(emb))])})
    if extra_kv:
        t = t.replace_schema_metadata(extra_kv)
    pq.write_table(t, path, compression=compression, row_group_size=max(1, len(emb) // 3))
    return t


def toy_blob(n, dim=4, c=3):
    rng = np.random.default_rng(n)
    assign = rng.integers(0, c, n)
    ids = np.concatenate([np.nonzero(assign == j)[0] for j in range(c)]).astype(np.uint32)
    offsets = np.concatenate([[0], np.cumsum([(assign == j).sum() for j in range(c)])]).astype(np.uint64)
    return O.index_to_bytes(dim, rng.random((c, dim), dtype=np.float32), offsets, ids)


def fnv1a(b: bytes) -> int:
    h = 0xcbf29ce484222325
    for x in b:
        h = ((h ^ x) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def test_index_written_by_cpp_is_read_by_python_and_back(tmp_path):
    from pq_vector_b200 import builders as B
    emb = np.random.default_rng(1).random((50, 4), dtype=np.float32)
    for comp in ("snappy", "NONE"):
        p = str(tmp_path / f"t_{comp}.parquet")
        before = write_table(p, emb, extra_kv={"user_key": "kept"}, compression=comp)
        assert cli("has-index", p).stdout.strip() == "0" and not B.has_pq_vector_index(p)
        blob = toy_blob(50)
        (tmp_path / "blob.bin").write_bytes(blob)
        size0, groups0 = os.path.getsize(p), pq.read_metadata(p).num_row_groups
        cli("append-index", p, tmp_path / "blob.bin", "embedding")
        assert cli("has-index", p).stdout.strip() == "1" and B.has_pq_vector_index(p)
        got, column = B.read_index_payload(p)                       # Python reads what C++ wrote
        assert got == blob and column == "embedding"
        assert B.read_index_metadata(p)[0] == size0 - 8             # payload sits where the 8-byte footer tail was
        after = pq.read_table(p)                                     # the data pages did not move
        assert after.equals(before.replace_schema_metadata(after.schema.metadata))
        assert pq.read_metadata(p).metadata[b"user_key"] == b"kept"
        assert pq.read_metadata(p).num_row_groups == groups0 > 1
        r = cli("read-index", p, tmp_path / "back.bin")             # C++ reads it back
        assert r.stdout.strip() == "embedding" and (tmp_path / "back.bin").read_bytes() == blob
        # a second build replaces the key-values (parquet.rs:578-580) and the newest payload wins
        blob2 = toy_blob(50, c=5)
        B.append_index_inplace(p, blob2, "embedding")               # Python writes ...
        r = cli("read-index", p, tmp_path / "back2.bin")            # ... C++ reads
        assert (tmp_path / "back2.bin").read_bytes() == blob2
        assert pq.read_table(p).num_rows == 50


def test_payload_errors_carry_the_crates_messages(tmp_path):
    p = str(tmp_path / "t.parquet")
    write_table(p, np.ones((5, 2), np.float32))
    r = cli("read-index", p, tmp_path / "x", ok=False)
    assert r.returncode == 1 and "Missing pq-vector index metadata in parquet footer" in r.stderr
    # metadata pointing at garbage
    write_table(p, np.ones((5, 2), np.float32), extra_kv={"pq_vector_index_offset": "4", "pq_vector_embedding_column": "embedding"})
    r = cli("read-index", p, tmp_path / "x", ok=False)
    assert "Failed to decode pq-vector index payload at offset 4: Invalid pq-vector index magic" in r.stderr
    write_table(p, np.ones((5, 2), np.float32), extra_kv={"pq_vector_index_offset": "12x", "pq_vector_embedding_column": "embedding"})
    assert "invalid digit found in string" in cli("has-index", p, ok=False).stderr
    write_table(p, np.ones((5, 2), np.float32), extra_kv={"pq_vector_index_offset": "4", "pq_vector_embedding_column": "  "})
    assert "Embedding column name cannot be empty" in cli("has-index", p, ok=False).stderr
    (tmp_path / "tiny").write_bytes(b"PAR1")
    assert cli("append-index", tmp_path / "tiny", p, "embedding", ok=False).returncode == 1


def test_read_embeddings_validates_like_the_crate(tmp_path):
    rng = np.random.default_rng(2)
    emb = rng.random((40, 6), dtype=np.float32)
    p = str(tmp_path / "f32.parquet")
    write_table(p, emb)
    n, dim, h = cli("read-embeddings", p, "embedding").stdout.split()
    assert (int(n), int(dim)) == (40, 6) and int(h) == fnv1a(emb.tobytes())
    # Float64 items are narrowed with `as f32` (parquet.rs:288-291)
    emb64 = rng.random((17, 3))
    p64 = str(tmp_path / "f64.parquet")
    write_table(p64, emb64, value_type=pa.float64())
    n, dim, h = cli("read-embeddings", p64, "embedding").stdout.split()
    assert (int(n), int(dim)) == (17, 3) and int(h) == fnv1a(emb64.astype(np.float32).tobytes())
    assert "Column 'nope' not found" in cli("read-embeddings", p, "nope", ok=False).stderr
    assert "Embedding column name cannot be empty" in cli("read-embeddings", p, " ", ok=False).stderr
    assert "Embedding column is not a list array" in cli("read-embeddings", p, "id", ok=False).stderr

    def bad(col, msg):
        q = str(tmp_path / "bad.parquet")
        pq.write_table(pa.table({"embedding": col}), q)
        assert msg in cli("read-embeddings", q, "embedding", ok=False).stderr

    bad(pa.array([[1.0, 2.0], None], type=pa.list_(pa.float32())), "Embedding column contains null rows")
    bad(pa.array([[1.0, None]], type=pa.list_(pa.float32())), "Embedding values contain nulls")
    bad(pa.array([[1.0, 2.0], [1.0]], type=pa.list_(pa.float32())), "Embedding vectors have inconsistent dimensions")
    bad(pa.array([[1.0, 2.0], []], type=pa.list_(pa.float32())), "Embedding row has zero length")
    bad(pa.array([[1, 2]], type=pa.list_(pa.int32())), "Embedding values are not float32/float64")
    bad(pa.array([], type=pa.list_(pa.float32())), "Embedding column has no rows")


def test_builder_argument_errors(tmp_path):
    p = str(tmp_path / "t.parquet")
    write_table(p, np.ones((5, 2), np.float32))
    q = tmp_path / "q.bin"
    q.write_bytes(np.zeros(2, np.float32).tobytes())
    assert "k must be > 0" in cli("search", p, 0, 1, q, ok=False).stderr            # search.rs:66-69
    assert "nprobe must be > 0" in cli("search", p, 1, 0, q, ok=False).stderr        # search.rs:71-74
    assert "k must be set" in cli("search", p, "-", 1, q, ok=False).stderr           # search.rs:77
    assert "nprobe must be set" in cli("search", p, 1, "-", q, ok=False).stderr      # search.rs:78
    assert "Missing pq-vector index metadata in parquet footer" in cli("search", p, 1, 1, q, ok=False).stderr
    assert "max_iters must be > 0" in cli("build-inplace", p, "embedding", "-", 0, ok=False).stderr   # parquet.rs:89-91
    assert "n_clusters must be > 0" in cli("build-inplace", p, "embedding", 0, ok=False).stderr       # parquet.rs:93


def test_no_gpu_means_loud_failure(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = str(tmp_path / "t.parquet")
    write_table(p, np.random.default_rng(3).random((30, 4), dtype=np.float32))
    r = cli("build-inplace", p, "embedding", ok=False)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr
    assert cli("has-index", p).stdout.strip() == "0"                # nothing was written


# ---------------------------------------------------------------------------------------------------------------- GPU
def _search(path, k, nprobe, query, tmp_path):
    q = tmp_path / "query.bin"
    q.write_bytes(np.asarray(query, np.float32).tobytes())
    out = cli("search", path, k, nprobe, q).stdout.split()
    return [int(x) for x in out[0::2]], [int(x) for x in out[1::2]]


@pytest.mark.gpu
def test_cpp_build_and_search_equal_the_python_mirror(tmp_path):
    from pq_vector_b200 import builders as B
    rng = np.random.default_rng(5)
    emb = rng.random((3000, 32), dtype=np.float32)
    a, b, c = (str(tmp_path / f"{n}.parquet") for n in "abc")
    write_table(a, emb)
    shutil.copy(a, b)
    cli("build-inplace", a, "embedding", 16, 10, 7)                                  # C++
    B.IndexBuilder(b, "embedding").n_clusters(16).max_iters(10).seed(7).build_inplace()   # Python mirror
    assert B.read_index_payload(a) == B.read_index_payload(b)                        # same blob, same column
    assert pq.read_table(a).equals(pq.read_table(b))
    cli("build-new", b, "embedding", c, 16, 10, 7)                                   # C++: new file from an indexed source
    assert B.read_index_payload(c)[0] == B.read_index_payload(a)[0]
    assert pq.read_table(c).column("id").equals(pq.read_table(a).column("id"))
    dim, cent, offs, ids = O.index_from_bytes(B.read_index_payload(a)[0])
    for qi in range(2):                      # every CLI call is a fresh process (CUDA start-up): keep the matrix small
        query = emb[qi * 17] + (0.01 if qi % 2 else 0.0)
        for k, nprobe in ((10, 4), (1, 1), (1024, 64)):
            want = B.TopkBuilder(b, query).k(k).nprobe(nprobe).search()
            for path in (a, c):
                rows, dbits = _search(path, k, nprobe, query, tmp_path)
                assert rows == [r.row_idx for r in want]
                assert dbits == np.array([r.distance for r in want], np.float32).view(np.uint32).tolist()
            # and the oracle's restatement of search.rs:83-142 over the same index
            cand = O.candidate_rows(query, cent, offs, ids, nprobe)
            er, ed = O.topk_rerank_gather(query, emb, cand, k, 0, True)
            assert [r.row_idx for r in want] == er.tolist()
    q = tmp_path / "q3.bin"
    q.write_bytes(np.zeros(3, np.float32).tobytes())
    assert "Query dimension mismatch: expected 32, got 3" in cli("search", a, 1, 1, q, ok=False).stderr
    B.drop_resident()


@pytest.mark.gpu
def test_cpp_vector_topk_is_the_operators_answer(tmp_path):
    """the rows of src/df_vector/tests.rs:31-39 through topk_from_batches (no filter): ids in operator order"""
    from pq_vector_b200 import builders as B
    emb = np.array([(0, 0), (1, 0), (0, 2), (5, 5), (2, 2), (0.1, 0.1)], np.float32)
    p = str(tmp_path / "toy.parquet")
    t = write_table(p, emb)
    q = tmp_path / "q.bin"
    q.write_bytes(np.zeros(2, np.float32).tobytes())
    out = cli("vector-topk", p, "embedding", 4, q, "id").stdout.split()
    assert [int(x) for x in out[0::2]] == [0, 5, 1, 2]
    want = B.vector_topk(t.to_batches(max_chunksize=3), "embedding", np.zeros(2, np.float32), 4)
    assert want.column("id").to_pylist() == [0, 5, 1, 2]
    er, ed = O.topk_rerank(np.zeros(2, np.float32), emb, None, 4, 1, False)
    assert [int(x) for x in out[1::2]] == ed.view(np.uint32).tolist()
    # Float64 items, a null row and a short row: skipped as exec.rs:496-498, 526-528
    col = pa.array([[0.5, 0.5], None, [1.0], [0.25, 0.0], [3.0, 3.0]], type=pa.list_(pa.float64()))
    p2 = str(tmp_path / "mixed.parquet")
    pq.write_table(pa.table({"id": pa.array(np.arange(5, dtype=np.int64)), "embedding": col}), p2)
    out = cli("vector-topk", p2, "embedding", 10, q, "id").stdout.split()
    assert [int(x) for x in out[0::2]] == [3, 0, 4]


@pytest.mark.gpu
@pytest.mark.parametrize("rows,min_id,want_ids,fetched", [
    ([(0, 0), (1, 0), (0, 2), (5, 5), (2, 2), (0.1, 0.1)], 2, [5, 2], 4),                       # df_vector/tests.rs:16-104
    ([(0, 0), (.05, .05), (.2, .2), (1, 1), (1.1, 1.1), (1.4, 1.4)], 3, [3, 4], 3),             # df_vector/tests.rs:151-241
])
def test_cpp_vector_topk_indexed_passes_the_crates_sql_tests(tmp_path, rows, min_id, want_ids, fetched):
    """The crate's two SQL tests at the operator: `WHERE id >= m ORDER BY array_distance(vec, [0, 0]) LIMIT 2` over the
    file IndexBuilder::build_new wrote, options {nprobe: 64, max_candidates: None}; ids and the snapshot counters
    (`candidate_rows: 6`, `embeddings_fetched`) as pinned by the crate."""
    emb = np.array(rows, np.float32)
    src, out = str(tmp_path / "source.parquet"), str(tmp_path / "indexed.parquet")
    t = pa.table({"id": pa.array(np.arange(6, dtype=np.int32)),
                  "vec": pa.array([r.tolist() for r in emb], type=pa.list_(pa.float32()))})
    pq.write_table(t, src)
    cli("build-new", src, "vec", out)                                   # IndexBuilder::new(source, "vec").build_new(indexed)
    assert cli("has-index", out).stdout.strip() == "1" and pq.read_table(out).column("id").to_pylist() == list(range(6))
    q = tmp_path / "q.bin"
    q.write_bytes(np.zeros(2, np.float32).tobytes())
    mask = tmp_path / "mask.bin"                                        # the FilterExec predicate over the file's rows
    mask.write_bytes(np.packbits(np.arange(6) >= min_id, bitorder="little").tobytes())
    lines = cli("vector-topk-indexed", out, 2, 64, "-", q, mask).stdout.split("\n")
    assert lines[0].split() == ["6", str(fetched)]
    got = [int(l.split()[0]) for l in lines[1:] if l]
    assert got == want_ids
    # without the filter the closest two rows win; a cap of 3 candidates (rank order) is reported in the counters
    lines = cli("vector-topk-indexed", out, 2, 64, "-", q).stdout.split("\n")
    assert lines[0].split() == ["6", "6"] and [int(l.split()[0]) for l in lines[1:] if l][0] == 0
    lines = cli("vector-topk-indexed", out, 2, 64, 3, q).stdout.split("\n")
    assert lines[0].split()[1] == "3"
    q3 = tmp_path / "q3.bin"
    q3.write_bytes(np.zeros(3, np.float32).tobytes())
    assert "Query dimension mismatch: expected 2, got 3" in cli("vector-topk-indexed", out, 2, 64, "-", q3, ok=False).stderr
