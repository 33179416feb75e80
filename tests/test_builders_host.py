"""CPU-only: the host half of the reference-interface mirror (pq_vector_b200/builders.py) -- footer key-values, index
payload framing (parquet.rs:105-112, 151-174, 542-611), argument validation with the reference's error text, the
embedding reader's checks (parquet.rs:216-305).  Nothing here computes a distance; the GPU half is in
test_gpu_builders.py."""
import os
import struct

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import oracle as O
from pq_vector_b200 import builders as B


def _write(path, rows, extra=None, typ=pa.float32()):
    cols = {"id": pa.array(list(range(len(rows))), pa.int32()), "embedding": pa.array(rows, pa.list_(typ))}
    if extra:
        cols.update(extra)
    pq.write_table(pa.table(cols), path, compression="NONE")


def test_append_index_inplace_round_trip(tmp_path):
    """mirror of test_build_index_inplace_appends_footer (parquet.rs:623-660) for the file surgery alone"""
    path = str(tmp_path / "a.parquet")
    _write(path, [[0.0, 0.0], [1.0, 0.0], [0.0, 2.0]])
    before = pq.read_table(path)
    size0 = os.path.getsize(path)
    assert not B.has_pq_vector_index(path)
    with pytest.raises(B.PqVectorError, match="Missing pq-vector index metadata in parquet footer"):
        B.read_index_payload(path)
    cent = np.arange(1, 7, dtype=np.float32).reshape(2, 3)           # the blob of index.rs:495-511
    blob = O.index_to_bytes(3, cent, np.array([0, 3, 5], np.uint64), np.array([0, 2, 4, 1, 3], np.uint32))
    B.append_index_inplace(path, blob, "embedding")
    assert os.path.getsize(path) > size0                             # parquet.rs:649-650
    assert B.has_pq_vector_index(path)
    off, col = B.read_index_metadata(path)
    assert (off, col) == (size0 - 8, "embedding")                    # parquet.rs:566-567: index_offset = metadata_end
    raw = open(path, "rb").read()
    assert raw[off:off + 10] == b"PQ_VECTOR1" and struct.unpack("<Q", raw[off + 10:off + 18])[0] == len(blob)
    got, col = B.read_index_payload(path)
    assert got == blob and col == "embedding"
    assert pq.read_table(path).equals(before)                        # data pages untouched, file still valid parquet
    # a second build replaces the key-values instead of stacking them (parquet.rs:573-575)
    B.append_index_inplace(path, blob, "embedding")
    kv = pq.read_metadata(path).metadata
    assert int(kv[b"pq_vector_index_offset"]) > off and B.read_index_payload(path)[0] == blob


def test_corrupt_payloads(tmp_path):
    path = str(tmp_path / "a.parquet")
    _write(path, [[1.0, 2.0]])
    B.append_index_inplace(path, b"x" * 16, "embedding")
    off, _ = B.read_index_metadata(path)
    raw = bytearray(open(path, "rb").read())
    raw[off:off + 10] = b"NOT_VECTOR"
    open(path, "wb").write(raw)
    with pytest.raises(B.PqVectorError, match=f"Failed to decode pq-vector index payload at offset {off}: Invalid pq-vector index magic"):
        B.read_index_payload(path)


def test_builder_argument_errors(tmp_path):
    path = str(tmp_path / "a.parquet")
    _write(path, [[1.0, 2.0], [3.0, 4.0]])
    q = [0.0, 0.0]
    with pytest.raises(B.PqVectorError, match="^k must be > 0$"):            # search.rs:66-69
        B.TopkBuilder(path, q).k(0)
    with pytest.raises(B.PqVectorError, match="^nprobe must be > 0$"):       # search.rs:71-74
        B.TopkBuilder(path, q).nprobe(0)
    with pytest.raises(B.PqVectorError, match="^k must be set$"):            # search.rs:77
        B.TopkBuilder(path, q).nprobe(1).search()
    with pytest.raises(B.PqVectorError, match="^nprobe must be set$"):       # search.rs:78
        B.TopkBuilder(path, q).k(1).search()
    with pytest.raises(B.PqVectorError, match="^Missing pq-vector index metadata in parquet footer$"):
        B.TopkBuilder(path, q).k(1).nprobe(1).search()
    with pytest.raises(B.PqVectorError, match="^max_iters must be > 0$"):    # parquet.rs:89-91
        B.IndexBuilder(path, "embedding").max_iters(0).build_inplace()
    with pytest.raises(B.PqVectorError, match="^n_clusters must be > 0$"):   # parquet.rs:92-94
        B.IndexBuilder(path, "embedding").n_clusters(0).build_inplace()
    with pytest.raises(B.PqVectorError, match="^Embedding column name cannot be empty$"):   # mod.rs:24-26
        B.IndexBuilder(path, "  ").build_inplace()
    with pytest.raises(B.PqVectorError, match="^Column 'nope' not found$"):  # parquet.rs:233-235
        B.IndexBuilder(path, "nope").build_inplace()
    with pytest.raises(B.PqVectorError, match="^k must be > 0$"):
        B.vector_topk([], "embedding", q, 0)


def test_embedding_reader_checks(tmp_path):
    p = str(tmp_path / "a.parquet")
    _write(p, [[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])
    table, emb = B.read_parquet_with_embeddings(p, "embedding")
    assert emb.dtype == np.float32 and emb.tolist() == [[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]] and table.num_rows == 3
    _write(p, [[1.0, 2.0], [3.0, 4.0]], typ=pa.float64())                  # f64 narrowed (parquet.rs:288-291)
    assert B.read_parquet_with_embeddings(p, "embedding")[1].dtype == np.float32
    _write(p, [[1.0, 2.0], None])
    with pytest.raises(B.PqVectorError, match="Embedding column contains null rows"):
        B.read_parquet_with_embeddings(p, "embedding")
    _write(p, [[1.0, None]])
    with pytest.raises(B.PqVectorError, match="Embedding values contain nulls"):
        B.read_parquet_with_embeddings(p, "embedding")
    _write(p, [[1.0, 2.0], [3.0]])
    with pytest.raises(B.PqVectorError, match="Embedding vectors have inconsistent dimensions"):
        B.read_parquet_with_embeddings(p, "embedding")
    _write(p, [[1.0, 2.0], []])
    with pytest.raises(B.PqVectorError, match="Embedding row has zero length"):
        B.read_parquet_with_embeddings(p, "embedding")
    _write(p, [])
    with pytest.raises(B.PqVectorError, match="Embedding column has no rows"):
        B.read_parquet_with_embeddings(p, "embedding")
    with pytest.raises(B.PqVectorError, match="Embedding column is not a list array"):
        B.read_parquet_with_embeddings(p, "id")
    pq.write_table(pa.table({"embedding": pa.array([[1, 2]], pa.list_(pa.int32()))}), p)
    with pytest.raises(B.PqVectorError, match="Embedding values are not float32/float64"):
        B.read_parquet_with_embeddings(p, "embedding")


def test_dense_rows_of_every_list_flavour():
    q = 2
    a = pa.array([[1.0, 2.0], None, [3.0], [4.0, 5.0]], pa.list_(pa.float32()))
    v, idx = B._dense_rows(a, q)
    assert v.tolist() == [[1.0, 2.0], [4.0, 5.0]] and idx.tolist() == [0, 3]      # null + wrong length skipped
    v, idx = B._dense_rows(a.slice(2, 2), q)
    assert v.tolist() == [[4.0, 5.0]] and idx.tolist() == [1]
    a = pa.array([[1.0, 2.0], [3.0, 4.0]], pa.large_list(pa.float64()))
    v, idx = B._dense_rows(a, q)
    assert v.dtype == np.float64 and v.tolist() == [[1.0, 2.0], [3.0, 4.0]] and idx.tolist() == [0, 1]
    a = pa.array([[1.0, 2.0], None, [5.0, 6.0]], pa.list_(pa.float32(), 2))       # FixedSizeList (exec.rs:504-510)
    v, idx = B._dense_rows(a, q)
    assert v.tolist() == [[1.0, 2.0], [5.0, 6.0]] and idx.tolist() == [0, 2]
    v, idx = B._dense_rows(a.slice(1, 2), q)
    assert v.tolist() == [[5.0, 6.0]] and idx.tolist() == [1]
    v, idx = B._dense_rows(a, 3)
    assert v.shape == (0, 3) and idx.size == 0
    with pytest.raises(B.PqVectorError, match="Vector column must be list or fixed-size list"):
        B._dense_rows(pa.array([1.0, 2.0]), q)
    with pytest.raises(B.PqVectorError, match="Vector column must be Float32 or Float64 list"):
        B._dense_rows(pa.array([[1, 2]], pa.list_(pa.int64())), q)


def _check_surgery_keeps_footer(path):
    """append_index_inplace on `path` keeps schema, created_by, row-group metadata and foreign key-values byte for byte"""
    md0 = pq.read_metadata(path)
    before = pq.read_table(path)
    kv0 = dict(md0.metadata or {})
    size0 = os.path.getsize(path)
    blob = bytes(range(64))
    for rep in range(2):                                             # the second build replaces the two keys (parquet.rs:573-575)
        B.append_index_inplace(path, blob, "embedding")
        md1 = pq.read_metadata(path)
        kv1 = dict(md1.metadata)
        assert md1.schema.equals(md0.schema) and md1.created_by == md0.created_by
        assert md1.num_rows == md0.num_rows and md1.num_row_groups == md0.num_row_groups
        for i in range(md0.num_row_groups):
            assert md1.row_group(i).to_dict() == md0.row_group(i).to_dict()
        assert {k: v for k, v in kv1.items() if not k.startswith(b"pq_vector_")} == kv0     # ARROW:schema included
        assert kv1[b"pq_vector_embedding_column"] == b"embedding"
        assert int(kv1[b"pq_vector_index_offset"]) == (size0 - 8 if rep == 0 else size1 - 8)
        assert B.read_index_payload(path) == (blob, "embedding")
        assert pq.read_table(path).equals(before)
        size1 = os.path.getsize(path)


def test_append_index_inplace_keeps_foreign_footers(tmp_path):
    """Files written by arrow-rs / DataFusion / the reference crate name list children `item` (pyarrow: `element`) and carry
    nested columns; the surgery must not re-derive the schema (pq.write_metadata + AppendRowGroups refused
    them).  use_compliant_nested_type=False gives the arrow-rs naming from pyarrow."""
    path = str(tmp_path / "rs_style.parquet")
    n = 40
    t = pa.table({"id": pa.array(range(n), pa.int64()),
                  "authors": pa.array([["a", "b"]] * n, pa.list_(pa.string())),
                  "embedding": pa.array([[float(i), 1.0, 2.0] for i in range(n)], pa.list_(pa.float32())),
                  "meta": pa.array([{"x": i, "y": str(i)} for i in range(n)])})
    pq.write_table(t, path, compression="NONE", use_compliant_nested_type=False, row_group_size=16)
    assert pq.read_metadata(path).schema.column(1).path == "authors.list.item"
    _check_surgery_keeps_footer(path)
    # no key-value metadata at all in the source footer (field 5 absent -> inserted, the following field's header re-based)
    bare = str(tmp_path / "bare.parquet")
    pq.write_table(t, bare, compression="SNAPPY", store_schema=False)
    assert not pq.read_metadata(bare).metadata
    _check_surgery_keeps_footer(bare)
    # more than 14 key-values: long-form thrift list header
    many = str(tmp_path / "many.parquet")
    pq.write_table(t.replace_schema_metadata({f"k{i}": "v" * i for i in range(20)}), many)
    _check_surgery_keeps_footer(many)


@pytest.mark.skipif(not os.path.exists("/root/reference/data/vldb_2025.parquet"), reason="reference checkout not present")
def test_append_index_inplace_on_the_reference_dataset(tmp_path):
    """the parquet-rs 55.2.0 file the crate ships (14 columns, List<Float32> embeddings)"""
    import shutil
    path = str(tmp_path / "vldb.parquet")
    shutil.copy("/root/reference/data/vldb_2025.parquet", path)
    assert b"parquet-rs" in pq.read_metadata(path).created_by.encode()
    _check_surgery_keeps_footer(path)
