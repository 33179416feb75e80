"""Host-side mirror of the reference's public interface for the hot path, same names / argument meaning / error text:

    IndexBuilder(source, embedding_column).n_clusters(n).max_iters(n).seed(s).build_inplace() / .build_new(output)
                                                                       src/ivf/parquet.rs:21-102
    TopkBuilder(parquet_path, query).k(k).nprobe(n).search() -> [SearchResult(row_idx, distance)]
                                                                       src/ivf/search.rs:39-80
    has_pq_vector_index(path)                                          src/ivf/parquet.rs:185-188
    vector_topk(batches, column, query, k)                             src/df_vector/exec.rs:257-277, 429-550
                                                                       (VectorTopKExec::topk_from_batches)

The reference is Rust (no toolchain in this image), so this mirror is Python over pyarrow; every distance, argmin and
top-k runs on the GPU through the C ABI (include/pqv.h) -- there is no CPU path here.  Parquet I/O is host work and not
the subject of this repository; it is kept format-compatible so that files written here are readable by the reference and
vice versa: index payload = b"PQ_VECTOR1" + u64 LE length + IvfIndex::to_bytes bytes, placed behind the data, located by
the footer key `pq_vector_index_offset`, embedding column in `pq_vector_embedding_column` (parquet.rs:105-112, 356-372,
542-611).

What differs operationally: the reference re-reads index and vectors from the file for
every query (search.rs:89, 102-110); here the embedding column and the index of a file stay resident in HBM, keyed by
(real path, size, mtime), and a query ships only its vector and gets k (row, distance) pairs back; vector_topk computes
on indices first and materialises only the k winning rows (exec.rs:472 materialises every candidate row).
"""
from __future__ import annotations

import os
import struct
from typing import Iterable, NamedTuple

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

from . import _native as N
from .api import Context, Dataset, IvfIndex, PqvError

PQ_VECTOR_INDEX_MAGIC = b"PQ_VECTOR1"                          # parquet.rs:105
PQ_VECTOR_INDEX_OFFSET_KEY = b"pq_vector_index_offset"         # parquet.rs:108
PQ_VECTOR_EMBEDDING_COLUMN_KEY = b"pq_vector_embedding_column"  # parquet.rs:111


class PqVectorError(RuntimeError):
    """The reference returns Box<dyn Error> built from these strings; the text is kept."""


class SearchResult(NamedTuple):  # search.rs:40-45
    row_idx: int
    distance: float


class VectorTopKOptions(NamedTuple):  # src/df_vector/options.rs:5-19
    nprobe: int = 5
    max_candidates: "int | None" = None


# ---------------------------------------------------------------------------------------------------------------
# GPU context + residency
# ---------------------------------------------------------------------------------------------------------------
_ctx: "Context | None" = None
_tables: "dict[tuple, tuple[Dataset, int, int]]" = {}
_indexes: "dict[tuple, tuple[IvfIndex, str]]" = {}


def set_context(ctx: "Context | None"):
    """Use `ctx` for the builders (default: a Context on the current device, created on first use)."""
    global _ctx
    drop_resident()
    _ctx = ctx


def context() -> Context:
    global _ctx
    if _ctx is None:
        _ctx = Context()  # raises PqvError(PQV_ENODEV) without a GPU: no CPU fallback
    return _ctx


def drop_resident():
    """Forget every resident table / index (their HBM is released)."""
    for ds, _, _ in _tables.values():
        try:
            ds.drop()
        except PqvError:
            pass
    for ix, _ in _indexes.values():
        try:
            ix.drop()
        except PqvError:
            pass
    _tables.clear()
    _indexes.clear()


def _file_key(path) -> tuple:
    st = os.stat(path)
    return (os.path.realpath(path), st.st_size, st.st_mtime_ns)


def _evict_stale(cache: dict, key: tuple):
    """a file that was rewritten leaves entries of its old (size, mtime) behind: release their HBM when the path comes back"""
    for old in [k for k in cache if k[0] == key[0] and k[1:3] != key[1:3]]:
        obj = cache.pop(old)[0]
        try:
            obj.drop()
        except PqvError:
            pass


# ---------------------------------------------------------------------------------------------------------------
# Parquet side (host I/O; format as the reference's)
# ---------------------------------------------------------------------------------------------------------------
def _embedding_column(name: str) -> str:  # src/ivf/mod.rs:21-27
    if not str(name).strip():
        raise PqVectorError("Embedding column name cannot be empty")
    return str(name)


def read_parquet_with_embeddings(path, embedding_column: str):
    """parquet.rs:216-305: the whole file as a table + the dense row-major f32 matrix of the embedding column."""
    table = pq.read_table(path)
    if embedding_column not in table.column_names:
        raise PqVectorError(f"Column '{embedding_column}' not found")
    col = table.column(embedding_column)
    if not pa.types.is_list(col.type):  # the reference accepts ListArray only here
        raise PqVectorError("Embedding column is not a list array")
    vt = col.type.value_type
    if not (pa.types.is_float32(vt) or pa.types.is_float64(vt)):
        raise PqVectorError("Embedding values are not float32/float64")
    if col.null_count > 0:
        raise PqVectorError("Embedding column contains null rows")
    dim = None
    parts = []
    for chunk in col.chunks:
        if len(chunk) == 0:
            continue
        flat = chunk.flatten()  # the values the chunk's rows reference, in row order
        if flat.null_count > 0:
            raise PqVectorError("Embedding values contain nulls")
        lens = np.diff(chunk.offsets.to_numpy())
        if (lens == 0).any():
            raise PqVectorError("Embedding row has zero length")
        if dim is None:
            dim = int(lens[0])
        if (lens != dim).any():
            raise PqVectorError("Embedding vectors have inconsistent dimensions")
        parts.append(flat.to_numpy(zero_copy_only=False).astype(np.float32, copy=False))  # f64 narrowed, parquet.rs:290
    if dim is None:
        raise PqVectorError("Embedding column has no rows")
    values = parts[0] if len(parts) == 1 else np.concatenate(parts)
    return table, np.ascontiguousarray(values).reshape(-1, dim)


def read_index_metadata(path):
    """parquet.rs:114-149, 176-183: (offset, embedding column) from the footer key-values, or None."""
    kv = pq.read_metadata(path).metadata or {}
    off, col = kv.get(PQ_VECTOR_INDEX_OFFSET_KEY), kv.get(PQ_VECTOR_EMBEDDING_COLUMN_KEY)
    if off is None or col is None:
        return None
    return int(off.decode()), _embedding_column(col.decode())


def has_pq_vector_index(path) -> bool:
    return read_index_metadata(path) is not None


def read_index_payload(path) -> "tuple[bytes, str]":
    """parquet.rs:151-174, 191-208: the IvfIndex bytes embedded in the file and the embedding column name."""
    meta = read_index_metadata(path)
    if meta is None:
        raise PqVectorError("Missing pq-vector index metadata in parquet footer")
    offset, column = meta
    with open(path, "rb") as f:
        f.seek(offset)
        payload = f.read()
    header_len = len(PQ_VECTOR_INDEX_MAGIC) + 8
    try:
        if len(payload) < header_len:
            raise PqVectorError("pq-vector index payload is truncated")
        if payload[:len(PQ_VECTOR_INDEX_MAGIC)] != PQ_VECTOR_INDEX_MAGIC:
            raise PqVectorError("Invalid pq-vector index magic")
        (index_len,) = struct.unpack("<Q", payload[len(PQ_VECTOR_INDEX_MAGIC):header_len])
        if len(payload) < header_len + index_len:
            raise PqVectorError("pq-vector index bytes are truncated")
    except PqVectorError as e:
        raise PqVectorError(f"Failed to decode pq-vector index payload at offset {offset}: {e}") from None
    return payload[header_len:header_len + index_len], column


# --- thrift compact protocol, just enough to edit FileMetaData.key_value_metadata (field 5) of an existing footer -------
def _tc_varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _tc_put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        if v < 0x80:
            out.append(v)
            return bytes(out)
        out.append((v & 0x7F) | 0x80)
        v >>= 7


def _tc_skip(buf, pos, t):
    """end position of a value of compact-protocol type t that starts at pos"""
    if t in (1, 2):          # bool carried by the field header
        return pos
    if t == 3:               # byte
        return pos + 1
    if t in (4, 5, 6):       # zigzag varints
        return _tc_varint(buf, pos)[1]
    if t == 7:               # double
        return pos + 8
    if t == 8:               # binary / string
        n, pos = _tc_varint(buf, pos)
        return pos + n
    if t in (9, 10):         # list / set
        head = buf[pos]
        pos += 1
        n, et = head >> 4, head & 0x0F
        if n == 15:
            n, pos = _tc_varint(buf, pos)
        for _ in range(n):
            pos = pos + 1 if et in (1, 2) else _tc_skip(buf, pos, et)
        return pos
    if t == 11:              # map
        n, pos = _tc_varint(buf, pos)
        if n:
            kt, vt = buf[pos] >> 4, buf[pos] & 0x0F
            pos += 1
            for _ in range(n):
                pos = pos + 1 if kt in (1, 2) else _tc_skip(buf, pos, kt)
                pos = pos + 1 if vt in (1, 2) else _tc_skip(buf, pos, vt)
        return pos
    if t == 12:              # struct
        for _fid, ft, _s, e in _tc_fields(buf, pos):
            pos = e
        return pos + 1       # STOP byte
    raise PqVectorError(f"Unsupported thrift type {t} in parquet footer")


def _tc_fields(buf, pos):
    """(field id, type, value start, value end) of every field of the struct that starts at pos"""
    last = 0
    while True:
        head = buf[pos]
        pos += 1
        if head == 0:
            return
        t, delta = head & 0x0F, head >> 4
        if delta:
            fid = last + delta
        else:
            z, pos = _tc_varint(buf, pos)
            fid = (z >> 1) ^ -(z & 1)
        end = _tc_skip(buf, pos, t)
        yield fid, t, pos, end
        last, pos = fid, end


def _tc_field_header(fid: int, last: int, t: int) -> bytes:
    d = fid - last
    if 0 < d <= 15:
        return bytes([(d << 4) | t])
    return bytes([t]) + _tc_put_varint((fid << 1) ^ (fid >> 31))


def _tc_string(b: bytes) -> bytes:
    return _tc_put_varint(len(b)) + b


def footer_with_key_values(footer: bytes, drop_keys, add: "list[tuple[bytes, bytes]]") -> bytes:
    """The FileMetaData thrift struct `footer` with key_value_metadata (field 5) = its old entries minus drop_keys plus
    `add`, every other field byte for byte as it was (schema, row groups, created_by, column orders: parquet.rs:566-583
    rebuilds FileMetaData from exactly those parts)."""
    fields = list(_tc_fields(footer, 0))
    entries = []
    for fid, t, s, e in fields:
        if fid != 5:
            continue
        if t != 9:
            raise PqVectorError("key_value_metadata is not a list in this parquet footer")
        head = footer[s]
        pos = s + 1
        n = head >> 4
        if n == 15:
            n, pos = _tc_varint(footer, pos)
        for _ in range(n):
            key = None
            for kf, kt, ks, ke in _tc_fields(footer, pos):
                if kf == 1 and kt == 8:
                    ln, kp = _tc_varint(footer, ks)
                    key = bytes(footer[kp:kp + ln])
            end = _tc_skip(footer, pos, 12)
            if key not in drop_keys:
                entries.append(bytes(footer[pos:end]))
            pos = end
    for k, v in add:  # KeyValue { 1: required string key, 2: optional string value }
        entries.append(b"\x18" + _tc_string(k) + b"\x18" + _tc_string(v) + b"\x00")
    kv_list = (bytes([(len(entries) << 4) | 12]) if len(entries) < 15 else bytes([0xF0 | 12]) + _tc_put_varint(len(entries)))
    kv_list += b"".join(entries)
    out = bytearray()
    last, placed = 0, False
    for fid, t, s, e in fields:
        if fid == 5:
            continue
        if fid > 5 and not placed:
            out += _tc_field_header(5, last, 9) + kv_list
            last, placed = 5, True
        out += _tc_field_header(fid, last, t) + footer[s:e]
        last = fid
    if not placed:
        out += _tc_field_header(5, last, 9) + kv_list
    out.append(0)
    return bytes(out)


def append_index_inplace(path, index_bytes: bytes, embedding_column: str):
    """parquet.rs:542-611: the payload goes where the 8-byte footer tail was (the old footer bytes stay behind as dead
    space, exactly as in the reference), followed by a new footer = the old one with the two key-values replaced /
    appended.  The footer is edited at the thrift level, so schema, row groups, created_by and every foreign key-value
    (ARROW:schema included) survive byte for byte whichever writer produced the file (parquet-rs, pyarrow, ...).  Data
    pages do not move, so every column-chunk offset stays valid."""
    size = os.path.getsize(path)
    if size < 8:
        raise PqVectorError("Parquet file too small to contain a footer")
    with open(path, "rb") as f:
        f.seek(size - 8)
        tail = f.read(8)
        if tail[4:] == b"PARE":
            raise PqVectorError("Encrypted parquet footers are not supported for in-place indexing")
        if tail[4:] != b"PAR1":
            raise PqVectorError("Invalid parquet footer magic")
        (old_len,) = struct.unpack("<I", tail[:4])
        if old_len + 8 > size:
            raise PqVectorError("Parquet footer length exceeds file size")
        f.seek(size - 8 - old_len)
        old_footer = f.read(old_len)
    index_offset = size - 8
    footer = footer_with_key_values(
        old_footer, (PQ_VECTOR_INDEX_OFFSET_KEY, PQ_VECTOR_EMBEDDING_COLUMN_KEY),
        [(PQ_VECTOR_INDEX_OFFSET_KEY, str(index_offset).encode()), (PQ_VECTOR_EMBEDDING_COLUMN_KEY, embedding_column.encode())])
    with open(path, "r+b") as f:
        f.seek(index_offset)
        f.write(PQ_VECTOR_INDEX_MAGIC)
        f.write(struct.pack("<Q", len(index_bytes)))
        f.write(index_bytes)
        f.write(footer)
        f.write(struct.pack("<I", len(footer)))
        f.write(b"PAR1")
        f.truncate()


# ---------------------------------------------------------------------------------------------------------------
# IndexBuilder (parquet.rs:21-102)
# ---------------------------------------------------------------------------------------------------------------
class IndexBuilder:
    def __init__(self, source, embedding_column: str):
        self._source = os.fspath(source)
        self._embedding_column = embedding_column
        self._n_clusters = None
        self._max_iters = 20
        self._seed = 42

    def n_clusters(self, n_clusters: int) -> "IndexBuilder":
        self._n_clusters = int(n_clusters)
        return self

    def max_iters(self, max_iters: int) -> "IndexBuilder":
        self._max_iters = int(max_iters)
        return self

    def seed(self, seed: int) -> "IndexBuilder":
        self._seed = int(seed)
        return self

    def _build(self):
        if self._max_iters == 0:  # parquet.rs:88-102
            raise PqVectorError("max_iters must be > 0")
        if self._n_clusters is not None and self._n_clusters == 0:
            raise PqVectorError("n_clusters must be > 0")
        column = _embedding_column(self._embedding_column)
        table, emb = read_parquet_with_embeddings(self._source, column)
        ctx = context()
        ds = ctx.dataset_from(emb)
        try:
            ix = ctx.ivf_build(ds, n_clusters=self._n_clusters, max_iters=self._max_iters, seed=self._seed)
        except PqvError as e:  # index.rs:157-170 wording comes through the ABI
            ds.drop()
            raise PqVectorError(str(e).split(": ", 1)[-1]) from None
        return table, column, ds, ix, emb.shape

    def _register(self, path, ds, ix, column, shape):
        key = _file_key(path)
        _evict_stale(_tables, key)
        _evict_stale(_indexes, key)
        _tables[key + (column,)] = (ds, shape[0], shape[1])
        _indexes[key] = (ix, column)

    def build_inplace(self) -> None:
        _, column, ds, ix, shape = self._build()
        append_index_inplace(self._source, ix.to_bytes(), column)
        self._register(self._source, ds, ix, column, shape)  # the freshly built table and index stay resident

    def build_new(self, output) -> None:
        table, column, ds, ix, shape = self._build()
        output = os.fspath(output)
        src = pq.read_metadata(self._source)
        comp = {}
        if src.num_row_groups:  # parquet.rs:379-470: keep each column's compression; no dictionary for the vectors
            rg = src.row_group(0)
            for i in range(rg.num_columns):
                c = rg.column(i)
                comp[c.path_in_schema.split(".")[0]] = "NONE" if c.compression == "UNCOMPRESSED" else c.compression
        use_dict = [n for n in table.column_names if n != column]
        pq.write_table(table, output, compression=comp or "NONE", use_dictionary=use_dict,
                       data_page_size=max(int(shape[1]) * 4, 1))
        append_index_inplace(output, ix.to_bytes(), column)
        self._register(output, ds, ix, column, shape)


# ---------------------------------------------------------------------------------------------------------------
# TopkBuilder (search.rs:47-142)
# ---------------------------------------------------------------------------------------------------------------
def _resident_index(path) -> "tuple[IvfIndex, str]":
    key = _file_key(path)
    hit = _indexes.get(key)
    if hit is None:
        _evict_stale(_indexes, key)
        blob, column = read_index_payload(path)
        try:
            hit = (context().ivf_from_bytes(blob), column)
        except PqvError as e:
            raise PqVectorError(str(e).split(": ", 1)[-1]) from None
        _indexes[key] = hit
    return hit


def _resident_table(path, column) -> "tuple[Dataset, int, int]":
    key = _file_key(path) + (column,)   # one resident block per (file state, vector column)
    hit = _tables.get(key)
    if hit is None:
        _evict_stale(_tables, key)
        _, emb = read_parquet_with_embeddings(path, column)
        hit = (context().dataset_from(emb), emb.shape[0], emb.shape[1])
        _tables[key] = hit
    return hit


class TopkBuilder:
    def __init__(self, parquet_path, query):
        self._path = os.fspath(parquet_path)
        self._query = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        self._k = None
        self._nprobe = None

    def k(self, k: int) -> "TopkBuilder":
        if int(k) <= 0:
            raise PqVectorError("k must be > 0")
        self._k = int(k)
        return self

    def nprobe(self, nprobe: int) -> "TopkBuilder":
        if int(nprobe) <= 0:
            raise PqVectorError("nprobe must be > 0")
        self._nprobe = int(nprobe)
        return self

    def search(self) -> "list[SearchResult]":
        if self._k is None:
            raise PqVectorError("k must be set")
        if self._nprobe is None:
            raise PqVectorError("nprobe must be set")
        ix, column = _resident_index(self._path)
        if self._query.size != ix.dim:
            raise PqVectorError(f"Query dimension mismatch: expected {ix.dim}, got {self._query.size}")
        ds, _, dim = _resident_table(self._path, column)
        if dim != ix.dim:
            raise PqVectorError(f"Embedding dimension mismatch: expected {ix.dim}, got {dim}")  # search.rs:224-231
        # through the coalescing front door: concurrent search() callers (the reference's is an async fn, search.rs:76-80)
        # are answered together by one batched pass; a lone caller gets the plain single-query pipeline
        rows, dist = ix.search_coalesced(ds, self._query, self._k, self._nprobe, N.PQV_SQRT)
        return [SearchResult(int(r), float(d)) for r, d in zip(rows, dist)]


# ---------------------------------------------------------------------------------------------------------------
# VectorTopKExec::topk_from_batches (exec.rs:257-277, 429-550)
# ---------------------------------------------------------------------------------------------------------------
def _dense_rows(array: pa.Array, dim: int):
    """Rows of a list column that the reference would score (exec.rs:494-546): not null, length == len(query).
    Returns (values [m, dim] f32 or f64, kept row indices)."""
    t = array.type
    if not (pa.types.is_list(t) or pa.types.is_large_list(t) or pa.types.is_fixed_size_list(t)):
        raise PqVectorError("Vector column must be list or fixed-size list")
    vt = t.value_type
    if not (pa.types.is_float32(vt) or pa.types.is_float64(vt)):
        raise PqVectorError("Vector column must be Float32 or Float64 list")
    n = len(array)
    valid = np.ones(n, dtype=bool) if array.null_count == 0 else np.asarray(array.is_valid())
    np_t = np.float32 if pa.types.is_float32(vt) else np.float64
    if pa.types.is_fixed_size_list(t):
        width = t.list_size
        flat = array.flatten().to_numpy(zero_copy_only=False) if array.null_count == 0 else None
        if width != dim:
            return np.empty((0, dim), np_t), np.empty(0, np.int64)
        if flat is None:  # null slots still occupy `width` child values
            child = array.values.to_numpy(zero_copy_only=False)[array.offset * width:(array.offset + n) * width]
            vals = child.reshape(n, width)[valid]
        else:
            vals = flat.reshape(n, width)
        return np.ascontiguousarray(vals, dtype=np_t), np.nonzero(valid)[0]
    offs = array.offsets.to_numpy()
    lens = np.diff(offs)
    keep = valid & (lens == dim)
    child = array.values.to_numpy(zero_copy_only=False)
    idx = np.nonzero(keep)[0]
    if idx.size == n and n and offs[-1] - offs[0] == n * dim:  # dense: one contiguous block
        vals = child[offs[0]:offs[-1]].reshape(n, dim)
    else:
        vals = np.stack([child[offs[i]:offs[i] + dim] for i in idx]) if idx.size else np.empty((0, dim), np_t)
    return np.ascontiguousarray(vals, dtype=np_t), idx


def vector_topk(batches: "Iterable[pa.RecordBatch]", column: str, query, k: int, schema: "pa.Schema | None" = None):
    """The k rows with the smallest squared-L2 distance (sequential f32 sum, exec.rs:529-533) between `column` and
    `query`, as ONE RecordBatch in ascending distance order -- what VectorTopKExec emits.  Null rows and rows whose
    length differs from the query are skipped (exec.rs:496-498, 526-528); Float64 values are narrowed to f32 before
    the subtraction (exec.rs:542).  Distances are computed on the GPU from the batches' values buffers as they
    arrive (pqv_topk_stream_*); only the k winners are materialised."""
    q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
    if int(k) <= 0:
        raise PqVectorError("k must be > 0")
    stream = context().topk_stream(q, int(k), N.PQV_SUM_SEQ)
    kept_batches, row_maps, base = [], [], []
    pushed = 0
    for batch in batches:
        if schema is None:
            schema = batch.schema
        vals, idx = _dense_rows(batch.column(batch.schema.get_field_index(column)), q.size)
        kept_batches.append(batch)
        row_maps.append(idx)
        base.append(pushed)
        if idx.size:
            stream.push(vals)
            pushed += idx.size
    rows, _ = stream.finish()
    if schema is None:
        raise PqVectorError("vector_topk needs at least one batch or a schema")
    if rows.size == 0:
        return pa.RecordBatch.from_pylist([], schema=schema)
    base_arr = np.asarray(base, dtype=np.int64)
    pieces = []
    # position in the pushed sequence -> (batch, row); batches that pushed nothing share their successor's base, and
    # searchsorted(side="right") lands on the last batch of such a run -- the only one that can own the position
    for pos in rows.astype(np.int64):
        b = int(np.searchsorted(base_arr, pos, side="right") - 1)
        pieces.append(kept_batches[b].slice(int(row_maps[b][pos - base[b]]), 1))
    return pa.Table.from_batches(pieces, schema=schema).combine_chunks().to_batches()[0]
