"""Host-side mirror of the reference's DataFusion surface for the hot path (src/df_vector/):

    SessionStateBuilder().with_pq_vector(VectorTopKOptions(nprobe, max_candidates)).build() -> SessionContext
                                                                       session.rs:23-35 (PqVectorSessionBuilderExt)
    ctx.register_parquet("t", path | [paths]);  ctx.sql("SELECT .. ORDER BY array_distance(col, [..]) LIMIT k").collect()

DataFusion itself (SQL front end, planner, Parquet scan) is out of scope; this module plans exactly
the query shape the reference's optimizer rule recognises (physical.rs:32-91, 134-229: ONE ascending sort key that is
`array_distance(Column, Literal)` in either argument order, a LIMIT, one Parquet table underneath, optional WHERE) and
executes it the way the reference's operators do, with every distance and top-k on the GPU:

  * rule registered (with_pq_vector)   -> VectorTopKExec (exec.rs:207-293): per file the index's candidate rows
        (index_exec.rs:83-188; a file without an index is an error, as there), CandidateCursor round-robin capped by
        max_candidates (access.rs:193-243), rows visited in file order (RowSelection), the scan subtree's filter applied
        BEFORE scoring (tests.rs:151-241), sequential-order f32 squared distances, bounded heap of k (exec.rs:257-277).
        The rows are scored from the file's HBM-resident column by row id (pqv_l2_topk_gather) -- nothing is re-read.
  * no rule (the bench's "no index" arm, benches/query.rs:76-103) -> stock SortExec(TopK) over the built-in UDF:
        Float64 array_distance of every row that passes the filter, k smallest (pqv_array_distance[_topk]).

`DataFrame.metrics` carries the counters the reference's plan snapshots show (candidate_rows, embeddings_fetched,
batches_fetched, files_scanned: src/df_vector/snapshots/*.snap)."""
from __future__ import annotations

import os
import re

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pyarrow.parquet as pq

from . import _native as N
from . import builders as B
from .builders import PqVectorError, VectorTopKOptions

_SQL = re.compile(
    r"^\s*SELECT\s+(?P<cols>.+?)\s+FROM\s+(?P<table>[A-Za-z_]\w*)(?:\s+WHERE\s+(?P<where>.+?))?"
    r"\s+ORDER\s+BY\s+array_distance\s*\(\s*(?P<args>.+?)\s*\)(?:\s+(?P<dir>ASC|DESC))?"
    r"\s+LIMIT\s+(?P<k>\d+)\s*;?\s*$", re.I | re.S)
_PRED = re.compile(r"^\s*(?P<col>[A-Za-z_]\w*)\s*(?P<op>>=|<=|<>|!=|=|>|<)\s*(?P<lit>'[^']*'|[-+0-9.eE]+)\s*$")
_OPS = {"=": pc.equal, "!=": pc.not_equal, "<>": pc.not_equal, "<": pc.less, "<=": pc.less_equal, ">": pc.greater,
        ">=": pc.greater_equal}


class ParsedQuery:
    """SELECT <cols> FROM <table> [WHERE p AND ..] ORDER BY array_distance(<column>, <literal>) LIMIT <k>"""

    def __init__(self, sql: str):
        m = _SQL.match(sql)
        if not m:
            raise PqVectorError("unsupported SQL: this mirror plans only `SELECT .. FROM t [WHERE ..] ORDER BY "
                                "array_distance(column, [literal]) LIMIT k` (the shape physical.rs:32-91 rewrites)")
        if (m.group("dir") or "ASC").upper() == "DESC":
            raise PqVectorError("unsupported SQL: descending array_distance order (physical.rs:143 leaves it to DataFusion)")
        cols = m.group("cols").strip()
        self.columns = None if cols == "*" else [c.strip() for c in cols.split(",")]
        self.table = m.group("table")
        self.k = int(m.group("k"))
        args = m.group("args")
        lit = re.search(r"\[(.*)\]", args, re.S)
        if not lit:
            raise PqVectorError("array_distance needs a column and a list literal")
        rest = (args[:lit.start()] + args[lit.end():]).strip().strip(",").strip()   # either argument order, physical.rs:204-211
        if not re.fullmatch(r"[A-Za-z_]\w*", rest):
            raise PqVectorError("array_distance needs a column and a list literal")
        self.column = rest
        body = lit.group(1).strip()
        try:
            self.literal = np.array([float(x) for x in body.split(",")] if body else [], dtype=np.float64)
        except ValueError:
            raise PqVectorError("array_distance literal must be a list of numbers") from None
        self.predicates = []
        if m.group("where"):
            for part in re.split(r"\s+AND\s+", m.group("where"), flags=re.I):
                pm = _PRED.match(part)
                if not pm:
                    raise PqVectorError(f"unsupported predicate '{part.strip()}' (column <op> literal [AND ..])")
                lit_s = pm.group("lit")
                value = lit_s[1:-1] if lit_s.startswith("'") else (float(lit_s) if re.search(r"[.eE]", lit_s) else int(lit_s))
                self.predicates.append((pm.group("col"), pm.group("op"), value))


class CandidateCursor:
    """access.rs:193-243: round-robin over the files' candidate lists, one row per file per turn."""

    def __init__(self, file_count: int):
        self.candidates = [[] for _ in range(file_count)]
        self.positions = [0] * file_count
        self.round_robin = 0

    def add_candidates(self, idx: int, candidates):
        if 0 <= idx < len(self.candidates):
            self.candidates[idx] = candidates

    def next_batch(self, batch_size: int):
        if batch_size == 0 or not self.candidates:
            return []
        nf = len(self.candidates)
        out, idx = [], self.round_robin
        while len(out) < batch_size:
            progressed = False
            for _ in range(nf):
                f = idx % nf
                idx += 1
                if self.positions[f] < len(self.candidates[f]):
                    out.append((f, int(self.candidates[f][self.positions[f]])))
                    self.positions[f] += 1
                    progressed = True
                    if len(out) >= batch_size:
                        break
            if not progressed:
                break
        self.round_robin = idx % nf
        return out


_host_tables: "dict[tuple, pa.Table]" = {}


def _host_table(path, columns) -> pa.Table:
    """The columns a query touches besides the distance itself (projection + predicates), cached per file; the vector
    column is only read here when the query SELECTs it -- distances come from the HBM-resident copy."""
    key = B._file_key(path) + (tuple(columns),)
    t = _host_tables.get(key)
    if t is None:
        t = pq.read_table(path, columns=list(columns))
        _host_tables[key] = t
    return t


def _filter_mask(table: pa.Table, predicates) -> "np.ndarray | None":
    mask = None
    for col, op, value in predicates:
        if col not in table.column_names:
            raise PqVectorError(f"Schema error: No field named {col}.")
        m = _OPS[op](table.column(col), pa.scalar(value))
        mask = m if mask is None else pc.and_kleene(mask, m)
    if mask is None:
        return None
    return np.asarray(pc.fill_null(mask, False).combine_chunks())   # NULL predicate = row dropped (FilterExec)


class DataFrame:
    def __init__(self, ctx: "SessionContext", q: ParsedQuery):
        self._ctx, self._q = ctx, q
        self.metrics: dict = {}

    def explain(self) -> dict:
        q = self._q
        if self._ctx.options is not None:
            return {"operator": "VectorTopKExec", "column": q.column, "k": q.k, "nprobe": self._ctx.options.nprobe,
                    "query_dim": int(q.literal.size), "children": ["VectorIndexScanExec", "FilterExec" if q.predicates else "DataSourceExec"]}
        return {"operator": "SortExec(TopK)", "expr": f"array_distance({q.column}, literal)", "fetch": q.k}

    def collect(self) -> "list[pa.RecordBatch]":
        t = self.to_table()
        return t.to_batches() if t.num_rows else []

    def to_table(self) -> pa.Table:
        q = self._q
        files = self._ctx._tables.get(q.table)
        if files is None:
            raise PqVectorError(f"table '{q.table}' not found")
        schemas = [pq.read_schema(p) for p in files]
        cols = q.columns or schemas[0].names
        needed = list(dict.fromkeys(list(cols) + [c for c, _, _ in q.predicates]))
        for sc in schemas:
            if q.column not in sc.names:
                raise PqVectorError(f"Vector column '{q.column}' not found in schema")        # exec.rs:247-255
            for c in needed:
                if c not in sc.names:
                    raise PqVectorError(f"Schema error: No field named {c}.")
            col_t = sc.field(q.column).type
            if not (pa.types.is_list(col_t) or pa.types.is_large_list(col_t) or pa.types.is_fixed_size_list(col_t)):
                raise PqVectorError("Vector column must be list or fixed-size list")           # exec.rs:519-521
        tables = [_host_table(p, needed) for p in files]
        masks = [_filter_mask(t, q.predicates) for t in tables]
        picked = self._vector_topk(files, masks) if self._ctx.options is not None else self._stock_topk(files, masks)
        pieces = [tables[f].select(cols).slice(r, 1) for f, r in picked]
        if not pieces:
            return tables[0].select(cols).slice(0, 0)
        return pa.concat_tables(pieces).combine_chunks()

    # ---- VectorTopKExec (exec.rs:207-293) ------------------------------------------------------------------------
    def _vector_topk(self, files, masks):
        q, opt = self._q, self._ctx.options
        query = q.literal.astype(np.float32)                                 # scalar_to_f32_list, physical.rs:229
        if q.k <= 0:
            raise PqVectorError("k must be > 0")
        indexes = []
        for path in files:
            if not B.has_pq_vector_index(path):
                raise PqVectorError(f"Missing pq-vector index metadata in '{path}'")                  # index_exec.rs:116-121
            ix, column = B._resident_index(path)
            if column != q.column:
                raise PqVectorError(f"IVF index column mismatch: expected '{q.column}', found '{column}'")  # :123-129
            if ix.dim != query.size:
                raise PqVectorError(f"Query dimension mismatch: expected {ix.dim}, got {query.size}")       # :152-158
            indexes.append(ix)
        fetched = batches = total = 0
        results = []        # (squared distance f32, scan position, file, row)
        if opt.max_candidates == 0:
            pass                                                             # Some(0): target_candidates = 0, exec.rs:222-223
        elif len(files) == 1:
            # one file: CandidateCursor is a prefix of the candidate list -> the whole operator is ONE device round trip
            ds, _, dim = B._resident_table(files[0], q.column)
            r, d, total, fetched = indexes[0].vector_topk(ds, query, q.k, opt.nprobe, N.PQV_SUM_SEQ, opt.max_candidates, masks[0])
            batches = 1 if fetched else 0
            results = [(float(dd), i, 0, int(rr)) for i, (dd, rr) in enumerate(zip(d, r))]
        else:
            cursor = CandidateCursor(len(files))
            for i, ix in enumerate(indexes):
                cand = ix.candidate_rows(query, opt.nprobe)
                cursor.add_candidates(i, cand)
                total += int(cand.size)
            target = min(opt.max_candidates if opt.max_candidates is not None else total, total)         # exec.rs:222-223
            per_file = [[] for _ in files]
            for f, r in cursor.next_batch(target):
                per_file[f].append(r)
            pos0 = 0
            for f, path in enumerate(files):
                rows = np.unique(np.asarray(per_file[f], dtype=np.uint32))   # RowSelection: file order, access.rs:107-176
                if masks[f] is not None and rows.size:
                    rows = rows[masks[f][rows]]                                # FilterExec above the scan
                if rows.size == 0:
                    continue
                ds, _, dim = B._resident_table(path, q.column)
                if dim != query.size:
                    continue                                                   # rows of another length are skipped, exec.rs:526-528
                r, d = ds.l2_topk_gather(query, rows, q.k, N.PQV_SUM_SEQ)     # squared distances, reference heap order
                fetched += int(rows.size)
                batches += 1
                where = np.searchsorted(rows, r)
                results += [(float(dd), pos0 + int(w), f, int(rr)) for dd, w, rr in zip(d, where, r)]
                pos0 += int(rows.size)
        self.metrics = {"candidate_rows": total, "files": len(files), "files_scanned": len(files),
                        "embeddings_fetched": fetched, "batches_fetched": batches, "k": q.k, "nprobe": opt.nprobe,
                        "query_dim": int(query.size), "column": q.column}
        if len(files) > 1:   # one heap over all files: k smallest, earlier scan position first among equal distances
            results.sort(key=lambda t: (np.float32(t[0]).view(np.uint32) if t[0] == t[0] else 0xFFFFFFFF, t[1]))
            results = results[:q.k]
        return [(f, r) for _, _, f, r in results]

    # ---- stock plan: SortExec(TopK) over the built-in array_distance (un-indexed arm) ----------------------------
    def _stock_topk(self, files, masks):
        q = self._q
        found = []          # (f64 distance, file, row)
        scanned = 0
        for f, path in enumerate(files):
            ds, n, dim = B._resident_table(path, q.column)
            if dim != q.literal.size:
                raise PqVectorError("Both arrays must have the same length")
            scanned += n
            r, d = ds.array_distance_topk(q.literal, q.k, row_mask=masks[f])      # the filter is applied on the device
            found += [(float(dd), f, int(rr)) for dd, rr in zip(d, r)]
        self.metrics = {"rows_scanned": scanned, "files": len(files), "k": q.k, "query_dim": int(q.literal.size),
                        "column": q.column}
        if len(files) > 1:
            found.sort(key=lambda t: (t[0] != t[0], t[0] if t[0] == t[0] else 0.0, t[1], t[2]))
            found = found[:q.k]
        return [(f, r) for _, f, r in found]


class SessionContext:
    """What the tests / examples of the reference use of datafusion::prelude::SessionContext."""

    def __init__(self, options: "VectorTopKOptions | None" = None):
        self.options = options
        self._tables: "dict[str, list[str]]" = {}

    def register_parquet(self, name: str, path) -> None:
        paths = [os.fspath(p) for p in (path if isinstance(path, (list, tuple)) else [path])]
        for p in paths:
            if not os.path.exists(p):
                raise PqVectorError(f"parquet file '{p}' not found")
        if not paths:
            raise PqVectorError("VectorTopKExec requires at least one indexed parquet file")     # exec.rs:214-218
        self._tables[name] = paths

    def sql(self, query: str) -> DataFrame:
        return DataFrame(self, ParsedQuery(query))


class SessionStateBuilder:
    """SessionStateBuilder + PqVectorSessionBuilderExt (session.rs:23-35): with_pq_vector registers the
    VectorTopK physical optimizer rule; without it the stock plan runs."""

    def __init__(self):
        self._options = None

    def with_pq_vector(self, options: "VectorTopKOptions | None" = None) -> "SessionStateBuilder":
        self._options = options if options is not None else VectorTopKOptions()
        if self._options.nprobe <= 0:
            raise PqVectorError("nprobe must be > 0")
        return self

    def with_physical_optimizer_rule(self, options: VectorTopKOptions) -> "SessionStateBuilder":  # tests.rs:56-61 spelling
        return self.with_pq_vector(options)

    def build(self) -> SessionContext:
        return SessionContext(self._options)


def drop_resident():
    _host_tables.clear()
    B.drop_resident()
