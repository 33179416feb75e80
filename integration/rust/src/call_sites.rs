//! The bodies that replace the reference's hot loops once `gpu.rs` is in the crate (INTEGRATION.md section 3 lists
//! the sites).  Written against pq-vector's own types (`crate::ivf::*`, arrow 57, datafusion 52) -- NOT COMPILED in the
//! authoring image (no cargo/rustc).  Everything outside these functions (Parquet I/O, index format, optimizer rule,
//! plan plumbing) stays as it is.
use crate::gpu::{Gpu, Index, Result, Table};
use crate::ivf::{EmbeddingColumn, SearchResult};
use arrow::array::{Array, ArrayRef, FixedSizeListArray, Float32Array, Float64Array, LargeListArray, ListArray, UInt32Array};
use arrow::record_batch::RecordBatch;
use parquet::arrow::arrow_reader::ParquetRecordBatchReaderBuilder;
use parquet::arrow::ProjectionMask;
use std::collections::HashMap;
use std::fs::File;
use std::path::{Path, PathBuf};
use std::sync::{Arc, Mutex, OnceLock};
use std::time::SystemTime;

/// A file's embedding column and (when embedded) its index, resident in HBM.  Loaded once per (path, size, mtime):
/// replaces `read_index_from_parquet` + `read_embeddings_for_rows` running for every query (src/ivf/search.rs:89, 102-110).
pub struct ResidentFile {
    pub table: Table,
    pub index: Option<Index>,
}

type FileKey = (PathBuf, u64, SystemTime);

fn cache() -> &'static Mutex<HashMap<FileKey, Arc<ResidentFile>>> {
    static CACHE: OnceLock<Mutex<HashMap<FileKey, Arc<ResidentFile>>>> = OnceLock::new();
    CACHE.get_or_init(|| Mutex::new(HashMap::new()))
}

pub fn resident_file(path: &Path, embedding_column: &EmbeddingColumn, index_blob: Option<&[u8]>) -> Result<Arc<ResidentFile>> {
    let meta = std::fs::metadata(path)?;
    let key = (path.to_path_buf(), meta.len(), meta.modified()?);
    if let Some(hit) = cache().lock().unwrap().get(&key) {
        return Ok(hit.clone());
    }
    let gpu = Gpu::global()?;
    let builder = ParquetRecordBatchReaderBuilder::try_new(File::open(path)?)?;
    let column = builder
        .parquet_schema()
        .columns()
        .iter()
        .position(|c| c.path().parts()[0] == embedding_column.as_str())
        .ok_or_else(|| format!("Column '{}' not found", embedding_column.as_str()))?;
    let rows = builder.metadata().file_metadata().num_rows() as usize;
    let mask = ProjectionMask::leaves(builder.parquet_schema(), [column]);
    let mut table: Option<Table> = None;
    for batch in builder.with_projection(mask).with_batch_size(1 << 16).build()? {
        let batch = batch?;
        // the checks of src/ivf/parquet.rs:236-272: no null rows, no null values, one dimension
        let list = batch.column(0).as_any().downcast_ref::<ListArray>().ok_or("Embedding column is not a list array")?;
        if list.null_count() > 0 {
            return Err("Embedding column contains null rows".into());
        }
        if list.is_empty() {
            continue;
        }
        let dim = list.value_length(0) as usize;
        if dim == 0 {
            return Err("Embedding row has zero length".into());
        }
        if (0..list.len()).any(|r| list.value_length(r) as usize != dim) {
            return Err("Embedding vectors have inconsistent dimensions".into());
        }
        let first = list.value_offsets()[0] as usize;
        let len = list.len() * dim;
        if table.is_none() {
            table = Some(gpu.create_table(dim, rows)?);
        }
        let t = table.as_ref().expect("created above");
        if let Some(values) = list.values().as_any().downcast_ref::<Float32Array>() {
            if values.null_count() > 0 {
                return Err("Embedding values contain nulls".into());
            }
            t.append(&values.values()[first..first + len])?; // zero-copy: the Arrow values buffer itself
        } else if let Some(values) = list.values().as_any().downcast_ref::<Float64Array>() {
            if values.null_count() > 0 {
                return Err("Embedding values contain nulls".into());
            }
            let narrowed: Vec<f32> = values.values()[first..first + len].iter().map(|&v| v as f32).collect(); // parquet.rs:288-291
            t.append(&narrowed)?;
        } else {
            return Err("Embedding values are not float32/float64".into());
        }
    }
    let table = table.ok_or("Embedding column has no rows")?;
    let index = index_blob.map(|blob| gpu.load_index(blob)).transpose()?;
    let file = Arc::new(ResidentFile { table, index });
    cache().lock().unwrap().insert(key, file.clone());
    Ok(file)
}

/// Site 1' -- src/ivf/search.rs:83-142 (`topk`), whole function.  `index_blob` = the payload bytes that
/// `read_index_from_parquet` (src/ivf/parquet.rs:191-208) has just read; they are parsed on the first query only.
/// Call from `TopkBuilder::search` through `tokio::task::spawn_blocking` (the call blocks on the GPU).
pub fn topk(parquet_path: &Path, embedding_column: &EmbeddingColumn, index_blob: &[u8], query: &[f32], k: usize,
            nprobe: usize) -> Result<Vec<SearchResult>> {
    let file = resident_file(parquet_path, embedding_column, Some(index_blob))?;
    let index = file.index.as_ref().expect("loaded with a blob");
    let hits = file.table.ivf_search(index, query, k, nprobe)?; // ranking, expansion, gathered scan, heap, sqrt, sort
    Ok(hits.row_idx.into_iter().zip(hits.distance).map(|(row_idx, distance)| SearchResult { row_idx, distance }).collect())
}

/// Site 1 -- src/ivf/search.rs:112-141 only (the host keeps `IvfIndex::candidate_rows`).
pub fn rerank(file: &ResidentFile, query: &[f32], rows_to_check: &[u32], k: usize) -> Result<Vec<SearchResult>> {
    let hits = file.table.topk_gather(query, rows_to_check, k)?;
    Ok(hits.row_idx.into_iter().zip(hits.distance).map(|(row_idx, distance)| SearchResult { row_idx, distance }).collect())
}

/// Site 2 -- src/df_vector/exec.rs:257-277 (`topk_from_batches`) + `update_topk_heap` (:457-482): indices first, then
/// `take` the <= k winners -- `row_to_scalar_values` no longer runs for every candidate row.
/// Returns, per winner in output order, (batch number, row in batch).
pub fn topk_from_batches(batches: &[RecordBatch], vector_idx: usize, query: &[f32], k: usize) -> Result<Vec<(usize, usize)>> {
    let mut stream = Gpu::global()?.topk_stream(query, k)?;
    let mut origin: Vec<(usize, usize)> = Vec::new(); // pushed sequence -> (batch, row)
    for (b, batch) in batches.iter().enumerate() {
        let (values, offsets, valid) = list_parts(batch.column(vector_idx))?;
        // rows the operator skips (exec.rs:496-498 null row, :526-528 / :537-539 length mismatch) never reach the device
        let keep: Vec<usize> = (0..batch.num_rows())
            .filter(|&r| valid(r) && (offsets(r + 1) - offsets(r)) == query.len())
            .collect();
        origin.extend(keep.iter().map(|&r| (b, r)));
        let dense = keep.len() == batch.num_rows();
        if let Some(f) = values.as_any().downcast_ref::<Float32Array>() {
            if dense {
                stream.push_f32(&f.values()[offsets(0)..offsets(batch.num_rows())])?;
            } else {
                let packed: Vec<f32> = keep.iter().flat_map(|&r| f.values()[offsets(r)..offsets(r + 1)].iter().copied()).collect();
                stream.push_f32(&packed)?;
            }
        } else if let Some(f) = values.as_any().downcast_ref::<Float64Array>() {
            let packed: Vec<f64> = keep.iter().flat_map(|&r| f.values()[offsets(r)..offsets(r + 1)].iter().copied()).collect();
            stream.push_f64(&packed)?; // narrowed to f32 on the device before the subtraction, as exec.rs:542
        } else {
            return Err("Vector column must be Float32 or Float64 list".into());
        }
    }
    let hits = stream.finish()?;
    Ok(hits.row_idx.iter().map(|&i| origin[i as usize]).collect())
}

/// (child values, offset of row r in the child, validity of row r) for the three list layouts of exec.rs:494-519.
fn list_parts(array: &ArrayRef) -> Result<(ArrayRef, Box<dyn Fn(usize) -> usize + '_>, Box<dyn Fn(usize) -> bool + '_>)> {
    if let Some(l) = array.as_any().downcast_ref::<ListArray>() {
        return Ok((l.values().clone(), Box::new(move |r| l.value_offsets()[r] as usize), Box::new(move |r| l.is_valid(r))));
    }
    if let Some(l) = array.as_any().downcast_ref::<LargeListArray>() {
        return Ok((l.values().clone(), Box::new(move |r| l.value_offsets()[r] as usize), Box::new(move |r| l.is_valid(r))));
    }
    if let Some(l) = array.as_any().downcast_ref::<FixedSizeListArray>() {
        let w = l.value_length() as usize;
        return Ok((l.values().clone(), Box::new(move |r| r * w), Box::new(move |r| l.is_valid(r))));
    }
    Err("Vector column must be list or fixed-size list".into())
}

/// Site 2' -- src/df_vector/exec.rs:207-277 (`execute_with_candidates` + `topk_from_batches`) over one resident,
/// indexed file: returns the winning file rows (operator order) as a take-index array plus the two plan counters of
/// the reference's snapshots (`candidate_rows`, `embeddings_fetched`).  `filter_mask`: the scan subtree's predicate
/// evaluated over the file's rows, as an Arrow boolean buffer (None: no FilterExec under the scan).
pub fn vector_topk_indexed(file: &ResidentFile, query: &[f32], k: usize, nprobe: usize, max_candidates: Option<usize>,
                           filter_mask: Option<&[u8]>) -> Result<(UInt32Array, u64, u64)> {
    let index = file.index.as_ref().ok_or("VectorTopKExec requires at least one indexed parquet file")?; // exec.rs:213-217
    let (hits, candidate_rows, embeddings_fetched) =
        file.table.vector_topk_indexed(index, query, k, nprobe, max_candidates, filter_mask)?;
    Ok((UInt32Array::from(hits.row_idx), candidate_rows, embeddings_fetched))
}

/// Sites 3 + 3' -- src/ivf/index.rs:152-214 (`build_ivf_index`): the blob `IvfIndex::to_bytes` would write, built on
/// the device from the resident column; `append_index_inplace` / `write_parquet_with_index` embed it unchanged.
pub fn build_index_blob(file: &ResidentFile, n_clusters: Option<usize>, max_iters: usize, seed: u64) -> Result<Vec<u8>> {
    file.table.build_index(n_clusters, max_iters, seed)?.to_bytes()
}

/// Site 3 alone -- the final assignment sweep src/ivf/index.rs:189-206 with host-trained centroids:
/// inverted lists in ascending row order, as the reference pushes them.
pub fn final_assignment(file: &ResidentFile, centroids: &[f32]) -> Result<Vec<Vec<u32>>> {
    let (assign, sizes) = file.table.kmeans_assign(centroids)?;
    let mut lists: Vec<Vec<u32>> = sizes.iter().map(|&s| Vec::with_capacity(s as usize)).collect();
    for (row, &c) in assign.iter().enumerate() {
        lists[c as usize].push(row as u32);
    }
    Ok(lists)
}

/// Site 6 -- the un-indexed arm: `array_distance(col, literal)` for the rows `[first_row, first_row + len)` of a
/// resident table, as the Float64 array a `ScalarUDFImpl::invoke_with_args` override returns (parity unpinned:
/// the reference's tests never reach DataFusion's built-in).
pub fn array_distance_column(file: &ResidentFile, literal: &[f64], first_row: usize, len: usize) -> Result<Float64Array> {
    let all = file.table.array_distance(literal)?;
    Ok(Float64Array::from(all[first_row..first_row + len].to_vec()))
}
