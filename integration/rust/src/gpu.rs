//! Safe wrappers over `pqv_sys` (the generated `extern "C"` block of include/pqv.h): the module a pq-vector
//! maintainer adds as `src/gpu.rs` (+ `src/pqv_sys.rs`) so that the loops of SURVEY section 8a run on the B200.
//!
//! NOT COMPILED in the authoring image (no cargo/rustc there): the ABI below is exercised through the same
//! prototypes from Python (`pq_vector_b200/_native.py`, `tests/test_gpu_*.py`), and `tests/test_abi.py` checks that
//! every `sys::pqv_*` call in this file names an exported function and passes the number of arguments the header
//! declares.  Nothing here computes: validation, ownership and error mapping only.
use crate::pqv_sys as sys;
use std::ffi::CStr;
use std::os::raw::c_int;
use std::sync::{Arc, OnceLock};

pub type Error = Box<dyn std::error::Error + Send + Sync>;
pub type Result<T> = std::result::Result<T, Error>;

/// `flags` of the top-k calls (include/pqv.h): summation order of the reference loop being replaced.
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum SumOrder {
    /// `squared_l2_distance`, src/ivf/index.rs:461-480
    Unroll4,
    /// `compute_distance_values`, src/df_vector/exec.rs:529-533
    Sequential,
}

impl SumOrder {
    fn bits(self) -> u32 {
        match self {
            SumOrder::Unroll4 => sys::PQV_SUM_UNROLL4,
            SumOrder::Sequential => sys::PQV_SUM_SEQ,
        }
    }
}

fn check(rc: c_int) -> Result<()> {
    if rc == sys::PQV_OK as c_int {
        return Ok(());
    }
    // thread-local message, valid until the next call on this thread: copy it out now
    let msg = unsafe { CStr::from_ptr(sys::pqv_last_error()) }.to_string_lossy().into_owned();
    Err(msg.into())
}

struct Ctx(*mut sys::PqvCtx);
// the library serialises access to a context internally (include/pqv.h, "Threading")
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { sys::pqv_destroy(self.0) }
    }
}

/// One per process.  `Gpu::global()` fails (it does not fall back to the CPU loops) when no B200 is visible.
#[derive(Clone)]
pub struct Gpu(Arc<Ctx>);

impl Gpu {
    pub fn new(device_ids: &[i32]) -> Result<Self> {
        let mut ctx = std::ptr::null_mut();
        let ids = if device_ids.is_empty() { std::ptr::null() } else { device_ids.as_ptr() };
        check(unsafe { sys::pqv_init(&mut ctx, ids, device_ids.len() as c_int) })?;
        Ok(Gpu(Arc::new(Ctx(ctx))))
    }

    pub fn global() -> Result<Self> {
        static GPU: OnceLock<std::result::Result<Gpu, String>> = OnceLock::new();
        GPU.get_or_init(|| Gpu::new(&[]).map_err(|e| e.to_string())).clone().map_err(Into::into)
    }

    fn raw(&self) -> *mut sys::PqvCtx {
        (self.0).0
    }

    /// Resident embedding column (replaces `Embeddings`, src/ivf/mod.rs:51-102, and the per-query re-read of
    /// src/ivf/search.rs:155-244).  Rows get ids in append order from 0, as the file's row numbers.
    pub fn create_table(&self, dim: usize, rows_hint: usize) -> Result<Table> {
        if dim == 0 {
            return Err("Embedding dimension must be > 0".into()); // src/ivf/mod.rs:57
        }
        let mut handle = 0u64;
        check(unsafe { sys::pqv_dataset_create(self.raw(), dim as u32, rows_hint as u64, &mut handle) })?;
        Ok(Table { gpu: self.clone(), handle, dim })
    }

    /// `IvfIndex::from_bytes`, src/ivf/index.rs:85-128, kept on the device (centroids + CSR lists).
    pub fn load_index(&self, blob: &[u8]) -> Result<Index> {
        let mut handle = 0u64;
        check(unsafe { sys::pqv_ivf_from_bytes(self.raw(), blob.as_ptr(), blob.len() as u64, &mut handle) })?;
        Ok(Index { gpu: self.clone(), handle })
    }

    /// `find_closest_centroids`, src/ivf/index.rs:130-149, for one query against host centroids.
    pub fn centroid_rank(&self, centroids: &[f32], dim: usize, query: &[f32], nprobe: usize) -> Result<Vec<usize>> {
        let n_clusters = centroids.len() / dim;
        let mut ids = vec![0u32; nprobe.min(n_clusters)];
        let mut n_eff = 0u32;
        check(unsafe {
            sys::pqv_centroid_rank(self.raw(), centroids.as_ptr(), n_clusters as u32, dim as u32, query.as_ptr(), 1,
                                   nprobe as u32, ids.as_mut_ptr(), &mut n_eff)
        })?;
        ids.truncate(n_eff as usize);
        Ok(ids.into_iter().map(|c| c as usize).collect())
    }

    /// `VectorTopKExec::topk_from_batches`, src/df_vector/exec.rs:257-277: begin / push per batch / finish.
    pub fn topk_stream(&self, query: &[f32], k: usize) -> Result<TopkStream> {
        let mut handle = 0u64;
        check(unsafe {
            sys::pqv_topk_stream_begin(self.raw(), query.len() as u32, query.as_ptr(), k as u32, sys::PQV_SUM_SEQ, &mut handle)
        })?;
        Ok(TopkStream { gpu: self.clone(), handle, k, dim: query.len(), done: false })
    }
}

pub struct Table {
    gpu: Gpu,
    handle: u64,
    dim: usize,
}

/// `(row_idx, distance)` pairs in the reference's output order.
pub struct Hits {
    pub row_idx: Vec<u32>,
    pub distance: Vec<f32>,
}

impl Hits {
    fn with_capacity(k: usize) -> (Vec<u32>, Vec<f32>, u32) {
        (vec![0u32; k], vec![0f32; k], 0u32)
    }
    fn finish(mut idx: Vec<u32>, mut dist: Vec<f32>, n: u32) -> Hits {
        idx.truncate(n as usize);
        dist.truncate(n as usize);
        Hits { row_idx: idx, distance: dist }
    }
}

impl Table {
    pub fn dim(&self) -> usize {
        self.dim
    }

    /// One record batch: `values` is the child buffer of the List<Float32> column (`list.values()`), dense row-major.
    pub fn append(&self, values: &[f32]) -> Result<()> {
        if values.len() % self.dim != 0 {
            return Err("Embedding data length must be a multiple of dim".into()); // src/ivf/mod.rs:83-86
        }
        check(unsafe { sys::pqv_dataset_append(self.gpu.raw(), self.handle, values.as_ptr(), (values.len() / self.dim) as u64) })
    }

    pub fn rows(&self) -> Result<usize> {
        let (mut rows, mut dim) = (0u64, 0u32);
        check(unsafe { sys::pqv_dataset_rows(self.gpu.raw(), self.handle, &mut rows, &mut dim) })?;
        Ok(rows as usize)
    }

    fn check_query(&self, query: &[f32]) -> Result<()> {
        if query.len() != self.dim {
            // src/ivf/search.rs:91-98
            return Err(format!("Query dimension mismatch: expected {}, got {}", self.dim, query.len()).into());
        }
        Ok(())
    }

    /// The re-rank loop of src/ivf/search.rs:112-141 over `rows_to_check` (candidate order): heap, `sqrt`, stable sort.
    pub fn topk_gather(&self, query: &[f32], rows_to_check: &[u32], k: usize) -> Result<Hits> {
        self.check_query(query)?;
        let (mut idx, mut dist, mut n) = Hits::with_capacity(k);
        check(unsafe {
            sys::pqv_l2_topk_gather(self.gpu.raw(), self.handle, query.as_ptr(), rows_to_check.as_ptr(),
                                    rows_to_check.len() as u64, k as u32, sys::PQV_SUM_UNROLL4 | sys::PQV_SQRT,
                                    idx.as_mut_ptr(), dist.as_mut_ptr(), &mut n)
        })?;
        Ok(Hits::finish(idx, dist, n))
    }

    /// The same loop when every row is a candidate (nprobe >= n_clusters is NOT this: list order differs from row order;
    /// this is the brute-force scan of BASELINE config 2).  Safe to call from many threads: concurrent calls are
    /// answered by one batched pass over the table.
    pub fn topk_all_rows(&self, query: &[f32], k: usize, order: SumOrder, sqrt: bool) -> Result<Hits> {
        self.check_query(query)?;
        let flags = order.bits() | if sqrt { sys::PQV_SQRT } else { 0 };
        let (mut idx, mut dist, mut n) = Hits::with_capacity(k);
        check(unsafe {
            sys::pqv_l2_topk_coalesced(self.gpu.raw(), self.handle, query.as_ptr(), k as u32, flags, idx.as_mut_ptr(),
                                       dist.as_mut_ptr(), &mut n)
        })?;
        Ok(Hits::finish(idx, dist, n))
    }

    /// `TopkBuilder::topk`, src/ivf/search.rs:83-142, with table and index resident: one host<->device round trip.
    pub fn ivf_search(&self, index: &Index, query: &[f32], k: usize, nprobe: usize) -> Result<Hits> {
        self.check_query(query)?;
        let (mut idx, mut dist, mut n) = Hits::with_capacity(k);
        check(unsafe {
            sys::pqv_ivf_search_coalesced(self.gpu.raw(), self.handle, index.handle, query.as_ptr(), k as u32, nprobe as u32,
                                          sys::PQV_SUM_UNROLL4 | sys::PQV_SQRT, idx.as_mut_ptr(), dist.as_mut_ptr(), &mut n)
        })?;
        Ok(Hits::finish(idx, dist, n))
    }

    /// `execute_with_candidates` + `topk_from_batches`, src/df_vector/exec.rs:207-277, over one indexed file.
    /// `row_mask`: the scan subtree's predicate as an Arrow boolean buffer over the file's rows (None = no filter).
    /// Returns the hits (squared distances, operator order) and the plan counters (candidate_rows, embeddings_fetched).
    pub fn vector_topk_indexed(&self, index: &Index, query: &[f32], k: usize, nprobe: usize, max_candidates: Option<usize>,
                               row_mask: Option<&[u8]>) -> Result<(Hits, u64, u64)> {
        self.check_query(query)?;
        if let Some(mask) = row_mask {
            if mask.len() < (self.rows()? + 7) / 8 {
                return Err("row mask shorter than the table".into());
            }
        }
        let (mut idx, mut dist, mut n) = Hits::with_capacity(k);
        let (mut candidates, mut scored) = (0u64, 0u64);
        check(unsafe {
            sys::pqv_vector_topk_indexed(self.gpu.raw(), self.handle, index.handle, query.as_ptr(), k as u32, nprobe as u32,
                                         sys::PQV_SUM_SEQ, max_candidates.map_or(sys::PQV_NO_CANDIDATE_CAP, |m| m as u64),
                                         row_mask.map_or(std::ptr::null(), |m| m.as_ptr()), idx.as_mut_ptr(),
                                         dist.as_mut_ptr(), &mut n, &mut candidates, &mut scored)
        })?;
        Ok((Hits::finish(idx, dist, n), candidates, scored))
    }

    /// Assignment sweeps of src/ivf/index.rs:193-201 (all rows) over the resident table: first minimum wins.
    pub fn kmeans_assign(&self, centroids: &[f32]) -> Result<(Vec<u32>, Vec<u64>)> {
        let n = self.rows()?;
        let n_clusters = centroids.len() / self.dim;
        let (mut assign, mut sizes) = (vec![0u32; n], vec![0u64; n_clusters]);
        check(unsafe {
            sys::pqv_kmeans_assign(self.gpu.raw(), self.handle, std::ptr::null(), n as u64, self.dim as u32, centroids.as_ptr(),
                                   n_clusters as u32, assign.as_mut_ptr(), sizes.as_mut_ptr())
        })?;
        Ok((assign, sizes))
    }

    /// One k-means++ sweep, src/ivf/index.rs:344-358, over the selected rows; the f32 sums of :359-370 stay with the caller.
    pub fn min_dist_update(&self, row_sel: &[u64], centroid: &[f32], init: bool, min_distances: &mut [f32]) -> Result<()> {
        assert_eq!(row_sel.len(), min_distances.len());
        check(unsafe {
            sys::pqv_min_dist_update(self.gpu.raw(), self.handle, std::ptr::null(), row_sel.as_ptr(), row_sel.len() as u64,
                                     self.dim as u32, centroid.as_ptr(), init as c_int, min_distances.as_mut_ptr())
        })
    }

    /// `build_ivf_index`, src/ivf/index.rs:152-214, on the device; `Index::to_bytes` gives `IvfIndex::to_bytes`' blob.
    pub fn build_index(&self, n_clusters: Option<usize>, max_iters: usize, seed: u64) -> Result<Index> {
        let mut handle = 0u64;
        check(unsafe {
            sys::pqv_ivf_build(self.gpu.raw(), self.handle, n_clusters.unwrap_or(0) as u32, max_iters as u32, seed, 0, &mut handle)
        })?;
        Ok(Index { gpu: self.gpu.clone(), handle })
    }

    /// DataFusion's built-in `array_distance` over the whole column (Float64 out), SURVEY row a10.
    pub fn array_distance(&self, literal: &[f64]) -> Result<Vec<f64>> {
        let mut out = vec![0f64; self.rows()?];
        check(unsafe {
            sys::pqv_array_distance(self.gpu.raw(), self.handle, literal.as_ptr(), literal.len() as u32, sys::PQV_METRIC_L2,
                                    out.as_mut_ptr())
        })?;
        Ok(out)
    }

    /// `array_distance` + `SortExec(fetch = k)` (+ the WHERE clause as a row mask) in one pass.
    pub fn array_distance_topk(&self, literal: &[f64], k: usize, row_mask: Option<&[u8]>) -> Result<(Vec<u32>, Vec<f64>)> {
        let (mut idx, mut dist, mut n) = (vec![0u32; k], vec![0f64; k], 0u32);
        check(unsafe {
            sys::pqv_array_distance_topk_filtered(self.gpu.raw(), self.handle, literal.as_ptr(), literal.len() as u32,
                                                  sys::PQV_METRIC_L2, k as u32, row_mask.map_or(std::ptr::null(), |m| m.as_ptr()),
                                                  idx.as_mut_ptr(), dist.as_mut_ptr(), &mut n)
        })?;
        idx.truncate(n as usize);
        dist.truncate(n as usize);
        Ok((idx, dist))
    }

    /// The k winning rows' vectors (what `build_batch_from_rows` needs for the vector column).
    pub fn read_rows(&self, row_ids: &[u32]) -> Result<Vec<f32>> {
        let mut out = vec![0f32; row_ids.len() * self.dim];
        check(unsafe { sys::pqv_dataset_read_rows(self.gpu.raw(), self.handle, row_ids.as_ptr(), row_ids.len() as u64, out.as_mut_ptr()) })?;
        Ok(out)
    }
}

impl Drop for Table {
    fn drop(&mut self) {
        unsafe { sys::pqv_dataset_drop(self.gpu.raw(), self.handle) };
    }
}

pub struct Index {
    gpu: Gpu,
    handle: u64,
}

impl Index {
    /// `(dim, n_clusters, total ids)`
    pub fn info(&self) -> Result<(usize, usize, u64)> {
        let (mut dim, mut clusters, mut ids) = (0u32, 0u32, 0u64);
        check(unsafe { sys::pqv_ivf_info(self.gpu.raw(), self.handle, &mut dim, &mut clusters, &mut ids) })?;
        Ok((dim as usize, clusters as usize, ids))
    }

    /// `IvfIndex::to_bytes`, src/ivf/index.rs:65-83 (same bytes).
    pub fn to_bytes(&self) -> Result<Vec<u8>> {
        let mut len = 0u64;
        // size query: cap 0 reports the length needed
        let _ = unsafe { sys::pqv_ivf_to_bytes(self.gpu.raw(), self.handle, std::ptr::null_mut(), 0, &mut len) };
        let mut out = vec![0u8; len as usize];
        check(unsafe { sys::pqv_ivf_to_bytes(self.gpu.raw(), self.handle, out.as_mut_ptr(), len, &mut len) })?;
        out.truncate(len as usize);
        Ok(out)
    }

    /// `IvfIndex::candidate_rows`, src/ivf/index.rs:57-63.
    pub fn candidate_rows(&self, query: &[f32], nprobe: usize) -> Result<Vec<u32>> {
        let (_, _, ids) = self.info()?;
        let mut out = vec![0u32; ids as usize];
        let mut n = 0u64;
        check(unsafe { sys::pqv_ivf_candidate_rows(self.gpu.raw(), self.handle, query.as_ptr(), nprobe as u32, out.as_mut_ptr(), ids, &mut n) })?;
        out.truncate(n as usize);
        Ok(out)
    }
}

impl Drop for Index {
    fn drop(&mut self) {
        unsafe { sys::pqv_ivf_drop(self.gpu.raw(), self.handle) };
    }
}

/// Streaming top-k over record batches; indices returned by `finish` count pushed rows in push order.
pub struct TopkStream {
    gpu: Gpu,
    handle: u64,
    k: usize,
    dim: usize,
    done: bool,
}

impl TopkStream {
    /// `values`: the batch's dense values with null / wrong-length rows already dropped (exec.rs:496-498, 526-528).
    pub fn push_f32(&mut self, values: &[f32]) -> Result<()> {
        check(unsafe { sys::pqv_topk_stream_push(self.gpu.raw(), self.handle, values.as_ptr(), (values.len() / self.dim) as u64) })
    }

    /// Float64 list items: narrowed to f32 before the subtraction, as exec.rs:542 does.
    pub fn push_f64(&mut self, values: &[f64]) -> Result<()> {
        check(unsafe { sys::pqv_topk_stream_push_f64(self.gpu.raw(), self.handle, values.as_ptr(), (values.len() / self.dim) as u64) })
    }

    pub fn finish(mut self) -> Result<Hits> {
        let (mut idx, mut dist, mut n) = Hits::with_capacity(self.k);
        self.done = true;
        check(unsafe { sys::pqv_topk_stream_finish(self.gpu.raw(), self.handle, idx.as_mut_ptr(), dist.as_mut_ptr(), &mut n) })?;
        Ok(Hits::finish(idx, dist, n))
    }
}

impl Drop for TopkStream {
    fn drop(&mut self) {
        if !self.done {
            // an abandoned stream (error between batches): finish into scratch to release the device buffers
            let (mut idx, mut dist, mut n) = Hits::with_capacity(self.k);
            unsafe { sys::pqv_topk_stream_finish(self.gpu.raw(), self.handle, idx.as_mut_ptr(), dist.as_mut_ptr(), &mut n) };
        }
    }
}
