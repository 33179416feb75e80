//! Stand-in crate root: in pq-vector these three lines go into src/lib.rs next to `pub mod ivf; pub mod df_vector;`.
pub mod call_sites;
pub mod gpu;
pub mod pqv_sys;
