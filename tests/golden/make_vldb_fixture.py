"""Generates tests/golden/vldb_2025_embeddings.npz from the reference's shipped dataset
(/root/reference/data/vldb_2025.parquet, 496 x 4096 f32, List<Float32>).  Run once in the
authoring container (the GPU box has no /root/reference); the output is committed.
Only the embedding column is kept (bit-exact f32); it is config C1 of BASELINE.json."""
import sys
import numpy as np
import pyarrow.parquet as pq

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/vldb_2025.parquet"
t = pq.read_table(src)
col = "embedding"
arr = t.column(col).combine_chunks()
vals = arr.values.to_numpy(zero_copy_only=False).astype(np.float32, copy=False)
n = len(arr)
emb = vals.reshape(n, -1)
print(col, emb.shape, emb.dtype)
np.savez_compressed("tests/golden/vldb_2025_embeddings.npz", embedding=emb, column=np.array(col))
