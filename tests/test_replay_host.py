"""CPU-only: pqv_replay_candidates (the host half of every top-k call: heap replay over the entrant keys, with the
no-replay shortcut when the result cannot depend on the heap layout) against the oracle's restatement of the
reference loop src/ivf/search.rs:112-141 / src/df_vector/exec.rs:264-275.  No GPU needed: the function is pure host."""
import numpy as np
import pytest

import oracle as O
import pq_vector_b200 as P


def _keys(dist, pos):
    return (dist.astype(np.float32).view(np.uint32).astype(np.uint64) << np.uint64(32)) | pos.astype(np.uint64)


def _entrants(dist, k):
    """rows the reference heap admits (push while len < k, else d < root), in position order"""
    import heapq
    h, keep = [], []
    for i, d in enumerate(dist):
        if len(h) < k:
            heapq.heappush(h, -d)
            keep.append(i)
        elif d < -h[0]:
            heapq.heapreplace(h, -d)
            keep.append(i)
    return np.array(keep, dtype=np.int64)


@pytest.mark.parametrize("levels", [0, 3, 17, 1000])       # 0 = continuous (no ties), small = many exact ties
@pytest.mark.parametrize("k", [1, 7, 100])
@pytest.mark.parametrize("do_sqrt", [False, True])
def test_replay_matches_oracle(levels, k, do_sqrt):
    rng = np.random.default_rng(levels * 1000 + k)
    n = 5000
    dist = rng.random(n).astype(np.float32) * 4 + 0.5
    if levels:
        dist = (np.floor(dist * levels) / levels).astype(np.float32)
    er, ed = O.heap_topk(dist, None, k, do_sqrt)
    flags = P.PQV_SQRT if do_sqrt else 0
    # (a) exactly the admitted rows, shuffled; (b) a superset: every row
    ent = _entrants(dist, k)
    for sel in (rng.permutation(ent), rng.permutation(n)):
        rows, d = P.replay_candidates(_keys(dist[sel], sel), k, flags)
        assert rows.tolist() == er.tolist()
        assert d.view(np.uint32).tolist() == ed.view(np.uint32).tolist()


def test_replay_fewer_candidates_than_k_and_row_map():
    dist = np.array([3.0, 1.0, 2.0, 1.0], dtype=np.float32)
    ids = np.array([40, 30, 20, 10], dtype=np.uint32)
    er, ed = O.heap_topk(dist, ids, 10, True)
    rows, d = P.replay_candidates(_keys(dist, np.arange(4)), 10, P.PQV_SQRT, row_ids=ids)
    assert rows.tolist() == er.tolist() and d.tolist() == ed.tolist()


def test_replay_sqrt_collision_takes_the_heap_path():
    # two distinct squared distances whose f32 square roots are equal: the output order is the heap's, not (d, pos)
    a = np.float32(1.0)
    b = np.nextafter(a, np.float32(2))
    assert a != b and np.sqrt(a) == np.sqrt(b)
    dist = np.array([b, 5.0, a, 4.0, 3.0], dtype=np.float32)
    for k in (2, 3, 5):
        er, ed = O.heap_topk(dist, None, k, True)
        rows, d = P.replay_candidates(_keys(dist, np.arange(5)), k, P.PQV_SQRT)
        assert rows.tolist() == er.tolist() and d.view(np.uint32).tolist() == ed.view(np.uint32).tolist()


def _chunked_threshold_superset(dist, k, chunk=2048):
    """numpy model of prefix_entrants_kernel (pq_vector_b200/csrc/pqv_tie.cuh): the first chunk whole, then every row
    below the exact k-th smallest distance of all rows BEFORE its chunk (the threshold in force at the chunk's start)."""
    n = dist.size
    keep = list(range(min(n, chunk)))
    for base in range(chunk, n, chunk):
        thr = np.partition(dist[:base], k - 1)[k - 1]          # k <= chunk <= base
        keep.extend((base + np.nonzero(dist[base:base + chunk] < thr)[0]).tolist())
    return np.array(keep, dtype=np.int64)


@pytest.mark.parametrize("levels", [0, 5, 200])
@pytest.mark.parametrize("k", [1, 10, 100, 1024])
@pytest.mark.parametrize("trend", ["random", "descending"])
def test_chunked_threshold_entrants_are_a_superset_of_the_heap_admissions(levels, k, trend):
    """The batched tie path replays the reference heap over the rows a per-chunk threshold lets
    through instead of over every row: that set must contain every row the reference heap admits, and the replay over it
    must return the reference answer (order among equal distances included)."""
    rng = np.random.default_rng(levels + 7 * k)
    n = 20000
    dist = rng.random(n).astype(np.float32) * 4 + 0.5
    if trend == "descending":                                   # every row is admitted: the superset is everything
        dist = (dist * 0.01 + np.linspace(9, 1, n)).astype(np.float32)
    if levels:
        dist = (np.floor(dist * levels) / levels).astype(np.float32)
    sup = _chunked_threshold_superset(dist, k)
    adm = _entrants(dist, k)
    assert np.isin(adm, sup).all()
    if trend == "random" and k <= 100:
        assert sup.size < 2048 + 40 * k                         # and it is small: first chunk + O(k log(n / chunk))
    for do_sqrt in (False, True):
        er, ed = O.heap_topk(dist, None, k, do_sqrt)
        rows, d = P.replay_candidates(_keys(dist[sup], sup), k, P.PQV_SQRT if do_sqrt else 0)
        assert rows.tolist() == er.tolist()
        assert d.view(np.uint32).tolist() == ed.view(np.uint32).tolist()


@pytest.mark.parametrize("levels", [0, 50])
@pytest.mark.parametrize("k", [1, 10, 100])
@pytest.mark.parametrize("do_sqrt", [False, True])
def test_short_prefix_plus_thresholded_candidates_replay_to_the_reference_answer(levels, k, do_sqrt):
    """Model of the plan in DESIGN.md section 8 for the tie queries of the masked IVF pass: the pass knows every candidate
    with d < tau (tau >= the final k-th smallest distance) but not the early rows the heap admits while its threshold is
    still above tau.  Those can only sit before the position P* where the k-th candidate of the sequence appears -- from
    P* on the heap's threshold is below tau -- so exact distances are needed for the prefix [0, P*] only:
    replay(admissions of the exact prefix + candidates behind it) must equal the reference loop over the whole sequence."""
    rng = np.random.default_rng(1000 * levels + 10 * k + do_sqrt)
    n = 30000
    dist = (rng.random(n).astype(np.float32) * 4 + 0.5)
    if levels:
        dist = (np.floor(dist * levels) / levels).astype(np.float32)
    er, ed = O.heap_topk(dist, None, k, do_sqrt)
    # tau as the batched pass derives it: the k-th smallest distance of a SAMPLE of the rows (an upper bound of the final one)
    sample = rng.choice(n, n // 16, replace=False)
    tau = np.partition(dist[sample], k - 1)[k - 1]
    cand = np.nonzero(dist < tau)[0]                       # strict: rows equal to tau are not guaranteed to be candidates
    if cand.size < k:                                      # (possible with heavy ties) -> the whole sequence is the prefix
        p_star = n - 1
    else:
        p_star = int(cand[k - 1])                          # position of the k-th candidate in sequence order
    prefix_adm = _entrants(dist[:p_star + 1], k)           # exact distances over the short prefix only
    later = cand[cand > p_star]
    sel = np.concatenate([prefix_adm, later])
    assert np.isin(_entrants(dist, k), sel).all()          # nothing the reference heap admits is missing
    if levels == 0 and k == 100:
        assert p_star < n // 8                             # and the prefix is short: ~ k / (candidates per row)
    rows, d = P.replay_candidates(_keys(dist[sel], sel), k, P.PQV_SQRT if do_sqrt else 0)
    assert rows.tolist() == er.tolist()
    assert d.view(np.uint32).tolist() == ed.view(np.uint32).tolist()


@pytest.mark.parametrize("levels", [0, 3, 40])
@pytest.mark.parametrize("k", [1, 10, 100])
@pytest.mark.parametrize("world", [2, 8])
def test_rank_threshold_prefilter_keeps_the_reference_answer(levels, k, world):
    """Model of peer_wait_pack_kernel (pqv_peer.cuh): rows are numbered rank by rank, every rank publishes the rows its LOCAL
    heap admits plus its final k-th smallest distance, and the reader drops rank r's entrants with
    d >= T_r = min over ranks j < r of (rank j's k-th smallest distance) -- when the global heap reaches rank r's rows it
    already holds the k best of the earlier ranks, and admission needs d < threshold.  The replay over what is left must be
    the reference loop over the whole table, ties included (levels: coarse grids -> thousands of bit-equal distances)."""
    rng = np.random.default_rng(100 * levels + 10 * k + world)
    per = 4000
    n = per * world
    dist = rng.random(n).astype(np.float32) * 4 + 0.5
    if levels:
        dist = (np.floor(dist * levels) / levels).astype(np.float32)
    dist[rng.integers(0, n, 5)] = np.float32(0.25)                     # a few global winners scattered over the ranks
    kept, T, raw = [], np.float32(np.inf), 0
    for r in range(world):
        lo = r * per
        local = dist[lo:lo + per]
        ent = lo + _entrants(local, k)                                  # what rank r's scan emits (global positions)
        raw += ent.size
        kept.append(ent[dist[ent] < T])                                 # the reader's filter
        kth = np.partition(local, k - 1)[k - 1] if per >= k else np.float32(np.inf)
        T = min(T, kth)
    sel = np.concatenate(kept)
    assert np.isin(_entrants(dist, k), sel).all()                       # nothing the global heap admits was dropped
    if levels == 0 and world == 8 and k == 100:
        assert sel.size < raw // 3                                      # and most of the later ranks' entrants are gone
    for do_sqrt in (False, True):
        er, ed = O.heap_topk(dist, None, k, do_sqrt)
        rows, d = P.replay_candidates(_keys(dist[sel], sel), k, P.PQV_SQRT if do_sqrt else 0)
        assert rows.tolist() == er.tolist()
        assert d.view(np.uint32).tolist() == ed.view(np.uint32).tolist()
