"""CPU-only model of the device-side k-means++ pick (pqv_kmeanspp.cuh).

The reference picks centroid i by walking the min-distance array with ONE serial f32 chain and stopping at the first slot
whose running sum reaches the threshold (src/ivf/index.rs:372-383).  A serial chain of 50 000 dependent adds per pick is what
kept the build on the host.  The kernel runs the SAME chain exactly, but in parallel, using what an f32 add does inside one
binade: with x = m * u (u = ulp of x's binade, 2^23 <= m < 2^24) and x + d < 2^24 u,

    fl(x + d) = (m + r(d / u)) * u ,   r = round to nearest, ties to the even neighbour of m + d/u

so every element acts on the integer m as  m -> m + a + tie * ((m + a) & 1)  -- a function that only depends on the parity of
m.  Such functions compose associatively (a pair of deltas, one per input parity), hence a block of elements is a prefix
scan; the first element whose result leaves the binade (m >= 2^24) is re-done with a real f32 add and the scan restarts
behind it in the new binade.  This file restates that block procedure in integer arithmetic (the kernel's formulas, one
"thread" per 4 elements) and checks it against the literal serial chain in numpy float32 on adversarial inputs."""
import numpy as np
import pytest

SAT = 1 << 25
TOP = 1 << 24


def serial_pick(md, thr):
    """the reference loop: first slot with cumsum >= thr (None: never reached), and the final cumsum"""
    x = np.float32(0.0)
    for s, d in enumerate(md):
        x = np.float32(x + d)
        if x >= thr:
            return s, x
    return None, x


def classify(db, E):
    """element bits -> (a, tie) in units of the state's ulp 2^(E-150); SAT = leaves the binade for sure"""
    db = int(db)
    if db & 0x7FFFFFFF == 0:
        return 0, 0
    Ed = db >> 23
    Md = db & 0x7FFFFF
    if Ed == 0:
        Ed = 1
    else:
        Md |= 0x800000
    sh = E - Ed
    if sh <= 0:
        return SAT, 0
    if sh >= 25:
        return 0, 0
    q = Md >> sh
    rem = Md & ((1 << sh) - 1)
    half = 1 << (sh - 1)
    return q + (1 if rem > half else 0), (1 if rem == half else 0)


def apply(fn, p):
    return fn[p]


def compose(g, f):
    """g first, then f: both are (delta for even input, delta for odd input), saturating"""
    out = []
    for p in (0, 1):
        a = g[p]
        out.append(min(a + f[(p + a) & 1], SAT))
    return tuple(out)


def elem_fn(a, tie):
    out = []
    for p in (0, 1):
        cur = a
        if tie:
            cur += (p + cur) & 1
        out.append(min(cur, SAT))
    return tuple(out)


def block_pick(md, thr, pro=16, threads=64, ept=4, serial_burst=32):
    """the kernel's procedure (small thread count so that short arrays cross many blocks)"""
    md = np.asarray(md, np.float32)
    n = md.size
    bits = md.view(np.uint32)
    x = np.float32(0.0)
    base = 0

    def serial(cnt):
        nonlocal x, base
        for _ in range(cnt):
            if base >= n:
                return None
            x = np.float32(x + md[base])
            base += 1
            if x >= thr:
                return base - 1
        return None

    hit = serial(min(pro, n))
    if hit is not None:
        return hit, x
    B = threads * ept
    while base < n:
        xb = int(np.float32(x).view(np.uint32))
        E = xb >> 23
        if E < 24 or E >= 254:          # zero / tiny / huge state: plain adds for a while
            hit = serial(serial_burst)
            if hit is not None:
                return hit, x
            continue
        m0 = (xb & 0x7FFFFF) | 0x800000
        u = np.float32(2.0) ** np.float32(E - 150)
        # per thread: element functions and their composition
        cls = []
        tf = []
        for t in range(threads):
            fs = (0, 0)
            row = []
            for e in range(ept):
                j = base + t * ept + e
                a, tie = classify(bits[j], E) if j < n else (0, 0)
                row.append((a, tie))
                fs = compose(fs, elem_fn(a, tie))
            cls.append(row)
            tf.append(fs)
        # exclusive scan over the threads (the kernel: warp shuffles + one pass over the warp totals)
        pre = [(0, 0)]
        for t in range(threads - 1):
            pre.append(compose(pre[-1], tf[t]))
        event = None   # (relative position, kind, m before the element)
        m_end = None
        for t in range(threads):
            m = m0 + pre[t][m0 & 1]
            if m >= TOP:
                ev = (t * ept, 0, None)
                if event is None or ev[0] < event[0]:
                    event = ev
                continue
            for e in range(ept):
                a, tie = cls[t][e]
                mn = m + a + (((m + a) & 1) if tie else 0)
                if mn >= TOP:
                    ev = (t * ept + e, 0, m)
                elif np.float32(mn) * u >= thr and base + t * ept + e < n:
                    ev = (t * ept + e, 1, m)
                else:
                    m = mn
                    continue
                if event is None or ev[0] < event[0]:
                    event = ev
                break
            else:
                if t == threads - 1:
                    m_end = m
        if event is None:
            x = np.float32(m_end) * u
            base += B
            continue
        pos, kind, mb = event
        assert mb is not None          # the first event always comes from a thread that started inside the binade
        if kind == 1:
            return base + pos, np.float32(np.float32(mb) * u + md[base + pos])
        x = np.float32(np.float32(mb) * u + md[base + pos])      # the crossing add itself: a real f32 add
        base += pos + 1
        if x >= thr:
            return base - 1, x
        if pos < 8:                     # slow progress (alternating magnitudes): plain adds for a while
            hit = serial(serial_burst)
            if hit is not None:
                return hit, x
    return None, x


def _cases():
    rng = np.random.default_rng(5)
    yield "uniform", (rng.random(3000).astype(np.float32) * 100 + 78)
    yield "tiny-to-huge", np.float32(2.0) ** rng.integers(-30, 30, 2000).astype(np.float32)
    yield "integers (ties once the ulp passes 1)", rng.integers(0, 7, 4000).astype(np.float32) * np.float32(4099)
    yield "halves", (rng.integers(0, 64, 4000).astype(np.float32) + np.float32(0.5)) * np.float32(1024)
    yield "zeros then values", np.concatenate([np.zeros(700, np.float32), rng.random(900).astype(np.float32)])
    yield "subnormals", (rng.random(500) * 1e-40).astype(np.float32)
    yield "one giant", np.concatenate([rng.random(300).astype(np.float32), [np.float32(3e9)], rng.random(800).astype(np.float32)])
    yield "all equal", np.full(5000, np.float32(127.99999), np.float32)
    yield "squared distances", ((rng.random((2500, 16)) - rng.random(16)) ** 2).sum(1).astype(np.float32)


@pytest.mark.parametrize("name,md", list(_cases()), ids=[c[0] for c in _cases()])
def test_block_procedure_equals_the_serial_chain(name, md):
    total_s, total = serial_pick(md, np.float32(np.inf))
    assert total_s is None
    got_s, got_x = block_pick(md, np.float32(np.inf))
    assert got_s is None and np.float32(got_x).view(np.uint32) == np.float32(total).view(np.uint32), name
    rng = np.random.default_rng(len(md))
    for r in list(rng.random(12)) + [0.0, 1.0 - 2 ** -24, 1.0]:
        thr = np.float32(np.float32(r) * total)
        es, ex = serial_pick(md, thr)
        gs, gx = block_pick(md, thr)
        assert gs == es, (name, r)
        if es is not None:
            assert np.float32(gx).view(np.uint32) == np.float32(ex).view(np.uint32)


def test_block_shapes():
    rng = np.random.default_rng(9)
    md = (rng.random(1777) * 50).astype(np.float32)
    _, total = serial_pick(md, np.float32(np.inf))
    for threads, ept, pro in [(32, 1, 1), (32, 4, 0), (1024, 4, 128), (7, 3, 5)]:
        for r in (0.1, 0.5, 0.93):
            thr = np.float32(np.float32(r) * total)
            assert block_pick(md, thr, pro=pro, threads=threads, ept=ept)[0] == serial_pick(md, thr)[0]


def classify_float(d, E):
    """the kernel's form of classify(): three f32 adds against a magic constant with the running sum's ulp"""
    f = np.float32
    u = f(2.0) ** f(E - 150)
    half_u = f(2.0) ** f(E - 151)
    big = f(1.5) * f(2.0) ** f(E - 127)
    dmax = f(2.0) ** f(E - 128)
    d = f(d)
    if d >= dmax:
        return SAT, 0
    r = f(f(big + d) - big)
    res = f(d - r)
    tie = 1 if abs(res) == half_u else 0
    af = f(r - u) if res == -half_u else r
    return int(f(af / u)), tie


def test_float_classification_equals_the_integer_one():
    rng = np.random.default_rng(77)
    for E in (25, 60, 127, 140, 150, 200, 253):
        scale = np.float32(2.0) ** np.float32(E - 127)
        vals = np.concatenate([
            (rng.random(4000).astype(np.float32) * scale * np.float32(2.0) ** rng.integers(-30, 1, 4000).astype(np.float32)),
            # exact ties and near-ties at every shift
            np.array([(q + 0.5) * 2.0 ** (E - 150) for q in range(0, 40)], np.float32),
            np.array([(q + 0.5) * 2.0 ** (E - 150) * (1 + s * 2.0 ** -20) for q in range(1, 20) for s in (-1, 1)], np.float32),
            np.array([0.0, 2.0 ** (E - 151), 2.0 ** (E - 152), 2.0 ** (E - 129), 2.0 ** (E - 128), 2.0 ** (E - 127)], np.float32),
        ]).astype(np.float32)
        dmax = np.float32(2.0) ** np.float32(E - 128)
        for d in vals:
            a1, t1 = classify(np.float32(d).view(np.uint32), E)
            a2, t2 = classify_float(d, E)
            if d >= dmax:
                assert a2 == SAT and a1 >= (1 << 22)     # both lead to "leaves the binade" or a huge step; the kernel re-adds it
            else:
                assert (a1, t1) == (a2, t2), (E, float(d))


# ---- the kernel's element classification and threshold, as it computes them (round 2: no float -> int conversions) ----------
def kernel_classify(d, E):
    """pqv_kmeanspp.cuh: t = big + d is d rounded to a multiple of u on top of big = 1.5 * 2^(E-127); a is read off the BIT
    PATTERNS (bits(t) - bits(big)), the residual d - (t - big) is exact and marks the tie; a = the round-down choice of a tie"""
    d = np.float32(d)
    big_bits = np.uint32((E << 23) | 0x400000)
    big = big_bits.view(np.float32)
    half_u = np.uint32((E - 24) << 23).view(np.float32)
    dmax = np.uint32((E - 1) << 23).view(np.float32)
    if d >= dmax:
        return SAT, 0
    t = np.float32(big + d)
    res = np.float32(d - np.float32(t - big))
    tie = int(abs(res) == half_u)
    a = int(t.view(np.uint32)) - int(big_bits) - (1 if res == -half_u else 0)
    return a, tie


def test_bit_pattern_classification_equals_the_integer_model():
    rng = np.random.default_rng(11)
    n_checked = 0
    for E in [25, 26, 40, 100, 127, 128, 150, 200, 253]:
        u_exp = E - 150
        cases = []
        # random elements over 40 binades below the state's, exact half-way cases, multiples of u, sub-normals, zero
        for sh in range(1, 40):
            Ed = E - sh
            if Ed < 1:
                break
            for _ in range(40):
                cases.append(np.uint32((Ed << 23) | int(rng.integers(0, 1 << 23))).view(np.float32))
            if sh <= 23:
                half = 1 << (sh - 1)
                for q in (0, 1, 2, 3, 12345):  # mantissa = q * 2^sh + half: a rounding tie in the state's units
                    Md = ((q << sh) | half) & 0x7FFFFF
                    cases.append(np.uint32((Ed << 23) | Md).view(np.float32))
                    cases.append(np.uint32((Ed << 23) | ((q << sh) & 0x7FFFFF)).view(np.float32))
        cases += [np.float32(0.0), np.uint32(1).view(np.float32), np.uint32(0x7FFFFF).view(np.float32)]
        for d in cases:
            if not (d >= 0):
                continue
            a_ref, tie_ref = classify(np.float32(d).view(np.uint32), E)
            a, tie = kernel_classify(d, E)
            if a_ref >= (1 << 22):  # a quarter of the binade or more: the kernel sends it to the real add (SAT) -- allowed
                assert a == SAT or (a, tie) == (a_ref, tie_ref)
                continue
            assert (a, tie) == (a_ref, tie_ref), (E, u_exp, float(d), a, tie, a_ref, tie_ref)
            n_checked += 1
    assert n_checked > 5000


def test_integer_threshold_equals_the_float_compare():
    """m u >= thr  <=>  m >= ceil(thr / u), with thr / u computed as an f32 multiply by the power of two 1 / u"""
    rng = np.random.default_rng(12)
    for E in [25, 60, 127, 150, 230, 253]:
        u = np.uint32((E - 23) << 23).view(np.float32)
        inv_u = np.uint32((277 - E) << 23).view(np.float32)
        ms = [1 << 23, (1 << 23) + 1, (1 << 24) - 1] + [int(v) for v in rng.integers(1 << 23, 1 << 24, 200)]
        with np.errstate(over="ignore", under="ignore"):
            thrs = [np.float32(0.0), np.float32(np.inf)]
            for m in ms[:40]:
                x = np.float32(np.float32(m) * u)
                thrs += [x, np.nextafter(x, np.float32(np.inf)), np.nextafter(x, np.float32(0))]
            thrs += [np.float32(float(u) * 0.3), np.float32(float(u) * 2 ** 30) if E < 200 else np.float32(3e38)]
            for thr in thrs:
                tq = np.float32(thr * inv_u)
                thr_m = int(np.ceil(tq)) if tq < np.float32(33554432.0) else SAT
                for m in ms:
                    lhs = bool(np.float32(np.float32(m) * u) >= thr)
                    assert lhs == (m >= thr_m), (E, m, float(thr), thr_m)
