"""GPU parity tests of the device-resident BestAssociationsHeap set (kg_select_*): the heaps the GPU keeps must be the
reference's std::priority_queue state -- same k-mers, same score bits, same rows, same pop order under ties -- for
both scan engines, across rounds / tiles, after overflow recovery, export / import, and the multi-shard log merge."""
import numpy as np
import pytest

import support as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kg(gpu_device):
    import kmersgwas_b200 as kg
    return kg


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _oracle_heaps(table, keep, scores, kbest):
    out = []
    for j in range(scores.shape[0]):
        k = kbest[j] if np.ndim(kbest) else kbest
        out.append(S.oracle_topk(table, keep, scores[j], int(k)).dump())
    return out


def _pop_order(layout):
    """(kmers, scores, rows) in libstdc++ layout order -> pop order, by pushing them in layout order into the oracle's
    restatement of std::priority_queue (a valid heap array pushed front to back keeps its layout)."""
    k, s, r = layout
    h = S.OracleHeap(max(len(k), 1))
    h.add_many(k, s, r)
    return h.dump()


def _assert_heaps_equal(dev_layouts, want):
    assert len(dev_layouts) == len(want)
    for j, (lay, w) in enumerate(zip(dev_layouts, want)):
        # the exported array must be a valid min-heap on the score
        s = lay[1]
        idx = np.arange(1, len(s))
        assert not np.any(s[(idx - 1) // 2] > s[idx]), f"phenotype {j}: exported layout is not a heap"
        got = _pop_order(lay)
        assert len(got[0]) == len(w[0]), f"phenotype {j}: size {len(got[0])} != {len(w[0])}"
        assert np.array_equal(got[0], w[0]), f"phenotype {j}: k-mers differ"
        assert np.array_equal(_bits(got[1]), _bits(w[1])), f"phenotype {j}: scores differ"
        assert np.array_equal(got[2], w[2]), f"phenotype {j}: rows differ"


def _case(n_file, n_pheno, n_rows, seed, tie_patterns=0):
    table = S.synth_table(seed, n_rows, n_file)
    if tie_patterns:
        table[:, 1:] = table[np.arange(n_rows) % tie_patterns, 1:].copy()
    y = S.synth_phenotypes(seed + 1, n_file, n_pheno)
    mc = S.min_count_of(n_file, 0.05, 5)
    idx = np.arange(n_file)
    keep, scores, kept = S.oracle_scan(table, n_file, idx // 64, idx % 64, y, mc)
    return table, y, mc, keep, scores, kept


def _run_select(kg, table, y, mc, kbest, engine, n_file, tiles, growth=None, device_rows=False, cand_cap=None):
    ctx = kg.Context.identity(n_file)
    ctx.set_option(kg.OPT_SCAN_ENGINE, engine)
    if growth is not None:
        ctx.set_option(kg.OPT_SELECT_GROWTH_PERMILLE, growth)
    if cand_cap is not None:
        ctx.set_option(kg.OPT_SELECT_CAND_CAP, cand_cap)
    ctx.set_phenotypes(y, mc)
    ctx.select_begin(kbest)
    keepalive = []
    r0 = 0
    for n in tiles:
        t = np.ascontiguousarray(table[r0:r0 + n])
        if device_rows:
            import torch
            d = torch.from_numpy(t.view(np.int64).copy()).cuda()
            keepalive.append(d)
            ctx.scan_submit(d.data_ptr(), n, r0)
        else:
            keepalive.append(t)
            ctx.scan_submit(t, n, r0)
        r0 += n
    return ctx, keepalive


@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("n_file,n_pheno,n_rows,kbest", [(241, 8, 30000, 500), (1135, 101, 12000, 301), (96, 3, 20000, 1000),
                                                         (64, 1, 5000, 1)])
def test_select_heaps_equal_oracle_heaps(kg, engine, n_file, n_pheno, n_rows, kbest):
    table, y, mc, keep, scores, kept = _case(n_file, n_pheno, n_rows, 900 + n_file)
    want = _oracle_heaps(table, keep, scores, kbest)
    tiles = [n_rows // 3, n_rows // 3 + 7, n_rows - 2 * (n_rows // 3) - 7]
    ctx, ka = _run_select(kg, table, y, mc, kbest, engine, n_file, tiles, growth=300)
    applied, dkept = ctx.select_sync()
    assert applied == n_rows and dkept == kept
    _assert_heaps_equal(ctx.select_heaps(), want)
    thr = ctx.select_thresholds()
    for j in range(n_pheno):
        full = len(want[j][1]) == kbest
        assert (thr[j] == want[j][1][0]) if full else (thr[j] == -1.0)
    ctx.close()


@pytest.mark.parametrize("engine", [1, 2])
def test_select_ties_and_device_rows(kg, engine):
    """identical presence patterns -> identical scores: which equal minimum is evicted and the pop order of equal
    scores follow libstdc++'s heap layout"""
    n_file, n_pheno, n_rows, kbest = 96, 4, 6000, 37
    table, y, mc, keep, scores, kept = _case(n_file, n_pheno, n_rows, 77, tie_patterns=23)
    want = _oracle_heaps(table, keep, scores, kbest)
    ctx, ka = _run_select(kg, table, y, mc, kbest, engine, n_file, [1000, 2500, 2500], growth=200, device_rows=True)
    applied, dkept = ctx.select_sync()
    assert applied == n_rows and dkept == kept
    _assert_heaps_equal(ctx.select_heaps(), want)
    ctx.close()


def test_select_per_phenotype_capacity_and_underfull(kg):
    """--first_phenotype_best: heap 0 has its own capacity; a capacity above the kept rows never fills"""
    n_file, n_pheno, n_rows = 131, 3, 4000
    table, y, mc, keep, scores, kept = _case(n_file, n_pheno, n_rows, 31)
    kbest = np.array([50, 700, 10000], dtype=np.uint64)
    want = _oracle_heaps(table, keep, scores, kbest)
    ctx, ka = _run_select(kg, table, y, mc, kbest, 0, n_file, [n_rows])
    applied, dkept = ctx.select_sync()
    assert applied == n_rows and dkept == kept
    _assert_heaps_equal(ctx.select_heaps(), want)
    assert len(want[2][0]) == kept
    ctx.close()


@pytest.mark.parametrize("tie_patterns", [0, 53])
def test_select_tiny_and_odd_capacities(kg, tie_patterns):
    """Heap capacities 1 .. 33 around the powers of two: the sift's counted double steps, its bounded tail, the
    single-left-child case and the record / slot-warp hand-off at every small heap shape; with tie_patterns the table
    has 53 distinct presence patterns only, so most scores tie and the layout decides what is evicted."""
    caps = [1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 31, 32, 33]
    n_file, n_pheno, n_rows = 96, len(caps), 5000
    table, y, mc, keep, scores, kept = _case(n_file, n_pheno, n_rows, 77, tie_patterns=tie_patterns)
    kbest = np.array(caps, dtype=np.uint64)
    want = _oracle_heaps(table, keep, scores, kbest)
    for engine in (1, 2):
        ctx, ka = _run_select(kg, table, y, mc, kbest, engine, n_file, [1500, 3500], growth=400)
        applied, dkept = ctx.select_sync()
        assert applied == n_rows and dkept == kept
        _assert_heaps_equal(ctx.select_heaps(), want)
        ctx.close()


def test_select_overflow_recovery(kg):
    """a candidate segment that is too small poisons the round; sync reports how far the heaps got and the caller
    resubmits the rest (the library shortens its rounds)"""
    n_file, n_pheno, n_rows, kbest = 130, 2, 30000, 64
    table, y, mc, keep, scores, kept = _case(n_file, n_pheno, n_rows, 5)
    # rising scores along the table: sort the rows by the first phenotype's score, so nearly every row is a candidate
    order = np.argsort(scores[0], kind="stable")
    table = np.ascontiguousarray(table[order])
    idx = np.arange(n_file)
    keep, scores, kept = S.oracle_scan(table, n_file, idx // 64, idx % 64, y, mc)
    want = _oracle_heaps(table, keep, scores, kbest)
    ctx = kg.Context.identity(n_file)
    ctx.set_option(kg.OPT_SELECT_CAND_CAP, 512)
    ctx.set_option(kg.OPT_SELECT_GROWTH_PERMILLE, 4000)
    ctx.set_phenotypes(y, mc)
    ctx.select_begin(kbest)
    done, overflows = 0, 0
    while done < n_rows:
        ctx.scan_submit(np.ascontiguousarray(table[done:]), n_rows - done, done)
        try:
            applied, _ = ctx.select_sync()
        except kg.KgError as e:
            assert e.status == kg.KG_ERR_HITS_OVERFLOW
            applied = e.rows_applied
            overflows += 1
            assert overflows < 50
        assert applied >= done
        done = applied
    assert overflows >= 1
    applied, dkept = ctx.select_sync()
    assert applied == n_rows and dkept == kept
    _assert_heaps_equal(ctx.select_heaps(), want)
    ctx.close()


def test_select_export_import_digest(kg):
    n_file, n_pheno, n_rows, kbest = 241, 6, 16000, 400
    table, y, mc, keep, scores, kept = _case(n_file, n_pheno, n_rows, 11)
    half = 9000
    a, ka = _run_select(kg, table, y, mc, kbest, 2, n_file, [half])
    applied, k1 = a.select_sync()
    state = a.select_export()
    b = kg.Context.identity(n_file)
    b.set_phenotypes(y, mc)
    b.select_begin(kbest)
    assert b.select_digest() != a.select_digest()
    b.select_import(state, applied, k1)
    assert b.select_digest() == a.select_digest()
    assert np.array_equal(b.select_thresholds(), a.select_thresholds())
    rest = np.ascontiguousarray(table[half:])
    for c in (a, b):
        c.scan_submit(rest, n_rows - half, half)
    ra, rb = a.select_sync(), b.select_sync()
    assert ra == rb == (n_rows, kept)
    assert a.select_digest() == b.select_digest()
    _assert_heaps_equal(b.select_heaps(), _oracle_heaps(table, keep, scores, kbest))
    a.close()
    b.close()


@pytest.mark.parametrize("use_floor", [False, True])
def test_select_shard_log_merge(kg, use_floor):
    """two row shards: shard 1 warm-starts from a shared prefix, logs what its heaps admit, and shard 0 (exact for
    its own rows) replays that log -> the sequential reference heaps"""
    import torch
    n_file, n_pheno, n_rows, kbest = 241, 5, 24000, 300
    table, y, mc, keep, scores, kept = _case(n_file, n_pheno, n_rows, 21, tie_patterns=4001)
    want = _oracle_heaps(table, keep, scores, kbest)
    cut, prefix = 13000, 3000
    s0, ka0 = _run_select(kg, table[:cut], y, mc, kbest, 2, n_file, [cut])
    s1 = kg.Context.identity(n_file)
    s1.set_option(kg.OPT_SCAN_ENGINE, 2)
    s1.set_phenotypes(y, mc)
    s1.select_begin(kbest, kg.SELECT_LOG)
    pre = np.ascontiguousarray(table[:prefix])
    s1.scan_submit(pre, prefix, 0)
    s1.select_log_reset()
    if use_floor:
        # threshold exchange: shard 0's heaps (rows before shard 1's block) + shard 1's own rows (none yet; its prefix
        # entries are masked out by min_row so that no row counts twice)
        s0.select_sync()
        s1.select_sync()
        sc = torch.empty((2, n_pheno, kbest), dtype=torch.float64, device="cuda")
        s0.select_export_scores(0, sc[0].data_ptr())
        s1.select_export_scores(cut, sc[1].data_ptr())
        s0.sync()
        s1.sync()
        assert bool((sc[1] < 0).all())
        s1.select_set_floor(sc.data_ptr(), 2)
        thr_before = s0.select_thresholds()
    tail = np.ascontiguousarray(table[cut:])
    s1.scan_submit(tail, n_rows - cut, cut)
    applied1, kept1 = s1.select_sync()
    assert applied1 == prefix + (n_rows - cut)
    applied0, kept0 = s0.select_sync()
    off, ent = s1.select_log()
    assert np.all(np.diff(off.astype(np.int64)) > 0)
    if use_floor:
        assert np.all(s1.select_thresholds() >= thr_before)      # the floor lifted shard 1 to shard 0's thresholds
    s0.select_replay(ent, off, n_rows - cut, kept1)
    applied, dkept = s0.select_sync()
    assert applied == n_rows and dkept == kept
    _assert_heaps_equal(s0.select_heaps(), want)
    s0.close()
    s1.close()


def test_select_mode_rejects_host_threshold_calls(kg):
    n_file = 64
    table, y, mc, keep, scores, kept = _case(n_file, 1, 500, 3)
    ctx = kg.Context.identity(n_file)
    ctx.set_phenotypes(y, mc)
    ctx.select_begin(10)
    with pytest.raises(kg.KgError):
        ctx.set_thresholds(np.zeros(1))
    with pytest.raises(kg.KgError):
        ctx.scan_fetch()
    ctx.select_end()
    ctx.set_thresholds(np.zeros(1))
    with pytest.raises(kg.KgError):
        ctx.select_begin(10 ** 6)      # does not fit shared memory: host replay path
    ctx.close()


def test_context_reuse_with_more_phenotypes(kg):
    """the filter's list buffers are sized per phenotype count; a second kg_scan_set_phenotypes with more
    16-column groups on the same context must reallocate them"""
    n_file, n_rows = 241, 9000
    table = S.synth_table(55, n_rows, n_file)
    idx = np.arange(n_file)
    mc = S.min_count_of(n_file, 0.05, 5)
    ctx = kg.Context.identity(n_file)
    ctx.set_option(kg.OPT_SCAN_ENGINE, 2)
    for n_pheno in (5, 101):
        y = S.synth_phenotypes(56 + n_pheno, n_file, n_pheno)
        keep_o, scores_o, kept_o = S.oracle_scan(table, n_file, idx // 64, idx % 64, y, mc)
        thr = np.array([np.quantile(scores_o[j][keep_o], 0.99) for j in range(n_pheno)])
        ctx.set_phenotypes(y, mc)
        ctx.set_thresholds(thr)
        ctx.scan_submit(table, n_rows, 0)
        hits, seen, kept = ctx.scan_fetch()
        assert seen == n_rows and kept == kept_o
        for j in range(n_pheno):
            sel = keep_o & (scores_o[j] > thr[j])
            assert np.array_equal(hits["row"][hits["pheno"] == j], np.nonzero(sel)[0])
    ctx.close()


def test_host_replay_recovers_from_hit_buffer_overflow(kg):
    """a device hit interval that overflows is dropped as a whole; the association driver redoes that
    round in shorter rounds -- also when it was the last round of a call (kgh_associate_finish) -- instead of failing"""
    n_file, n_pheno, n_rows, kbest = 130, 40, 20000, 2000
    table, y, mc, keep, scores, kept = _case(n_file, n_pheno, n_rows, 61)
    want = _oracle_heaps(table, keep, scores, kbest)
    idx = np.arange(n_file)
    sess = kg.Session(n_file, (idx // 64).astype(np.uint32), (idx % 64).astype(np.uint32), y, mc, kbest)
    sess.set_option(kg.OPT_HIT_CAPACITY, 3000)      # far below n_pheno x kbest: the fill rounds and the first warm rounds overflow
    sess.associate(table, n_rows, 0)
    assert sess.tested(0) == kept
    for j in range(n_pheno):
        k, s, r = sess.heap(j)
        assert np.array_equal(k, want[j][0]) and np.array_equal(_bits(s), _bits(want[j][1])) and np.array_equal(r, want[j][2])
    sess.close()
