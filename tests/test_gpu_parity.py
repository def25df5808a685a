"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs and against the committed golden fixtures (reference outputs).  Bit-exact for scores
(IEEE double bit patterns), k-mer identities, kept counts and kinship integers."""
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


def _torch_rows(table):
    import torch
    return torch.from_numpy(table.view(np.int64).copy()).cuda()


# ------------------------------------------------------------------------------ generator
def test_synth_device_matches_host(kg):
    import torch
    for n_file in (64, 131, 1135):
        ctx = kg.Context.identity(n_file)
        n_rows, first = 3000, 12345
        buf = torch.zeros(n_rows * (ctx.w_file + 1), dtype=torch.int64, device="cuda")
        ctx.synth_rows_device(77, first, n_rows, buf.data_ptr())
        ctx.sync()
        dev = buf.cpu().numpy().view(np.uint64).reshape(n_rows, ctx.w_file + 1)
        assert np.array_equal(dev, S.synth_table(77, n_rows, n_file, first_row=first))
        ctx.close()


# ------------------------------------------------------------------------------ dense scores
@pytest.mark.parametrize("n_file,n_pheno", [(64, 1), (65, 2), (127, 3), (128, 4), (129, 5), (241, 8), (1135, 11)])
def test_dense_scores_bit_exact_identity(kg, n_file, n_pheno):
    n_rows = 2500
    table = S.synth_table(n_file, n_rows, n_file)
    y = S.synth_phenotypes(n_file + 1, n_file, n_pheno)
    mc = S.min_count_of(n_file, 0.05, 5)
    names = [f"s{i}" for i in range(n_file)]
    mw, mb = S.column_map(names, names)
    keep_o, scores_o, kept_o = S.oracle_scan(table, n_file, mw, mb, y, mc)
    ctx = kg.Context(n_file, mw, mb)
    ctx.set_phenotypes(y, mc)
    keep, scores = ctx.scores_dense(table, n_rows)
    assert np.array_equal(keep, keep_o)
    assert np.array_equal(_bits(scores[:, keep]), _bits(scores_o[:, keep_o]))
    # same tile already resident on the device
    dev = _torch_rows(table)
    keep2, scores2 = ctx.scores_dense(dev.data_ptr(), n_rows)
    assert np.array_equal(keep2, keep_o)
    assert np.array_equal(_bits(scores2[:, keep2]), _bits(scores_o[:, keep_o]))
    ctx.close()


@pytest.mark.parametrize("n_file,n_used,n_pheno", [(300, 211, 2), (1135, 1000, 9), (130, 64, 1)])
def test_dense_scores_subset_permuted_columns(kg, n_file, n_used, n_pheno):
    rng = np.random.default_rng(n_file)
    names = [f"s{i}" for i in range(n_file)]
    used = [names[i] for i in rng.permutation(n_file)[:n_used]]
    n_rows = 1800
    table = S.synth_table(5, n_rows, n_file)
    y = S.synth_phenotypes(6, n_used, n_pheno)
    mc = S.min_count_of(n_used, 0.05, 5)
    mw, mb = S.column_map(names, used)
    keep_o, scores_o, _ = S.oracle_scan(table, n_file, mw, mb, y, mc)
    ctx = kg.Context(n_file, mw, mb)
    ctx.set_phenotypes(y, mc)
    keep, scores = ctx.scores_dense(table, n_rows)
    assert np.array_equal(keep, keep_o)
    assert np.array_equal(_bits(scores[:, keep]), _bits(scores_o[:, keep_o]))
    ctx.close()


def test_dense_edge_rows(kg):
    """empty tile, single row, all-zero / all-one rows (fail the MAC filter), ragged tail."""
    n_file = 131
    ctx = kg.Context.identity(n_file)
    y = S.synth_phenotypes(1, n_file, 3)
    mc = S.min_count_of(n_file, 0.05, 5)
    ctx.set_phenotypes(y, mc)
    ctx.scan_submit(np.zeros((0, 4), dtype=np.uint64), 0)      # empty tile is a no-op
    hits, seen, kept = ctx.scan_fetch()
    assert len(hits) == 0 and seen == 0 and kept == 0
    table = S.synth_table(9, 515, n_file)
    table[0, 1:] = 0
    table[1, 1:] = np.uint64(0xFFFFFFFFFFFFFFFF)
    table[1, -1] = np.uint64((1 << (n_file % 64)) - 1)
    idx = np.arange(n_file)
    keep_o, scores_o, _ = S.oracle_scan(table, n_file, idx // 64, idx % 64, y, mc)
    assert not keep_o[0] and not keep_o[1]
    for n in (1, 2, 63, 64, 65, 515):
        keep, scores = ctx.scores_dense(np.ascontiguousarray(table[:n]), n)
        assert np.array_equal(keep, keep_o[:n])
        assert np.array_equal(_bits(scores[:, keep]), _bits(scores_o[:, :n][:, keep_o[:n]]))
    ctx.close()


def test_non_finite_and_extreme_phenotypes(kg):
    """inf / huge / denormal phenotype values take the same predicated-add path as the reference's blend."""
    n_file = 96
    y = S.synth_phenotypes(3, n_file, 4)
    y[0, 5] = np.inf
    y[1, 7] = 3.0e38
    y[1, 9] = 3.0e38
    y[2, :] *= 1e-42   # denormals
    y[3, 11] = np.nan
    table = S.synth_table(4, 1200, n_file)
    mc = 5
    idx = np.arange(n_file)
    keep_o, scores_o, _ = S.oracle_scan(table, n_file, idx // 64, idx % 64, y, mc)
    ctx = kg.Context.identity(n_file)
    ctx.set_phenotypes(y, mc)
    keep, scores = ctx.scores_dense(table, 1200)
    assert np.array_equal(keep, keep_o)
    a, b = scores[:, keep], scores_o[:, keep_o]
    both_nan = np.isnan(a) & np.isnan(b)
    assert np.array_equal(_bits(a)[~both_nan], _bits(b)[~both_nan])
    ctx.close()


# ------------------------------------------------------------------------------ hits + heap replay
def _associate(ctx, table, y, min_count, kbest, tile_rows, on_device=False):
    """Reference associate loop (associate_kmers.cpp:123-148) over the C ABI: tiles -> hits -> replay
    through the oracle's BestAssociationsHeap restatement, thresholds fed back per tile."""
    n_rows, p = table.shape[0], y.shape[0]
    ctx.set_phenotypes(y, min_count)
    heaps = [S.OracleHeap(kbest) for _ in range(p)]
    thr = np.full(p, -1.0)
    dev = _torch_rows(table) if on_device else None
    stride = table.shape[1]
    r0 = 0
    kept = 0
    while r0 < n_rows:
        n = min(tile_rows, n_rows - r0)
        ctx.set_thresholds(thr)
        if on_device:
            ctx.scan_submit(dev.data_ptr() + r0 * stride * 8, n, r0)
        else:
            ctx.scan_submit(np.ascontiguousarray(table[r0:r0 + n]), n, r0)
        hits, seen, kept = ctx.scan_fetch()
        for j in range(p):
            h = hits[hits["pheno"] == j]
            heaps[j].add_many(h["kmer"], h["score"], h["row"])
            k, s, _ = heaps[j].dump()
            if len(k) >= kbest:
                thr[j] = s[0]
        r0 += n
    return heaps, kept


@pytest.mark.parametrize("name", S.GOLDEN_CASES)
@pytest.mark.parametrize("on_device", [False, True])
def test_topk_matches_reference_golden(kg, name, on_device):
    g = S.Golden(name)
    ctx = kg.Context(g.n_file, g.map_word, g.map_bit)
    heaps, kept = _associate(ctx, g.table, g.y, g.min_count, g.kbest, tile_rows=257, on_device=on_device)
    assert kept == int(g.z["cli_tested"])
    for j in range(g.n_pheno):
        k, s, _ = heaps[j].dump()
        assert np.array_equal(k, g.z["top_kmers"][j])
        assert np.array_equal(_bits(s), _bits(g.z["top_scores"][j]))
    ctx.close()


def test_hits_all_rows_when_threshold_negative(kg):
    n_file, n_rows = 241, 4000
    table = S.synth_table(31, n_rows, n_file)
    y = S.synth_phenotypes(32, n_file, 6)
    mc = S.min_count_of(n_file, 0.05, 5)
    idx = np.arange(n_file)
    keep_o, scores_o, kept_o = S.oracle_scan(table, n_file, idx // 64, idx % 64, y, mc)
    ctx = kg.Context.identity(n_file)
    ctx.set_phenotypes(y, mc)
    ctx.scan_submit(table, n_rows, first_row_id=1000)
    hits, seen, kept = ctx.scan_fetch()
    assert seen == n_rows and kept == kept_o and len(hits) == kept_o * 6
    for j in range(6):
        h = hits[hits["pheno"] == j]
        assert np.array_equal(h["row"], np.nonzero(keep_o)[0] + 1000)       # sorted by row
        assert np.array_equal(h["kmer"], table[keep_o, 0])
        assert np.array_equal(_bits(h["score"]), _bits(scores_o[j][keep_o]))
    ctx.close()


def test_hit_overflow_is_reported_and_recoverable(kg):
    n_file, n_rows = 64, 3000
    table = S.synth_table(41, n_rows, n_file)
    y = S.synth_phenotypes(42, n_file, 2)
    ctx = kg.Context.identity(n_file)
    ctx.set_option(kg.OPT_HIT_CAPACITY, 1000)
    ctx.set_phenotypes(y, 5)
    ctx.scan_submit(table, n_rows)
    with pytest.raises(kg.KgError) as ei:
        ctx.scan_fetch()
    assert ei.value.status == 4
    # resubmit in smaller pieces: state was rolled back
    total = 0
    for r0 in range(0, n_rows, 250):
        n = min(250, n_rows - r0)
        ctx.scan_submit(np.ascontiguousarray(table[r0:r0 + n]), n, r0)
        hits, seen, kept = ctx.scan_fetch()
        total += len(hits)
    idx = np.arange(n_file)
    _, _, kept_o = S.oracle_scan(table, n_file, idx // 64, idx % 64, y, 5)
    assert seen == n_rows and kept == kept_o and total == 2 * kept_o
    ctx.close()


# ------------------------------------------------------------------------------ kinship
@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("n_file", [64, 65, 129, 241, 300, 1135])
def test_kinship_matches_oracle(kg, n_file, engine):
    import math
    n_rows = 5000
    table = S.synth_table(50 + n_file, n_rows, n_file)
    idx = np.arange(n_file)
    mc = int(math.ceil(n_file * 0.05))
    K_o, cnt_o = S.oracle_kinship(table, n_file, idx // 64, idx % 64, mc)
    ctx = kg.Context.identity(n_file)
    ctx.set_option(kg.OPT_KINSHIP_ENGINE, engine)
    ctx.kinship_begin(mc)
    for r0 in range(0, n_rows, 1700):   # several tiles accumulate
        n = min(1700, n_rows - r0)
        ctx.kinship_submit(np.ascontiguousarray(table[r0:r0 + n]), n)
    K, cnt = ctx.kinship_fetch()
    assert cnt == cnt_o
    assert np.array_equal(K, K_o)
    ctx.close()


@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("name", ["identity_n131", "plumbing_n64", "ties_n96", "thaliana_n1135"])
def test_kinship_matches_reference_golden(kg, name, engine):
    import hashlib
    g = S.Golden(name)
    ctx = kg.Context.identity(g.n_file)
    ctx.set_option(kg.OPT_KINSHIP_ENGINE, engine)
    ctx.kinship_begin(int(g.z["kin_min_count"]))
    ctx.kinship_submit(g.table, g.n_rows)
    K, cnt = ctx.kinship_fetch()
    assert cnt == int(g.z["kin_cnt"])
    if "kin" in g.z:
        assert np.array_equal(K, g.z["kin"])
    else:
        assert hashlib.sha256(K.tobytes()).digest() == g.z["kin_sha256"].tobytes()
    ctx.close()


@pytest.mark.parametrize("engine", [1, 2])
def test_kinship_subset_columns(kg, engine):
    n_file, n_used = 300, 150
    rng = np.random.default_rng(2)
    names = [f"s{i}" for i in range(n_file)]
    used = [names[i] for i in rng.permutation(n_file)[:n_used]]
    mw, mb = S.column_map(names, used)
    table = S.synth_table(61, 3000, n_file)
    K_o, cnt_o = S.oracle_kinship(table, n_file, mw, mb, 8)
    ctx = kg.Context(n_file, mw, mb)
    ctx.set_option(kg.OPT_KINSHIP_ENGINE, engine)
    ctx.kinship_begin(8)
    ctx.kinship_submit(table, 3000)
    K, cnt = ctx.kinship_fetch()
    assert cnt == cnt_o and np.array_equal(K, K_o)
    ctx.close()


def test_kinship_external_accumulator(kg):
    """Caller-owned torch accumulator (what the multi-GPU path all-reduces with NCCL)."""
    import torch
    n_file = 129
    table = S.synth_table(71, 2000, n_file)
    idx = np.arange(n_file)
    K_o, cnt_o = S.oracle_kinship(table, n_file, idx // 64, idx % 64, 7)
    ctx = kg.Context.identity(n_file)
    acc = torch.ones(ctx.kinship_accum_len(), dtype=torch.int64, device="cuda")
    ctx.kinship_begin(7, acc.data_ptr())
    # two "shards" accumulated separately then summed == one pass
    ctx.kinship_submit(np.ascontiguousarray(table[:900]), 900)
    ctx.sync()
    part = acc.clone()
    ctx.kinship_begin(7, acc.data_ptr())
    ctx.kinship_submit(np.ascontiguousarray(table[900:]), 1100)
    ctx.sync()
    acc += part
    K, cnt = ctx.kinship_fetch()
    assert cnt == cnt_o and np.array_equal(K, K_o)
    ctx.close()


# ------------------------------------------------------------------------------ shapes beyond the tensor engines
def _check_session_heaps(kg, sess, table, keep_o, scores_o, kbest, phenos):
    for j in phenos:
        h = S.oracle_topk(table, keep_o, scores_o[j], kbest)
        ko, so, ro = h.dump()
        k, s, r = sess.heap(j)
        assert np.array_equal(k, ko) and np.array_equal(_bits(s), _bits(so)) and np.array_equal(r, ro)


@pytest.mark.parametrize("engine", [0, 2])
def test_wide_table_on_the_tensor_path(kg, engine):
    """BASELINE config 5 shape (4096 samples): the B tile of a pass shrinks to 16 columns and the raw ring to 2 stages so
    that the int8 filter still runs (engine 2 forced, or the auto engine once the heaps are full); the kinship operand
    stages do not fit next to 520-byte rows, so the auto kinship engine stays on the popcount Gram.  All exact."""
    n_file, n_rows, n_pheno, kbest = 4096, 3000, 20, 50
    table = S.synth_table(91, n_rows, n_file)
    y = S.synth_phenotypes(92, n_file, n_pheno)
    mc = S.min_count_of(n_file, 0.05, 5)
    idx = np.arange(n_file)
    mw, mb = idx // 64, idx % 64
    keep_o, scores_o, kept_o = S.oracle_scan(table, n_file, mw, mb, y, mc)
    sess = kg.Session(n_file, mw, mb, y, mc, kbest, scan_engine=engine)
    for r0 in range(0, n_rows, 1100):
        n = min(1100, n_rows - r0)
        sess.associate(np.ascontiguousarray(table[r0:r0 + n]), n, r0)
    assert sess.tested(0) == kept_o
    _check_session_heaps(kg, sess, table, keep_o, scores_o, kbest, range(n_pheno))
    kt_rows = sess.kernel_times()
    sess.close()
    # the filter's integer sums over two passes of <= 15 columns
    ctx = kg.Context.identity(n_file)
    ctx.set_phenotypes(y, mc)
    q, yq = ctx.filter_sums(np.ascontiguousarray(table[:300]), 300)
    bits = np.unpackbits(np.ascontiguousarray(table[:300, 1:]).view(np.uint8), axis=1, bitorder="little").astype(np.int64)
    assert np.array_equal(q.astype(np.int64), bits @ yq.astype(np.int64).T)
    # device heaps on the same shape
    ctx.set_option(kg.OPT_SCAN_ENGINE, engine)
    ctx.select_begin(kbest)
    ctx.scan_submit(table, n_rows, 0)
    assert ctx.select_sync() == (n_rows, kept_o)
    for j, (k_, s_, r_) in enumerate(ctx.select_heaps()):
        h = S.OracleHeap(kbest)
        h.add_many(k_, s_, r_)
        ko, so, ro = S.oracle_topk(table, keep_o, scores_o[j], kbest).dump()
        kd, sd, rd = h.dump()
        assert np.array_equal(kd, ko) and np.array_equal(_bits(sd), _bits(so)) and np.array_equal(rd, ro)
    ctx.close()
    ctx = kg.Context.identity(n_file)
    ctx.kinship_begin(mc)
    ctx.kinship_submit(np.ascontiguousarray(table[:200]), 200)
    K, cnt = ctx.kinship_fetch()
    K_o, cnt_o = S.oracle_kinship(table[:200], n_file, mw, mb, mc)
    assert cnt == cnt_o and np.array_equal(K, K_o)
    ctx.close()


@pytest.mark.parametrize("engine", [0, 2])
def test_many_phenotypes_in_passes(kg, engine):
    """P = 300 > 127 columns: the filter scans the tile in three passes of 100 phenotype columns (BASELINE config 5 has
    1001 phenotypes); host replay path and device heaps, exact either way."""
    n_file, n_rows, n_pheno, kbest = 241, 9000, 300, 30
    table = S.synth_table(93, n_rows, n_file)
    y = S.synth_phenotypes(94, n_file, n_pheno)
    mc = S.min_count_of(n_file, 0.05, 5)
    idx = np.arange(n_file)
    keep_o, scores_o, kept_o = S.oracle_scan(table, n_file, idx // 64, idx % 64, y, mc)
    sess = kg.Session(n_file, idx // 64, idx % 64, y, mc, kbest, scan_engine=engine)
    sess.set_option(kg.OPT_KERNEL_TIMING, 1)
    for r0 in range(0, n_rows, 2500):
        n = min(2500, n_rows - r0)
        sess.associate(np.ascontiguousarray(table[r0:r0 + n]), n, r0)
    assert sess.tested(0) == kept_o
    _check_session_heaps(kg, sess, table, keep_o, scores_o, kbest, (0, 69, 139, 200, 299))
    kt = sess.kernel_times()
    assert kt["scan_filter"][1] > 0 and kt["scan_filter"][1] % 3 == 0      # three filter passes per filtered tile
    sess.close()
    ctx = kg.Context.identity(n_file)
    ctx.set_option(kg.OPT_SCAN_ENGINE, engine)
    ctx.set_phenotypes(y, mc)
    q, yq = ctx.filter_sums(np.ascontiguousarray(table[:500]), 500)
    bits = np.unpackbits(np.ascontiguousarray(table[:500, 1:]).view(np.uint8), axis=1, bitorder="little").astype(np.int64)
    assert np.array_equal(q.astype(np.int64), bits @ yq.astype(np.int64).T)
    ctx.select_begin(kbest)
    ctx.scan_submit(table, n_rows, 0)
    assert ctx.select_sync() == (n_rows, kept_o)
    heaps = ctx.select_heaps()
    for j in (0, 99, 100, 199, 299):
        k_, s_, r_ = heaps[j]
        h = S.OracleHeap(kbest)
        h.add_many(k_, s_, r_)
        ko, so, ro = S.oracle_topk(table, keep_o, scores_o[j], kbest).dump()
        kd, sd, rd = h.dump()
        assert np.array_equal(kd, ko) and np.array_equal(_bits(sd), _bits(so)) and np.array_equal(rd, ro)
    ctx.close()


def test_ecoli_shape_through_the_pipeline(kg):
    """BASELINE config 4 shape (241 samples, 101 phenotypes): auto engine, several batches, device-resident rows."""
    n_file, n_rows, n_pheno, kbest = 241, 60000, 101, 100
    table = S.synth_table(95, n_rows, n_file)
    y = S.synth_phenotypes(96, n_file, n_pheno)
    mc = S.min_count_of(n_file, 0.05, 5)
    idx = np.arange(n_file)
    keep_o, scores_o, kept_o = S.oracle_scan(table, n_file, idx // 64, idx % 64, y, mc)
    dev = _torch_rows(table)
    sess = kg.Session(n_file, idx // 64, idx % 64, y, mc, kbest)
    stride = table.shape[1]
    for r0 in range(0, n_rows, 17000):
        n = min(17000, n_rows - r0)
        sess.associate(dev.data_ptr() + r0 * stride * 8, n, r0)
    assert sess.tested(0) == kept_o
    st = sess.stats()
    assert st["rows_scored"] == n_rows
    for j in (0, 50, 100):
        h = S.oracle_topk(table, keep_o, scores_o[j], kbest)
        ko, so, ro = h.dump()
        k, s, r = sess.heap(j)
        assert np.array_equal(k, ko) and np.array_equal(_bits(s), _bits(so)) and np.array_equal(r, ro)
    sess.close()


def test_mac_filter_matches_numpy(kg):
    """kg_mac_filter = load_kmers' MAC filter alone (reference :117-121), subset of the columns in shuffled order."""
    n_file, n_rows = 300, 5000
    rng = np.random.default_rng(11)
    used = rng.permutation(n_file)[:211]
    table = S.synth_table(21, n_rows, n_file)
    ctx = kg.Context(n_file, (used // 64).astype(np.uint32), (used % 64).astype(np.uint32))
    for mc in (0, 1, 11, 105, 106):
        keep, kept = ctx.mac_filter(table, n_rows, mc)
        bits = np.unpackbits(table[:, 1:].copy().view(np.uint8), axis=1, bitorder="little")[:, used]
        cnt = bits.sum(axis=1)
        want = (cnt >= mc) & (cnt <= len(used) - mc)
        assert np.array_equal(keep, want) and kept == int(want.sum())
    ctx.close()


# ------------------------------------------------------------------------------ distinct patterns on the device
def _pattern_hashes(table, n_file, used, min_count):
    """reference hash of every kept row (kmers_multiple_databases.cpp:367-374; Hash64 = kmer_general.h:32-41), numpy"""
    bits = np.unpackbits(np.ascontiguousarray(table[:, 1:]).view(np.uint8), axis=1, bitorder="little")[:, used]
    cnt = bits.sum(axis=1)
    keep = (cnt >= min_count) & (cnt <= len(used) - min_count)
    w_mem = S.w_mem_of(len(used))
    padded = np.zeros((table.shape[0], 64 * w_mem), dtype=np.uint8)
    padded[:, :len(used)] = bits
    words = np.packbits(padded, axis=1, bitorder="little").view(np.uint64)
    seed = np.zeros(table.shape[0], dtype=np.uint64)
    with np.errstate(over="ignore"):
        for w in range(w_mem):
            k = words[:, w].copy()
            k = (k ^ (k >> np.uint64(33))) * np.uint64(0xff51afd7ed558ccd)
            k = (k ^ (k >> np.uint64(33))) * np.uint64(0xc4ceb9fe1a85ec53)
            k = k ^ (k >> np.uint64(33))
            seed ^= k + np.uint64(0x9e3779b97f4a7c15) + (seed << np.uint64(6)) + (seed >> np.uint64(2))
    return seed[keep], int(keep.sum())


@pytest.mark.parametrize("subset", [False, True])
def test_pattern_counter_on_device(kg, subset):
    n_file, n_rows, mc = 300, 40000, 11
    rng = np.random.default_rng(5)
    used = rng.permutation(n_file)[:211] if subset else np.arange(n_file)
    table = S.synth_table(23, n_rows, n_file)
    table[:, 1:] = table[np.arange(n_rows) % 9001, 1:].copy()           # many repeated patterns
    want, kept = _pattern_hashes(table, n_file, used, mc)
    ctx = kg.Context(n_file, (used // 64).astype(np.uint32), (used % 64).astype(np.uint32))
    ctx.patterns_begin(0)                                                # grow on demand: several rehashes
    for r0 in range(0, n_rows, 7000):
        n = min(7000, n_rows - r0)
        ctx.patterns_submit(np.ascontiguousarray(table[r0:r0 + n]), n, mc)
    distinct, dkept = ctx.patterns_count()
    assert dkept == kept and distinct == len(np.unique(want))
    keys = ctx.patterns_export()
    assert np.array_equal(np.sort(keys), np.unique(want))
    # union with another shard's keys; attached to a scan (the scan's own device copy of the rows is hashed)
    other = kg.Context(n_file, (used // 64).astype(np.uint32), (used % 64).astype(np.uint32))
    other.patterns_begin(0)
    extra = S.synth_table(29, 5000, n_file)
    y = S.synth_phenotypes(7, len(used), 3)
    other.set_phenotypes(y, mc)
    other.select_begin(20)
    other.patterns_attach(mc, 5000)
    other.scan_submit(extra, 5000, n_rows)
    other.select_sync()
    want2, kept2 = _pattern_hashes(extra, n_file, used, mc)
    assert other.patterns_count() == (len(np.unique(want2)), kept2)
    ctx.patterns_insert(other.patterns_export())
    assert ctx.patterns_count()[0] == len(np.unique(np.concatenate([want, want2])))
    ctx.close()
    other.close()


def test_bind_host_to_device_reports_a_node_or_nothing(gpu_device):
    """kg_bind_host_to_device: -1 (nothing changed: single-node box, no sysfs entry) or the NUMA node of the GPU's PCIe
    slot with at least one CPU left to run on; the library keeps working afterwards either way."""
    import os
    import kmersgwas_b200 as kg
    before = os.sched_getaffinity(0)
    node, cpus = kg._abi.bind_host_to_device(0)
    after = os.sched_getaffinity(0)
    try:
        if node < 0:
            assert cpus == 0 and after == before
        else:
            assert 1 <= cpus == len(after) and after <= before
        c = kg.Context.identity(64)
        c.close()
    finally:
        os.sched_setaffinity(0, before)
