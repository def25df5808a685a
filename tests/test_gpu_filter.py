"""GPU parity tests of scan engine 2: the int8 tensor-core (tcgen05 / TMEM) filter + exact refine.
The filter's integer sums must equal the integer dot products bit for bit, and the hits it lets through
must be exactly the hits of the exact engine (same rows, same k-mers, same IEEE score bits)."""
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


def _unpack_bits(table, n_file):
    """[n_rows, 64*W] 0/1 matrix in file column order."""
    words = np.ascontiguousarray(table[:, 1:])
    b = np.unpackbits(words.view(np.uint8), axis=1, bitorder="little")
    return b.astype(np.int64)


@pytest.mark.parametrize("n_file,n_pheno,n_rows", [(64, 1, 300), (65, 2, 129), (131, 3, 1000), (241, 8, 5000),
                                                   (1135, 101, 3001), (1135, 17, 128 * 148 * 2 + 5)])
def test_filter_sums_equal_integer_dot_products(kg, n_file, n_pheno, n_rows):
    table = S.synth_table(100 + n_file, n_rows, n_file)
    y = S.synth_phenotypes(200 + n_file, n_file, n_pheno)
    ctx = kg.Context.identity(n_file)
    ctx.set_phenotypes(y, S.min_count_of(n_file, 0.05, 5))
    q, yq = ctx.filter_sums(table, n_rows)
    assert np.abs(yq).max() <= 127 and np.abs(yq).max() > 100      # symmetric int8, full range used
    bits = _unpack_bits(table, n_file)
    want = bits @ yq.astype(np.int64).T
    assert np.array_equal(q.astype(np.int64), want)
    ctx.close()


def test_filter_sums_subset_columns_and_device_rows(kg):
    import torch
    n_file, n_used, n_pheno, n_rows = 300, 211, 5, 2000
    rng = np.random.default_rng(3)
    names = [f"s{i}" for i in range(n_file)]
    used = [names[i] for i in rng.permutation(n_file)[:n_used]]
    mw, mb = S.column_map(names, used)
    table = S.synth_table(7, n_rows, n_file)
    y = S.synth_phenotypes(8, n_used, n_pheno)
    ctx = kg.Context(n_file, mw, mb)
    ctx.set_phenotypes(y, 11)
    dev = torch.from_numpy(table.view(np.int64).copy()).cuda()
    q, yq = ctx.filter_sums(dev.data_ptr(), n_rows)
    cols = mw.astype(np.int64) * 64 + mb
    unused = np.setdiff1d(np.arange(yq.shape[1]), cols)
    assert not yq[:, unused].any()                                   # unused file columns carry weight 0
    want = _unpack_bits(table, n_file) @ yq.astype(np.int64).T
    assert np.array_equal(q.astype(np.int64), want)
    # odd (8-byte but not 16-byte aligned) device pointer: one row further in
    q2, _ = ctx.filter_sums(dev.data_ptr() + table.shape[1] * 8, n_rows - 1)
    assert np.array_equal(q2, q[1:])
    ctx.close()


# pair_limit: -1 = default (every list of these small tiles goes through the per-column re-test and pair mode),
# 0 = list mode only (16 phenotypes per listed row), 1500 = mixed: the cold column's group is "dense" (every kept row
# of a 7001-row tile is listed for it) and stays in list mode, the other groups go through pair mode
@pytest.mark.parametrize("pair_limit", [-1, 0, 1500])
@pytest.mark.parametrize("n_file,n_pheno", [(241, 8), (1135, 101), (96, 3)])
def test_filter_engine_hits_identical_to_exact_engine(kg, n_file, n_pheno, pair_limit):
    n_rows = 20000
    table = S.synth_table(300 + n_file, n_rows, n_file)
    y = S.synth_phenotypes(400 + n_file, n_file, n_pheno)
    mc = S.min_count_of(n_file, 0.05, 5)
    idx = np.arange(n_file)
    keep_o, scores_o, kept_o = S.oracle_scan(table, n_file, idx // 64, idx % 64, y, mc)
    # thresholds: the 99.5 % quantile of each phenotype's kept scores (a warm heap), one cold column
    thr = np.array([np.quantile(scores_o[j][keep_o], 0.995) for j in range(n_pheno)])
    thr[n_pheno // 2] = -1.0
    res = []
    for engine in (1, 2):
        ctx = kg.Context.identity(n_file)
        ctx.set_option(kg.OPT_SCAN_ENGINE, engine)
        ctx.set_option(kg.OPT_FILTER_PAIR_LIMIT, pair_limit)
        ctx.set_phenotypes(y, mc)
        ctx.set_thresholds(thr)
        for r0 in range(0, n_rows, 7001):     # several ragged tiles between two fetches
            n = min(7001, n_rows - r0)
            ctx.scan_submit(np.ascontiguousarray(table[r0:r0 + n]), n, r0)
        hits, seen, kept = ctx.scan_fetch()
        assert seen == n_rows and kept == kept_o
        res.append(hits)
        ctx.close()
    a, b = res
    assert len(a) == len(b) and len(a) > 0
    for f in ("row", "kmer", "pheno"):
        assert np.array_equal(a[f], b[f])
    assert np.array_equal(_bits(a["score"]), _bits(b["score"]))
    # and both equal the oracle's strict '>' selection
    for j in range(n_pheno):
        sel = keep_o & ((scores_o[j] > thr[j]) | (thr[j] < 0))
        assert np.array_equal(a["row"][a["pheno"] == j], np.nonzero(sel)[0])


def test_filter_engine_degenerate_phenotypes(kg):
    """constant, non-finite and huge phenotype columns cannot be bounded: every kept row is re-scored exactly."""
    n_file, n_rows = 130, 3000
    y = S.synth_phenotypes(5, n_file, 5)
    y[0, :] = 2.5
    y[1, 3] = np.inf
    y[2, 7] = 1e35
    y[3, :] *= 1e-30
    table = S.synth_table(6, n_rows, n_file)
    idx = np.arange(n_file)
    keep_o, scores_o, kept_o = S.oracle_scan(table, n_file, idx // 64, idx % 64, y, 6)
    thr = np.array([0.0, 1.0, 1.0, 0.0, np.quantile(scores_o[4][keep_o], 0.9)])
    out = []
    for engine in (1, 2):
        ctx = kg.Context.identity(n_file)
        ctx.set_option(kg.OPT_SCAN_ENGINE, engine)
        ctx.set_phenotypes(y, 6)
        ctx.set_thresholds(thr)
        ctx.scan_submit(table, n_rows)
        hits, _, kept = ctx.scan_fetch()
        assert kept == kept_o
        out.append(hits)
        ctx.close()
    a, b = out
    assert np.array_equal(a["row"], b["row"]) and np.array_equal(a["pheno"], b["pheno"])
    sa, sb = _bits(a["score"]), _bits(b["score"])
    nan = np.isnan(a["score"]) & np.isnan(b["score"])
    assert np.array_equal(sa[~nan], sb[~nan])


@pytest.mark.parametrize("pair_limit", [-1, 0, 40])
@pytest.mark.parametrize("name", ["subset_n300", "ties_n96", "thaliana_n1135"])
def test_topk_golden_through_filter_engine(kg, name, pair_limit):
    """Reference top-K (golden fixtures) through the product's host driver with the filter engine forced on; pair mode,
    list mode and their mix (subset_n300: shuffled subset of the columns, i.e. squeezed copies of the listed rows)."""
    g = S.Golden(name)
    sess = kg.Session(g.n_file, g.map_word, g.map_bit, g.y, g.min_count, g.kbest, scan_engine=2)
    sess.set_option(kg.OPT_FILTER_PAIR_LIMIT, pair_limit)
    for r0 in range(0, g.n_rows, 1500):
        n = min(1500, g.n_rows - r0)
        sess.associate(np.ascontiguousarray(g.table[r0:r0 + n]), n, r0)
    assert sess.tested(0) == int(g.z["cli_tested"])
    for j in range(g.n_pheno):
        k, s, _ = sess.heap(j)
        assert np.array_equal(k, g.z["top_kmers"][j])
        assert np.array_equal(_bits(s), _bits(g.z["top_scores"][j]))
    sess.close()


def test_one_context_reused_with_more_phenotype_groups(kg):
    """kg_scan_set_phenotypes may be called again on one context (MultipleKmersDataBases::ensure_phenotypes does): going
    from 5 to 101 phenotypes needs seven 16-column groups where one was enough, so the filter's list buffers must be
    re-sized for the new group count; then back to 3 phenotypes (the larger buffers are kept)."""
    n_file, n_rows = 1135, 20000
    table = S.synth_table(911, n_rows, n_file)
    mc = S.min_count_of(n_file, 0.05, 5)
    idx = np.arange(n_file)
    ctx = kg.Context.identity(n_file)
    ctx.set_option(kg.OPT_SCAN_ENGINE, 2)
    for n_pheno in (5, 101, 3):
        y = S.synth_phenotypes(920 + n_pheno, n_file, n_pheno)
        keep_o, scores_o, kept_o = S.oracle_scan(table, n_file, idx // 64, idx % 64, y, mc)
        thr = np.array([np.quantile(scores_o[j][keep_o], 0.99) for j in range(n_pheno)])
        ctx.set_phenotypes(y, mc)
        ctx.set_thresholds(thr)
        ctx.scan_submit(table, n_rows, 0)
        hits, seen, kept = ctx.scan_fetch()
        assert seen == n_rows and kept == kept_o
        for j in range(n_pheno):
            sel = keep_o & (scores_o[j] > thr[j])
            got = hits[hits["pheno"] == j]
            order = np.argsort(got["row"], kind="stable")
            assert np.array_equal(got["row"][order], np.nonzero(sel)[0])
            assert np.array_equal(_bits(got["score"][order]), _bits(scores_o[j][sel]))
    ctx.close()
