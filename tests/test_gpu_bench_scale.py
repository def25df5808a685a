"""Parity at the tile size bench.py times: one full 2^23-row tile (N = 1135, P = 101, K = 10001) at warm thresholds
through scan engine 2 (int8 tensor filter + exact re-scoring: pair lists at capacity, group lists, kernel-argument
constants, all at bench scale) and through scan engine 1 (the exact kernel on every row) must leave identical device
heaps; and the shard protocol of the multi-GPU job must reproduce one sequential scan at the same scale."""
import numpy as np
import pytest

import support as S

pytestmark = pytest.mark.gpu

N, P, K = 1135, 101, 10001
TILE = 1 << 23


@pytest.fixture(scope="module")
def kg(gpu_device):
    import kmersgwas_b200 as kg
    return kg


def _ctx(kg, y, mc, engine, flags=0):
    c = kg.Context.identity(N)
    c.set_option(kg.OPT_SCAN_ENGINE, engine)
    c.set_phenotypes(y, mc)
    c.select_begin(K, flags)
    return c


def test_full_tile_filter_engine_equals_exact_engine_at_warm_thresholds(kg):
    import torch
    y = np.ascontiguousarray(np.random.default_rng(4242).standard_normal((P, N)).astype(np.float32))
    mc = S.min_count_of(N, 0.05, 5)
    stride = (N + 63) // 64 + 1
    warm = _ctx(kg, y, mc, 0)
    buf = torch.empty(TILE * stride, dtype=torch.int64, device="cuda")
    for t in range(3):                                    # 2.5e7 rows: heaps full, thresholds warm
        warm.synth_rows_device(77, t * TILE, TILE, buf.data_ptr())
        warm.scan_submit(buf.data_ptr(), TILE, t * TILE)
    applied, kept = warm.select_sync()
    assert applied == 3 * TILE and 0.85 * applied < kept < applied
    assert np.all(warm.select_thresholds() > 0)
    state = torch.empty(warm.select_state_len(), dtype=torch.int64, device="cuda")
    warm.select_export(dev_ptr=state.data_ptr())
    warm.synth_rows_device(77, 3 * TILE, TILE, buf.data_ptr())          # the tile under test: rows the heaps have not seen
    res = []
    for engine in (2, 1):
        c = _ctx(kg, y, mc, engine)
        c.set_option(kg.OPT_KERNEL_TIMING, 1)
        c.select_import(state.data_ptr(), applied, kept)
        assert c.select_digest() == warm.select_digest()
        c.scan_submit(buf.data_ptr(), TILE, 3 * TILE)
        a2, k2 = c.select_sync()
        kt = c.kernel_times()
        res.append((c.select_digest(), a2, k2, c.select_stats()["admitted"]))
        if engine == 2:
            assert kt["scan_filter"][2] == TILE and kt["scan_exact"][2] == 0      # one 2^23-row filter launch, no dense fallback
        else:
            assert kt["scan_exact"][2] == TILE and kt["scan_filter"][1] == 0
        c.close()
    warm.close()
    assert res[0] == res[1]
    assert res[0][1] == 4 * TILE and res[0][0] != 0


@pytest.mark.parametrize("n,p", [(241, 101), (64, 101), (128, 30)])
def test_narrow_table_role_splits_equal_exact_engine_at_scale(kg, n, p):
    """The narrow-table role splits of the filter kernel (<8,3> for W <= 4 presence words, <4,4> for W <= 2) on a full
    2^22-row tile at warm thresholds: same device heaps as the exact engine."""
    import torch
    y = np.ascontiguousarray(np.random.default_rng(4300 + n).standard_normal((p, n)).astype(np.float32))
    mc = S.min_count_of(n, 0.05, 5)
    stride = (n + 63) // 64 + 1
    tile = 1 << 22

    def ctx(engine):
        c = kg.Context.identity(n)
        c.set_option(kg.OPT_SCAN_ENGINE, engine)
        c.set_phenotypes(y, mc)
        c.select_begin(K, 0)
        return c

    warm = ctx(0)
    buf = torch.empty(tile * stride, dtype=torch.int64, device="cuda")
    for t in range(3):
        warm.synth_rows_device(79, t * tile, tile, buf.data_ptr())
        warm.scan_submit(buf.data_ptr(), tile, t * tile)
    applied, kept = warm.select_sync()
    assert applied == 3 * tile
    state = torch.empty(warm.select_state_len(), dtype=torch.int64, device="cuda")
    warm.select_export(dev_ptr=state.data_ptr())
    warm.synth_rows_device(79, 3 * tile, tile, buf.data_ptr())
    res = []
    for engine in (2, 1):
        c = ctx(engine)
        c.set_option(kg.OPT_KERNEL_TIMING, 1)
        c.select_import(state.data_ptr(), applied, kept)
        c.scan_submit(buf.data_ptr(), tile, 3 * tile)
        a2, k2 = c.select_sync()
        kt = c.kernel_times()
        if engine == 2:
            assert kt["scan_filter"][2] == tile and kt["scan_exact"][2] == 0
        res.append((c.select_digest(), a2, k2, c.select_stats()["admitted"]))
        c.close()
    warm.close()
    assert res[0] == res[1]
    assert res[0][1] == 4 * tile and res[0][0] != 0


def test_shard_protocol_equals_sequential_scan_at_scale(kg):
    import torch
    y = np.ascontiguousarray(np.random.default_rng(4243).standard_normal((P, N)).astype(np.float32))
    mc = S.min_count_of(N, 0.05, 5)
    stride = (N + 63) // 64 + 1
    rows = 1 << 22
    tiles = []
    seq = _ctx(kg, y, mc, 0)
    for r in range(3):
        b = torch.empty(rows * stride, dtype=torch.int64, device="cuda")
        seq.synth_rows_device(78, r * rows, rows, b.data_ptr())
        tiles.append(b)
        seq.scan_submit(b.data_ptr(), rows, r * rows)
    want = (seq.select_sync(), seq.select_digest())
    seq.close()
    s0 = _ctx(kg, y, mc, 0)
    s0.scan_submit(tiles[0].data_ptr(), rows, 0)
    for r in (1, 2):
        sr = _ctx(kg, y, mc, 0, kg.SELECT_LOG)
        sr.scan_submit(tiles[0].data_ptr(), rows // 8, 0)       # shared prefix
        sr.select_log_reset()
        sr.scan_submit(tiles[r].data_ptr(), rows, r * rows)
        _, kept_r = sr.select_sync()
        off, ent = sr.select_log()
        assert off[-1] > 0
        s0.select_replay(ent, off, rows, kept_r)
        sr.close()
    got = (s0.select_sync(), s0.select_digest())
    s0.close()
    assert got == want
