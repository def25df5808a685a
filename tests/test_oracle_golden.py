"""CPU oracle vs the committed golden fixtures (reference outputs, tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest

import support as S


@pytest.mark.parametrize("name", S.GOLDEN_CASES)
def test_oracle_scores_match_reference(name):
    g = S.Golden(name)
    keep, scores, kept = S.oracle_scan(g.table, g.n_file, g.map_word, g.map_bit, g.y, g.min_count)
    assert kept == int(g.z["tested"]) == int(g.z["cli_tested"])
    for j in range(g.n_pheno):
        assert np.array_equal(g.table[keep, 0], g.z["ref_kmers"][j])
        assert np.array_equal(scores[j][keep].view(np.uint64), g.z["ref_scores"][j].view(np.uint64))


@pytest.mark.parametrize("name", S.GOLDEN_CASES)
def test_oracle_heap_matches_reference_cli(name):
    g = S.Golden(name)
    keep, scores, _ = S.oracle_scan(g.table, g.n_file, g.map_word, g.map_bit, g.y, g.min_count)
    for j in range(g.n_pheno):
        k, s, _ = S.oracle_topk(g.table, keep, scores[j], g.kbest).dump()
        assert np.array_equal(k, g.z["top_kmers"][j])
        assert np.array_equal(s.view(np.uint64), g.z["top_scores"][j].view(np.uint64))


@pytest.mark.parametrize("name", S.GOLDEN_CASES)
def test_oracle_kinship_matches_reference(name):
    g = S.Golden(name)
    names = g.names
    mw, mb = S.column_map(names, names)
    K, cnt = S.oracle_kinship(g.table, g.n_file, mw, mb, int(g.z["kin_min_count"]))
    assert cnt == int(g.z["kin_cnt"])
    if "kin" in g.z:
        assert np.array_equal(K, g.z["kin"])
    else:
        assert hashlib.sha256(K.tobytes()).digest() == g.z["kin_sha256"].tobytes()
        assert np.array_equal(K[:64, :64], g.z["kin_corner"])


def test_synth_generator_properties():
    t = S.synth_table(1, 20000, 131)
    assert np.all(np.diff(t[:, 0].astype(np.int64)) > 0)          # k-mers strictly increase
    assert np.all((t[:, -1] >> np.uint64(131 % 64)) == 0)          # padding bits are zero
    a = S.synth_table(1, 100, 131, first_row=500)
    assert np.array_equal(a, t[500:600])                           # counter-based: any window
    dup = np.all(t[1:, 1:] == t[:-1, 1:], axis=1).mean()
    assert 0.005 < dup < 0.04                                      # ~1.6 % exact duplicate patterns
