"""world_size = 2 (gloo, CPU) tests of the N > 1 host logic: k-mer-block sharding, per-shard heaps with their own
(lower) thresholds, gather of the shard hit logs, exact merge on rank 0; and the kinship accumulator all-reduce.
The device is replaced by the oracle's scores here -- what is under test is the sharding / merge / collective
plumbing that bench.py and the CLI use, not the kernels (those are covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import support as S

N_FILE, N_ROWS, N_PHENO, KBEST, WORLD = 96, 6000, 3, 25, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    table = S.synth_table(11, N_ROWS, N_FILE)
    table[:, 1:] = table[np.arange(N_ROWS) % 700, 1:].copy()   # many identical patterns -> ties at the heap boundary
    y = S.synth_phenotypes(12, N_FILE, N_PHENO)
    idx = np.arange(N_FILE)
    return table, y, (idx // 64).astype(np.uint32), (idx % 64).astype(np.uint32), S.min_count_of(N_FILE, 0.05, 5)


def _worker(rank, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    import kmersgwas_b200 as kg
    table, y, mw, mb, mc = _inputs()
    lo, hi = N_ROWS * rank // WORLD, N_ROWS * (rank + 1) // WORLD
    keep, scores, kept = S.oracle_scan(table[lo:hi], N_FILE, mw, mb, y, mc)

    # ---- scan shard: local heaps supply (lower) thresholds; every admitted candidate is logged
    local = kg.HeapSet(KBEST, N_PHENO)
    log = []
    for j in range(N_PHENO):
        thr = -1.0
        for r in np.nonzero(keep)[0]:
            s = scores[j, r]
            if thr < 0 or s > thr:
                local.add(j, int(table[lo + r, 0]), float(s), lo + int(r))
                log.append((lo + int(r), int(table[lo + r, 0]), float(s), j, 0))
                k, sc, _ = local.heap(j)
                thr = sc[0] if len(k) >= KBEST else -1.0
    log = np.array(log, dtype=kg.HIT_DTYPE)
    sizes = [None] * WORLD
    dist.all_gather_object(sizes, (len(log), kept))
    cap = max(s[0] for s in sizes)
    mine = torch.zeros(max(cap, 1) * kg.HIT_DTYPE.itemsize, dtype=torch.uint8)
    mine[: len(log) * kg.HIT_DTYPE.itemsize] = torch.from_numpy(log.view(np.uint8).copy())
    gathered = [torch.empty_like(mine) for _ in range(WORLD)] if rank == 0 else None
    dist.gather(mine, gathered, dst=0)

    # ---- kinship shard: Gram counts are plain sums over rows -> one all-reduce
    bits = np.unpackbits(np.ascontiguousarray(table[lo:hi, 1:]).view(np.uint8), axis=1, bitorder="little")[:, :N_FILE]
    bits = bits[keep].astype(np.int64)
    acc = torch.zeros(N_FILE * N_FILE + 1, dtype=torch.int64)
    acc[:-1] = torch.from_numpy((bits.T @ bits).reshape(-1))
    acc[-1] = int(keep.sum())
    dist.all_reduce(acc)

    if rank == 0:
        parts = [g[: sizes[i][0] * kg.HIT_DTYPE.itemsize].numpy().view(kg.HIT_DTYPE) for i, g in enumerate(gathered)]
        merged = kg.HeapSet(KBEST, N_PHENO)
        merged.merge(np.concatenate(parts), sum(s[1] for s in sizes))
        res = {"tested": merged.tested(0), "acc": acc.numpy()}
        for j in range(N_PHENO):
            k, s, r = merged.heap(j)
            res[f"k{j}"], res[f"s{j}"], res[f"r{j}"] = k, s, r
        np.savez(out_path, **res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_merge_to_the_single_process_result(tmp_path):
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(_free_port(), out), nprocs=WORLD, join=True)
    z = np.load(out)
    table, y, mw, mb, mc = _inputs()
    keep, scores, kept = S.oracle_scan(table, N_FILE, mw, mb, y, mc)
    assert int(z["tested"]) == kept
    for j in range(N_PHENO):
        h = S.oracle_topk(table, keep, scores[j], KBEST)      # the sequential reference heap over ALL rows
        k, s, r = h.dump()
        assert np.array_equal(z[f"k{j}"], k)
        assert np.array_equal(z[f"s{j}"].view(np.uint64), s.view(np.uint64))
        assert np.array_equal(z[f"r{j}"], r)
    # kinship: all-reduced Gram -> IBS counts of the reference
    K_o, cnt_o = S.oracle_kinship(table, N_FILE, mw, mb, mc)
    acc = z["acc"]
    G, M = acc[:-1].reshape(N_FILE, N_FILE), int(acc[-1])
    assert M == cnt_o
    c = np.diag(G)
    ibs = np.tril(M - c[:, None] - c[None, :] + 2 * G, -1).astype(np.uint64)
    assert np.array_equal(ibs, K_o)
