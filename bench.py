#!/usr/bin/env python
"""bench.py -- k-mers scored/sec of the association scan (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference [--gpus N] [--steps K] ...    the unmodified reference CPU binary

Workload (config.workload): BASELINE.json configs[1] -- 2.3e9 k-mers x 1135 samples, 1 phenotype + 100 permutations
(P = 101), best K = 10001, maf 0.05 / mac 5 -- on synthetic rows of the counter-based generator documented in
oracle/oracle.c.  The timed region is the WHOLE JOB from empty heaps: --job-rows rows per GPU (default 2.3e9) cut into
K steps, every step one kg_scan_submit of its batch; the device-resident BestAssociationsHeap set (kg_select_*) takes
the candidates, so the cold phase, the heap work and (N > 1) the exact merge of the shards are all inside the timing.
The table (350 GB per GPU) is larger than HBM: the steps are generated into an HBM-resident ring segment by segment
(untimed) and each segment is then scanned (timed); the job time is the sum of the segments.

  value  : rows of the job / job time, batches resident in HBM (device pointers handed to the C ABI)
  e2e    : rows/s with the batches in pinned HOST memory; H2D of the rows and the D2H read of the result inside the timing
  N > 1  : weak scaling, every rank scans its own --job-rows k-mer block; no data-path collective in the scan; ranks > 0
           warm-start from a shared prefix and log what their heaps admit; the logs are all-gathered over NCCL and
           replayed on rank 0 through the same device heaps (merge_ms, inside the timed region)
  parity : bench-scale checks (untimed): tensor-filter engine == exact engine on a full 2^23-row tile at the job's final
           thresholds, and the shard protocol == one sequential scan (heap digests)
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

N_SAMPLES, N_PHENO, K_BEST, MAF, MAC = 1135, 101, 10001, 0.05, 5
SEED_TABLE, SEED_PHENO = 20260117, 4242


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def min_count_of(n, maf, mac):
    return max(int(math.ceil(float(n) * maf)), mac)   # associate_kmers.cpp:99-102


def phenotypes(n, p):
    rng = np.random.default_rng(SEED_PHENO)
    return np.ascontiguousarray(rng.standard_normal((p, n)).astype(np.float32))


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            d = json.loads(f.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, device_index: int):
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            import torch
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.nv = pynvml
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.nv = None
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ reference CPU arm
def _ref_paths():
    ref = ROOT / "oracle" / "_ref"
    return ref / "associate_kmers", ref / "emma_kinship_kmers"


def _write_ref_inputs(d: Path, n_rows: int, n: int, p: int):
    sys.path.insert(0, str(ROOT / "tests"))
    import support as S   # oracle-side helpers: synthetic generator + file writers (allowed in this leg only)
    table = S.synth_table(SEED_TABLE, n_rows, n)
    names = [f"s{i}" for i in range(n)]
    S.write_table(d / "t", table, n, names)
    S.write_pheno(d / "p.tsv", names, phenotypes(n, p))
    return d / "t", d / "p.tsv"


def _parse_minutes(stderr: str, key: str):
    tot = 0.0
    for line in stderr.splitlines():
        at = line.find(key + " [")   # "Associations [i]" follows the progress dots on the same line
        if at >= 0:
            tot += float(line[at:].split("\t")[1].replace("min", "")) * 60.0
    return tot


def run_reference_scan(n_rows: int, threads: int, workdir: Path, n=N_SAMPLES, p=N_PHENO):
    """One run of the UNMODIFIED reference associate_kmers (oracle/_ref) on n_rows synthetic rows.
    Returns (pass-1 seconds = its own Load + Associations timers, wall seconds)."""
    exe, _ = _ref_paths()
    base, pheno = _write_ref_inputs(workdir, n_rows, n, p)
    out = workdir / "out"
    out.mkdir(exist_ok=True)
    t0 = time.perf_counter()
    r = subprocess.run([str(exe), "-p", str(pheno), "-b", "ref", "-o", str(out), "--kmers_table", str(base),
                        "-n", str(K_BEST), "--parallel", str(threads), "--kmer_len", "31", "--maf", str(MAF),
                        "--mac", str(MAC), "--batch_size", str(max(n_rows, 1))],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("reference associate_kmers failed: " + r.stderr[-500:])
    load_s = _parse_minutes(r.stderr, "Load")
    assoc_s = _parse_minutes(r.stderr, "Associations")
    return load_s + assoc_s, wall, load_s, assoc_s


def reference_arm(args):
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return 0
    exe, _ = _ref_paths()
    if not exe.exists():
        emit({"impl": "reference", "unavailable": "oracle/_ref/associate_kmers not built"})
        return 0
    threads = os.cpu_count() or 1
    n_rows = args.ref_rows
    times = []
    with tempfile.TemporaryDirectory(prefix="kgref_") as td:
        td = Path(td)
        for i in range(args.warmup_ref + args.steps):
            t_pass1, wall, load_s, assoc_s = run_reference_scan(n_rows, threads, td)
            if i >= args.warmup_ref:
                times.append((t_pass1, wall, load_s, assoc_s))
    t_total = sum(t[0] for t in times)
    value = n_rows * len(times) / t_total
    sample = (f"{n_rows} rows x {N_SAMPLES} samples x {N_PHENO} phenotypes per step through oracle/_ref/associate_kmers "
              f"--parallel {threads}; time = its own pass-1 stderr timers (Load {np.mean([t[2] for t in times]):.3f} s 1 thread + "
              f"Associations {np.mean([t[3] for t in times]):.3f} s); whole binary incl. pass 2: {np.mean([t[1] for t in times]):.2f} s wall")
    line = {
        "impl": "reference", "metric": "k-mers scored/sec", "value": value, "unit": "k-mers/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup_ref, "ms_per_step": 1e3 * t_total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 lanes + f64 epilogue (SSE4.1)",
        "data": "synthetic",
        "config": {"workload": f"configs[1] shape: {N_SAMPLES} samples x {N_PHENO} phenotypes, K={K_BEST}; bounded sample of "
                               f"{n_rows} rows per step (reference CPU path; cost is linear in rows)"},
        "cpu_baseline": {"value": value, "unit": "k-mers/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------ our arm
def our_arm(args):
    import torch
    import torch.distributed as dist

    import kmersgwas_b200 as kg

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    # this rank's threads and the pinned row buffers it allocates below live on its GPU's NUMA node
    numa_node, numa_cpus = (-1, 0) if args.no_numa_bind else kg._abi.bind_host_to_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n, p = args.samples, args.phenos
    w_file = (n + 63) // 64
    stride = w_file + 1
    row_bytes = 8 * stride
    W, K = args.warmup, args.steps
    J = args.job_rows                                   # rows per GPU (the whole config-2 table at N = 1)
    R = (J + K - 1) // K                                # rows per step
    y = phenotypes(n, p)
    mc = min_count_of(n, MAF, MAC)
    row0 = rank * J                                     # first global row of this rank's k-mer block

    stream = torch.cuda.Stream()
    ctx = kg.Context.identity(n, device=local, stream=stream.cuda_stream)
    ctx.set_option(kg.OPT_SCAN_ENGINE, args.scan_engine)
    ctx.set_option(kg.OPT_KERNEL_TIMING, 1)
    if args.max_round:
        ctx.set_option(kg.OPT_SELECT_MAX_ROUND, args.max_round)
    if args.growth_permille:
        ctx.set_option(kg.OPT_SELECT_GROWTH_PERMILLE, args.growth_permille)
    ctx.set_phenotypes(y, mc)
    fshape = ctx.filter_shape()
    int8_peak = ctx.probe_int8_peak()

    # ---- batch ring: as many steps resident in HBM as fit; the job runs in segments of that many steps, each segment
    # generated (untimed) and then scanned (timed); the heaps carry over, so the timed regions add up to the whole job
    free_b, _tot = torch.cuda.mem_get_info()
    seg_steps = max(1, min(K, int(free_b * args.hbm_fraction) // (R * row_bytes)))
    bufs = [torch.empty(R * stride, dtype=torch.int64, device="cuda") for _ in range(seg_steps)]

    def step_rows(s):
        return min(R, J - s * R)

    def fill(buf, first_global_row, rows):
        ctx.synth_rows_device(SEED_TABLE, first_global_row, rows, buf.data_ptr())

    sampler = ClockSampler(local)
    with torch.cuda.stream(stream):
        # ---- warm-up: W untimed steps of the same loop on other rows (allocations, first launches), then fresh heaps
        ctx.select_begin(args.kbest)
        wrows = min(R, 1 << 24)
        for s in range(W):
            fill(bufs[0], (1 << 40) + s * wrows, wrows)
            ctx.scan_submit(bufs[0].data_ptr(), wrows, (1 << 40) + s * wrows)
        ctx.select_sync()
        flags = kg.SELECT_LOG if (world > 1 and rank > 0) else 0
        ctx.select_begin(args.kbest, flags)
        merge_cap = args.merge_log_cap * p
        if world > 1:
            # merge buffers (entries of 3 x u64) and a warm-up of the collectives the merge uses: the first large
            # all-gather on a communicator pays NCCL's lazy channel set-up, which is not part of the job
            merge_mine = torch.zeros(merge_cap * 3, dtype=torch.int64, device="cuda")
            merge_all = torch.empty(world * merge_cap * 3, dtype=torch.int64, device="cuda")
            for _ in range(2):
                dist.all_gather_into_tensor(merge_all, merge_mine)
                dist.all_gather_into_tensor(torch.zeros(world * p, dtype=torch.int64, device="cuda"), torch.zeros(p, dtype=torch.int64, device="cuda"))
            torch.cuda.synchronize()
        ctx.kernel_times_reset()
        launches0 = ctx.launches
        step_ms, seg_ms = [], []
        prefix_ms = 0.0
        timed_ms = 0.0
        barrier()
        with sampler:
            # ---- ranks > 0 warm-start from the shared prefix: the first --prefix-rows rows of the table (rank 0's own
            # first rows).  Their thresholds are then never above the sequential heap's, and their logs stay short.
            if world > 1 and rank > 0 and args.prefix_rows > 0:
                pre = min(args.prefix_rows, R)
                fill(bufs[0], 0, pre)
                stream.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ctx.scan_submit(bufs[0].data_ptr(), pre, 0)
                ctx.select_log_reset()
                e1.record(stream)
                stream.synchronize()
                prefix_ms = e0.elapsed_time(e1)
            s = 0
            while s < K:
                seg = list(range(s, min(K, s + seg_steps)))
                for i, st in enumerate(seg):
                    fill(bufs[i], row0 + st * R, step_rows(st))
                torch.cuda.synchronize()
                barrier()
                evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(seg) + 1)]
                evs[0].record(stream)
                for i, st in enumerate(seg):
                    ctx.scan_submit(bufs[i].data_ptr(), step_rows(st), row0 + st * R)
                    evs[i + 1].record(stream)
                if seg[-1] == K - 1:
                    applied, kept = ctx.select_sync()      # the heaps are final before the clock stops
                torch.cuda.synchronize()
                seg_ms.append(evs[0].elapsed_time(evs[-1]))
                step_ms += [evs[i].elapsed_time(evs[i + 1]) for i in range(len(seg))]
                s += len(seg)
            timed_ms = sum(seg_ms) + prefix_ms
            kt = ctx.kernel_times()
            launches_timed = ctx.launches - launches0

            # ---- N > 1: exact merge, inside the job: the logs of ranks 1 .. N-1 (what their heaps admitted, in row order)
            # are gathered on every rank and replayed in rank order through rank 0's heaps, which are exact for its
            # own block -> the sequential reference heaps of the whole N x J-row table (kg_select_replay)
            merge_ms, log_entries, merge_phases = None, None, None
            if world > 1:
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                tph = [time.perf_counter()]
                e0.record(stream)
                counts = torch.zeros(p, dtype=torch.int64, device="cuda")
                if rank > 0:
                    counts = torch.from_numpy(ctx.select_log_counts().astype(np.int64)).cuda()
                all_counts = torch.zeros(world * p, dtype=torch.int64, device="cuda")
                dist.all_gather_into_tensor(all_counts, counts)
                all_counts = all_counts.cpu().numpy().reshape(world, p)
                totals = all_counts.sum(axis=1)
                mx = int(totals.max())
                tph.append(time.perf_counter())
                assert mx <= merge_cap, f"shard log of {mx} entries exceeds the merge buffer ({merge_cap}): raise --merge-log-cap"
                if rank > 0 and totals[rank] > 0:
                    ctx.select_log(dev_ptr=merge_mine.data_ptr())
                kept_t = torch.tensor([kept, applied], dtype=torch.int64, device="cuda")
                kept_all = torch.zeros(2 * world, dtype=torch.int64, device="cuda")
                stream.synchronize()
                tph.append(time.perf_counter())
                # one all-gather of the (padded) logs: 24 B per entry over NVLink
                n_send = max(mx, 1) * 3
                dist.all_gather_into_tensor(merge_all[: world * n_send], merge_mine[:n_send])
                dist.all_gather_into_tensor(kept_all, kept_t)
                torch.cuda.synchronize()
                tph.append(time.perf_counter())
                kept_all = kept_all.cpu().numpy().reshape(world, 2)
                if rank == 0:
                    for r in range(1, world):
                        off = np.zeros(p + 1, dtype=np.uint64)
                        off[1:] = np.cumsum(all_counts[r]).astype(np.uint64)
                        ctx.select_replay(merge_all.data_ptr() + r * n_send * 8, off, J, int(kept_all[r, 0]))
                    applied, kept = ctx.select_sync()
                e1.record(stream)
                torch.cuda.synchronize()
                tph.append(time.perf_counter())
                merge_ms = max_over_ranks(e0.elapsed_time(e1))
                merge_phases = {k_: round(1e3 * (tph[i + 1] - tph[i]), 3) for i, k_ in enumerate(("counts", "pack", "all_gather", "replay"))}
                log_entries = int(totals.sum())
                timed_ms += merge_ms
            barrier()
        job_ms = max_over_ranks(timed_ms)
        sel_stats = ctx.select_stats()
        job_digest = ctx.select_digest() if rank == 0 else 0
        thr_end = ctx.select_thresholds()
        status_rows = (applied, kept)

        # ---- parity at bench scale (untimed): (1) one full 2^23-row tile at the job's final thresholds through the tensor
        # filter engine and through the exact engine -> equal heap digests; (2) the N-shard protocol against one sequential scan
        parity = parity_block(args, kg, torch, dist, ctx, rank, world, local, n, p, y, mc, stride, stream)

        # ---- e2e: K steps of 2^23 rows from pinned HOST memory through the same C-ABI calls (H2D inside the timing, one
        # D2H read of the step's result = rows applied / kept), continuing on the job's heaps
        Re = min(args.e2e_rows, R)
        n_e2e = min(K, args.e2e_buffers)
        host = [torch.empty(Re * stride, dtype=torch.int64).pin_memory() for _ in range(n_e2e)]
        e2e_row0 = (world + rank) * J + (1 << 36)
        for i in range(n_e2e):
            fill(bufs[0], e2e_row0 + i * Re, Re)
            stream.synchronize()
            host[i].copy_(bufs[0][: Re * stride])
        for i in range(min(2, n_e2e)):   # untimed: first touch of the pinned buffers by the DMA engine
            ctx.scan_submit(host[i].data_ptr(), Re, e2e_row0 + i * Re)
        ctx.select_sync()
        torch.cuda.synchronize()
        ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        sampler2 = ClockSampler(local)
        sampler2.samples, sampler2.reasons = sampler.samples, sampler.reasons
        t_host0 = time.perf_counter()
        with sampler2:
            ev2.record(stream)
            for i in range(K):
                ctx.scan_submit(host[i % n_e2e].data_ptr(), Re, e2e_row0 + (2 + i) * Re)
                ctx.select_sync()
            ev3.record(stream)
            torch.cuda.synchronize()
        t_host1 = time.perf_counter()
        barrier()
        ms_e2e = max_over_ranks(max(ev2.elapsed_time(ev3), 1e3 * (t_host1 - t_host0)))

    # ---- kinship leg (config 3 shape, bounded rows): GB/s of table consumed
    kin = kinship_leg(args, kg, torch, dist, rank, world, local, n, stride, row_bytes)

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        rows_total = J * world
        value = rows_total / (job_ms * 1e-3)
        e2e_value = Re * K * world / (ms_e2e * 1e-3)
        dom = max(("scan_exact", "scan_filter", "scan_refine", "scan_select"), key=lambda k_: kt[k_][0])
        dom_ms, dom_launches, dom_rows = kt[dom]
        hbm_achieved = (dom_rows * row_bytes) / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        traffic = None
        try:
            tr = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
            if dom in tr and (n, p) == (N_SAMPLES, N_PHENO):
                traffic = tr[dom]["dram_bytes_per_row"] * dom_rows / max(dom_launches, 1)
        except Exception:
            pass
        k_pad = fshape["k_pad"] or 128 * ((64 * w_file + 127) // 128)
        p_pad = fshape["p_pad"] or 16 * ((p + 1 + 15) // 16)              # accumulator columns of ONE pass
        n_pass = max(fshape["n_pass"], 1)
        f_ms, f_launches, f_rows = kt["scan_filter"]                       # rows summed over launches: every pass counts its rows
        tensor_ops = 2.0 * k_pad * p_pad * f_rows                          # int8 multiply-adds x 2 the filter issues
        tensor_tops = tensor_ops / (f_ms * 1e-3) / 1e12 if f_ms > 0 else 0.0
        t_hbm_row = row_bytes / (hbm_peak * 1e9)
        t_tensor_row = 2.0 * k_pad * p_pad * n_pass / (int8_peak * 1e12) if int8_peak > 0 else 0.0
        tensor_binds = dom == "scan_filter" and t_tensor_row > t_hbm_row
        steady = step_ms[len(step_ms) // 2:]
        roofline = {
            "bound": "tensor_int8" if tensor_binds else "hbm", "kernel": dom,
            "achieved": tensor_tops if tensor_binds else hbm_achieved,
            "peak": int8_peak if tensor_binds else hbm_peak,
            "unit": "TOP/s" if tensor_binds else "GB/s",
            "frac": (tensor_tops / int8_peak if int8_peak > 0 else None) if tensor_binds else hbm_achieved / hbm_peak,
            "traffic": traffic,
            "peak_source": ("measured live: kg_probe_int8_peak (back-to-back tcgen05.mma kind::i8 M128 N256 K32 on all SMs, CUDA events)"
                            if tensor_binds else peak_src),
            "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak, "peak_source": peak_src,
                    "algorithmic_bytes_per_row": row_bytes},
            "tensor": {"achieved_int8_tops": tensor_tops, "measured_int8_peak_tops": int8_peak, "nominal_dense_int8_tops": 4500.0,
                       "frac_of_measured": tensor_tops / int8_peak if int8_peak > 0 else None,
                       "mma_shape_per_128_rows": f"M=128 N={p_pad} K={k_pad} x {n_pass} pass(es)",
                       "per_row_bounds_ps": {"hbm": t_hbm_row * 1e12, "tensor_int8": t_tensor_row * 1e12}},
            "launches": dom_launches, "avg_launch_ms": dom_ms / max(dom_launches, 1),
            "kernel_ms_share_of_job": {k_: v[0] / job_ms for k_, v in kt.items() if v[1]},
            "kernels": {k_: {"ms_total": v[0], "launches": v[1], "rows": v[2]} for k_, v in kt.items() if v[1]},
            "note": (f"achieved = algorithmic work of the dominant kernel's launches / their CUDA-event time on the launching stream; HBM: "
                     f"{row_bytes} B/row; tensor: 2 x K_pad x P_pad int8 ops per row and pass.  At P={p} the int8 tensor pipe binds the scan "
                     f"(SURVEY 7 hard part 2), so frac is against the live-measured int8 peak; the hbm block gives the GB/s view"),
        }
        line = {
            "metric": "k-mers scored/sec", "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": job_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 lane sums in reference order + f64 epilogue (int8 tensor filter in front)",
            "data": "synthetic",
            "config": {
                "workload": (f"BASELINE configs[1]: {J} rows x {n} samples x {p} phenotypes (1 + {p - 1} permutations) per GPU from EMPTY heaps, "
                             f"best K={args.kbest}, maf {MAF}/mac {MAC}; whole job = {K} steps of {R} rows ({R * row_bytes / 1e9:.1f} GB each, > 126 MB L2), "
                             f"cold phase, device heaps and (N>1) shard merge inside the timed region"),
                "segments": (f"table {J * row_bytes / 1e9:.0f} GB per GPU > HBM: {len(seg_ms)} segments of <= {seg_steps} HBM-resident steps; the timed "
                             f"region is the sum of the segments (generation between them untimed), heaps carry over"),
                "rows_per_gpu": J, "rows_per_step_per_gpu": R, "row_bytes": row_bytes,
                "l2": "inputs larger than L2, every step reads a different batch",
                "scan_engine": args.scan_engine,
                "parallelism": (f"k-mer-block shards x{world} (weak: {J} rows per GPU), no data-path collective in the scan; ranks > 0 warm-start from a "
                                f"{args.prefix_rows}-row shared prefix; logs all-gathered and replayed on rank 0 (exact merge)"),
                "selection": "device-resident BestAssociationsHeap set (kg_select_*): no host in the scan loop",
                "host_placement": {"numa_node": numa_node, "cpus": numa_cpus, "how": "kg_bind_host_to_device (rank 0 shown; -1 = not bound)"},
            },
            "timed_region_s": job_ms * 1e-3,
            "job": {"rows": rows_total, "seconds": job_ms * 1e-3, "rows_applied_rank0": status_rows[0], "rows_kept_rank0": status_rows[1],
                    "segment_ms": [round(v, 3) for v in seg_ms], "step_ms": [round(v, 3) for v in step_ms], "prefix_ms": prefix_ms,
                    "merge_ms": merge_ms, "merge_phases_ms_rank0": merge_phases, "log_entries": log_entries, "selection": sel_stats, "heap_digest": f"{job_digest:016x}",
                    "threshold_phenotype0": float(thr_end[0])},
            "steady_state": {"value": R * len(steady) * world / (sum(steady) * 1e-3) if steady else None, "unit": "k-mers/s",
                             "note": "secondary: the second half of the job's steps only (rank 0's device time)"},
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": "k-mers/s", "h2d_bytes_per_step": Re * row_bytes, "d2h_bytes_per_step": 128,
                    "ms_per_step": ms_e2e / K, "rows_per_step_per_gpu": Re,
                    "path": "pinned host rows -> kg_scan_submit (C ABI, H2D sub-tiles under the kernels) -> device heaps -> kg_select_sync per step"},
            "gpu_launches": launches_timed,
            "clocks": sampler.summary(),
            "parity": parity,
            "kinship": kin,
        }
        if merge_ms is not None:
            line["merge_ms"] = merge_ms
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def parity_block(args, kg, torch, dist, ctx, rank, world, local, n, p, y, mc, stride, stream):
    """Bench-scale parity (untimed, rank 0): (1) one full 2^23-row tile at the job's final thresholds, tensor filter
    engine vs exact engine -> identical heap digests; (2) the shard protocol (prefix + log + replay) vs one sequential scan."""
    out = {"engine_full_tile": None, "shard_merge": None}
    if rank != 0 or args.no_parity:
        return out
    T = args.parity_tile_rows
    buf = torch.empty(T * stride, dtype=torch.int64, device="cuda")
    first = (1 << 38) + 12345
    ctx.synth_rows_device(SEED_TABLE, first, T, buf.data_ptr())
    state = torch.empty(ctx.select_state_len(), dtype=torch.int64, device="cuda")
    ctx.select_export(dev_ptr=state.data_ptr())
    applied, kept = ctx.select_sync()
    digs, kepts = [], []
    for engine in (2, 1):
        c2 = kg.Context.identity(n, device=local, stream=stream.cuda_stream)
        c2.set_option(kg.OPT_SCAN_ENGINE, engine)
        c2.set_phenotypes(y, mc)
        c2.select_begin(args.kbest)
        c2.select_import(state.data_ptr(), applied, kept)
        c2.scan_submit(buf.data_ptr(), T, first)
        a2, k2 = c2.select_sync()
        digs.append(c2.select_digest())
        kepts.append((a2, k2))
        c2.close()
    out["engine_full_tile"] = {"ok": digs[0] == digs[1] and kepts[0] == kepts[1], "rows": T, "digest_filter": f"{digs[0]:016x}",
                               "digest_exact": f"{digs[1]:016x}", "at_rows_scanned": applied}
    # (2) two shards on this GPU with the bench's multi-GPU protocol vs the sequential scan of the same rows
    S_rows, pre = args.parity_shard_rows, args.parity_shard_rows // 8
    shards = max(2, world)
    seq = kg.Context.identity(n, device=local, stream=stream.cuda_stream)
    seq.set_phenotypes(y, mc)
    seq.select_begin(args.kbest)
    tiles = []
    for r in range(shards):
        b = torch.empty(S_rows * stride, dtype=torch.int64, device="cuda")
        seq.synth_rows_device(SEED_TABLE + 7, r * S_rows, S_rows, b.data_ptr())
        tiles.append(b)
        seq.scan_submit(b.data_ptr(), S_rows, r * S_rows)
    a_seq, k_seq = seq.select_sync()
    d_seq = seq.select_digest()
    seq.close()
    s0 = kg.Context.identity(n, device=local, stream=stream.cuda_stream)
    s0.set_phenotypes(y, mc)
    s0.select_begin(args.kbest)
    s0.scan_submit(tiles[0].data_ptr(), S_rows, 0)
    for r in range(1, shards):
        sr = kg.Context.identity(n, device=local, stream=stream.cuda_stream)
        sr.set_phenotypes(y, mc)
        sr.select_begin(args.kbest, kg.SELECT_LOG)
        sr.scan_submit(tiles[0].data_ptr(), pre, 0)
        sr.select_log_reset()
        sr.scan_submit(tiles[r].data_ptr(), S_rows, r * S_rows)
        a_r, k_r = sr.select_sync()
        off, ent = sr.select_log()
        s0.select_replay(ent, off, S_rows, k_r)
        sr.close()
    a_m, k_m = s0.select_sync()
    d_m = s0.select_digest()
    s0.close()
    out["shard_merge"] = {"ok": d_m == d_seq and (a_m, k_m) == (a_seq, k_seq), "shards": shards, "rows_per_shard": S_rows,
                          "digest_merged": f"{d_m:016x}", "digest_sequential": f"{d_seq:016x}"}
    return out


def kinship_leg(args, kg, torch, dist, rank, world, local, n, stride, row_bytes):
    """emma_kinship_kmers Gram pass on --kinship-rows rows per GPU per step, resident in HBM; N > 1 adds the path's one
    exchange step, the NCCL all-reduce of the u64 accumulator, issued by the LIBRARY (kg_kinship_allreduce over its own
    communicator) and timed separately after a warm-up all-reduce."""
    Rk = args.kinship_rows
    if Rk <= 0:
        return None
    from kmersgwas_b200 import _abi
    stream = torch.cuda.Stream()
    ctx = kg.Context.identity(n, device=local, stream=stream.cuda_stream)
    ctx.set_option(kg.OPT_KERNEL_TIMING, 1)
    ctx.set_option(kg.OPT_KINSHIP_ENGINE, args.kinship_engine)
    mc = int(math.ceil(n * MAF))
    if world > 1:
        ids = [_abi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init_rank(ids[0], world, rank)
    steps, warm = max(1, min(args.steps, 5)), 2
    with torch.cuda.stream(stream):
        bufs = []
        for s in range(2):
            b = torch.empty(Rk * stride, dtype=torch.int64, device="cuda")
            ctx.synth_rows_device(SEED_TABLE + 1, (s * world + rank) * Rk, Rk, b.data_ptr())
            bufs.append(b)
        ctx.kinship_begin(mc)
        for s in range(warm):
            ctx.kinship_submit(bufs[s % 2].data_ptr(), Rk)
        if world > 1:
            ctx.kinship_allreduce()           # warm-up: the first collective on a communicator pays its lazy set-up
        ctx.sync()
        ctx.kinship_begin(mc)
        ctx.kernel_times_reset()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0.record(stream)
        for s in range(steps):
            ctx.kinship_submit(bufs[s % 2].data_ptr(), Rk)
        ev1.record(stream)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        kt = ctx.kernel_times()
        ar_ms = None
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize()
            e0.record(stream)
            ctx.kinship_allreduce()
            e1.record(stream)
            torch.cuda.synchronize()
            ar_ms = e0.elapsed_time(e1)
            t = torch.tensor([ms, ar_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, ar_ms = float(t[0].item()), float(t[1].item())
        _, kept = ctx.kinship_fetch(want_matrix=False)
    int8_peak = ctx.probe_int8_peak()
    ctx.close()
    hbm_peak, _ = measured_peaks()
    rows = Rk * steps * world
    gbs = rows * row_bytes / (ms * 1e-3) / 1e9
    k_ms, k_launches, k_rows = kt["kinship"]
    k_pad = 128 * ((64 * ((n + 63) // 64) + 127) // 128)
    # tensor work the engine issues per row: lower-triangle tiles of 128 x 256 samples, 2 x 128 x 256 int8 ops each
    n_i = k_pad // 128
    tiles = sum(i // 2 + 1 for i in range(n_i))
    ops_row = 2.0 * 128 * 256 * tiles
    tops = ops_row * k_rows / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    out = {"metric": "kinship Gram table GB/s", "rows_per_s": rows / (ms * 1e-3), "gb_per_s": gbs,
           "frac_of_hbm_peak": gbs / hbm_peak / world, "rows_per_step_per_gpu": Rk, "steps": steps,
           "ms_per_step": ms / steps, "kernel_ms": k_ms, "aux_ms": kt["aux"][0],
           "int_ops_per_row": 2 * n * n, "allreduce_ms": ar_ms, "allreduce_bytes": 8 * (n * n + 1) if world > 1 else None,
           "allreduce_path": "kg_kinship_allreduce (library-owned NCCL communicator, u64 sum)" if world > 1 else None,
           "engine": args.kinship_engine, "rows_kept": kept,
           "roofline": {"bound": "tensor_int8", "achieved": tops, "peak": int8_peak, "unit": "TOP/s",
                        "frac": tops / int8_peak if int8_peak > 0 else None, "issued_int8_ops_per_row": ops_row,
                        "peak_source": "measured live: kg_probe_int8_peak"}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = kinship_cpu_baseline(args)
    return out


def kinship_cpu_baseline(args):
    """The unmodified reference emma_kinship_kmers (oracle/_ref, single-threaded by construction) on a bounded sample."""
    _, exe = _ref_paths()
    if not exe.exists():
        return {"value": None, "unit": "rows/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref/emma_kinship_kmers missing"}
    n_rows = args.kinship_cpu_rows
    with tempfile.TemporaryDirectory(prefix="kgkin_") as td:
        base, _ = _write_ref_inputs(Path(td), n_rows, args.samples, 1)
        t0 = time.perf_counter()
        r = subprocess.run([str(exe), "-t", str(base), "-k", "31", "--maf", str(MAF)], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        wall = time.perf_counter() - t0
    if r.returncode != 0:
        return {"value": None, "unit": "rows/s", "cores": 1, "kind": "reference", "sample": "emma_kinship_kmers failed: " + r.stderr[-200:]}
    return {"value": n_rows / wall, "unit": "rows/s", "cores": 1, "kind": "reference",
            "sample": f"{n_rows} rows x {args.samples} samples through oracle/_ref/emma_kinship_kmers --maf {MAF}: {wall:.2f} s wall (load + accumulate + print)"}


def cpu_baseline(args):
    """The unmodified reference binary (oracle/_ref) on this box's host cores, bounded sample."""
    exe, _ = _ref_paths()
    threads = os.cpu_count() or 1
    if not exe.exists():
        return {"value": None, "unit": "k-mers/s", "cores": threads, "kind": "reference",
                "sample": "oracle/_ref/associate_kmers missing (built only where /root/reference exists)"}
    with tempfile.TemporaryDirectory(prefix="kgcpu_") as td:
        t_pass1, wall, load_s, assoc_s = run_reference_scan(args.cpu_rows, threads, Path(td), args.samples, args.phenos)
    return {"value": args.cpu_rows / t_pass1, "unit": "k-mers/s", "cores": threads, "kind": "reference",
            "sample": (f"{args.cpu_rows} rows x {args.samples} samples x {args.phenos} phenotypes, oracle/_ref/associate_kmers --parallel "
                       f"{threads} -n {K_BEST}; pass-1 time from its own stderr timers: Load {load_s:.2f} s (1 thread) + "
                       f"Associations {assoc_s:.2f} s; whole binary incl. pass 2 and outputs {wall:.2f} s wall"),
            "load_rows_per_s": args.cpu_rows / load_s if load_s > 0 else None,
            "associate_rows_per_s": args.cpu_rows / assoc_s if assoc_s > 0 else None}


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL banner, torchrun children) may write to fd 1; the driver wants exactly ONE JSON line there.
    Route fd 1 to stderr for the run and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--job-rows", type=int, default=2_300_000_000, help="rows per GPU of the timed job (config 2: 2.3e9)")
    ap.add_argument("--hbm-fraction", type=float, default=0.72, help="share of the free HBM the resident batch ring may use")
    ap.add_argument("--max-round", type=int, default=0, help="longest selection round in rows (0 = library default)")
    ap.add_argument("--growth-permille", type=int, default=0, help="selection round growth (0 = library default)")
    ap.add_argument("--merge-log-cap", type=int, default=1 << 17, help="N > 1: log entries per phenotype the merge buffers hold")
    ap.add_argument("--prefix-rows", type=int, default=1 << 23, help="N > 1: rows of the shared prefix ranks > 0 warm-start from")
    ap.add_argument("--e2e-rows", type=int, default=1 << 23, help="rows per step of the e2e (pinned host memory) leg")
    ap.add_argument("--parity-tile-rows", type=int, default=1 << 23)
    ap.add_argument("--parity-shard-rows", type=int, default=1 << 22)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--samples", type=int, default=N_SAMPLES)
    ap.add_argument("--phenos", type=int, default=N_PHENO)
    ap.add_argument("--kbest", type=int, default=K_BEST)
    ap.add_argument("--scan-engine", type=int, default=0)
    ap.add_argument("--kinship-engine", type=int, default=0)
    ap.add_argument("--kinship-rows", type=int, default=1 << 20)
    ap.add_argument("--e2e-buffers", type=int, default=3)
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--cpu-rows", type=int, default=400000, help="rows of the cpu_baseline sample")
    ap.add_argument("--kinship-cpu-rows", type=int, default=15000, help="rows of the kinship cpu_baseline sample")
    ap.add_argument("--ref-rows", type=int, default=200000, help="rows per step of --impl reference")
    ap.add_argument("--warmup-ref", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    quiet_stdout()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        args.steps = max(1, min(args.steps, 10))
        return reference_arm(args)
    return our_arm(args)


if __name__ == "__main__":
    sys.exit(main())
