#!/usr/bin/env python
"""bench.py -- k-mers scored/sec of the association scan (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference [--gpus N] [--steps K] ...    the unmodified reference CPU binary

Workload (config.workload): the shape of BASELINE.json configs[1] -- 1135 samples, 1 phenotype + 100
permutations (P = 101), best K = 10001, maf 0.05 / mac 5 -- on synthetic rows of the counter-based
generator documented in oracle/oracle.c.  The full 2.3e9-row table (350 GB) does not fit one GPU's HBM, so a
"step" is one batch of --rows-per-step rows (default 2^23 = 1.27 GB, larger than the 126 MB L2; the reference's default
--batch_size is 10^7 rows) pushed through
the product's associate loop (kmersgwas_b200/host/association_driver.cpp: device scan -> candidate hits ->
exact replay through BestAssociationsHeap); every step scans rows no earlier step saw, and the heaps carry over.

  value  : rows/s with the batches already resident in HBM (device pointers handed to the C ABI)
  e2e    : rows/s with the batches in pinned HOST memory; H2D of the rows and D2H of the hits inside the timing
  N > 1  : rows sharded by k-mer block across ranks ("weak": every rank scans --rows-per-step rows per step), no
           data-path collective; after the timed steps the per-rank hit logs are exchanged (one all-to-all by
           phenotype) and merged exactly, every rank merging its share of the phenotypes (merge_ms).
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    n, p = args.samples, args.phenos
    w_file = (n + 63) // 64
    stride = w_file + 1
    row_bytes = 8 * stride
    R = args.rows_per_step
    W, K = args.warmup, args.steps
    y = phenotypes(n, p)
    mc = min_count_of(n, MAF, MAC)
    idx = np.arange(n)
    mw, mb = (idx // 64).astype(np.uint32), (idx % 64).astype(np.uint32)

    stream = torch.cuda.Stream()
    sess = kg.Session(n, mw, mb, y, mc, args.kbest, device=local, stream=stream.cuda_stream,
                      scan_engine=args.scan_engine, log_hits=(world > 1))
    abi = kg.load()
    h = sess.ctx_handle
    sess.set_option(kg.OPT_KERNEL_TIMING, 1)

    # ---- resident batches: step s, rank r scans global rows [(s*world + r) * R, +R)
    n_e2e = min(K, args.e2e_buffers)
    total_steps = W + K + n_e2e
    free_b, _tot = torch.cuda.mem_get_info()
    resident = min(W + K, max(2, int(free_b * 0.6) // (R * row_bytes)))
    bufs = [torch.empty(R * stride, dtype=torch.int64, device="cuda") for _ in range(resident)]

    # Scan position.  The timed steps should look like the bulk of the 2.3e9-row job, not like its first 2 %
    # (where the heaps have seen few rows, thresholds are low and most of the time goes into exact re-scoring of
    # candidates): --prefill-rows rows per GPU are scanned first, untimed, through the same associate loop.
    prefill_steps = (args.prefill_rows + R - 1) // R

    def first_row(step):
        return ((prefill_steps + step) * world + rank) * R

    def fill(buf, step):
        st = abi.kg_synth_rows_device(h, SEED_TABLE, first_row(step), R, buf.data_ptr())
        assert st == 0, abi.kg_last_error(h)

    launches0 = None
    with torch.cuda.stream(stream):
        t_pre0 = time.perf_counter()
        for ps in range(prefill_steps):
            b = bufs[ps % 2]
            st = abi.kg_synth_rows_device(h, SEED_TABLE, (ps * world + rank) * R, R, b.data_ptr())
            assert st == 0, abi.kg_last_error(h)
            sess.associate(b.data_ptr(), R, (ps * world + rank) * R)   # refills of b are stream-ordered behind its scan
        sess.finish()
        prefill_s = time.perf_counter() - t_pre0
        for s in range(min(resident, W + K)):
            fill(bufs[s], s)
        stream.synchronize()
        # ---- warm-up (fills the heaps: the cold phase of the scan happens here)
        for s in range(W):
            sess.associate(bufs[s % resident].data_ptr(), R, first_row(s))
        sess.finish()
        # batches beyond the resident ring are regenerated into freed slots before the timing starts
        for s in range(resident, W + K):
            fill(bufs[s % resident], s)
        stream.synchronize()
        sess.kernel_times_reset()
        launches0 = sess.launches()
        io0 = sess.io_bytes()
        stats0 = sess.stats()
        hostms0 = sess.host_ms()
        sampler = ClockSampler(local)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        with sampler:   # (the e2e region below is sampled as well)
            ev0.record(stream)
            for s in range(W, W + K):
                sess.associate(bufs[s % resident].data_ptr(), R, first_row(s))
            sess.finish()          # the last round's hits are in the heaps before the clock stops
            ev1.record(stream)
            torch.cuda.synchronize()
        barrier()
        ms_dev = ev0.elapsed_time(ev1)
        kt = sess.kernel_times()
        launches_timed = sess.launches() - launches0
        io1 = sess.io_bytes()
        stats1 = sess.stats()
        hostms1 = sess.host_ms()

        # ---- e2e: same loop from pinned host memory (H2D of the rows + D2H of the hits inside the timing)
        host = [torch.empty(R * stride, dtype=torch.int64).pin_memory() for _ in range(n_e2e)]
        for i in range(n_e2e):
            fill(bufs[0], W + K + i)
            stream.synchronize()
            host[i].copy_(bufs[0])
        for i in range(min(2, n_e2e)):   # untimed: first touch of the pinned buffers by the DMA engine
            sess.associate(host[i].data_ptr(), R, first_row(W + K + i))
        sess.finish()
        torch.cuda.synchronize()
        io2 = sess.io_bytes()
        ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        sampler2 = ClockSampler(local)
        sampler2.samples, sampler2.reasons = sampler.samples, sampler.reasons
        t_host0 = time.perf_counter()
        with sampler2:
            ev2.record(stream)
            for i in range(K):
                sess.associate(host[i % n_e2e].data_ptr(), R, first_row(W + K + (i % n_e2e)))
            sess.finish()
            ev3.record(stream)
            torch.cuda.synchronize()
        t_host1 = time.perf_counter()
        barrier()
        ms_e2e = max(ev2.elapsed_time(ev3), 1e3 * (t_host1 - t_host0))
        io3 = sess.io_bytes()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_dev = max_over_ranks(ms_dev)
    ms_e2e = max_over_ranks(ms_e2e)

    # ---- N > 1: exact merge of the shards' hit logs (host side, once per job).  Phenotypes are independent, so rank r
    # merges the phenotypes p = r (mod N): one all-to-all of the logs, then every rank replays its phenotypes' hits
    # in global row order through fresh heaps (kgh_merge_hit_log) -- the sequential reference heap state, ties included.
    merge_ms = None
    if world > 1:
        t0 = time.perf_counter()
        log = sess.hit_log()
        kept = sess.stats()["rows_kept"]
        dest = (log["pheno"] % world).astype(np.int64)
        order = np.argsort(dest, kind="stable")
        send_counts = np.bincount(dest, minlength=world).astype(np.int64)
        send = torch.from_numpy(log[order].view(np.uint8).reshape(-1).copy()).cuda()
        counts_all = torch.zeros(world * world, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(counts_all, torch.from_numpy(send_counts).cuda())
        counts_all = counts_all.cpu().numpy().reshape(world, world)          # [src][dst]
        recv_counts = counts_all[:, rank]
        isz = kg.HIT_DTYPE.itemsize
        recv = torch.empty(int(recv_counts.sum()) * isz, dtype=torch.uint8, device="cuda")
        dist.all_to_all_single(recv, send, output_split_sizes=[int(c) * isz for c in recv_counts],
                               input_split_sizes=[int(c) * isz for c in send_counts])
        kept_all = torch.tensor([kept], dtype=torch.int64, device="cuda")
        dist.all_reduce(kept_all)
        mine = recv.cpu().numpy().view(kg.HIT_DTYPE)
        hs = kg.HeapSet(args.kbest, p)
        hs.merge(mine, int(kept_all.item()))
        my_phenos = [j for j in range(p) if j % world == rank]
        assert all(hs.tested(j) == int(kept_all.item()) for j in my_phenos)
        merge_local = 1e3 * (time.perf_counter() - t0)
        merge_ms = max_over_ranks(merge_local)

    # ---- cold start (the first rows of the job: empty heaps, thresholds -1, then low): same loop, fresh session
    cold = None
    if args.cold_steps > 0:
        cs = kg.Session(n, mw, mb, y, mc, args.kbest, device=local, stream=stream.cuda_stream, scan_engine=args.scan_engine)
        nc_ = min(args.cold_steps, len(bufs))
        with torch.cuda.stream(stream):
            for s_ in range(nc_):
                st = abi.kg_synth_rows_device(h, SEED_TABLE, (s_ * world + rank) * R, R, bufs[s_].data_ptr())
                assert st == 0, abi.kg_last_error(h)
            stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            torch.cuda.synchronize()
            e0.record(stream)
            step_ev = []
            t_h0 = time.perf_counter()
            host_s = []
            for s_ in range(nc_):
                cs.associate(bufs[s_].data_ptr(), R, (s_ * world + rank) * R)
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                step_ev.append(ev)
                host_s.append(time.perf_counter() - t_h0)
            cs.finish()
            e1.record(stream)
            torch.cuda.synchronize()
        ms_cold = max_over_ranks(e0.elapsed_time(e1))
        cold_steps_ms = [e0.elapsed_time(step_ev[0])] + [step_ev[i - 1].elapsed_time(step_ev[i]) for i in range(1, nc_)]
        cold_stats = cs.stats()
        cold_host = cs.host_ms()
        cs.close()
        cold = {"value": R * nc_ * world / (ms_cold * 1e-3), "unit": "k-mers/s", "steps": nc_, "ms_per_step": ms_cold / nc_,
                "device_ms_by_step": [round(v, 3) for v in cold_steps_ms], "host_s_at_step": [round(v, 4) for v in host_s],
                "rounds": cold_stats["rounds"], "hits_replayed": cold_stats["hits_replayed"], "host_ms": cold_host,
                "note": f"rows [0, {R * nc_ * world}) of the job from empty heaps, no warm-up: exact engine until the heaps are full, then the "
                        f"filter with low thresholds; the headline value is measured at rows >= {prefill_steps * R * world}"}

    # ---- kinship leg (config 3 shape, bounded rows): GB/s of table consumed
    kin = kinship_leg(args, kg, torch, dist, rank, world, local, n, stride, row_bytes)

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        rows_total = R * K * world
        value = rows_total / (ms_dev * 1e-3)
        e2e_value = rows_total / (ms_e2e * 1e-3)
        dom = max(("scan_exact", "scan_filter", "scan_refine"), key=lambda k_: kt[k_][0])
        dom_ms, dom_launches, dom_rows = kt[dom]
        achieved = (dom_rows * row_bytes) / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        t_fp32_row = (128 * ((n + 127) // 128)) * p / (148 * 128 * 1.965e9)
        traffic = None
        try:
            tr = json.loads((ROOT / "profiles" / "r01_ncu_traffic.json").read_text())
            if dom in tr and (n, p) == (N_SAMPLES, N_PHENO):
                traffic = tr[dom]["dram_bytes_per_row"] * dom_rows / max(dom_launches, 1)
        except Exception:
            pass
        n_pad_cols = 128 * ((64 * w_file + 127) // 128)
        p_pad = 16 * ((p + 1 + 15) // 16)
        tensor_ops = 2.0 * n_pad_cols * p_pad * kt["scan_filter"][2]          # int8 MACs x 2 issued by the filter
        tensor_tops = tensor_ops / (kt["scan_filter"][0] * 1e-3) / 1e12 if kt["scan_filter"][0] > 0 else 0.0
        line = {
            "metric": "k-mers scored/sec", "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 lane sums in reference order + f64 epilogue (int8 tensor filter when enabled)",
            "data": "synthetic",
            "config": {
                "workload": f"BASELINE configs[1] shape: {n} samples x {p} phenotypes (1 + {p - 1} permutations), best K={args.kbest}, "
                            f"maf {MAF}/mac {MAC}; step = one {R}-row batch ({R * row_bytes / 1e6:.0f} MB, > 126 MB L2) per GPU through the "
                            f"associate loop; every step scans new rows, heaps carry over (full table = {2.3e9 / R:.0f} such steps streamed)",
                "scan_position": f"timed steps cover rows [{(prefill_steps + W) * R * world}, {(prefill_steps + W + K) * R * world}) of the job: "
                                 f"{prefill_steps * R} rows per GPU were scanned untimed first ({prefill_s:.1f} s) so that heap thresholds "
                                 f"are those of the bulk of the 2.3e9-row scan; --prefill-rows 0 times the cold start instead",
                "rows_per_step_per_gpu": R, "row_bytes": row_bytes, "l2": "inputs larger than L2, each step reads a different batch",
                "scan_engine": args.scan_engine, "parallelism": f"k-mer-block shards x{world}, no data-path collective",
                "hits_replayed_per_step": (stats1["hits_replayed"] - stats0["hits_replayed"]) / K,
                "threshold_rounds_per_step": (stats1["rounds"] - stats0["rounds"]) / K,
                "filter_listed_rows_per_step": kt["scan_refine"][2] / K,
                "host_ms_per_step": {k_: (hostms1[k_] - hostms0[k_]) / K for k_ in hostms1},
                "rows_per_step_by_engine": {"exact": kt["scan_exact"][2] / K, "tensor_filter": kt["scan_filter"][2] / K},
            },
            "roofline": {
                "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "tensor": {"achieved_int8_tops": tensor_tops, "nominal_dense_int8_tops": 4500.0, "frac_of_nominal": tensor_tops / 4500.0,
                           "note": f"filter MMA shape per 128-row block: M=128, N={p_pad}, K={n_pad_cols}; at P={p} the int8 tensor pipe, not HBM, is "
                                   f"the binding roofline of the scan (SURVEY 7, hard part 2)"},
                "launches": dom_launches, "avg_launch_ms": dom_ms / max(dom_launches, 1),
                "kernel_ms_share_of_step": {k_: v[0] / ms_dev for k_, v in kt.items() if v[1]},
                "kernels": {k_: {"ms_per_launch": v[0] / v[1], "launches": v[1], "rows_per_launch": v[2] / v[1]} for k_, v in kt.items() if v[1]},
                "note": (f"achieved = {row_bytes} algorithmic B/row x rows of the launch / CUDA-event time of the launch; the exact fp32-order "
                         f"kernel (scan_exact / scan_refine) is bound by the FP32 add pipe ({t_fp32_row * 1e9:.2f} ns/row at P={p}), not by HBM"),
            },
            "e2e": {"value": e2e_value, "unit": "k-mers/s", "h2d_bytes_per_step": R * row_bytes + (io3[0] - io2[0]) // K,
                    "d2h_bytes_per_step": (io3[1] - io2[1]) // K, "ms_per_step": ms_e2e / K,
                    "path": "pinned host rows -> kgh_session_associate (C ABI kg_scan_submit/kg_scan_fetch) -> host heaps"},
            "gpu_launches": launches_timed,
            "clocks": sampler.summary(),
            "kinship": kin,
            "cold_start": cold,
        }
        if merge_ms is not None:
            line["merge_ms"] = merge_ms
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        emit(line)
    sess.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def kinship_leg(args, kg, torch, dist, rank, world, local, n, stride, row_bytes):
    """emma_kinship_kmers Gram pass on --kinship-rows rows per GPU per step, resident in HBM; N > 1 adds the
    one NCCL all-reduce of the accumulator (timed separately)."""
    Rk = args.kinship_rows
    if Rk <= 0:
        return None
    stream = torch.cuda.Stream()
    ctx = kg.Context.identity(n, device=local, stream=stream.cuda_stream)
    ctx.set_option(kg.OPT_KERNEL_TIMING, 1)
    ctx.set_option(kg.OPT_KINSHIP_ENGINE, args.kinship_engine)
    mc = int(math.ceil(n * MAF))
    acc = torch.zeros(ctx.kinship_accum_len(), dtype=torch.int64, device="cuda")
    steps, warm = max(1, min(args.steps, 5)), 2
    with torch.cuda.stream(stream):
        bufs = []
        for s in range(2):
            b = torch.empty(Rk * stride, dtype=torch.int64, device="cuda")
            ctx.synth_rows_device(SEED_TABLE + 1, (s * world + rank) * Rk, Rk, b.data_ptr())
            bufs.append(b)
        ctx.kinship_begin(mc, acc.data_ptr())
        for s in range(warm):
            ctx.kinship_submit(bufs[s % 2].data_ptr(), Rk)
        ctx.sync()
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
            stream.wait_stream(torch.cuda.current_stream())
            e0.record(stream)
            dist.all_reduce(acc)
            e1.record(stream)
            torch.cuda.synchronize()
            ar_ms = e0.elapsed_time(e1)
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        _, kept = ctx.kinship_fetch(want_matrix=False)
    ctx.close()
    hbm_peak, _ = measured_peaks()
    rows = Rk * steps * world
    gbs = rows * row_bytes / (ms * 1e-3) / 1e9
    return {"metric": "kinship Gram table GB/s", "rows_per_s": rows / (ms * 1e-3), "gb_per_s": gbs,
            "frac_of_hbm_peak": gbs / hbm_peak / world, "rows_per_step_per_gpu": Rk, "steps": steps,
            "ms_per_step": ms / steps, "kernel_ms": kt["kinship"][0], "aux_ms": kt["aux"][0],
            "int_ops_per_row": 2 * n * n, "allreduce_ms": ar_ms, "engine": args.kinship_engine}


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
    ap.add_argument("--rows-per-step", type=int, default=1 << 23)
    ap.add_argument("--samples", type=int, default=N_SAMPLES)
    ap.add_argument("--phenos", type=int, default=N_PHENO)
    ap.add_argument("--kbest", type=int, default=K_BEST)
    ap.add_argument("--scan-engine", type=int, default=0)
    ap.add_argument("--kinship-engine", type=int, default=0)
    ap.add_argument("--kinship-rows", type=int, default=1 << 20)
    ap.add_argument("--e2e-buffers", type=int, default=3)
    ap.add_argument("--prefill-rows", type=int, default=1 << 28, help="rows per GPU scanned untimed before the warm-up steps")
    ap.add_argument("--cold-steps", type=int, default=10, help="steps of the cold-start leg (0 = skip)")
    ap.add_argument("--cpu-rows", type=int, default=400000, help="rows of the cpu_baseline sample")
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
