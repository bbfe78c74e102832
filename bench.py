#!/usr/bin/env python3
"""Headline benchmark: GCUPS of the database scan hot path on BASELINE.json config[1]
(the reference's "peak benchmark": PseudoDB of 1,000,000 equal-length subjects (256 aa) vs the 20 queries of
allqueries.fasta, BLOSUM62 -11/-1, database resident, `runpeakbenchmark.sh:26-54`).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                       (CPU arm: the oracle port on the host cores)

A *step* = one pass of the hot path over the whole query batch: all 20 queries scanned against the resident database
(20 sw4_scan calls through the C ABI, each: query H2D -> profile -> score kernels -> top-k -> k results D2H).
Weak scaling: every rank holds 1,000,000 subjects (shard `rank` of a N x 1M-subject database, global ids); the only
exchange is an all_gather of k (score, id) pairs per rank per query, merged on rank 0.

Two more legs run after the timed region of the headline metric (they never change `value`):
  c4           STRONG scaling on ONE fixed UniRef50-shaped database (C4: 65,000,000 subjects / 17.0 G residues, seed 4,
               sw4_set_pseudo_database_lengths): every rank generates and uploads only its own shard (blocks of 256
               subjects, rank, rank+N, ...), queries 0/5/9/14/19, one pass, per-query GCUPS = cells / max-over-ranks
               device time; `checked` = planted copies found with their self score and a >= 3,000-subject sample of
               every rank's scores equal to the CPU oracle (the reference's multi-GPU use, rununiref50benchmark.sh).
  ref_gpu      the reference's OWN align binary (oracle/_ref/align, built for sm_100a from /root/reference), --dpx and
               default half2 kernels, on the same PseudoDB (runpeakbenchmark.sh:27,44-50), timed on the same GPU
               right after ours (rank 0, N = 1 only).

  value        GCUPS from the device-timed scan regions (CUDA events inside libsw4b200.so on its own stream, the
               reference's own "Scan time" definition, src/cudasw4.cuh:707-726), max over ranks
  e2e          GCUPS from host wall-clock around the same C-ABI calls with HOST query buffers (pinned staging, H2D,
               D2H of the results inside the timed region), max over ranks
  roofline     score kernels only, against the ALU(DPX)-pipe roofline measured with tools/ubench
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SUBJECTS = 1_000_000
SUBJECT_LEN = 256
TOP_K = 10
CONFIG_NAME = "peak benchmark: PseudoDB 1,000,000 x 256 aa per GPU vs allqueries.fasta (20 queries, 41,752 aa), BLOSUM62 gop=-11 gex=-1, top-10, DB resident"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def _traffic():
    """dram__bytes_read+write of the dominant kernel per launch, from the committed ncu capture (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


class _ReferenceCpuScan:
    """The reference's OWN scalar CPU Gotoh (src/cudasw4.cuh:2331-2392, BLOSUM62), compiled from /root/reference into
    oracle/_ref/libref_harness.so by oracle/Makefile and driven one OpenMP task per subject like its
    computeAllScoresCPU_blosum62 (src/cudasw4.cuh:767-796). Present when the harness travelled with the repository."""

    def __init__(self, path):
        import ctypes
        self.lib = ctypes.CDLL(path)
        self.fn = self.lib.ref_cpu_scan_blosum62
        self.fn.restype = ctypes.c_int
        self.last_threads = 0

    def scan(self, blosum, q, db, gop, gex, threads=0):
        import ctypes
        import numpy as np
        assert blosum == 62
        q = np.ascontiguousarray(q, dtype=np.uint8)
        chars = np.ascontiguousarray(db.chars, dtype=np.uint8)
        offsets = np.ascontiguousarray(db.offsets, dtype=np.uint64)
        lengths = np.ascontiguousarray(db.lengths, dtype=np.int32)
        out = np.empty(len(lengths), dtype=np.int32)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        self.last_threads = self.fn(p(q), ctypes.c_int(len(q)), p(chars), p(offsets), p(lengths), ctypes.c_long(len(lengths)),
                                    ctypes.c_int(gop), ctypes.c_int(gex), p(out), ctypes.c_int(threads))
        return out


def _cpu_scanner():
    """(scanner, kind): the reference's own CPU routine when oracle/_ref holds it, else the oracle port."""
    path = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")
    if os.path.exists(path) and not os.environ.get("SW4_BENCH_CPU_PORT"):
        try:
            return _ReferenceCpuScan(path), "reference"
        except (OSError, AttributeError):
            pass
    from tests import oracle_lib
    return oracle_lib.load(), "port"


def host_threads() -> int:
    """Threads for the CPU arm: every core this process may run on. (torchrun exports OMP_NUM_THREADS=1, which would
    silently turn an `omp_get_max_threads()` default into a 1-core baseline.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(seconds_budget: float = 20.0, threads: int = 0):
    """CPU Gotoh (the reference's own scalar routine if built, else the oracle port; OpenMP over subjects) on a bounded
    sample of the same workload."""
    import numpy as np
    threads = threads or host_threads()
    from cudasw4_b200 import dbformat, synth
    oracle, kind = _cpu_scanner()
    queries = [dbformat.encode(s) for _, s in synth.load_queries()]
    subj = synth.pseudo_subject(SUBJECT_LEN, 42)
    # calibrate: 256 subjects x shortest query
    db = dbformat.from_equal_length_matrix(np.broadcast_to(subj, (256, SUBJECT_LEN)).copy())
    t0 = time.perf_counter()
    oracle.scan(62, queries[0], db, -11, -1, threads=threads)
    dt = max(time.perf_counter() - t0, 1e-4)
    rate = 256 * SUBJECT_LEN * len(queries[0]) / dt
    total_q = sum(len(q) for q in queries)
    n = int(max(64, min(N_SUBJECTS, seconds_budget * rate / (SUBJECT_LEN * total_q))))
    db = dbformat.from_equal_length_matrix(np.broadcast_to(subj, (n, SUBJECT_LEN)).copy())
    t0 = time.perf_counter()
    for q in queries:
        oracle.scan(62, q, db, -11, -1, threads=threads)
    dt = time.perf_counter() - t0
    cells = float(n) * SUBJECT_LEN * total_q
    return {"value": cells / 1e9 / dt, "unit": "GCUPS", "cores": int(oracle.last_threads), "kind": kind,
            "sample": f"first {n} of {N_SUBJECTS} subjects x all 20 queries ({cells:.3g} cells, {dt:.1f} s)"}, dt, cells


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps_dt, steps_cells = [], []
    base = None
    for i in range(args.warmup + args.steps):
        base, dt, cells = cpu_baseline(seconds_budget=max(2.0, 60.0 / max(1, args.warmup + args.steps)))
        if i >= args.warmup:
            steps_dt.append(dt)
            steps_cells.append(cells)
    value = sum(steps_cells) / 1e9 / sum(steps_dt)
    base["value"] = value
    line = {"impl": "reference", "metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(steps_dt) / len(steps_dt), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": CONFIG_NAME, "note": "CPU arm: the reference ships no CPU aligner; this is its own scalar "
                       "checking routine (src/cudasw4.cuh:2331-2392, built into oracle/_ref) or, without it, the oracle port, "
                       "on all host cores, bounded sample per step (see cpu_baseline.kind)"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import cudasw4_b200 as sw
    from cudasw4_b200 import synth

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    queries = synth.load_queries()
    total_q = sum(len(q) for _, q in queries)
    eng = sw.CudaSW4(deviceIds=[local_rank], numTop=TOP_K, blosumType=62)  # raises if libsw4b200.so / the GPU is missing
    eng.setShard(rank, world)
    eng.setPseudoDatabase(N_SUBJECTS * world, SUBJECT_LEN, 42)
    t0 = time.perf_counter()
    eng.prefetchDBToGpus()
    upload_s = time.perf_counter() - t0
    info = eng.dbInfo()
    shard_residues = int(info.shard_residues)

    from cudasw4_b200.distributed import gather_topk

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_step():
        """20 scans; returns (device seconds, kernel seconds, launches, merged top-k of the last query)."""
        dev_s = ker_s = 0.0
        launches = 0
        merged = None
        for _, q in queries:
            res = eng.scan(q)
            dev_s += res.stats.seconds
            ker_s += res.stats.kernelSeconds
            launches += res.stats.kernelLaunches
            # the only exchange step: k (score, id) pairs per rank (one NCCL all_gather of 80 bytes), merged on every rank
            merged = gather_topk(res.scores, res.referenceIds, TOP_K, device=dev)
        return dev_s, ker_s, launches, merged

    for _ in range(args.warmup):
        one_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    wall0 = time.perf_counter()
    dev_s = ker_s = 0.0
    launches = 0
    merged = None
    for _ in range(args.steps):
        d, k, l, merged = one_step()
        dev_s += d
        ker_s += k
        launches += l
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()

    # max over ranks
    t = torch.tensor([dev_s, ker_s, wall], dtype=torch.float64, device=dev)
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    dev_s, ker_s, wall = [float(x) for x in t.cpu()]
    launches = int(lt.item())

    if rank == 0:
        cells_per_step = float(shard_residues) * total_q * world
        cells = cells_per_step * args.steps
        value = cells / 1e9 / dev_s
        e2e = cells / 1e9 / wall
        kernel_gcups = cells / 1e9 / ker_s
        peaks, peak_src = _peaks()
        sm_hz = (clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)) * 1e6
        # ALU(DPX)-pipe roofline: 3.5 ALU-pipe instructions per cell-pair (1 max3-relu, 2 add-max, 1/2 max3), pipe issues one
        # warp instruction per 2 clocks per scheduler (measured, tools/ubench/pipes.cu): 64 cells per 7 clocks per scheduler
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        peak_gcups = sms * 4 * 64 / 7.0 * sm_hz / 1e9 * world
        # HBM side (for the record): per scan the kernel streams the shard once, 1 byte per residue (u16 fused pair code)
        hbm_bytes = float(shard_residues) * len(queries) * args.steps * world
        line = {
            "metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s16x2 (DPX) with exact s32 re-scoring", "data": "synthetic",
            "config": {"workload": CONFIG_NAME, "subjects_per_gpu": N_SUBJECTS, "subject_length": SUBJECT_LEN,
                       "queries": len(queries), "query_residues": total_q, "cells_per_step": cells_per_step,
                       "l2_note": "each scan streams 256 MB of database per GPU (> 126 MB L2) and scans alternate 20 different "
                                  "query profiles, so no timed iteration re-reads L2-resident inputs",
                       "db_upload_s": upload_s, "top1": merged[0] if merged else None},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "GCUPS", "h2d_bytes_per_step": total_q + 441 * len(queries),
                    "d2h_bytes_per_step": len(queries) * (2 * TOP_K * 4 + 32)},
            "gpu_launches": launches,
            "roofline": {"bound": "int_dpx_pipe", "achieved": kernel_gcups, "peak": peak_gcups, "unit": "GCUPS",
                         "frac": kernel_gcups / peak_gcups,
                         "peak_model": f"{sms} SMs x 4 schedulers x 64 cells / 7 clk (3.5 DPX ALU-pipe ops per s16x2 cell-pair at "
                                       f"1 warp-inst / 2 clk, measured) x {sm_hz/1e6:.0f} MHz (nvidia-smi under load)",
                         "traffic": (_traffic() or {}).get("dram_bytes_read_plus_write_per_launch"),
                         "traffic_detail": _traffic(),
                         "hbm": {"achieved": hbm_bytes / 1e9 / ker_s, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                 "frac": hbm_bytes / 1e9 / ker_s / peaks.get("hbm_gbs", 6650.0), "peak_source": peak_src,
                                 "note": "database streaming only; the path is compute bound (1/len_query bytes per cell)"}},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"], _, _ = cpu_baseline()
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
