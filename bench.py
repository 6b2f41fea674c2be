#!/usr/bin/env python3
"""bench.py — GB/s scanned by the B200 buffer-scan path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2]

One "step" = one pass of the hot path over one batch.  Headline workload (every N): config c2,
`ugrep -c -F -f words.txt` (the 1,000-literal alternation) over the DENSE 4 GiB-per-GPU corpus of
SURVEY.md 8(d)-2.  N > 1 is launched by torch.distributed.run, one rank per GPU; the corpus shards by
line-aligned piece, NCCL only all-gathers the per-shard {matches, newlines}.

value     whole-job GB/s with the corpus resident in HBM (CUDA events, max over ranks)
e2e       the same through the C ABI with a HOST buffer: H2D + scan + D2H of the result inside the timed region
          (pinned memory; `e2e.pageable` repeats it from ordinary pageable memory)
roofline  the dominant kernel: algorithmic bytes (1 per corpus byte) / its own event-timed duration / measured HBM peak
others    every other BASELINE config, timed the same way at its named size, each checked against the ORACLE on one
          block (and by count additivity over the line-aligned tiling at full size).  c5 is the sharded config: the
          logical corpus of N x 8 GiB is cut with sharding.tiled_cuts (forward to the next newline), rank r scans shard
          r, and the all-gather of the counts sits INSIDE its timed region.
--impl reference times the reference's own CPU implementation (oracle/_ref/ugrep, built unmodified from
/root/reference) with all host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GIB = 1 << 30

# config -> (pattern file, corpus generator name, scan mode, reference CLI args, GiB per GPU the config names)
CONFIGS = {
    "c1": ("c1", "c1", "lines", ["-c", "-F", "Sherlock Holmes"], 1),
    "c2": ("c2", "c2", "lines", ["-c", "-F", "-f", "@WORDS@"], 4),
    "c2s": ("c2", "c2s", "lines", ["-c", "-F", "-f", "@WORDS@"], 4),
    "c3": ("c3", "c3", "list", ["-n", "-b", "-o", "[A-Z][a-z]+ing\\b"], 4),
    "c3b": ("c3b", "c3", "list", ["-n", "-b", "-o", "[A-Z][a-z]+ing"], 4),
    "c4": ("c4", "c4", "lines", ["-i", "-c", "\\p{Greek}+|naïve\\w*"], 8),
    "c5": ("c5", "c5", "matches", ["-c", "-o", "-e", "ERROR|WARN", "-e", "\\d{3}-\\d{4}"], 8),
}
WORKLOAD_TEXT = {
    "c1": "ugrep -c -F 'Sherlock Holmes' over synthetic ASCII text",
    "c2": "ugrep -c -F -f words.txt (1,000-literal alternation) over the dense corpus (words drawn from the 1,000 needles and 20,000 like-made distractors: ~79% of lines match, ~23% of positions pass the prefilter)",
    "c2s": "ugrep -c -F -f words.txt over the sparse corpus (<3% matching lines; distractors never contain a needle)",
    "c3": "ugrep -n -b -o '[A-Z][a-z]+ing\\b' (config 3 as named: the reference's prefilter admits no candidate, empty output) over synthetic text",
    "c3b": "ugrep -n -b -o '[A-Z][a-z]+ing' (match records) over synthetic text",
    "c4": "ugrep -i -c '\\p{Greek}+|naïve\\w*' over mixed UTF-8",
    "c5": "ugrep -c -o -e 'ERROR|WARN' -e '\\d{3}-\\d{4}' over a synthetic log corpus",
}
PAT_DIR = os.path.join(ROOT, "ugrep_b200", "patterns")
REF_UGREP = os.path.join(ROOT, "oracle", "_ref", "ugrep")
BLOCK_BYTES = 64 << 20
HEADLINE = "c2"
OTHERS_ORDER = ["c1", "c2s", "c3", "c3b", "c4", "c5"]


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.samples = []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.samples.append([x.strip() for x in r.stdout.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def capture_for(cfg: str, kernel: str, nbytes: int, kernel_ms: float):
    """The committed `ncu --set full` capture of this config's dominant kernel (profiles/traffic.json), or None.  A
    capture is only used when it was taken on this kernel, this workload size, and a launch whose duration under ncu
    is consistent with today's event-timed one: ncu serialises and runs cold, so it may be slower, but a capture that
    is more than 5 % FASTER than the live kernel, or more than 35 % slower, describes another kernel build."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(cfg)
        if not t or t["kernel"] != kernel or t["nbytes"] != nbytes:
            return None
        ratio = t["gpu_time_ms"] / kernel_ms
        if ratio < 0.95 or ratio > 1.35:
            return None
        return t
    except Exception:
        return None


def make_block(cfg: str):
    from ugrep_b200 import corpus
    return corpus.block(CONFIGS[cfg][1], BLOCK_BYTES)


# ---------------------------------------------------------------- reference arm (CPU)

def reference_run(cfg: str, sample_bytes: int, steps: int, warmup: int):
    """Time `ugrep -J<cores>` over the sample split into line-aligned files (ugrep parallelises per file)."""
    import numpy as np
    if not os.access(REF_UGREP, os.X_OK):
        return None
    cores = os.cpu_count() or 1
    block = make_block(cfg)
    reps = max(1, sample_bytes // block.size)
    shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    d = tempfile.mkdtemp(prefix="ugx_ref_", dir=shm)
    try:
        nfiles = min(cores, 64)
        # cut the tiled sample into nfiles line-aligned pieces: each piece = some whole blocks + a line-aligned slice
        total = block.size * reps
        nl = np.flatnonzero(block == 10)
        files = []
        pos = 0
        for i in range(nfiles):
            target = total * (i + 1) // nfiles
            b, off = divmod(target, block.size)
            if i == nfiles - 1:
                endpos = total
            else:
                j = np.searchsorted(nl, off)
                endpos = b * block.size + (int(nl[j]) + 1 if j < len(nl) else block.size)
            path = os.path.join(d, "shard_%03d.txt" % i)
            with open(path, "wb") as f:
                p = pos
                while p < endpos:
                    o = p % block.size
                    take = min(block.size - o, endpos - p)
                    f.write(block[o:o + take].tobytes())
                    p += take
            files.append(path)
            pos = endpos
        args = [a if a != "@WORDS@" else os.path.join(PAT_DIR, "words.txt") for a in CONFIGS[cfg][3]]
        cmd = [REF_UGREP, "--no-config", "-J%d" % cores, *args, *files]
        times = []
        out = b""
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            r = subprocess.run(cmd, capture_output=True)
            dt = time.perf_counter() - t0
            if r.returncode not in (0, 1):
                raise RuntimeError("reference ugrep failed: %s" % r.stderr[:200])
            if it >= warmup:
                times.append(dt)
            out = r.stdout
        count = 0
        for line in out.splitlines():
            tail = line.rsplit(b":", 1)[-1]
            if tail.isdigit():
                count += int(tail)
        sec = sum(times) / len(times)
        return {"gbs": total / sec / 1e9, "seconds": sec, "cores": cores, "files": nfiles, "bytes": total,
                "count": count, "best_gbs": total / min(times) / 1e9, "warmup": warmup, "steps": steps}
    finally:
        shutil.rmtree(d, ignore_errors=True)


# ---------------------------------------------------------------- our arm (GPU)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=HEADLINE, choices=list(CONFIGS))
    ap.add_argument("--gib", type=float, default=None, help="GiB per GPU (default: the config's named size)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other configs (quick runs, profiling)")
    ap.add_argument("--others-gib", type=float, default=None, help="cap the other configs' size (GiB per GPU)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = args.config
    gib = args.gib if args.gib is not None else CONFIGS[cfg][4]
    peak, peak_kind = peaks()

    if args.impl == "reference":
        if rank != 0:
            return 0
        # a bounded sample (1 GiB) of the same workload; the warm-up and step counts are the ones asked for
        sample = int(min(gib * GIB, 1 * GIB))
        r = reference_run(cfg, sample, max(1, args.steps), max(0, args.warmup))
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ugrep is not built"}))
            return 0
        line = {
            "impl": "reference", "metric": "GB/s scanned", "value": round(r["gbs"], 3), "unit": "GB/s",
            "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"],
            "ms_per_step": round(r["seconds"] * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "%s, %.2f GiB sample in %d line-aligned files" % (WORKLOAD_TEXT[cfg], r["bytes"] / GIB, r["files"]),
                       "config": cfg},
            "cpu_baseline": {"value": round(r["gbs"], 3), "unit": "GB/s", "cores": r["cores"], "kind": "reference",
                             "sample": "%.2f GiB of the %s corpus, ugrep -J%d over %d files, page cache warm"
                                       % (r["bytes"] / GIB, cfg, r["cores"], r["files"])},
            "e2e": {"value": round(r["gbs"], 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "count": r["count"],
        }
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    from ugrep_b200 import api, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the scan path is CUDA-only; there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        tm = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return float(tm.item())

    stream = torch.cuda.current_stream().cuda_stream
    sc = api.Scanner(local, stream)

    def stepper(pat, mode):
        def step(data):
            if mode == "lines":
                return sc.count_lines(pat, data)
            if mode == "matches":
                return sc.count_matches(pat, data)
            return sc.find_all_device(pat, data)
        return step

    # ---- corpus: one seeded block, tiled to the named size.  N = 1: the tiles as they are.  N > 1: ONE logical corpus of
    # N x reps + 1 tiles, cut into N shards by sharding.tiled_cuts (the nominal cut n*r/N falls inside a line and moves
    # forward to the next newline), rank r materialises and scans shard r; the path's one exchange — all-gather of the
    # per-shard {matches, newlines} -> totals and line-number bases — sits inside the timed region of every step.
    block = make_block(cfg)
    reps = max(1, int(gib * GIB) // block.size)
    dblock = torch.from_numpy(block).cuda()
    pat = api.Pattern.load(os.path.join(PAT_DIR, CONFIGS[cfg][0] + ".ugxp"), local)
    mode = CONFIGS[cfg][2]
    step = stepper(pat, mode)
    t1 = step(dblock)   # the block alone: its count is checked against the oracle below (rank 0)
    if world > 1:
        reps_total = world * reps + 1
        cuts = sharding.tiled_cuts(block, reps_total, world)
        corpus_dev = sharding.materialize_tiled(dblock, cuts[rank], cuts[rank + 1])
    else:
        reps_total = reps
        cuts = [0, block.size * reps]
        corpus_dev = dblock.repeat(reps)
    del dblock
    nbytes = int(corpus_dev.numel())               # this rank's shard
    logical_bytes = block.size * reps_total         # the whole job
    expect_total = t1.matches * reps_total          # counts are additive over line-aligned pieces

    gather = sharding.CountExchange("cuda") if world > 1 else None

    def exchange(t):
        if world == 1:
            return t.matches
        nl = t.newlines if t.flags & 1 else shard_newlines
        return sum(c[0] for c in gather(t.matches, nl))

    # (the streaming `-c` kernels do not count newlines unless asked: the shard's newline count comes from nlcount, once)
    shard_newlines = sc.count_newlines(corpus_dev).newlines if world > 1 else 0
    for _ in range(args.warmup):
        tot = step(corpus_dev)
        total_matches = exchange(tot)
    if total_matches != expect_total:
        raise SystemExit("bench.py: wrong count %d != %d x %d" % (total_matches, t1.matches, reps_total))
    expect = tot.matches                            # this rank's own count (the e2e legs must reproduce it)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    kernel_ms = []
    launches = 0
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tot = step(corpus_dev)
        total_matches = exchange(tot)
        kernel_ms.append(tot.kernel_ms)
        launches += tot.launches
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    if total_matches != expect_total:
        raise SystemExit("bench.py: wrong count in the timed region")
    ms_per_step = ms / args.steps
    value = logical_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: host buffer through the C ABI, H2D + scan + D2H of the result inside the timed region
    e2e = None
    if not args.no_e2e:
        def time_host(hb, esteps):
            for _ in range(2):
                th = step(hb)
            if th.matches != expect:
                raise SystemExit("bench.py: wrong e2e count")
            barrier()
            t0 = time.perf_counter()
            for _ in range(esteps):
                th = step(hb)
            torch.cuda.synchronize()
            return max_over_ranks(time.perf_counter() - t0)

        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        host.copy_(corpus_dev)          # this rank's shard, byte for byte, in page-locked host memory
        hb = host.numpy()
        esteps = max(3, min(args.steps, 5))
        dt = time_host(hb, esteps)
        e2e = {"value": round(logical_bytes * esteps / dt / 1e9, 3), "unit": "GB/s",
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 32, "steps": esteps, "host_memory": "pinned"}
        # the same from ordinary pageable memory (what an mmap'ing caller hands over without registering it): the
        # library's feeder threads copy it through pinned slots (capi.cu feed_pageable)
        pg = np.empty(nbytes, dtype=np.uint8)
        pg[:] = hb
        del host, hb
        dtp = time_host(pg, 2)
        e2e["pageable"] = {"value": round(logical_bytes * 2 / dtp / 1e9, 3), "unit": "GB/s", "steps": 2}
        del pg

    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)
    del corpus_dev
    torch.cuda.empty_cache()

    # ---- the other configs (same timing method; each checked against the oracle on one block)
    oracle_check = None
    others = {}
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O  # the checker: CPU restatement of the path (test infrastructure)

        def oracle_block(pfile, blk, m):
            op = O.OraclePattern(os.path.join(PAT_DIR, pfile + ".ugxp"))
            if m == "lines":
                return op.count_lines(blk)
            if m == "matches":
                return op.count_matches(blk)
            return op.find_all(blk)

        want = oracle_block(CONFIGS[cfg][0], block, mode)
        oracle_check = {"block_bytes": int(block.size), "gpu": int(t1.matches),
                        "oracle": int(want if not hasattr(want, "__len__") else len(want))}
        oracle_check["ok"] = oracle_check["gpu"] == oracle_check["oracle"]

    def run_other(oc):
        pfile, _, om, _, ogib = CONFIGS[oc]
        if args.others_gib is not None:
            ogib = min(ogib, args.others_gib)
        ob = make_block(oc)
        opat = api.Pattern.load(os.path.join(PAT_DIR, pfile + ".ugxp"), local)
        ostep = stepper(opat, om)
        dob = torch.from_numpy(ob).cuda()
        res = {"workload": WORKLOAD_TEXT[oc]}
        sharded = oc == "c5"
        if sharded:
            # one logical corpus of world x ogib GiB (+ one block, so that the nominal cuts fall inside lines), cut
            # forward to the next newline; rank r materialises and scans shard r only
            reps_total = world * max(1, int(ogib * GIB) // ob.size) + (1 if world > 1 else 0)
            cuts = sharding.tiled_cuts(ob, reps_total, world)
            lo, hi = cuts[rank], cuts[rank + 1]
            od = sharding.materialize_tiled(dob, lo, hi)
            res["sharding"] = {"logical_bytes": int(ob.size) * reps_total, "cuts": [int(c) for c in cuts],
                               "cut_rule": "n*r/N forward to the next newline (sharding.tiled_cuts)"}
        else:
            if world > 1 and rank != 0:
                return None
            reps_total = max(1, int(ogib * GIB) // ob.size)
            od = dob.repeat(reps_total)
        tb = ostep(dob)   # the block alone: compared with the oracle on rank 0
        rec_block = sc.fetch(0, tb.matches) if om == "list" else None
        del dob
        for _ in range(2):
            t = ostep(od)
        kms = []
        nsteps = 5
        ogather = sharding.CountExchange("cuda") if sharded else None
        if sharded:
            barrier()
        else:
            torch.cuda.synchronize()
        a0 = torch.cuda.Event(enable_timing=True)
        a1 = torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(nsteps):
            t = ostep(od)
            kms.append(t.kernel_ms)
            if sharded:
                # the path's one exchange step, inside the timed region: per-shard {matches, newlines} -> totals, bases
                counts = ogather(t.matches, t.newlines)
        a1.record()
        torch.cuda.synchronize()
        oms = a0.elapsed_time(a1)
        if sharded:
            oms = max_over_ranks(oms)
            total = sum(c[0] for c in counts)
            logical = int(ob.size) * reps_total
            bases = sharding.bases_from_counts(counts, cuts)
            res["sharding"]["line_bases"] = [int(b[1]) for b in bases]
        else:
            total = t.matches
            logical = int(od.numel())
        k = sum(kms) / len(kms)
        res.update({"gib_per_gpu": round(od.numel() / GIB, 3), "n_gpus": world if sharded else 1,
                    "value": round(logical / (oms / nsteps * 1e-3) / 1e9, 2), "unit": "GB/s",
                    "kernel": t.kernel, "kernel_ms": round(k, 4), "launches_per_step": t.launches,
                    "frac": round(od.numel() / (k * 1e-3) / 1e9 / peak, 4), "count": int(total)})
        if rank == 0:
            want = oracle_block(pfile, ob, om)
            if om == "list":
                ok_block = len(want) == len(rec_block) and bool(np.all(want == rec_block))
                want_n = len(want)
            else:
                ok_block = int(want) == int(tb.matches)
                want_n = int(want)
            res["oracle"] = {"block_bytes": int(ob.size), "block_count": want_n, "gpu_block_count": int(tb.matches),
                             "tiles": reps_total, "ok": bool(ok_block and total == want_n * reps_total)}
            if not res["oracle"]["ok"]:
                raise SystemExit("bench.py: %s disagrees with the oracle: %r" % (oc, res))
        del od
        torch.cuda.empty_cache()
        return res

    if not args.no_others:
        for oc in OTHERS_ORDER:
            if oc == cfg:
                continue
            if world > 1 and oc != "c5":
                continue  # under torchrun only the sharded config is timed besides the headline
            r = run_other(oc)
            if r is not None and rank == 0:
                others[oc] = r

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            r = reference_run(cfg, min(nbytes, 1 * GIB), 2, 1)
            if r is not None:
                cpu = {"value": round(r["gbs"], 3), "unit": "GB/s", "cores": r["cores"], "kind": "reference",
                       "sample": "%.2f GiB of the %s corpus in %d line-aligned files, oracle/_ref/ugrep -J%d, page cache warm, mean of 2"
                                 % (r["bytes"] / GIB, cfg, r["files"], r["cores"])}
        except Exception as ex:  # the baseline is reported, never required
            cpu = {"value": None, "unit": "GB/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % ex}

    if rank == 0:
        if oracle_check is not None and not oracle_check["ok"]:
            raise SystemExit("bench.py: the headline block count disagrees with the oracle: %r" % (oracle_check,))
        k_ms = sum(kernel_ms) / len(kernel_ms)
        achieved = nbytes / (k_ms * 1e-3) / 1e9
        info = pat.info
        cap = capture_for(cfg, tot.kernel, nbytes, k_ms)
        smem = None
        if cap and "smem_wavefronts_per_launch" in cap:
            wf = cap["smem_wavefronts_per_launch"]
            smem = {"wavefronts_per_clk_per_sm": round(wf / 148.0 / cap["sm_cycles_per_launch"], 3), "peak": 1.0,
                    "bank_conflict_share": round(cap["smem_bank_conflict_wavefronts_per_launch"] / wf, 3),
                    "wavefronts_per_kib": round(wf * 1024.0 / nbytes, 1), "source": cap["source"]}
        line = {
            "metric": "GB/s scanned", "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "%s, %.2f GiB per GPU (%d x %d MiB seeded block, line-aligned)"
                                   % (WORKLOAD_TEXT[cfg], nbytes / GIB, reps, block.size >> 20),
                       "config": cfg, "l2": "input (%.1f GiB) larger than L2" % (nbytes / GIB),
                       "sharding": ("one logical corpus of %d tiles (%.2f GiB) cut into %d line-aligned shards by "
                                    "sharding.tiled_cuts; NCCL all-gather of {matches, newlines} inside every timed step"
                                    % (reps_total, logical_bytes / GIB, world)) if world > 1 else "single shard",
                       "cuts": [int(c) for c in cuts],
                       "dfa_states": info["states"], "byte_classes": info["classes"], "prefilter": info["advance_name"],
                       "table_in_smem": bool(info["table_in_smem"])},
            "hbm_frac": round(value / world / peak, 4),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4),
                         "traffic": int(cap["dram_bytes_per_launch"]) if cap else None,
                         "algorithmic_bytes": nbytes, "smem": smem, "peak_kind": peak_kind,
                         "kernel": tot.kernel, "kernel_ms": round(k_ms, 4)},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "result": {"count": total_matches, "expected": expect_total, "rank0_count": expect, "oracle": oracle_check},
            "others": others,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
