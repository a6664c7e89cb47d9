#!/usr/bin/env python
"""bench.py — BASELINE.json's metric (GCUPS, sequences/s) on config 2 (gene families, 20 x 1 kb, 5% edits).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--families F] [--impl ours|reference]

One "step" = one progressive-MSA pass of the hot path over a batch of F families (19 alignment rounds,
each one device batch), called through the public C-ABI entry `poasta_msa_batch` with HOST buffers.
  value   GCUPS over the device kernels only (CUDA events around the fill/traceback launches, inputs of
          each launch already resident in HBM) — explains the kernel;
  e2e     GCUPS of the whole call: FASTA text in host memory -> final graphs, including flattening,
          H2D/D2H copies and the host-side graph insertion between rounds — the headline.
1 CU = the three states (M, I, D) of one (real node, query position) pair; cells = sum V_real * Q
(SURVEY.md §8d). Under torchrun each rank takes a disjoint shard of families (no collective on the
data path; weak scaling: F families per GPU).

--impl reference times the CPU restatement of the reference's A* path (the oracle; the Rust reference
cannot be built in this image) on all host cores, one family per task, on a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_CU = 1.0      # traceback budget 1 B/CU; scores stay on chip
ALG_INT_OPS_PER_CU = 12.0   # 6*d_v + 6 at in-degree 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


def make_families(first, count, args):
    from poasta_b200 import synth
    return [synth.config2_family(first + f, n_seqs=args.seqs, root_len=args.length, rate=args.rate) for f in range(count)]


def cpu_reference(fastas, threads):
    """The CPU arm: oracle (restated reference A*) progressive MSA, one family per task over `threads` threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    n = len(fastas)
    ptr = (C.c_char_p * n)(*fastas)
    lens = (C.c_size_t * n)(*[len(f) for f in fastas])
    cells, seqs, err = C.c_uint64(), C.c_uint64(), C.c_int()
    secs = H.oracle().oracle_msa_many(ptr, lens, n, 4, 6, 2, 0, 1, threads, C.byref(cells), C.byref(seqs), C.byref(err))
    if err.value:
        raise RuntimeError(f"oracle failed: {err.value}")
    return secs, cells.value, seqs.value


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    cores = os.cpu_count() or 1
    sample = max(8, min(args.families, 2 * cores))
    fastas = make_families(0, sample, args)
    for _ in range(args.warmup and 1):
        cpu_reference(fastas[:max(1, min(cores, sample) // 4)], cores)
    tot_s = tot_cells = tot_seqs = 0
    for _ in range(args.steps):
        s, c, q = cpu_reference(fastas, cores)
        tot_s += s
        tot_cells += c
        tot_seqs += q
    gcups = tot_cells / tot_s / 1e9
    line = {
        "impl": "reference", "metric": "GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "seqs_per_s": tot_seqs / tot_s,
        "config": workload_config(args, sample),
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": cores, "kind": "port",
                         "sample": f"{sample} of {args.families} families per step, one family per task, {cores} threads; "
                                   "equivalent GCUPS (A* visits a fraction of the cells)"},
        "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, families):
    return {"workload": f"config2: {families} synthetic gene families x {args.seqs} seqs x {args.length} bp, "
                        f"{args.rate:.0%} subs+indels, global, costs 4/6/2, progressive MSA ({args.seqs - 1} rounds)",
            "families_per_gpu": families, "l2": "inputs larger than L2 (trace store >> 126 MB per round)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--families", type=int, default=int(os.environ.get("BENCH_FAMILIES", "10000")))
    ap.add_argument("--seqs", type=int, default=20)
    ap.add_argument("--length", type=int, default=1000)
    ap.add_argument("--rate", type=float, default=0.05)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: poasta_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from poasta_b200 import aligner, msa_batch

    fastas = make_families(rank * args.families, args.families, args)
    warm = fastas[:max(64, args.families // 20)]

    def step(batch):
        t0 = time.perf_counter()
        graphs, scores, cells = msa_batch(batch, device=local)
        wall = time.perf_counter() - t0
        kms, launches, _ = aligner.last_batch_stats()
        h2d, d2h = aligner.last_batch_io()
        nseq = sum(len(s) for s in scores)
        del graphs
        return wall, kms, launches, cells, nseq, h2d, d2h

    for w in range(args.warmup):
        step(warm if w < args.warmup - 1 else fastas[:max(64, args.families // 4)])

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sync()
    t_begin = time.perf_counter()
    tot = [0.0, 0.0, 0, 0, 0, 0, 0]
    for _ in range(args.steps):
        r = step(fastas)
        tot = [a + b for a, b in zip(tot, r)]
    sync()
    wall = time.perf_counter() - t_begin
    clocks = sampler.stop() if rank == 0 else None

    wall_t = torch.tensor([wall, tot[1]], dtype=torch.float64, device="cuda")
    sums = torch.tensor([tot[3], tot[4], tot[2], tot[5], tot[6]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(wall_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    wall_max, kms_max = wall_t.tolist()
    cells, nseq, launches, h2d, d2h = sums.tolist()

    if rank == 0:
        peak, which = peaks()
        gcups_dev = cells / (kms_max / 1e3) / 1e9
        gcups_e2e = cells / wall_max / 1e9
        # roofline of the dominant kernel (dp_fill): algorithmic bytes = 1 B per cell update
        per_rank_cells = cells / world
        achieved = per_rank_cells * ALG_BYTES_PER_CU / (kms_max / 1e3) / 1e9
        line = {
            "metric": "GCUPS", "value": gcups_dev, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall_max / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s32", "data": "synthetic",
            "config": workload_config(args, args.families),
            "seqs_per_s": nseq / wall_max, "kernel_ms_per_step": kms_max / args.steps,
            "e2e": {"value": gcups_e2e, "unit": "GCUPS", "h2d_bytes_per_step": h2d / world / args.steps,
                    "d2h_bytes_per_step": d2h / world / args.steps, "seqs_per_s": nseq / wall_max},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": which,
                         "note": "algorithmic bytes = 1 B/CU (traceback budget); int-op view: "
                                 f"{per_rank_cells * ALG_INT_OPS_PER_CU / (kms_max / 1e3) / 1e12:.3f} Tint-op/s at 12 ops/CU"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
            cores = os.cpu_count() or 1
            sample = max(4, min(args.families, cores))
            s, c, q = cpu_reference(fastas[:sample], cores)
            line["cpu_baseline"] = {"value": c / s / 1e9, "unit": "GCUPS", "cores": cores, "kind": "port",
                                    "sample": f"first {sample} families of the step, one family per task, {cores} threads, "
                                              f"{s:.1f} s wall; equivalent GCUPS (A* visits a fraction of the cells)",
                                    "seqs_per_s": q / s}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
