#!/usr/bin/env python
"""bench.py — BASELINE.json's metric (GCUPS, sequences/s) on config 2 (gene families, 20 x 1 kb, 5% edits).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--families F] [--impl ours|reference]

One "step" = one progressive-MSA pass of the hot path over one batch of F families (19 alignment rounds,
each round one device batch), called through the public C-ABI entry `poasta_msa_batch` with HOST buffers.
Step k aligns families [k*F, (k+1)*F) with F = 5,000 by default: the default 2 steps are BASELINE.json's config 2
(10,000 families); warm-up steps use further families of the same shape.
  value   GCUPS over the device kernels only (CUDA events on the launching streams around the DP fill /
          traceback and the order-exact search launches; inputs of each launch resident in HBM). The pipeline
          lanes of a rank run their kernels concurrently, so the device time of a step is the largest per-lane
          sum of kernel time (the device-side critical path), not the sum over lanes;
  e2e     GCUPS of the whole call: FASTA text in host memory -> final graphs, including flattening,
          H2D/D2H copies and the host-side graph insertion between rounds — the headline.
--tiebreak exact (default) returns the reference's alignments bit for bit (DP score + order-exact search);
--tiebreak dp returns the optimal score and a co-optimal alignment from the DP traceback alone.
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


def measured_traffic():
    """DRAM bytes per cell update of each kernel from the committed ncu captures (profiles/r01_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


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
    """Families [first, first + count) of config 2 (seeded per family, identical for both bench arms)."""
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
    sample = max(8, min(args.families, 2 * cores))  # bounded sample of every step's batch (about 5-10 s of CPU work each)
    for _ in range(1 if args.warmup else 0):
        cpu_reference(make_families(args.steps * args.families, max(1, cores // 2), args), cores)
    tot_s = tot_cells = tot_seqs = 0
    for k in range(args.steps):
        s, c, q = cpu_reference(make_families(k * args.families, sample, args), cores)
        tot_s += s
        tot_cells += c
        tot_seqs += q
    gcups = tot_cells / tot_s / 1e9
    line = {
        "impl": "reference", "metric": "GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "seqs_per_s": tot_seqs / tot_s,
        "config": workload_config(args, args.families),
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": cores, "kind": "port",
                         "sample": f"first {sample} of the {args.families} families of every step, one family per task, {cores} threads; "
                                   "equivalent GCUPS (A* visits a fraction of the cells)"},
        "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, families, world=1):
    total = families * args.steps * world
    return {"workload": f"config2: {total} synthetic gene families x {args.seqs} seqs x {args.length} bp, "
                        f"{args.rate:.0%} subs+indels, global, costs 4/6/2, mingap; progressive MSA ({args.seqs - 1} rounds) in "
                        f"{args.steps} steps of {families} families per GPU",
            "families_per_step_per_gpu": families, "tiebreak": getattr(args, "tiebreak", "exact"),
            "l2": "every step reads different families; per-round working set (tables, trace) is far larger than the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--families", type=int, default=int(os.environ.get("BENCH_FAMILIES", "5000")), help="families per step per GPU")
    ap.add_argument("--tiebreak", default=os.environ.get("BENCH_TIEBREAK", "exact"), choices=["exact", "dp"])
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
    # the ranks of one node share its host cores: give every pipeline lane (4 per rank) its share of them
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    os.environ.setdefault("POASTA_THREADS_PER_LANE", str(max(1, min(16, (os.cpu_count() or 1) // (local_world * 4)))))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from poasta_b200 import aligner, msa_batch

    from poasta_b200 import TIEBREAK_DP, TIEBREAK_EXACT
    tiebreak = TIEBREAK_DP if args.tiebreak == "dp" else TIEBREAK_EXACT
    F = args.families
    # step k of rank r aligns families [(k * world + r) * F, ...); warm-up batches lie beyond the timed ones
    from poasta_b200.sharding import step_family_range
    batches = [make_families(step_family_range(k, rank, world, F)[0], F, args) for k in range(args.steps)]
    warm_batches = [make_families(step_family_range(args.steps + w, rank, world, F)[0], F, args) for w in range(min(args.warmup, 2))]

    def step(batch):
        t0 = time.perf_counter()
        graphs, scores, cells = msa_batch(batch, device=local, tiebreak=tiebreak)
        wall = time.perf_counter() - t0
        kms, launches, _ = aligner.last_batch_stats()
        sms = aligner.last_batch_search_ms()
        crit = aligner.last_batch_device_ms()
        h2d, d2h = aligner.last_batch_io()
        nseq = sum(len(s) for s in scores)
        del graphs
        return [wall, kms, sms, launches, cells, nseq, h2d, d2h, crit]

    for w in range(args.warmup):
        step(warm_batches[w % len(warm_batches)])

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sync()
    t_begin = time.perf_counter()
    tot = [0.0] * 9
    for k in range(args.steps):
        tot = [a_ + b_ for a_, b_ in zip(tot, step(batches[k]))]
    sync()
    wall = time.perf_counter() - t_begin
    clocks = sampler.stop() if rank == 0 else None

    mx = torch.tensor([wall, tot[1], tot[2], tot[8]], dtype=torch.float64, device="cuda")
    sums = torch.tensor(tot[3:8], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    wall_max, dp_ms, search_ms, dev_ms = mx.tolist()
    launches, cells, nseq, h2d, d2h = sums.tolist()

    if rank == 0:
        peak, which = peaks()
        # dev_ms = device-side critical path: per step the largest per-lane sum of kernel time (CUDA events on the lane's
        # stream; the pipeline lanes of a rank run their kernels concurrently), max over ranks. dp_ms / search_ms are the
        # sums over all lanes and launches, i.e. launch durations added up although launches overlap.
        gcups_dev = cells / (dev_ms / 1e3) / 1e9
        gcups_e2e = cells / wall_max / 1e9
        per_rank_cells = cells / world
        # roofline of the dominant kernel of the step (by device time)
        dominant = "replay_kernel (order-exact search)" if search_ms > dp_ms else "dp_fill kernels"
        dom_ms = max(search_ms, dp_ms)
        # contract: algorithmic bytes per launch / average launch duration = bytes of the step / summed launch durations
        achieved = per_rank_cells * ALG_BYTES_PER_CU / (dom_ms / 1e3) / 1e9
        tr = measured_traffic().get("replay_kernel" if search_ms > dp_ms else "dp_fill_warp_kernel")
        n_dom_launches = max(1, int(launches / world) // (3 if tiebreak == TIEBREAK_EXACT else 4))
        traffic = tr["dram_bytes_per_cu"] * per_rank_cells / n_dom_launches if tr else None
        dp_gbs = per_rank_cells * ALG_BYTES_PER_CU / (dp_ms / 1e3) / 1e9 if dp_ms else None
        line = {
            "metric": "GCUPS", "value": gcups_dev, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall_max / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s32", "data": "synthetic",
            "config": workload_config(args, F, world),
            "seqs_per_s": nseq / wall_max,
            "device_ms_per_step": {"dp_fill+traceback": dp_ms / args.steps, "order_exact_search": search_ms / args.steps,
                                   "critical_path": dev_ms / args.steps,
                                   "note": "the first two are launch durations summed over the concurrent pipeline lanes; "
                                           "critical_path = the largest per-lane sum, which `value` is quoted on"},
            "e2e": {"value": gcups_e2e, "unit": "GCUPS", "h2d_bytes_per_step": h2d / world / args.steps,
                    "d2h_bytes_per_step": d2h / world / args.steps, "seqs_per_s": nseq / wall_max},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": which,
                         "traffic_note": (f"ncu dram__bytes_read+write = {tr['dram_bytes_per_cu']:.2f} B per cell update ({tr['capture']}), "
                                          "scaled to the average launch of this run") if tr else "no ncu capture",
                         "note": "algorithmic bytes = 1 B per cell update (SURVEY.md 8d traceback budget) per launch / average launch "
                                 "duration (CUDA events); up to 4 launches (pipeline lanes) share the device at a time, so the "
                                 "device-level rate is `value`. dp_fill_kernel alone: "
                                 + (f"{dp_gbs:.1f} GB/s = {dp_gbs / peak:.4f} of peak, "
                                    f"{per_rank_cells * ALG_INT_OPS_PER_CU / (dp_ms / 1e3) / 1e12:.3f} Tint-op/s at 12 ops/CU" if dp_ms else "n/a")},
            "clocks": clocks,
        }
        # integer issue peak of this device, measured after the timed region (the DP fill is integer-bound)
        try:
            add_t, min_t, addmin_t = aligner.measure_int_peak(local)
            int_peak = max(add_t, min_t, addmin_t)
            dp_tops = per_rank_cells * ALG_INT_OPS_PER_CU / (dp_ms / 1e3) / 1e12 if dp_ms else None
            line["int_roofline"] = {"kernel": "dp_fill kernels", "bound": "integer issue", "unit": "Tint-op/s",
                                    "achieved": dp_tops, "peak": int_peak, "frac": dp_tops / int_peak if dp_tops else None,
                                    "peaks_measured": {"add": add_t, "xor+min": min_t, "fused add+min (DPX)": addmin_t},
                                    "note": "achieved = 12 algorithmic ops per cell update (6*d_v + 6 at in-degree 1) x cells / DP kernel "
                                            "time summed over the concurrent pipeline lanes; peak = the fastest of three micro-kernels "
                                            "(8 independent chains per thread, 8 CTAs of 256 threads per SM)"}
        except Exception as exc:  # the probe is informative only
            line["int_roofline"] = {"error": str(exc)}
        if world == 1 and tiebreak == TIEBREAK_EXACT and not args.no_cpu_baseline:
            # the same workload through the DP-only mode (optimal scores, co-optimal alignments; not bit-exact at ties)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            g2, s2, c2 = msa_batch(batches[0], device=local, tiebreak=TIEBREAK_DP)
            w2 = time.perf_counter() - t0
            k2, _, _ = aligner.last_batch_stats()
            del g2
            line["dp_only_mode"] = {"e2e_gcups": c2 / w2 / 1e9, "seqs_per_s": sum(len(x) for x in s2) / w2, "ms_per_step": w2 * 1e3,
                                    "kernel_gcups": c2 / (min(k2, w2 * 1e3) / 1e3) / 1e9,
                                    "note": "tiebreak=dp on step 0's batch after the timed region; kernel time is the sum over "
                                            "concurrent pipeline lanes"}
        if world == 1 and not args.no_cpu_baseline:
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
            cores = os.cpu_count() or 1
            sample = max(4, min(F, cores))
            s_, c_, q_ = cpu_reference(batches[0][:sample], cores)
            line["cpu_baseline"] = {"value": c_ / s_ / 1e9, "unit": "GCUPS", "cores": cores, "kind": "port",
                                    "sample": f"first {sample} families of step 0, one family per task, {cores} threads, "
                                              f"{s_:.1f} s wall; equivalent GCUPS (A* visits a fraction of the cells)",
                                    "seqs_per_s": q_ / s_}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
