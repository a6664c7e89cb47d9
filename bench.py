#!/usr/bin/env python
"""bench.py — BASELINE.json's metric (GCUPS, sequences/s) on the reference's configs; default = config 2 (gene families,
20 x 1 kb, 5 % edits), the configuration the metric is quoted on.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--families F] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input through the public C-ABI entry with HOST buffers:
  config 2/3/4  `poasta_msa_batch`: progressive MSA of F families (FASTA text in host memory -> final graphs in host memory).
                In reference-exact mode (default) the whole build runs on the device: sequences in, ONE persistent kernel per
                batch (per-round graph index, order-exact search, graph update, all on the GPU), finished graphs out.
  config 5      `poasta_align_batch`: R reads against one fixed graph, ends-free, -H path (library semantics).
  value   GCUPS over the device kernels only (CUDA events on the launching stream around the engine / search launches; the
          step's inputs are resident in HBM when they start);
  e2e     GCUPS of the whole call with host buffers, H2D / D2H copies and the host-side conversion of the results inside
          the timed region — the headline.
1 CU = the three states (M, I, D) of one (real node, query position) pair; cells = sum V_real * Q.
Under torchrun each rank takes a disjoint shard of the units (no collective on the data path; weak scaling).

--impl reference times the CPU restatement of the reference's A* path (oracle/; the Rust reference cannot be built in this
image, so cpu_baseline.kind is "port") on all host cores, one family / read per task, on a bounded sample of every step.
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


def measured_traffic(name):
    """DRAM bytes per cell update of a kernel from the committed ncu captures (profiles/r02_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    return json.load(open(p)).get(name) if os.path.exists(p) else None


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


# ---- workloads (seeded per unit, identical for both bench arms) ---------------------------------------
CONFIGS = {
    2: dict(name="config2", seqs=20, length=1000, rate=0.05, units=6000, unit="families",
            what="synthetic gene families x 20 seqs x 1 kb, 5% subs+indels, global, costs 4/6/2, mingap; progressive MSA (19 rounds)"),
    3: dict(name="config3", seqs=30, length=500, rate=0.10, units=12000, unit="windows",
            what="long-read consensus windows x 30 reads x 500 bp, 10% error, global, costs 4/6/2, mingap; progressive MSA (29 rounds)"),
    4: dict(name="config4", seqs=16, length=1500, rate=0.15, units=1, unit="family",
            what="ONE family of 16S-like 1.5 kb sequences at 15% divergence (a single sequential chain: latency-bound, replicas only); "
                 "bounded to the first `seqs` members of the 1,000"),
    5: dict(name="config5", seqs=50, length=20000, rate=0.01, units=20000, unit="reads", read_len=2000, read_rate=0.05,
            what="reads of 2 kb (5% error) against ONE fixed graph of 50 haplotypes x 20 kb (1% divergence), ends-free, -H path, "
                 "library semantics (every read against the fixed graph)"),
}


def make_units(cfg, first, count, args):
    from poasta_b200 import synth
    c = CONFIGS[cfg]
    seqs, length, rate = args.seqs or c["seqs"], args.length or c["length"], args.rate if args.rate is not None else c["rate"]
    if cfg == 2:
        return [synth.config2_family(first + f, n_seqs=seqs, root_len=length, rate=rate) for f in range(count)]
    if cfg == 3:
        return [synth.config3_window(first + f, n_reads=seqs, template_len=length, rate=rate) for f in range(count)]
    if cfg == 4:
        return [synth.config4_family(n_seqs=seqs, root_len=length, rate=rate)]
    raise ValueError(cfg)


def config5_inputs(args, n_reads, first):
    """(haplotype FASTA, reads [first, first + n_reads)). The graph is built from the haplotypes once, outside the timed region."""
    from poasta_b200 import synth
    c = CONFIGS[5]
    return synth.config5_reads(args.seqs or c["seqs"], args.length or c["length"], args.rate if args.rate is not None else c["rate"],
                               first, n_reads, c["read_len"], c["read_rate"])


def cpu_reference_msa(fastas, threads, prm=(4, 6, 2, 0, 1)):
    """The CPU arm: oracle (restated reference A*) progressive MSA, one family per task over `threads` threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    n = len(fastas)
    ptr = (C.c_char_p * n)(*fastas)
    lens = (C.c_size_t * n)(*[len(f) for f in fastas])
    cells, seqs, err = C.c_uint64(), C.c_uint64(), C.c_int()
    secs = H.oracle().oracle_msa_many(ptr, lens, n, *prm, threads, C.byref(cells), C.byref(seqs), C.byref(err))
    if err.value:
        raise RuntimeError(f"oracle failed: {err.value}")
    return secs, cells.value, seqs.value


def cpu_reference_reads(graph, reads, threads):
    """Config 5 CPU arm: every read against the fixed graph with the oracle, ends-free, -H path, `threads` threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    return H.oracle_align_many(graph, reads, threads, mode=1, heuristic=2)


def workload_config(cfg, args, units, world=1):
    c = CONFIGS[cfg]
    d = {"workload": f"{c['name']}: {c['what']}; {args.steps} steps of {units} {c['unit']} per GPU",
         f"{c['unit']}_per_step_per_gpu": units, "tiebreak": getattr(args, "tiebreak", "exact"),
         "l2": "every step reads different units; the per-alignment working sets (visited tables, graphs) total tens of GB per GPU, "
               "far larger than the 126 MB L2"}
    for k in ("seqs", "length", "rate"):
        if getattr(args, k) is not None:
            d[k] = getattr(args, k)
    return d


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    cores = os.cpu_count() or 1
    cfg = args.config
    units = args.families or CONFIGS[cfg]["units"]
    tot_s = tot_cells = tot_seqs = 0
    if cfg == 5:
        sample = max(8, min(units, 64 * cores))
        haps, _ = config5_inputs(args, 0, 0)
        graph = build_config5_graph(haps, device=None)
        for k in range(args.steps):
            _, reads = config5_inputs(args, sample, k * units)
            s, c = cpu_reference_reads(graph, reads, cores)
            tot_s += s
            tot_cells += c
            tot_seqs += len(reads)
        what = f"first {sample} of the {units} reads of every step, one read per task, {cores} threads"
    else:
        sample = max(1, min(units, 2 * cores)) if cfg != 3 else max(1, min(units, 8 * cores))
        if args.warmup:
            cpu_reference_msa(make_units(cfg, args.steps * units, max(1, min(sample, cores // 2)), args), cores)
        for k in range(args.steps):
            s, c, q = cpu_reference_msa(make_units(cfg, k * units, sample, args), cores)
            tot_s += s
            tot_cells += c
            tot_seqs += q
        what = f"first {sample} of the {units} {CONFIGS[cfg]['unit']} of every step, one per task, {cores} threads"
    gcups = tot_cells / tot_s / 1e9
    line = {
        "impl": "reference", "metric": "GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "seqs_per_s": tot_seqs / tot_s,
        "config": workload_config(cfg, args, units),
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": cores, "kind": "port",
                         "sample": what + "; equivalent GCUPS (A* visits a fraction of the cells); oracle built -O3 -march=native"},
        "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def build_config5_graph(haps, device):
    """The fixed graph of config 5: a benchmark INPUT, built once per process outside the timed region with the product's DP-only
    mode (optimal alignments of the 50 haplotypes; the oracle's A* needs ~13 minutes for this graph, so both arms use the same
    builder and therefore the same graph). Without a CUDA device (CPU-only test runs at toy sizes) the oracle builds it."""
    from poasta_b200 import POAGraph
    import torch
    if torch.cuda.is_available():
        from poasta_b200 import TIEBREAK_DP, msa_batch
        return msa_batch([haps], device=device or 0, tiebreak=TIEBREAK_DP)[0][0]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    rc, data, _, _ = H.oracle_msa(haps, fmt=3)
    assert rc == 0
    return POAGraph.load(data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=int(os.environ.get("BENCH_CONFIG", "2")), choices=[2, 3, 4, 5])
    ap.add_argument("--families", type=int, default=int(os.environ.get("BENCH_FAMILIES", "0")), help="units (families / windows / reads) per step per GPU")
    ap.add_argument("--tiebreak", default=os.environ.get("BENCH_TIEBREAK", "exact"), choices=["exact", "dp"])
    ap.add_argument("--seqs", type=int, default=None)
    ap.add_argument("--length", type=int, default=None)
    ap.add_argument("--rate", type=float, default=None)
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
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    cores = os.cpu_count() or 1
    # the ranks of one node share its host cores (only the FASTA parsing and the final graph conversion run there)
    os.environ.setdefault("POASTA_HOST_THREADS", str(max(1, min(16, cores // local_world))))
    os.environ.setdefault("POASTA_THREADS_PER_LANE", str(max(1, min(16, cores // (local_world * 4)))))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from poasta_b200 import TIEBREAK_DP, TIEBREAK_EXACT, AffinePathAware, AlignmentType, POAGraph, aligner, align_batch, msa_batch
    from poasta_b200.sharding import step_family_range

    cfg = args.config
    tiebreak = TIEBREAK_DP if args.tiebreak == "dp" else TIEBREAK_EXACT
    F = args.families or CONFIGS[cfg]["units"]
    n_warm_batches = min(args.warmup, 2)

    if cfg == 5:
        haps, _ = config5_inputs(args, 0, 0)
        graph = build_config5_graph(haps, device=local)
        batches = [config5_inputs(args, F, step_family_range(k, rank, world, F)[0])[1] for k in range(args.steps)]
        # (a step takes well under a second: full-size warm-up batches, so that every buffer has its final size before the timed steps)
        warm_batches = [config5_inputs(args, F, step_family_range(args.steps + w, rank, world, F)[0])[1] for w in range(n_warm_batches)]
        view = graph.view()
        real_nodes = graph.node_count()

        def step(batch):
            t0 = time.perf_counter()
            res = align_batch([graph], batch, [0] * len(batch), AffinePathAware(), AlignmentType.EndsFree, device=local, tiebreak=TIEBREAK_EXACT)
            wall = time.perf_counter() - t0
            kms, launches, _ = aligner.last_batch_stats()
            sms = aligner.last_batch_search_ms()
            h2d, d2h = aligner.last_batch_io()
            cells = sum(real_nodes * len(q) for q in batch)
            del res
            return [wall, kms, sms, launches, cells, len(batch), h2d, d2h, sms + kms]
    else:
        batches = [make_units(cfg, step_family_range(k, rank, world, F)[0], F, args) for k in range(args.steps)]
        # warm-up batches: the same number of units (same worker pool and buffer sizes as the timed steps) cut to their first three
        # records, so a warm-up step costs ~1/10 of a timed one
        def head(fa, n=3):
            parts = fa.split(b">")
            return b">".join(parts[:n + 1])
        # (DP-only mode: a step takes about a second, so its warm-up batches are full ones and size every buffer the timed steps use)
        cut = head if tiebreak == TIEBREAK_EXACT else (lambda fa: fa)
        warm_batches = [[cut(f) for f in make_units(cfg, step_family_range(args.steps + w, rank, world, F)[0], F, args)]
                        for w in range(n_warm_batches)] if cfg != 4 else [[head(batches[0][0], 4)]]

        def step(batch):
            t0 = time.perf_counter()
            graphs, scores, cells = msa_batch(batch, device=local, tiebreak=tiebreak)
            wall = time.perf_counter() - t0
            kms, launches, _ = aligner.last_batch_stats()
            sms = aligner.last_batch_search_ms()
            crit = aligner.last_batch_device_ms()
            h2d, d2h = aligner.last_batch_io()
            nseq = sum(len(s) for s in scores)
            fill = aligner.last_batch_fill_ms() if aligner.last_msa_path() == 2 else (0.0, 0)
            dp_fill[0] += fill[0]
            dp_fill[1] += fill[1]
            POAGraph.free_many(graphs)  # the finished graphs go back on several host threads (inside the timed region)
            return [wall, kms, sms, launches, cells, nseq, h2d, d2h, crit]

    dp_fill = [0.0, 0]  # DP-only device pipeline: milliseconds and number of the fill launches alone (this rank)
    for w in range(args.warmup):
        step(warm_batches[w % len(warm_batches)])
    dp_fill[0], dp_fill[1] = 0.0, 0

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
        gcups_dev = cells / (dev_ms / 1e3) / 1e9
        gcups_e2e = cells / wall_max / 1e9
        per_rank_cells = cells / world
        exact = tiebreak == TIEBREAK_EXACT or cfg == 5
        if exact:
            dominant = "msa_engine_kernel (persistent: index build + order-exact search + graph update)" if cfg != 5 else "replay_kernel (order-exact search)"
            dom_ms, tr = search_ms, measured_traffic("msa_engine_kernel" if cfg != 5 else "replay_kernel")
            n_dom = args.steps  # one engine launch per step (plus one per extra pass for suspended families)
        elif dp_fill[1]:
            # device-resident DP rounds: the fill launches are timed alone (CUDA events around each launch, this rank's)
            dominant, dom_ms, tr = "dp_fill_packed_kernel (device-resident DP rounds; fill launches timed alone)", dp_fill[0], measured_traffic("dp_fill_packed_kernel")
            n_dom = dp_fill[1]
        else:
            dominant, dom_ms, tr = "dp_fill_packed_kernel + traceback_rows_kernel", dp_ms, measured_traffic("dp_fill_packed_kernel")
            n_dom = max(1, int(launches / world) // 4)
        # algorithmic bytes per cell update: 1 B (traceback budget of SURVEY.md 8d) for the search; the packed DP fill stores the M and D
        # rows as u16 (4 B per cell update) instead of per-cell decision bits, and those rows ARE its traceback store
        alg_bytes = ALG_BYTES_PER_CU if exact else 4.0
        achieved = per_rank_cells * alg_bytes / (dom_ms / 1e3) / 1e9
        traffic = tr["dram_bytes_per_cu"] * per_rank_cells / n_dom if tr else None
        line = {
            "metric": "GCUPS", "value": gcups_dev, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall_max / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16" if exact else ("s16x2" if dp_fill[1] else "s32"), "data": "synthetic",
            "config": workload_config(cfg, args, F, world),
            "seqs_per_s": nseq / wall_max,
            "device_ms_per_step": {"dp_fill+traceback": dp_ms / args.steps, "engine_or_search": search_ms / args.steps,
                                   "critical_path": dev_ms / args.steps},
            "e2e": {"value": gcups_e2e, "unit": "GCUPS", "h2d_bytes_per_step": h2d / world / args.steps,
                    "d2h_bytes_per_step": d2h / world / args.steps, "seqs_per_s": nseq / wall_max},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": which,
                         "traffic_note": (f"ncu dram__bytes_read+write = {tr['dram_bytes_per_cu']:.2f} B per cell update ({tr['capture']}), "
                                          "scaled to the average launch of this run") if tr else "no ncu capture",
                         "algorithmic_bytes_per_cu": alg_bytes,
                         "note": ("algorithmic bytes = 1 B per cell update (SURVEY.md 8d traceback budget) x cells of the launch / launch duration "
                                  "(CUDA events). In reference-exact mode the kernel is the order-exact search, a latency-bound state machine "
                                  "(one dependent chain of memory round trips per in-flight alignment), not a streaming kernel: its HBM fraction "
                                  "is small by construction (DESIGN.md section 4)") if exact else
                                 ("algorithmic bytes = 4 B per cell update (M and D rows as u16, the traceback store of the packed kernel) x cells "
                                  "/ duration of " + ("the fill launches alone (CUDA events around each launch)" if dp_fill[1] else
                                                      "fill + traceback + scan (CUDA events, summed over the concurrent pipeline lanes)"))},
            "clocks": clocks,
        }
        if exact and cfg != 5:
            q_, v_, p_ = aligner.last_batch_search_counts()
            line["search_counts_last_step"] = {"num_queued": q_, "num_visited": v_, "num_pruned": p_}
        try:
            add_t, min_t, addmin_t = aligner.measure_int_peak(local)
            int_peak = max(add_t, min_t, addmin_t)
            path_tops = gcups_dev / world * ALG_INT_OPS_PER_CU / 1e3
            line["int_roofline"] = {"kernel": "whole device path", "bound": "integer issue", "unit": "Tint-op/s", "achieved": path_tops,
                                    "peak": int_peak, "frac": path_tops / int_peak,
                                    "peaks_measured": {"add": add_t, "xor+min": min_t, "fused add+min (DPX)": addmin_t},
                                    "note": "achieved = value (per GPU) x 12 algorithmic integer ops per cell update (6*d_v + 6 at in-degree 1); peak = "
                                            "the fastest of three micro-kernels (8 independent chains per thread, 8 CTAs of 256 threads per SM)"}
        except Exception as exc:  # the probe is informative only
            line["int_roofline"] = {"error": str(exc)}
        if world == 1 and exact and cfg in (2, 3) and not args.no_cpu_baseline:
            # the same workload through the DP-only mode (optimal scores, co-optimal alignments; not bit-exact at ties)
            torch.cuda.synchronize()
            sub = batches[0][:max(1, len(batches[0]) // 2)]
            POAGraph.free_many(msa_batch(sub, device=local, tiebreak=TIEBREAK_DP)[0])  # untimed: sizes this mode's buffers
            t0 = time.perf_counter()
            g2, s2, c2 = msa_batch(sub, device=local, tiebreak=TIEBREAK_DP)
            w2 = time.perf_counter() - t0
            k2, _, _ = aligner.last_batch_stats()
            POAGraph.free_many(g2)
            line["dp_only_mode"] = {"e2e_gcups": c2 / w2 / 1e9, "seqs_per_s": sum(len(x) for x in s2) / w2, "ms": w2 * 1e3,
                                    "kernel_gcups": c2 / (min(k2, w2 * 1e3) / 1e3) / 1e9, "units": len(sub),
                                    "kernel_int_roofline_frac": (c2 / (min(k2, w2 * 1e3) / 1e3) / 1e12 * ALG_INT_OPS_PER_CU / line["int_roofline"]["peak"]
                                                                 if "peak" in line.get("int_roofline", {}) else None),
                                    "kernel_hbm_gbs": 4.0 * c2 / (min(k2, w2 * 1e3) / 1e3) / 1e9,
                                    "pipeline": {1: "engine", 2: "device-resident DP rounds", 3: "per-round host pipeline"}.get(aligner.last_msa_path()),
                                    "note": "tiebreak=dp (packed s16x2 DP fill + traceback kernels produce the alignments; graph packing and update on "
                                            "the device too) on half of step 0's batch after the timed region; kernel time = all rounds on the device"}
        if world == 1 and not args.no_cpu_baseline:
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
            if cfg == 5:
                sample = max(8, min(F, 32 * cores))
                s_, c_ = cpu_reference_reads(graph, batches[0][:sample], cores)
                q_ = sample
                what = f"first {sample} reads of step 0, one read per task"
            else:
                sample = max(1, min(F, cores if cfg != 3 else 4 * cores))
                s_, c_, q_ = cpu_reference_msa(batches[0][:sample], cores)
                what = f"first {sample} {CONFIGS[cfg]['unit']} of step 0, one per task"
            line["cpu_baseline"] = {"value": c_ / s_ / 1e9, "unit": "GCUPS", "cores": cores, "kind": "port",
                                    "sample": f"{what}, {cores} threads, {s_:.1f} s wall; equivalent GCUPS (A* visits a fraction of the cells)",
                                    "seqs_per_s": q_ / s_}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
