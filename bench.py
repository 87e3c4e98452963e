#!/usr/bin/env python3
"""bench.py -- sentences/sec of the Juman++ lattice-scoring hot path.

    python bench.py --gpus N --steps K --warmup W            (product arm, CUDA)
    python bench.py --impl reference --gpus N --steps K ...   (reference CPU arm)

A *step* is one pass of the hot path (AnalyzerImpl::computeScores for every
sentence: T0 features, global-beam scoring, beams, RNN rescoring, re-rank) over
the whole synthetic workload.  The workload is BASELINE.json configs[1]:
Jumandic spec, beam=5 (global beam 6, right check 1, right beam 5), RNN on,
100 000 synthetic sentences of 8-128 codepoints; per rank under torchrun (weak
scaling: every rank scores its own 100 000 sentences).

Untimed preparation (cached under --cache): the synthetic dictionary / corpus /
RNN model are generated from seeds, and the lattices are BUILT ON THE HOST by
the reference's own lattice builder (oracle/_ref/jpp_ref_harness stage
expected=0 -- resetForInput/prepareNodeSeeds/buildLattice only; host lattice
construction is out of scope and stays the reference's job, SURVEY.md section 2
rows 12-13).  No scoring of the product arm goes through oracle/: the harness
is used again only in the cpu_baseline leg (timing the reference's
computeScores on a bounded sample, and as the 1-best checker).

Printed JSON (rank 0): see README "bench contract".  `value` = device-resident
throughput, `e2e` = the same work through jpp_b200_score_batch with pinned host
buffers (H2D of the staged lattices + D2H of the beams inside the timed region).
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from jumanpp_b200 import abi, jppb, shard, synth  # noqa: E402

HARNESS = ROOT / "oracle" / "_ref" / "jpp_ref_harness"

WORKLOADS = {
    # name: scoring flags, rnn layer size, RNN on
    "jumandic_beam5": dict(beam=5, gbeam=6, rcheck=1, rbeam=5, rnn=True, H=128,
                           desc="BASELINE configs[1]: Jumandic beam=5 (gbeam 6, right-check 1, right-beam 5) + RNN H=128"),
    "jumandic_beam20": dict(beam=20, gbeam=20, rcheck=1, rbeam=5, rnn=True, H=256,
                            desc="BASELINE configs[2]: Jumandic beam=20 gbeam=20 wide beam + RNN H=256"),
    "jumandic_beam5_linear": dict(beam=5, gbeam=6, rcheck=1, rbeam=5, rnn=False, H=0,
                                  desc="BASELINE configs[3]: Jumandic beam=5, RNN disabled (linear features only)"),
}


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------ workload --
def prepare_workload(args, rank: int, world: int = 1, dist=None) -> dict:
    wl = WORKLOADS[args.workload]
    cache = Path(args.cache)
    sharded = getattr(args, "shard_corpus", False) and world > 1
    # weak scaling: every rank its own corpus (seed + rank); --shard-corpus: ONE corpus, rank r stages its shard
    seed = args.seed if sharded else args.seed + rank
    tag = f"{args.workload}_d{args.dict_rows}_n{args.sentences}_l{args.min_len}-{args.max_len}_s{seed}_w{args.log2w}"
    if sharded:
        tag += f"_shard{rank}of{world}"
    d = cache / tag
    d.mkdir(parents=True, exist_ok=True)
    dic, corpus, rnn = d / "dict.csv", d / "corpus.txt", d / "rnn"
    done = d / "staged.done"
    info = dict(dir=d, dict=dic, corpus=corpus, rnn=rnn if wl["rnn"] else None, wl=wl, shard=None)
    if sharded and (d / "shard.json").exists():
        info["shard"] = json.loads((d / "shard.json").read_text())
    if done.exists():
        return info
    if not HARNESS.exists():
        raise SystemExit(f"{HARNESS} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where "
                         "/root/reference is mounted (the host lattice builder is the reference's)")
    t0 = time.time()
    surfaces = synth.write_dictionary(dic, args.dict_rows, 20260923)
    if sharded:
        # ONE corpus = `world` chunks generated side by side (rank c writes chunk c, seed 1000 * seed + c), concatenated
        # in chunk order; every rank then reads the chunks and keeps its contiguous shard, balanced by codepoints
        common = cache / (tag.rsplit("_shard", 1)[0] + "_chunks")
        common.mkdir(parents=True, exist_ok=True)
        per = [args.sentences // world + (1 if c < args.sentences % world else 0) for c in range(world)]
        synth.write_corpus(common / f"chunk_{rank}.txt", per[rank], args.min_len, args.max_len, 1000 * seed + rank, surfaces)
        (common / f"chunk_{rank}.done").write_text("ok")
        if dist is not None:
            dist.barrier()
        while not all((common / f"chunk_{c}.done").exists() for c in range(world)):
            time.sleep(0.2)
        lines = []
        for c in range(world):
            lines += (common / f"chunk_{c}.txt").read_text(encoding="utf-8").splitlines()
        bounds = shard.shard_bounds(np.array([len(x) for x in lines], np.int64), world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        corpus.write_text("\n".join(lines[lo:hi]) + "\n", encoding="utf-8")
        info["shard"] = {"lo": lo, "hi": hi, "total": len(lines), "bounds": [int(x) for x in bounds]}
        (d / "shard.json").write_text(json.dumps(info["shard"]))
    else:
        synth.write_corpus(corpus, args.sentences, args.min_len, args.max_len, seed, surfaces)
    if wl["rnn"]:
        words = synth.rnn_vocabulary(dic, 50000)
        synth.write_rnn_model(rnn, words, wl["H"], 3, 1 << 22, 9.0, 2)
    log(f"rank {rank}: synthetic dictionary/corpus/rnn in {time.time() - t0:.1f}s")
    t0 = time.time()
    cmd = [str(HARNESS), "stage", f"dict={dic}", f"corpus={corpus}", f"out={d / 'staged.jppb'}", "expected=0",
           f"chunk={args.stage_chunk}", f"threads={max(1, cpu_budget()['cores_effective'] // max(1, args.gpus))}",
           f"log2w=10", *harness_model_args(args, wl, rnn, log2w=10)]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    log(f"rank {rank}: lattices built by the reference host pipeline in {time.time() - t0:.1f}s")
    done.write_text("ok")
    return info


def harness_model_args(args, wl, rnn, log2w=None) -> list[str]:
    a = [f"beam={wl['beam']}", f"gbeam={wl['gbeam']}", f"rcheck={wl['rcheck']}", f"rbeam={wl['rbeam']}",
         f"wseed={args.wseed}", "wscale=0.5"]
    if log2w is None:
        a.append(f"log2w={args.log2w}")
    if wl["rnn"]:
        a.append(f"rnn={rnn}")
    return a


def load_staged(info, args, pin) -> tuple[abi.ModelSpec, abi.ScoringConfig, list[abi.StagedBatch], abi.EntryTable | None]:
    """The staged lattices as a list of windows (one for the default 100 000 sentences; --window cuts larger
    workloads so that one window's lattices + results fit the device comfortably), in the batch form asked
    for: compact (one row per distinct entry of the window), resident (dictionary rows in a device table,
    SURVEY 8f1) or inline."""
    files = sorted(info["dir"].glob("staged.jppb*"))
    files = [f for f in files if not f.name.endswith(".done")]
    first = jppb.load(files[0])
    model = abi.ModelSpec.from_jppb(first, str(info["rnn"]) if info["rnn"] else None)
    model.log2_weights, model.weight_seed, model.weight_scale = args.log2w, args.wseed, 0.5
    cfg = abi.ScoringConfig.from_jppb(first)
    per_file = max(1, args.stage_chunk)
    files_per_window = max(1, (args.window + per_file - 1) // per_file)
    windows, tables = [], []
    for w0 in range(0, len(files), files_per_window):
        parts = [abi.StagedBatch.from_jppb(jppb.load(f)) for f in files[w0:w0 + files_per_window]]
        form = args.batch_form
        if form == "inline":
            batch = abi.StagedBatch.concat(parts)
        elif form == "resident":
            batch, table = abi.StagedBatch.concat(parts).resident_form()
            tables.append(table)
        else:  # one entry row per distinct entry of the window (jpp_b200_batch.node_entry_row / entry_rows)
            batch = abi.StagedBatch.concat([p.compact_entries() for p in parts]).compact_entries()
        batch = dataclasses.replace(batch, **{f.name: pin(getattr(batch, f.name)) for f in dataclasses.fields(batch)
                                              if getattr(batch, f.name) is not None})
        windows.append(batch)
    table = None
    if tables:   # the union over the windows: uploaded once, before any timed region
        ep = np.concatenate([t.entry_ptr for t in tables])
        uniq, firsti = np.unique(ep, return_index=True)
        rows = np.concatenate([t.rows for t in tables])[firsti]
        rid = None if tables[0].rnn_id is None else np.concatenate([t.rnn_id for t in tables])[firsti]
        table = abi.EntryTable(np.ascontiguousarray(uniq), np.ascontiguousarray(rows), None if rid is None else np.ascontiguousarray(rid))
    return model, cfg, windows, table


# -------------------------------------------------------------------- clocks --
class ClockSampler:
    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------- reference arm --
def cpu_budget() -> dict:
    """Host cores this process may really use: the scheduler affinity mask, cut by the cgroup CPU
    quota (a 1-GPU lease of a 128-thread host is typically a slice of it; os.cpu_count() is not)."""
    import math
    aff = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    quota = None
    try:
        q, per = Path("/sys/fs/cgroup/cpu.max").read_text().split()[:2]          # cgroup v2
        if q != "max":
            quota = float(q) / float(per)
    except Exception:
        try:                                                                      # cgroup v1
            q = float(Path("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read_text())
            per = float(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
            if q > 0:
                quota = q / per
        except Exception:
            pass
    eff = aff if quota is None else max(1, min(aff, int(math.ceil(quota))))
    model = ""
    try:
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    try:
        load = os.getloadavg()[0]
    except OSError:
        load = None
    return {"cores_effective": eff, "affinity": aff, "cgroup_cpu_quota": quota, "os_cpu_count": os.cpu_count(),
            "loadavg_1m": load, "cpu_model": model}


def reference_runs(info, args, steps: int = 2, warmup: int = 1) -> dict:
    """The reference's computeScores on the host: at cores_effective threads and at twice that
    (the path is gather-latency bound, SMT / oversubscription can pay), the better one is the
    baseline; one thread on a smaller sample gives the scaling."""
    bud = cpu_budget()
    eff = bud["cores_effective"]
    runs = []
    for threads in (eff, 2 * eff):
        sample = min(args.sentences, args.cpu_sample_per_core * threads)
        o = run_harness_bench(info, args, sample, threads, steps=steps, warmup=warmup)
        runs.append({"threads": threads, "sentences": o["sentences"], "value": o["score_path_sentences_per_s"],
                     "score_path_s_max_thread": o["score_path_s_max_thread"], "wall_s": o["wall_s"],
                     "analyze_value": o["analyze_sentences_per_s"]})
    best = max(runs, key=lambda r: r["value"])
    o1 = run_harness_bench(info, args, min(args.sentences, args.cpu_sample_per_core), 1, steps=1, warmup=1)
    single = o1["score_path_sentences_per_s"]
    return {"budget": bud, "runs": runs, "best": best, "single_thread_value": single,
            "speedup_over_1thread": best["value"] / single if single else None}


def cpu_baseline_record(rr: dict) -> dict:
    best, bud = rr["best"], rr["budget"]
    return {"value": best["value"], "unit": "sentences/s", "cores": bud["cores_effective"], "threads": best["threads"],
            "kind": "reference", "cores_effective": bud["cores_effective"], "affinity": bud["affinity"],
            "cgroup_cpu_quota": bud["cgroup_cpu_quota"], "os_cpu_count": bud["os_cpu_count"],
            "loadavg_1m_before": bud["loadavg_1m"], "cpu_model": bud["cpu_model"],
            "single_thread_value": rr["single_thread_value"], "speedup_over_1thread": rr["speedup_over_1thread"],
            "runs": [{"threads": r["threads"], "value": r["value"], "sentences": r["sentences"]} for r in rr["runs"]],
            "build": "unmodified reference sources, g++ -O3 -march=x86-64-v3 (AVX2+FMA; the binary is built where the "
                     "sources are mounted and must run on this host), Eigen stand-in with AVX2 dot products",
            "sample": f"first {best['sentences']} sentences of the corpus, {best['threads']} threads each owning an "
                      f"AnalyzerImpl over one CoreHolder; timed region = AnalyzerImpl::computeScores only "
                      f"({best['score_path_s_max_thread']:.2f}s, lattice build excluded); whole Analyzer::analyze: "
                      f"{best['analyze_value']:.0f} sentences/s"}


def run_harness_bench(info, args, sample: int, threads: int, steps: int = 1, warmup: int = 0) -> dict:
    cmd = [str(HARNESS), "bench", f"dict={info['dict']}", f"corpus={info['corpus']}", f"max_sent={sample}",
           f"threads={threads}", f"steps={steps}", f"warmup={warmup}", *harness_model_args(args, info["wl"], info["rnn"])]
    r = subprocess.run(cmd, check=True, capture_output=True, text=True)
    return json.loads(r.stdout.strip().splitlines()[-1])


def reference_arm(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    info = prepare_workload(args, 0, 1)
    # one process per thread count: model loaded once, W untimed + K timed passes over the sample
    rr = reference_runs(info, args, steps=args.steps, warmup=max(1, args.warmup))
    best = rr["best"]
    value = best["value"]
    line = {
        "impl": "reference", "metric": "sentences/sec Jumandic beam=5 (lattice-scoring path)", "value": value,
        "unit": "sentences/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * best["score_path_s_max_thread"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 hash / f32 sums", "data": "synthetic",
        "config": workload_config(args, info["wl"], 1),
        "cpu_baseline": cpu_baseline_record(rr),
        "e2e": {"value": value, "unit": "sentences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_result(line)


RESULT_OUT = None
FAILED: list = []


def emit_result(line: dict) -> None:
    out = RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def walk_kernel_name(wl) -> str:
    """The dominant kernel of the step as the library will run it for this configuration
    (jpp_b200.cu fused_walk_ok / jpp_b200_walk.cuh)."""
    if wl["rcheck"] > 0 and wl["gbeam"] > 0 and "JPP_B200_WAVEFRONT" not in os.environ:
        narrow = wl["beam"] <= 8 and wl["gbeam"] <= 8
        static = "JPP_B200_NO_STATIC" not in os.environ
        return (f"k_walk_fused<GW={8 if narrow else 32}>: a block of 256 threads walks 32 sentences through all their "
                "boundaries (global beam, prescore, cutoff, n-gram items, beam rows: n-gram hash + weight gather + beams), "
                + ("compiled-in Jumandic n-gram program, pattern rows in registers" if static else "table-driven n-gram program"))
    return ("wave loop: k_w_gbeam_warp + k_w_prescore + k_w_cutoff_group + k_w_items + k_w_final per boundary number "
            "(n-gram hash + weight gather + beams)")


def workload_config(args, wl, world: int = 1) -> dict:
    c = {"workload": wl["desc"], "sentences_per_gpu": args.sentences, "length_codepoints": [args.min_len, args.max_len],
         "dictionary_rows": args.dict_rows, "weights": f"float[2^{args.log2w}] counter-based synthetic",
         "beam": wl["beam"], "global_beam": wl["gbeam"], "right_check": wl["rcheck"], "right_beam": wl["rbeam"],
         "rnn_layer": wl["H"], "batch_form": args.batch_form, "window_sentences": min(args.window, args.sentences),
         "l2_policy": "inputs of one step (several GB of staged lattices + 256 MiB table) exceed the 126 MB L2"}
    if args.shard_corpus and world > 1:
        c["sentences_per_gpu"] = None
        c["sentences_total"] = args.sentences
        c["sharding"] = f"ONE corpus of {args.sentences} sentences cut into {world} contiguous shards balanced by codepoints (jumanpp_b200/shard.py)"
    return c


def bind_to_gpu_numa(device: int) -> dict:
    """Keeps this rank's threads (and so its first-touched and pinned host pages) on the NUMA node the GPU
    hangs off: eight ranks uploading ~1 GB each per step across the socket link is what made e2e creep at
    N = 8.  Returns what was done; never fails the run."""
    import torch
    info = {"bound": False}
    try:
        p = torch.cuda.get_device_properties(device)
        addr = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{addr}/numa_node").read_text())
        info.update(pci=addr, node=node)
        if node < 0:
            return info
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        mine = os.sched_getaffinity(0) & cpus
        if mine:
            os.sched_setaffinity(0, mine)
            info.update(bound=True, cpus=len(mine))
    except Exception as e:
        info["error"] = repr(e)
    return info


# --------------------------------------------------------------- product arm --
def product_arm(args, rank: int, world: int, local_rank: int) -> None:
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the scored path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from jumanpp_b200.scorer import LatticeScorer, pinned_copy

    info = prepare_workload(args, rank, world, dist)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa(local_rank) if not args.no_numa_bind else {"bound": False}
    log(f"rank {rank}: NUMA {numa}")
    t0 = time.time()
    model, cfg, windows, table = load_staged(info, args, pin=pinned_copy)
    n_sent = sum(w.num_sentences for w in windows)
    n_nodes = sum(w.num_nodes for w in windows)
    log(f"rank {rank}: staged lattices in pinned memory: {n_sent} sentences, {n_nodes} nodes, {len(windows)} window(s), "
        f"form {args.batch_form} ({time.time() - t0:.1f}s)")
    scorer = LatticeScorer(model, cfg, device=local_rank)
    if table is not None:
        scorer.add_entry_rows(table)   # the pre-decoded dictionary: once, like the weights
    results = []
    for w in windows:
        r = scorer.allocate_result(w)
        r.best_path_len, r.best_path_nodes = pinned_copy(r.best_path_len), pinned_copy(r.best_path_nodes)
        r.sentence_status = pinned_copy(r.sentence_status)
        results.append(r)
    h2d = sum(w.h2d_bytes() for w in windows)
    d2h = sum(r.best_path_len.nbytes + r.best_path_nodes.nbytes + r.sentence_status.nbytes for r in results)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`): lattices of a window already in HBM when its timed region
    # starts; the step's time = CUDA events around all kernels, summed over the windows
    scorer.upload(windows[0])
    for _ in range(args.warmup):
        scorer.score_resident()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    kms, counters = [], None
    one_lane_paths = None
    for step in range(args.steps):
        km, cn = None, None
        for wi, w in enumerate(windows):
            if len(windows) > 1 or step == 0:
                if not (wi == 0 and step == 0):
                    scorer.upload(w)             # untimed (outside the events of score_resident)
            scorer.score_resident()              # synchronises its own stream (CUDA events inside)
            c = scorer.counters()
            km = c["kernel_ms"] if km is None else {k: km[k] + v for k, v in c["kernel_ms"].items()}
            cn = c if cn is None else {k: (cn[k] + v if isinstance(v, int) else v) for k, v in c.items()}
            if step == 0 and wi == 0:
                scorer.download(results[0], paths_only=True)   # 1-best paths of the resident (one-lane) run
                one_lane_paths = (results[0].best_path_len.copy(), results[0].best_path_nodes.copy())
        kms.append(km)
        counters = cn
    torch.cuda.synchronize()
    dt_res = time.perf_counter() - t0
    clocks = sampler.stop()
    cnt = counters
    dev_ms = [k["score_total"] for k in kms]
    t_res = torch.tensor([sum(dev_ms) / 1000.0, dt_res], dtype=torch.float64, device="cuda")
    # ---- end to end through the C ABI with host buffers (`e2e`): staged lattices in pinned host memory
    # -> H2D -> all kernels -> D2H of the device-resolved 1-best paths + per-sentence status (what the
    # CLI's default output consumes), one jpp_b200_score_batch call per window
    for _ in range(min(2, args.warmup)):
        scorer.compute_scores(windows[0], results[0], paths_only=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for w, r in zip(windows, results):
            scorer.compute_scores(w, r, paths_only=True)       # H2D + kernels + D2H
    torch.cuda.synchronize()
    dt_e2e = time.perf_counter() - t0
    lanes_agree = bool(np.array_equal(one_lane_paths[0], results[0].best_path_len) and
                       np.array_equal(one_lane_paths[1], results[0].best_path_nodes))
    failed = int(sum(int(r.sentence_status.astype(bool).sum()) for r in results))
    t_e2e = torch.tensor([dt_e2e], dtype=torch.float64, device="cuda")
    counts = torch.tensor([n_sent, cnt["gathers"], n_nodes, failed], dtype=torch.float64, device="cuda")
    gathered = None
    per_gpu = None
    if dist:
        mine = torch.tensor([n_sent, float(t_res[0].item()), dt_e2e], dtype=torch.float64, device="cuda")
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        per_gpu = [{"rank": r, "sentences": int(x[0].item()), "sentences_per_s_resident": x[0].item() * args.steps / x[1].item(),
                    "sentences_per_s_e2e": x[0].item() * args.steps / x[2].item()} for r, x in enumerate(every)]
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)   # NCCL: gather the per-rank counters
        if args.shard_corpus:
            lens, nodes = shard.relative_paths(windows, results)
            gathered = shard.gather_in_corpus_order(lens, nodes, info["shard"]["lo"], info["shard"]["hi"], info["shard"]["total"],
                                                    dist, torch.device("cuda", local_rank))
    total_sent = float(counts[0].item())
    dev_s, wall_s, e2e_s = float(t_res[0].item()), float(t_res[1].item()), float(t_e2e[0].item())
    value = total_sent * args.steps / dev_s
    e2e_value = total_sent * args.steps / e2e_s

    if rank == 0:
        walk_ms = statistics.mean(k["beam_walk"] for k in kms)
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = cnt["walk_bytes"] / (walk_ms / 1000.0) / 1e9
        traffic = None
        tf = ROOT / "profiles" / "walk_traffic.json"
        if tf.exists():
            try:
                t = json.loads(tf.read_text()).get(args.workload)
                if t and t.get("sentences") == args.sentences and t.get("kernel", "") in walk_kernel_name(info["wl"]):
                    traffic = t["dram_bytes_per_launch"]
            except Exception:
                pass
        # the ceiling of the access pattern itself on this device: independent random 4-byte reads of the same
        # table (jpp_b200_gather_probe, after the timed region), best of three in-flight depths
        probe = None
        if not args.no_gather_probe:
            try:
                rates = {ilp: scorer.gather_probe(1 << 31, ilp, 8) for ilp in (4, 8, 16)}
                best = max(rates.values())
                issued = float(cnt["walk_gathers_issued"])
                probe = {"peak_gathers_per_s": best, "by_ilp": rates,
                         "achieved_gathers_per_s": issued / (walk_ms / 1000.0),
                         "frac": issued / (walk_ms / 1000.0) / best,
                         "what": "table reads the walk issues per second / random 4-byte reads per second the device serves "
                                 "from the same table with nothing else to do (8 blocks of 256 threads per SM)"}
            except Exception as e:   # a diagnostic never fails the bench line
                probe = {"error": str(e)}
        line = {
            "metric": "sentences/sec Jumandic beam=5 (lattice-scoring path)", "value": value, "unit": "sentences/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dev_s / args.steps,
            "higher_is_better": True, "scaling": "strong" if (args.shard_corpus and world > 1) else "weak",
            "vs_baseline": None, "dtype": "u64 hash / f32 sums",
            "data": "synthetic", "config": workload_config(args, info["wl"], world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "sentences/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1000.0 * e2e_s / args.steps,
                    "how": f"one jpp_b200_score_batch call per window ({len(windows)} per step) from pinned host buffers, six internal "
                           f"lanes (JPP_B200_LANES); batch form '{args.batch_form}'"
                           + (f" with {len(table.entry_ptr)} dictionary rows resident on the device" if table is not None else "")
                           + "; D2H = device-resolved 1-best paths + per-sentence status",
                    "paths_equal_to_one_lane_run": lanes_agree},
            "gpu_launches": int(cnt["launches"]) * args.steps,
            "wall_ms_per_step_resident": 1000.0 * wall_s / args.steps,
            "stage_ms": {k: statistics.mean(x[k] for x in kms) for k in kms[0]},
            "roofline": {"bound": "hbm", "kernel": walk_kernel_name(info["wl"]),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "algorithmic_bytes_per_launch": int(cnt["walk_bytes"]), "kernel_ms": walk_ms,
                         "gathers_per_launch": int(cnt["walk_gathers"]),
                         "gathers_issued_per_launch": int(cnt["walk_gathers_issued"]), "traffic": traffic,
                         "per": "step (all windows)"},
            "gather_roofline": probe,
            "work": {"sentences": int(total_sent), "nodes": int(counts[2].item()), "gathers": int(counts[1].item()),
                     "rnn_nodes": int(cnt["rnn_nodes"]), "rnn_scores": int(cnt["rnn_scores"]),
                     "failed_sentences": int(counts[3].item())},
        }
        line["numa"] = numa
        if per_gpu is not None:
            line["per_gpu"] = per_gpu
        if gathered is not None:
            line["sharded"] = {k: v for k, v in gathered.items() if k not in ("lens", "nodes")}
        os.sched_setaffinity(0, all_cpus)   # the CPU legs below use every core this process was given
        if not args.no_cpu_baseline:
            line.update(cpu_baseline_leg(args, info, model, cfg, windows[0], scorer, results[0]))
        if args.e2e_full and world == 1:
            line["e2e_full"] = e2e_full_leg(args, info, local_rank, line.get("cpu_baseline"))
        # a run whose own correctness checks fail is not a result: say so and exit non-zero
        problems = []
        if not lanes_agree:
            problems.append("1-best paths of the multi-lane call differ from the one-lane resident run")
        if counts[3].item():
            problems.append(f"{int(counts[3].item())} sentences failed validation on the device")
        par = line.get("parity")
        if par is not None:
            if "error" in par:
                problems.append("parity check failed to run: " + par["error"])
            else:
                problems += [f"parity: {k} = {v}" for k, v in par.items() if k.endswith("_mismatches") and v]
        if gathered is not None and not gathered["in_corpus_order"]:
            problems.append("gathered shards do not cover the corpus in order")
        line["valid"] = not problems
        if problems:
            line["invalid_because"] = problems
        emit_result(line)
        if problems:
            log("INVALID RESULT:", "; ".join(problems))
            FAILED.append(1)
    scorer.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


GPU_CLI = ROOT / "binding" / "_build" / "jumanpp_v2_gpu"


def e2e_full_leg(args, info, device: int, cpu_baseline: dict | None) -> dict:
    """SURVEY 8(d)'s metric through the reference-side binding: text in -> lattices built on host threads ->
    staged -> jpp_b200_score_batch -> un-staged into the reference's Lattice -> formatted by the reference's
    JumanFormat -> text out (binding/jumanpp_v2_gpu), beside the reference doing the same work per sentence
    (Analyzer::analyze + JumanFormat) on the same host cores."""
    out = {}
    try:
        wl = info["wl"]
        modelf = info["dir"] / "cli.model"
        if not modelf.exists():
            cmd = [str(HARNESS), "model", f"dict={info['dict']}", f"out={modelf}", *harness_model_args(args, wl, info["rnn"])]
            subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        n = min(args.sentences, args.e2e_full_sentences)
        sub = info["dir"] / f"corpus_{n}.txt"
        if not sub.exists():
            with open(info["corpus"], encoding="utf-8") as f, open(sub, "w", encoding="utf-8") as g:
                for i, l in enumerate(f):
                    if i >= n:
                        break
                    g.write(l)
        bud = cpu_budget()
        flags = [f"--beam={wl['beam']}", f"--global-beam={wl['gbeam']}", f"--right-check={wl['rcheck']}", f"--right-beam={wl['rbeam']}"]
        cmd = [str(GPU_CLI), "--gpu-stats", f"--gpu-device={device}", f"--gpu-batch={args.e2e_full_batch}",
               f"--gpu-threads={bud['cores_effective']}", f"--model={modelf}", *flags, "--output=/dev/null", str(sub)]
        r = subprocess.run(cmd, capture_output=True, text=True, check=True)
        import re
        st = json.loads(re.findall(r'\{"gpu_stats".*\}\}', r.stderr)[-1])["gpu_stats"]
        out = {"value": st["sentences"] / st["total_s"], "unit": "sentences/s", "sentences": st["sentences"],
               "host_threads": st["threads"], "batch_sentences": st["batch"], "stages_s": {k: st[k] for k in st if k.endswith("_s")},
               "h2d_bytes": st["h2d_bytes"], "d2h_bytes": st["d2h_bytes"], "entry_rows_added": st["entry_rows_added"],
               "compute_scores_calls": st["compute_scores_calls"], "failed": st["failed"],
               "how": "binding/_build/jumanpp_v2_gpu: text -> buildLattice on host threads -> resident-form staging in pinned "
                      "buffers -> jpp_b200_score_batch -> un-stage into the reference Lattice -> JumanFormat -> /dev/null; "
                      "total_s excludes model load and analyzer construction (init_s)"}
        # the reference doing the same per sentence: Analyzer::analyze + JumanFormat, all effective cores
        best = None
        for threads in (bud["cores_effective"], 2 * bud["cores_effective"]):
            cmdr = [str(HARNESS), "bench", f"dict={info['dict']}", f"corpus={sub}", f"threads={threads}", "steps=1", "warmup=1",
                    "format=1", f"max_sent={min(n, args.cpu_sample_per_core * threads)}", *harness_model_args(args, wl, info["rnn"])]
            o = json.loads(subprocess.run(cmdr, check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1])
            if best is None or o["analyze_sentences_per_s"] > best["value"]:
                best = {"value": o["analyze_sentences_per_s"], "threads": threads, "sentences": o["sentences"]}
        out["reference"] = {**best, "unit": "sentences/s", "cores_effective": bud["cores_effective"],
                            "how": "oracle/_ref/jpp_ref_harness bench format=1: Analyzer::analyze + JumanFormat per sentence"}
        out["ratio"] = out["value"] / best["value"]
    except Exception as e:  # pragma: no cover
        out = {"error": repr(e)[:2000]}
    return out


def cpu_baseline_leg(args, info, model, cfg, batch, scorer, res) -> dict:
    """Rank 0 only: (1) time the reference's computeScores on the host cores over a
    bounded sample; (2) check the GPU beams of a bounded sample against the
    reference's own post-state (1-best + every beam element)."""
    out = {}
    try:
        out["cpu_baseline"] = cpu_baseline_record(reference_runs(info, args))
    except Exception as e:  # pragma: no cover
        out["cpu_baseline"] = {"error": repr(e)}
    # parity sample: reference post-state for the first sentences of the same corpus
    try:
        n = min(args.parity_sample, batch.num_sentences)
        exp_file = info["dir"] / f"parity_{n}.jppb"
        if not exp_file.exists():
            cmd = [str(HARNESS), "stage", f"dict={info['dict']}", f"corpus={info['corpus']}", f"out={exp_file}",
                   "expected=1", f"max_sent={n}", *harness_model_args(args, info["wl"], info["rnn"])]
            subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        d = jppb.load(exp_file)
        sub = batch.select(0, n)
        got = scorer.compute_scores(sub)
        exp_beam = np.zeros(got.node_beam.shape, abi.BEAM_EL_DT)
        exp_beam["left"], exp_beam["prev_beam"] = d["exp_beam_lp"][:, :, 0], d["exp_beam_lp"][:, :, 1]
        exp_beam["total_score"] = d["exp_beam_total"]
        exp_res = abi.ScoreResult(exp_beam, d["exp_beam_scores"], d["exp_patterns"], got.node_t0_score)
        pg, pe = abi.best_paths(sub, got), abi.best_paths(sub, exp_res)
        mism = sum(1 for a, b in zip(pg, pe) if a != b)
        links = int(((exp_beam["left"] != got.node_beam["left"]) | (exp_beam["prev_beam"] != got.node_beam["prev_beam"])).sum())
        out["parity"] = {"checked_sentences": n, "onebest_mismatches": mism, "beam_link_mismatches": links,
                         "pattern_mismatches": int((d["exp_patterns"] != got.node_patterns).sum()),
                         "checker": "reference post-state (oracle/_ref) on the same lattices"}
    except Exception as e:  # pragma: no cover
        out["parity"] = {"error": repr(e)}
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="jumandic_beam5", choices=sorted(WORKLOADS))
    ap.add_argument("--sentences", type=int, default=100_000)
    ap.add_argument("--min-len", type=int, default=8)
    ap.add_argument("--max-len", type=int, default=128)
    ap.add_argument("--dict-rows", type=int, default=200_000)
    ap.add_argument("--log2w", type=int, default=26)
    ap.add_argument("--wseed", type=int, default=1)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--stage-chunk", type=int, default=2048)
    ap.add_argument("--cache", default=os.environ.get("JPP_B200_CACHE", "/tmp/jpp_b200_cache"))
    ap.add_argument("--cpu-sample-per-core", type=int, default=1500)
    ap.add_argument("--parity-sample", type=int, default=10000)
    ap.add_argument("--no-gather-probe", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true")
    ap.add_argument("--window", type=int, default=131072,
                    help="sentences per jpp_b200_score_batch call (larger workloads are streamed window by window)")
    ap.add_argument("--shard-corpus", action="store_true",
                    help="N>1: ONE corpus of --sentences cut into contiguous shards (strong scaling, BASELINE configs[4]) "
                         "instead of one corpus per rank")
    ap.add_argument("--batch-form", default="compact", choices=["compact", "resident", "inline"])
    ap.add_argument("--e2e-full", action="store_true", help="also run the reference-side binding end to end (text in, text out)")
    ap.add_argument("--e2e-full-sentences", type=int, default=100_000)
    ap.add_argument("--e2e-full-batch", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    # banner to stdout at init): keep the real stdout aside for the result and point fd 1 at stderr.
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3 and args.impl == "b200":
        log("warning: fewer than 3 warm-up steps")
    if args.impl == "reference":
        reference_arm(args, rank, world)
    else:
        product_arm(args, rank, world, local_rank)
    if FAILED:
        sys.exit(1)


if __name__ == "__main__":
    main()
