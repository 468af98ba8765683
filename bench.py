#!/usr/bin/env python
"""Headline benchmark: train samples/sec (fwd+bwd, gather and scatter-add included, optimizer
excluded) for DeepFM / xDeepFM on synthetic Criteo-shaped data (BASELINE.json), plus the roofline
of the dominant kernel and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model both|deepfm|xdeepfm|fm|dcn|pnn|lr]
    python bench.py --impl reference ...        # the CPU baseline of the reference path

Default (--model both): the line's value / e2e / roofline / cpu_baseline are DeepFM (BASELINE configs[1] at
N = 1, configs[4] -- the 100 M-row row-sharded table -- at N > 1) and the key "xdeepfm" holds the same
record for xDeepFM (configs[2]) measured in the same invocation: BASELINE.json's metric names both.

One "step" = ParRecModel.optimize for one batch (rec/model/ParRecModel.scala:439-478) without the PS
RPC: lookup (gather) -> forward -> backward -> per-id gradient scatter-add.
  value : whole-job samples/s with the ids already resident in HBM (b200rec_step_dev)
  e2e   : the same step through the host-facing C-ABI call b200rec_step -- ids and labels copied
          from pinned host memory every step, the loss read back every step
N > 1 (torchrun, one rank per GPU): row-sharded table, NCCL all-to-all for lookups and gradients,
NCCL allreduce for the dense gradients; weak scaling (per-GPU batch fixed).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODELS = {
    # name: (kind, fc, cin, depth)   -- BASELINE.json configs[1..3]
    "deepfm": ("deepfm", [400, 400, 400], [], 0),
    "xdeepfm": ("xdeepfm", [400, 400, 400], [200, 200, 200], 0),
    "fm": ("fm", [], [], 0),
    "lr": ("lr", [], [], 0),
    "dcn": ("dcn", [400, 400, 400], [], 6),
    "pnn": ("pnn", [400, 400, 400], [], 0),
}
F, K = 39, 16
SEED_DATA, SEED_PARAMS = 1234, 42


def gather_ceiling_us(n_rows):
    """Measured time of a pure random-row copy of n_rows rows (scripts/ubench/gather_rate.cu), or None."""
    path = os.path.join(ROOT, "profiles", "r02x_gather_rate_ubench.txt")
    if not os.path.exists(path):
        return None
    power_law = False
    for line in open(path):
        if line.startswith("== ids:"):
            power_law = "power-law" in line
        if power_law and line.startswith("copy U=5  rows + separate weights") and f"n={n_rows:8d}" in line:
            try:
                return float(line.split("avg")[1].split("us")[0])
            except (IndexError, ValueError):
                return None
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm),
                    reasons=sorted(reasons))


def make_batches(synth, n, B, rows, first_step=0):
    out = []
    for s in range(n):
        _, feats = synth.make_feats(SEED_DATA, first_step + s, B, F, rows)
        out.append((feats, synth.make_targets(SEED_DATA, feats, B, F)))
    return out


# ------------------------------------------------------------------------------------------------
# algorithmic work per launch (SURVEY.md 8d / DESIGN.md) for the roofline lines
# ------------------------------------------------------------------------------------------------
def algorithmic_work(kind, fc, cin, depth, B, U):
    """-> {phase: (bound, amount per step, unit)} ; bytes for HBM-bound phases, flops for GEMMs."""
    N = B * F
    D = F * K
    w = {}
    has_dense = kind in ("deepfm", "xdeepfm", "dcn", "pnn")
    second = kind in ("fm", "deepfm")
    if kind != "lr":
        # ids 4 + row 64 + weight 4 (+ row written once for the dense branch / backward) ; per sample out
        w["gather_fm_fwd"] = ("hbm", N * (4 + 4 * K + 4 + 4 * K) + B * (8 + (4 * K if second else 0)), "B")
        # rows + dense dX in, per-nnz grads out, dw out
        w["emb_grad"] = ("hbm", N * (4 * K * (3 if has_dense else 2) + 4) + B * (4 + 4 * K), "B")
        w["scatter_add"] = ("hbm", N * (4 + 4 * K + 4) + U * (4 * K + 8), "B")
    mlp_in = {"deepfm": D, "xdeepfm": D, "dcn": D, "pnn": None}.get(kind)
    flops = 0
    if kind == "pnn":
        P = F * (F - 1) // 2
        flops += 2 * (D + P) * fc[0]
        d = fc[0]
        for o in fc[1:] + [1]:
            flops += 2 * d * o
            d = o
    elif has_dense:
        d = mlp_in
        for o in fc + ([1] if kind == "deepfm" else []):
            flops += 2 * d * o
            d = o
    if kind == "xdeepfm":
        h = F
        for c in cin:
            flops += 2 * K * F * h * c
            h = c
    if flops:
        w["dense_fwd"] = ("tensor", flops * B, "FLOP")
        w["dense_bwd"] = ("tensor", 2 * flops * B, "FLOP")
    return w


def default_rows(args):
    """BASELINE configs[1] (N = 1): 39 * 2^18 ids; configs[4] (row-sharded, N > 1): 100 M ids."""
    if args.rows:
        return args.rows
    return 39 * (1 << 18) if args.gpus == 1 else 100_000_000


def model_list(args):
    return ["deepfm", "xdeepfm"] if args.model == "both" else [args.model]


def workload(name, args, rows):
    """config.workload -- the SAME string in the B200 arm and the reference arm."""
    kind, fc, cin, depth = MODELS[name]
    base = f"{name} k={K} F={F} fc={fc} cin={cin} depth={depth}"
    if args.gpus == 1:
        return f"{base} batch={args.batch} table_rows={rows}"
    return f"{base} per-GPU batch={args.batch} table_rows={rows} row-sharded over {args.gpus} GPUs"


def run_b200(args):
    import torch
    import __graft_entry__ as g
    pkg = g.load_package()
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1 or args.force_sharded:
        from recommendation_models_b200 import sharded  # noqa: F401  (registered by load_package)
        return sharded.bench(args, pkg)
    names = model_list(args)
    out = measure_1gpu(args, names[0], pkg, torch, local)
    for extra in names[1:]:
        out[extra] = measure_1gpu(args, extra, pkg, torch, local)
    print(json.dumps(out))


def measure_1gpu(args, name, pkg, torch, local):
    """One model on one GPU: device-timed value, e2e through the host-facing calls, per-kernel roofline,
    the literal drop-in call with host arrays, CPU baseline."""
    L = pkg._lib
    synth = pkg.synth
    kind, fc, cin, depth = MODELS[name]
    B, rows = args.batch, default_rows(args)
    model = pkg.make_model(kind, F, K, fc, cin, depth, device=local)
    if args.gemm_mode is not None:
        model.setGemmMode(args.gemm_mode)
    if args.no_graph:
        L.check(L.lib().b200rec_model_set_graph(model.handle, 0))
    table = pkg.EmbeddingTable(rows, K if kind != "lr" else 0, device=local)
    table.init_uniform(SEED_PARAMS)
    mats = synth.init_mats(SEED_PARAMS, model.getMatsSize())
    ps = pkg.ParRecModel(model, table)
    ps.setParams(np.array([0.1], np.float32), mats)
    lib = L.lib()
    sp = C.c_void_p()
    L.check(lib.b200rec_model_stream(model.handle, C.byref(sp)))
    stream = torch.cuda.ExternalStream(sp.value)

    W, Ksteps = args.warmup, args.steps
    nb = min(W + Ksteps, 64)
    batches = make_batches(synth, nb, B, rows)
    dev = [(torch.from_numpy(f).cuda(), torch.from_numpy(t).cuda()) for f, t in batches]
    pin = [(torch.from_numpy(f).pin_memory(), torch.from_numpy(t).pin_memory()) for f, t in batches]
    torch.cuda.synchronize()

    def step_dev(i):
        f, t = dev[i % nb]
        L.check(lib.b200rec_step_dev(model.handle, table.handle, B, f.data_ptr(), t.data_ptr(), None))

    loss = C.c_float(0)

    def stage_host(i):
        f, t = pin[i % nb]
        L.check(lib.b200rec_stage_batch(model.handle, B, f.data_ptr(), t.data_ptr()))

    def step_host(i, first=False):
        # batch i was staged during the previous step: start the copy of batch i + 1, enqueue step i and
        # its loss read-back, then wait for and read the loss of step i - 1.  Every step has one H2D of a
        # batch and one D2H of a loss; the host reads each loss one step late, so the GPU never idles.
        stage_host(i + 1)
        L.check(lib.b200rec_step_staged_async(model.handle, table.handle))
        if not first:
            L.check(lib.b200rec_step_wait(model.handle, table.handle, C.byref(loss)))
        return loss.value

    # ---- device-resident timing ------------------------------------------------------------------
    for i in range(W):
        step_dev(i)
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = pkg.launch_count()
    torch.cuda.synchronize()
    e0.record(stream)
    for i in range(Ksteps):
        step_dev(W + i)
    e1.record(stream)
    torch.cuda.synchronize()
    launches = pkg.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    L.check(lib.b200rec_model_sync(model.handle))
    res = ps.stepResults()
    U = len(res["unique"])

    # ---- end to end through the host-facing call ---------------------------------------------------
    nw = min(W, 5)
    stage_host(0)
    for i in range(nw):
        step_host(i, first=(i == 0))
    L.check(lib.b200rec_step_wait(model.handle, table.handle, C.byref(loss)))   # drain: nothing in flight
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    last = 0.0
    for i in range(Ksteps):
        last = step_host(nw + i, first=(i == 0))
    L.check(lib.b200rec_step_wait(model.handle, table.handle, C.byref(loss)))   # the last step's loss
    last = loss.value
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clk = clocks.stop()

    # ---- the literal drop-in call: Internal<M>Model.backward with HOST arrays (DeepFM.scala:83-124) ---
    # index / weights / bias / embedding / mats / targets go in over PCIe, the four gradient arrays and the
    # loss come back, every call (the reference's PS-worker flow keeps all of them on the host)
    dropin = None
    if kind != "lr":
        nd = max(3, min(Ksteps, 10))
        feats0 = batches[0][0]
        index = np.repeat(np.arange(B, dtype=np.int32), F)
        emb0 = synth.table_rows(SEED_PARAMS, feats0, K).reshape(-1)
        w0 = synth.wtable_rows(SEED_PARAMS, feats0)
        tg = batches[0][1]
        mt = mats if mats.size else None
        times = []
        for i in range(nd + 2):
            a = (w0.copy(), np.array([0.1], np.float32), emb0.copy(), None if mt is None else mt.copy())
            t0 = time.perf_counter()
            model.backward(B, index, a[0], a[1], a[2], a[3], tg)
            if i >= 2:
                times.append(time.perf_counter() - t0)
        nbytes_in = index.nbytes + w0.nbytes + 4 + emb0.nbytes + mats.nbytes + tg.nbytes
        nbytes_out = w0.nbytes + 4 + emb0.nbytes + mats.nbytes + 4
        dropin = {"value": round(B / statistics.median(times), 1), "unit": "samples/s",
                  "ms_per_call": round(1e3 * statistics.median(times), 4), "calls": nd,
                  "h2d_bytes_per_step": int(nbytes_in), "d2h_bytes_per_step": int(nbytes_out),
                  "call": "b200rec_backward(batchSize, index, weights, bias, embedding, mats, targets) with pageable host "
                          "arrays in and the gradients written back over them (INTEGRATION.md section 2, row 1); rows "
                          "already gathered by the caller, as in the reference's PS-worker flow"}

    # ---- per-kernel pass for the roofline (separate from the timed region above) -------------------
    L.profile_begin()
    for i in range(Ksteps):
        step_dev(W + i)
    prof = L.profile_end()
    # an event pair costs the kernel it brackets a few microseconds (pipeline drain + relaunch); measured
    # on an empty kernel, minus ~1 us for the empty kernel itself, and subtracted per launch below so that
    # the times are kernel durations (they then agree with ncu's gpu__time_duration, profiles/*launches*)
    ev_us = max(0.0, L.profile_overhead_us(local) - 1.0)
    pk = peaks()
    work = algorithmic_work(kind, fc, cin, depth, B, U)
    by_phase = {}
    for tag, kname, cnt, ms in prof:
        d = by_phase.setdefault(tag, dict(ms=0.0, launches=0, kernels={}))
        d["ms"] += ms
        d["launches"] += cnt
        k = d["kernels"].setdefault(kname, [0, 0.0])
        k[0] += cnt
        k[1] += ms
    kernels = []
    tf32_peak = pk["bf16_sus"] / 2.0  # dense TF32 = half of dense bf16 on tcgen05
    for tag, d in sorted(by_phase.items(), key=lambda kv: -kv[1]["ms"]):
        raw_step = d["ms"] / Ksteps
        ms_step = max(raw_step - ev_us * 1e-3 * d["launches"] / Ksteps, 0.25 * raw_step)
        row = dict(phase=tag, ms_per_step=round(ms_step, 5), ms_per_step_with_event_overhead=round(raw_step, 5),
                   launches_per_step=d["launches"] / Ksteps,
                   kernels={n: round(max(v[1] - ev_us * 1e-3 * v[0], 0.25 * v[1]) / Ksteps, 5)
                            for n, v in d["kernels"].items()})
        if tag in work and ms_step > 0:
            bound, amount, unit = work[tag]
            if bound == "hbm":
                ach = amount / (ms_step * 1e-3) / 1e9
                row.update(bound="hbm", achieved=round(ach, 1), peak=pk["hbm"], unit="GB/s",
                           frac=round(ach / pk["hbm"], 4), algorithmic_bytes=amount)
                if ach > pk["hbm"]:
                    row["note"] = "above the HBM copy peak: part of the operands was still in L2 from the producing kernel"
                if tag == "gather_fm_fwd":
                    c = gather_ceiling_us(B * F)
                    if c:
                        row["pattern_ceiling"] = dict(
                            us=c, frac_of_ceiling=round(min(1.0, c * 1e-3 / ms_step), 3),
                            source="profiles/r02x_gather_rate_ubench.txt: a stand-alone kernel that only copies the same number of "
                                   "random 64-B rows (+ 4-B weights) out of a 654 MB table, power-law ids, timed with CUDA events: "
                                   "at one batch the gather is latency- / launch-bound, not HBM-bound (2.9-4.7 TB/s at 4 M rows)")
            else:
                ach = amount / (ms_step * 1e-3) / 1e12
                row.update(bound="tensor", achieved=round(ach, 2), peak=tf32_peak, unit="TFLOP/s",
                           frac=round(ach / tf32_peak, 4), algorithmic_flops=amount,
                           peak_note=f"TF32 dense = 1/2 x {pk['src']} sustained bf16 {pk['bf16_sus']}")
        kernels.append(row)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    dom = next((r for r in kernels if "bound" in r), None)
    if os.path.exists(tpath) and dom:
        traffic = json.load(open(tpath)).get(name, {}).get(dom["phase"])
    roofline = None
    if dom:
        roofline = dict(kernel=dom["phase"], bound=dom["bound"], achieved=dom["achieved"], peak=dom["peak"],
                        unit=dom["unit"], frac=dom["frac"], traffic=traffic, peak_source=pk["src"],
                        launches_per_step=dom["launches_per_step"], ms_per_step=dom["ms_per_step"],
                        event_overhead_us_per_launch=round(ev_us, 2),
                        note="kernel durations from a separate pass with a CUDA-event pair around every launch on one stream: "
                             "no dependent-launch overlap, no auxiliary-stream overlap -- the phases sum to more than the "
                             "graph-replayed step (ms_per_step of the line), which has both")

    out = {
        "metric": "train samples/sec (fwd+bwd)", "value": round(B * Ksteps / (ms_total * 1e-3), 1),
        "unit": "samples/s", "n_gpus": 1, "steps": Ksteps, "warmup": W,
        "ms_per_step": round(ms_total / Ksteps, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(name, args, rows),
                   "baseline_config": {"deepfm": "configs[1]", "xdeepfm": "configs[2]", "fm": "configs[0]"}.get(name, "configs[3]"),
                   "global_batch": B, "ids": "power-law, one per field, new batch every step",
                   "l2": f"table {rows * K * 4 / 1e6:.0f} MB > 126 MB L2; distinct ids per step; no explicit flush",
                   "parallelism": "1 GPU", "gemm_mode": args.gemm_mode, "cuda_graph": not args.no_graph, "seed_data": SEED_DATA,
                   "seed_params": SEED_PARAMS, "distinct_ids_last_step": U},
        "clocks": clk,
        "e2e": {"value": round(B * Ksteps / e2e_s, 1), "unit": "samples/s", "h2d_bytes_per_step": B * F * 4 + B * 4,
                "d2h_bytes_per_step": 32, "call": "b200rec_stage_batch (next batch, pinned host ids + labels) + b200rec_step_staged_async + b200rec_step_wait (every loss read by the host, one step late)",
                "last_loss": round(float(last), 6)},
        "e2e_dropin": dropin,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "kernels": kernels,
    }
    model.close()
    table.close()
    if not args.no_cpu:
        out["cpu_baseline"] = cpu_reference(args, name, budget_s=args.cpu_seconds)["cpu_baseline"]
    return out


# ------------------------------------------------------------------------------------------------
# The reference arm: the CPU baseline of the reference's path (oracle/cbaseline.{cpp,py}; BASELINE.md
# section 3): fastutil-style open-addressing hash maps for the gather and the scatter-add, the BigDL module graph
# pass by pass with MKL sgemm (torch-CPU), per-step copies of the dense params.  Two variants: 1
# thread (the reference's own configuration) and all usable host threads; the line's value is the
# FASTER one.  The reference itself is Scala on a JVM with un-vendored BigDL/Angel jars and cannot be
# built or run here.
# ------------------------------------------------------------------------------------------------
def cpu_reference(args, name, budget_s=12.0, steps=None, warmup=1):
    import __graft_entry__ as g
    pkg = g.load_package()
    synth = pkg.synth
    from oracle import cbaseline
    kind, fc, cin, depth = MODELS[name]
    B, rows = args.batch, default_rows(args)
    # bounded sample: full batches for the MLP-only models; CIN at B=8192 needs a 4 GB Z per layer
    # on the CPU (the reference materialises it, CINEncoder.scala:152), so it is sampled at B/16.
    Bs = B if kind != "xdeepfm" else max(64, B // 16)
    ncores = cbaseline.host_threads()
    common = (kind, F, K, fc, cin, depth, Bs, rows, synth, SEED_DATA, SEED_PARAMS)
    r_all = cbaseline.run_steps(*common, threads=ncores, budget_s=budget_s * 0.65, max_steps=steps, warmup=warmup)
    r_one = cbaseline.run_steps(*common, threads=1, budget_s=budget_s * 0.35, max_steps=None if steps is None else max(2, steps // 4),
                                warmup=1)
    best = r_all if r_all["value"] >= r_one["value"] else r_one
    def slim(r):
        return dict(value=round(r["value"], 1), ms_per_step=round(r["ms_per_step"], 3), steps=r["steps"],
                    threads=r["threads"], phases_ms=r["phases_ms"])
    return dict(value=best["value"], ms_per_step=best["ms_per_step"], steps=best["steps"], Bs=Bs,
                cpu_baseline=dict(value=round(best["value"], 1), unit="samples/s", cores=best["threads"], kind="port",
                                  sample=f"{best['steps']} steps of batch {Bs} ({name}; C++ open-addressing hash maps for "
                                         f"gather / scatter-add, BigDL module graph with {best['blas']}; the faster of the "
                                         f"1-thread and the {ncores}-thread variant)",
                                  host_threads_available=ncores,
                                  variants={"threads_1": slim(r_one), "threads_all": slim(r_all)}))


def reference_record(args, name):
    rows = default_rows(args)
    # each step is one bounded sample of the same workload; the run is bounded to ~1.5 minutes per model
    r = cpu_reference(args, name, steps=args.steps, warmup=min(args.warmup, 2), budget_s=90.0)
    return {
        "impl": "reference", "metric": "train samples/sec (fwd+bwd)", "value": round(r["value"], 1),
        "unit": "samples/s", "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup,
        "ms_per_step": round(r["ms_per_step"], 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(name, args, rows), "global_batch": args.batch * args.gpus,
                   "note": "CPU baseline of the reference path on this box's host cores (one process; a multi-GPU run is "
                           "compared with the same single host); the Scala/BigDL/Angel reference cannot be built or run "
                           "here (no JVM, un-vendored jars)"},
        "cpu_baseline": r["cpu_baseline"],
        "e2e": {"value": round(r["value"], 1), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    names = model_list(args)
    out = reference_record(args, names[0])
    for extra in names[1:]:
        out[extra] = reference_record(args, extra)
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="both", choices=["both"] + list(MODELS),
                    help="both = DeepFM as the headline record + an 'xdeepfm' sub-record (BASELINE.json's metric names both)")
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--rows", type=int, default=0, help="table rows; default 39*2^18 at N=1, 100 000 000 at N>1")
    ap.add_argument("--gemm-mode", type=int, default=None, help="0 fp32 SIMT, 1 3xTF32 tcgen05, 2 1xTF32")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--force-sharded", dest="force_sharded", action="store_true",
                    help="run the row-sharded step even on one rank (torchrun --nproc-per-node 1): the exchange "
                         "kernels then store into the rank's own buffers, which makes them profilable with ncu")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: NVLink peer-memory exchange fused into the kernels, or NCCL all-to-all")
    ap.add_argument("--no-graph", action="store_true", help="plain stream launches instead of the CUDA graph")
    ap.add_argument("--cpu-seconds", type=float, default=16.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
