#!/usr/bin/env python
"""Headline benchmark: train samples/sec (fwd+bwd, gather and scatter-add included, optimizer
excluded) for DeepFM / xDeepFM on synthetic Criteo-shaped data (BASELINE.json), plus the roofline
of the dominant kernel and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model deepfm|xdeepfm|fm|dcn|pnn|lr]
    python bench.py --impl reference ...        # the CPU restatement of the reference path

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


def run_b200(args):
    import torch
    import __graft_entry__ as g
    pkg = g.load_package()
    L = pkg._lib
    synth = pkg.synth
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1 or args.force_sharded:
        from recommendation_models_b200 import sharded  # noqa: F401  (registered by load_package)
        return sharded.bench(args, pkg)

    kind, fc, cin, depth = MODELS[args.model]
    B, rows = args.batch, args.rows
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
    for tag, name, cnt, ms in prof:
        d = by_phase.setdefault(tag, dict(ms=0.0, launches=0, kernels={}))
        d["ms"] += ms
        d["launches"] += cnt
        k = d["kernels"].setdefault(name, [0, 0.0])
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
        traffic = json.load(open(tpath)).get(args.model, {}).get(dom["phase"])
    roofline = None
    if dom:
        roofline = dict(kernel=dom["phase"], bound=dom["bound"], achieved=dom["achieved"], peak=dom["peak"],
                        unit=dom["unit"], frac=dom["frac"], traffic=traffic, peak_source=pk["src"],
                        launches_per_step=dom["launches_per_step"], ms_per_step=dom["ms_per_step"],
                        event_overhead_us_per_launch=round(ev_us, 2))

    out = {
        "metric": "train samples/sec (fwd+bwd)", "value": round(B * Ksteps / (ms_total * 1e-3), 1),
        "unit": "samples/s", "n_gpus": 1, "steps": Ksteps, "warmup": W,
        "ms_per_step": round(ms_total / Ksteps, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} k={K} F={F} fc={fc} cin={cin} depth={depth} batch={B} "
                               f"table_rows={rows} (BASELINE configs[{ {'deepfm': 1, 'xdeepfm': 2}.get(args.model, 3)}])",
                   "global_batch": B, "ids": "power-law, one per field, new batch every step",
                   "l2": f"table {rows * K * 4 / 1e6:.0f} MB > 126 MB L2; distinct ids per step; no explicit flush",
                   "parallelism": "1 GPU", "gemm_mode": args.gemm_mode, "cuda_graph": not args.no_graph, "seed_data": SEED_DATA,
                   "seed_params": SEED_PARAMS, "distinct_ids_last_step": U},
        "clocks": clk,
        "e2e": {"value": round(B * Ksteps / e2e_s, 1), "unit": "samples/s", "h2d_bytes_per_step": B * F * 4 + B * 4,
                "d2h_bytes_per_step": 32, "call": "b200rec_stage_batch (next batch, pinned host ids + labels) + b200rec_step_staged_async + b200rec_step_wait (every loss read by the host, one step late)",
                "last_loss": round(float(last), 6)},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "kernels": kernels,
    }
    if not args.no_cpu:
        out["cpu_baseline"] = cpu_reference(args, budget_s=args.cpu_seconds)["cpu_baseline"]
    print(json.dumps(out))
    model.close()
    table.close()


# ------------------------------------------------------------------------------------------------
# The reference arm: the CPU restatement of the reference's path (oracle/refport.py), all host
# threads (numpy/OpenBLAS for the sgemm BigDL would send to MKL).  The reference itself is Scala on
# a JVM with un-vendored BigDL/Angel jars and cannot be built or run here.
# ------------------------------------------------------------------------------------------------
def cpu_reference(args, budget_s=15.0, steps=None, warmup=1):
    import __graft_entry__ as g
    pkg = g.load_package()
    synth = pkg.synth
    from oracle import refport
    kind, fc, cin, depth = MODELS[args.model]
    B, rows = args.batch, args.rows
    # bounded sample: full batches for the MLP-only models; CIN at B=8192 needs a 4 GB Z per layer
    # on the CPU (the reference materialises it, CINEncoder.scala:152), so it is sampled at B/16.
    Bs = B if kind != "xdeepfm" else max(64, B // 16)
    o = refport.Model(kind, F, K, fc, cin, depth)
    mats = synth.init_mats(SEED_PARAMS, o.mats_size())
    times = []
    n_done = 0
    t_start = time.perf_counter()
    step = 0
    while True:
        index, feats = synth.make_feats(SEED_DATA, step, Bs, F, rows)
        targets = synth.make_targets(SEED_DATA, feats, Bs, F)
        ids = np.unique(feats)
        # the PS pull (rows of the distinct ids) is outside the path: untimed
        E = synth.table_rows(SEED_PARAMS, ids, K)
        wv = synth.wtable_rows(SEED_PARAMS, ids)
        t0 = time.perf_counter()
        pos = np.searchsorted(ids, feats)
        emb = refport.make_embeddings(E, pos) if kind != "lr" else None       # makeEmbeddings
        ww = refport.make_weights(wv, pos)                                     # makeWeights
        bias = np.array([0.1], np.float32)
        m = mats.copy() if mats.size else None                                 # makeMats (per-step copy)
        o.backward(Bs, index, ww, bias, emb, m, targets)                       # Internal<M>Model.backward
        if emb is not None:
            refport.make_embedding_grad(emb, feats, K)                         # makeEmbeddingGrad
        refport.make_weights_grad(ww, feats)                                   # makeWeightsGrad
        dt = time.perf_counter() - t0
        step += 1
        if step > warmup:
            times.append(dt)
            n_done += 1
        if steps is not None and n_done >= steps:
            break
        if time.perf_counter() - t_start > budget_s and n_done >= 2:
            break
    sec = sum(times)
    value = Bs * n_done / sec
    cores = os.cpu_count()
    return dict(value=value, ms_per_step=1e3 * sec / n_done, steps=n_done, Bs=Bs,
                cpu_baseline=dict(value=round(value, 1), unit="samples/s", cores=cores, kind="port",
                                  sample=f"{n_done} steps of batch {Bs} ({args.model}; numpy/OpenBLAS, all host threads)"))


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    kind, fc, cin, depth = MODELS[args.model]
    # each step is one full batch of the same workload; the run is bounded to ~2 minutes
    r = cpu_reference(args, steps=args.steps, warmup=min(args.warmup, 3), budget_s=120.0)
    out = {
        "impl": "reference", "metric": "train samples/sec (fwd+bwd)", "value": round(r["value"], 1),
        "unit": "samples/s", "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup,
        "ms_per_step": round(r["ms_per_step"], 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} k={K} F={F} fc={fc} cin={cin} depth={depth} batch={args.batch} "
                               f"table_rows={args.rows}", "global_batch": args.batch,
                   "note": "CPU restatement of the reference path (oracle port); the Scala/BigDL/Angel reference "
                           "cannot be built or run here (no JVM, un-vendored jars)"},
        "cpu_baseline": r["cpu_baseline"],
        "e2e": {"value": round(r["value"], 1), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="deepfm", choices=list(MODELS))
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--rows", type=int, default=39 * (1 << 18))
    ap.add_argument("--gemm-mode", type=int, default=None, help="0 fp32 SIMT, 1 3xTF32 tcgen05, 2 1xTF32")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--force-sharded", dest="force_sharded", action="store_true",
                    help="run the row-sharded step even on one rank (torchrun --nproc-per-node 1): the exchange "
                         "kernels then store into the rank's own buffers, which makes them profilable with ncu")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: NVLink peer-memory exchange fused into the kernels, or NCCL all-to-all")
    ap.add_argument("--no-graph", action="store_true", help="plain stream launches instead of the CUDA graph")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
