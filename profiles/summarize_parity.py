"""profiles/r02_parity_report.txt from gpurun_out/parity_report.jsonl (written by tests/common.py::assert_close
and tests/test_gpu_fullsize.py during `pytest -m gpu`):  python profiles/summarize_parity.py > profiles/r02_parity_report.txt"""
import collections
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")
rows = [json.loads(l) for l in open(path)]
cmp_rows = [r for r in rows if "tol_over_scale" in r]
print(f"# {len(cmp_rows)} tensor comparisons of one `pytest tests -m gpu` run on a B200 (GPU through the C ABI vs oracle/refport.py, fp32)")
print("# err = max|got - oracle32| / max|oracle32|;  bar = the tolerance applied (1e-5, + the oracle's own distance to its fp64 twin capped at 1e-5;")
print("# looser only where the test says so: optimizer chains 1e-3, encoder / replay tests 2e-5);  rel = element-wise |got - want| / |want| over the")
print("# entries with |want| >= 1e-3 max|want|;  rms_gpu / rms_o32 = rms error against the fp64 value, GPU vs the fp32 oracle itself")
print()
groups = collections.OrderedDict()
for r in cmp_rows:
    key = re.sub(r"\s+step \d+", "", r["what"]) or "(unnamed)"
    groups.setdefault(key, []).append(r)
print(f"{'comparison':58s} {'n':>4s} {'worst err':>10s} {'bar':>9s} {'rel p50':>9s} {'rel p99':>9s} {'rel max':>9s} {'rms_gpu/rms_o32':>16s}")
for key, rs in groups.items():
    worst = max(rs, key=lambda r: r["max_err_over_scale"])
    p50 = max((r.get("rel_p50", 0) for r in rs))
    p99 = max((r.get("rel_p99", 0) for r in rs))
    pm = max((r.get("rel_max", 0) for r in rs))
    ratio = [r["rms_err_gpu_vs_exact"] / r["rms_err_oracle32_vs_exact"] for r in rs
             if r.get("rms_err_oracle32_vs_exact", 0) > 0]
    rt = f"{max(ratio):.2f}" if ratio else "-"
    print(f"{key[:58]:58s} {len(rs):4d} {worst['max_err_over_scale']:10.2e} {worst['tol_over_scale']:9.1e} {p50:9.1e} {p99:9.1e} {pm:9.1e} {rt:>16s}")
print()
print("# per-sample comparisons at the benchmark batch and on unfiltered batches (tests/test_gpu_fullsize.py)")
for r in rows:
    if "tol_over_scale" not in r:
        print("  " + json.dumps(r))
