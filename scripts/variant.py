"""Build timing / A-B variants of libb200rec.so:  python scripts/variant.py NAME -DFOO [-DBAR=1 ...]  ->  libb200rec_NAME.so"""
import os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "recommendation-models_b200"))
import build as b
name, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(b.HERE, f"libb200rec_{name}.so")
r = subprocess.run([b.NVCC, *b.FLAGS, *defs, "-shared", "-o", out, *b._sources()], capture_output=True, text=True)
print(out, "rc", r.returncode, r.stderr[-400:] if r.returncode else "")
sys.exit(r.returncode)
