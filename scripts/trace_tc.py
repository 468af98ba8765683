"""Pipeline timeline of the tcgen05 GEMM (CTA 0) from the trace build: B200REC_LIB=.../libb200rec_trace.so"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
lib = pkg.lib()
rng = np.random.default_rng(0)

def dump(title, nkb, fn="b200rec_debug_tc_trace", first=0):
    buf = (C.c_longlong * (3 * 512 * 4))()
    getattr(lib, fn)(buf, 3 * 512 * 4)
    a = np.array(buf, dtype=np.int64).reshape(3, 512, 4)
    t0 = a[0, 0, 0]
    print("==", title, "nkb", nkb)
    print(" kb | MMA: start  full  bfull  issued | PROD: start  waited  stored  arrived | drained0 drained")
    for kb in range(first, min(nkb, first + 24)):
        m, p = a[0, kb] - t0, a[1, kb] - t0
        d = a[2, kb, 0] - t0 if a[2, kb, 0] else 0
        d0 = a[2, kb, 1] - t0 if a[2, kb, 1] else 0
        print(f"{kb:3d} | {m[0]:7d} {m[1]:7d} {m[2]:7d} {m[3]:7d} | {p[0]:7d} {p[1]:7d} {p[2]:7d} {p[3]:7d} | {d0:7d} {d:7d}")
    ph = a[2, 510] - t0
    if a[2, 510, 0]:
        print(f" staged epilogue: tile parked {ph[0]}, barrier passed {ph[1]}")
    e = a[2, 511] - t0
    print(f" epilogue: producers done {e[0]}, last MMA drained {e[1]}, epilogue stored {e[2]}, CTA end {e[3]}")
    m = a[0, :nkb] - t0
    print(" mean stage period (MMA issued->issued):", float(np.diff(m[:, 3]).mean()) if nkb > 1 else 0)

B, I, O = 8192, 624, 400
x = rng.standard_normal((B, I)).astype(np.float32)
w = (rng.standard_normal((O, I)) / np.sqrt(I)).astype(np.float32)
lin = pkg.Linear(I, O, False, w)
lin.updateOutput(x)
dump("linear fwd 1024x624x400", (I + 31) // 32)
gy = rng.standard_normal((B, O)).astype(np.float32)
lin.updateGradInput(x, gy)
dump("linear dx", (O + 31) // 32)
lin.accGradParameters(x, gy)
dump("linear dW (split-K)", 8)
# CIN forward through a tiny xdeepfm forward
F, K = 39, 16
m = pkg.make_model("xdeepfm", F, K, [16], [200, 200])
Bm = int(os.environ.get('TRACE_CIN_BATCH', '64'))
n = Bm * F
from oracle import refport
mats = pkg.synth.init_mats(1, refport.mats_size("xdeepfm", F, K, [16], [200, 200]))
m.forward(Bm, np.repeat(np.arange(Bm, dtype=np.int32), F), rng.standard_normal(n).astype(np.float32),
          np.zeros(1, np.float32), rng.standard_normal(n * K).astype(np.float32), mats)
dump("cin fwd layer 2 (H=200)", 39 * 7)

if os.environ.get('TRACE_CIN_BWD'):
    y = m.forward(Bm, np.repeat(np.arange(Bm, dtype=np.int32), F), rng.standard_normal(n).astype(np.float32),
                  np.zeros(1, np.float32), rng.standard_normal(n * K).astype(np.float32), mats)
    m.backward(Bm, np.repeat(np.arange(Bm, dtype=np.int32), F), rng.standard_normal(n).astype(np.float32),
               np.zeros(1, np.float32), rng.standard_normal(n * K).astype(np.float32), mats.copy(),
               (rng.random(Bm) < 0.5).astype(np.float32))
    dump("cin dW layer 1 (H=200), one split", 342, "b200rec_debug_tc_trace_dw", first=40)
