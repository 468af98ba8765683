"""Write the inputs of every golden case as raw little-endian arrays for scala/DumpFixtures.scala
(tests/golden/README.md).  Run from the repo root:  python tests/golden/export_inputs.py"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from common import CONFIGS, kind_of  # noqa: E402


def main():
    for path in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
        case = os.path.basename(path)[:-4]
        name, B, F, K = case.split("_")
        cfg = CONFIGS[name]
        g = np.load(path)
        d = os.path.join(HERE, "reference", case)
        os.makedirs(d, exist_ok=True)
        g["index"].astype("<i4").tofile(os.path.join(d, "in_index.bin"))
        for k in ("weights", "bias", "embedding", "mats", "targets"):
            if k in g:
                g[k].astype("<f4").tofile(os.path.join(d, f"in_{k}.bin"))
        with open(os.path.join(d, "meta.txt"), "w") as f:
            f.write(f"kind={kind_of(name)}\nbatchSize={B[1:]}\nnFields={F[1:]}\nembeddingDim={K[1:]}\n")
            f.write("fcDims=" + ",".join(map(str, cfg.get("fc_dims", ()))) + "\n")
            f.write("cinDims=" + ",".join(map(str, cfg.get("cin_dims", ()))) + "\n")
            f.write(f"crossDepth={cfg.get('cross_depth', 0)}\n")
        print("wrote", d)


if __name__ == "__main__":
    main()
