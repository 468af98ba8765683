"""The C-ABI library loads on a CPU-only box, exports every symbol include/b200rec.h declares,
and refuses to compute without a B200 (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200rec.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rec_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(pkg):
    names = _declared()
    assert len(names) >= 45
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and the Python binding table covers exactly the header
    assert sorted(pkg._lib.SIGNATURES) == names


def test_abi_version_and_error_string(pkg):
    lib = pkg.lib()
    assert lib.b200rec_abi_version() == 1
    assert isinstance(pkg._lib.last_error(), str)


def test_header_is_plain_c(tmp_path):
    """include/b200rec.h must compile as C (the JNA / cgo side sees a C header)."""
    import subprocess
    c = tmp_path / "t.c"
    c.write_text('#include "b200rec.h"\nint main(void){return B200REC_OK;}\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c),
                        "-o", str(tmp_path / "t.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def _has_gpu(pkg):
    try:
        return pkg.device_count() > 0
    except Exception:
        return False


def test_no_cpu_fallback(pkg):
    """Without an sm_100 device every compute entry point fails loudly (status -4)."""
    if _has_gpu(pkg):
        pytest.skip("a B200 is present")
    with pytest.raises(pkg.B200RecError, match="no CPU fallback|no CUDA device|sm_100"):
        pkg.make_model("fm", 3, 4)
    with pytest.raises(pkg.B200RecError):
        pkg.EmbeddingTable(10, 4)
    with pytest.raises(pkg.B200RecError):
        pkg.Scatter(2, 1).updateOutput((np.ones(2, np.float32), np.zeros(2, np.int32)))
    with pytest.raises(pkg.B200RecError):
        pkg.scatter_add(np.zeros(2, np.int32), np.ones((2, 4), np.float32), np.ones(2, np.float32), dim=4)


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (test infrastructure only)."""
    pkgdir = os.path.join(ROOT, "recommendation-models_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text , f
