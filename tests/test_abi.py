"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/treepl_b200.h declares, and FAILS LOUDLY (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from treepl_b200 import build as tbuild
from treepl_b200 import engine as eng

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "treepl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tpl_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = tbuild.build()
    assert os.path.exists(path)
    L = C.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 25
    for nm in names:
        assert hasattr(L, nm), "missing export: " + nm
    assert set(names) == set(eng.EXPORTED_SYMBOLS)


def test_library_is_sm100a_only():
    out = os.popen("cuobjdump -lelf %s 2>/dev/null" % tbuild.build()).read()
    if not out.strip():
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_param_order_vector_host_helper_matches_reference_layout():
    free = np.array([0, 1, 0, 0, 1, 1, 0, 1, 0, 0, 1, 0, 0], dtype=np.int32)
    fp, P = eng.generate_param_order_vector(free, lf=False)
    assert P == 17
    assert fp.tolist() == [-1] + list(range(12)) + [-1, 12, -1, -1, 13, 14, -1, 15, -1, -1, 16, -1, -1]
    fp, P = eng.generate_param_order_vector(free, lf=True)
    assert P == 6 and fp[:13].tolist() == [-1] + [0] * 12
    cv = np.zeros(13, dtype=np.int32)
    cv[12] = 1
    fp, P = eng.generate_param_order_vector(free, lf=False, cvnodes=cv)
    assert P == 16 and fp[12] == -1 and fp[11] == 10 and fp[14] == 11


def test_no_cpu_fallback_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from golden_util import Golden
    gd = Golden("test")
    with pytest.raises(eng.EngineError):
        eng.PLCalcParallel.from_flat(gd.flat)


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under treepl_b200/ may import, link or load it."""
    pkg = os.path.join(ROOT, "treepl_b200")
    banned = ("import oracle", "from oracle", "pl_oracle", "libpl_oracle", "libtreepl_ref", "refbind", "oracle.plo")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, fn)).read()
                for b in banned:
                    assert b not in src, (fn, b)
