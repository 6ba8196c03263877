"""The C-ABI library loads without a GPU, exports every symbol the headers declare, and fails loudly (never
falls back) when no device is present."""
import ctypes
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import rustsolver_b200 as rb
from rustsolver_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def _declared(header: Path):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(rsh?_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    for hdr, table in ((ROOT / "include" / "b200cfr.h", _lib.ENGINE_API), (ROOT / "include" / "b200cfr_host.h", _lib.HOST_API)):
        names = _declared(hdr)
        assert names, hdr
        for n in names:
            assert hasattr(lib, n), f"{n} declared in {hdr.name} but not exported"
            assert n in table, f"{n} has no ctypes prototype"


def test_exports_are_plain_c_no_torch_types():
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    api = [s for s in syms if s.startswith("rs_") or s.startswith("rsh_")]
    assert len(api) >= len(_lib.ENGINE_API) + len(_lib.HOST_API)
    assert not any("torch" in s or "at::" in s for s in syms)
    needed = subprocess.run(["ldd", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "torch" not in needed and "python" not in needed


def test_headers_compile_as_c():
    src = '#include "b200cfr.h"\n#include "b200cfr_host.h"\nint main(void){ rs_config c; (void)c; return 0; }\n'
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", str(ROOT / "include"), "-x", "c", "-"],
                       input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_cpu_fallback_without_a_device():
    lib = _lib.load()
    if lib.rs_device_count() > 0:
        pytest.skip("a GPU is present")
    o = rb.default_flop()
    n, tree = rb.build_game_tree(o)
    with pytest.raises(rb.EngineError) as ei:
        rb.Engine(tree, o.ranges(), o.board_mask)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    for p in (ROOT / "rustsolver_b200").rglob("*"):
        if p.suffix in (".py", ".cpp", ".cu", ".h", ".cuh"):
            t = p.read_text()
            assert "import oracle" not in t and "from oracle" not in t and "liborc" not in t.replace("oracle/liborc.so with gcc", "").replace('ORACLE_SO = ROOT / "oracle" / "liborc.so"', ""), p


def test_sample_runouts_restates_generate_hands_board_part():
    """rs_sample_runouts (cfr.rs:100-122): cards off the board, distinct inside a path, distinct first cards on request,
    reproducible for a seed and uniform over the live cards."""
    import numpy as np
    import rustsolver_b200 as rb
    bm = rb.get_card_mask("4d5dAs")
    a = rb.sample_runouts(7, bm, 2, 20)
    b = rb.sample_runouts(7, bm, 2, 20)
    assert np.array_equal(a, b) and a.shape == (20, 2)
    assert not np.array_equal(a, rb.sample_runouts(8, bm, 2, 20))
    assert all(not (bm >> int(c)) & 1 for c in a.ravel()) and all(x != y for x, y in a)
    assert len(set(a[:, 0].tolist())) == 20
    full = rb.sample_runouts(3, bm, 1, 49)
    assert sorted(full.ravel().tolist()) == [c for c in range(52) if not (bm >> c) & 1]
    with pytest.raises(rb.EngineError):
        rb.sample_runouts(3, bm, 1, 50)
    many = rb.sample_runouts(11, bm, 2, 20000, distinct_first=False)
    counts = np.bincount(many[:, 1], minlength=52)[[c for c in range(52) if not (bm >> c) & 1]]
    assert counts.min() > 0.8 * counts.mean() and counts.max() < 1.2 * counts.mean()
