"""CPU: the C-ABI library builds, loads and exports every symbol include/ttr_b200.h declares."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not (ROOT / "twotowermlretrieval_b200" / "libttr_b200.so").exists():
        g.build()
    from twotowermlretrieval_b200 import _lib
    return _lib.load()


def header_symbols():
    text = (ROOT / "include" / "ttr_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ttr_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
    syms = header_symbols()
    assert "ttr_score_topk" in syms and "ttr_gru_recurrence_fwd" in syms and len(syms) >= 20


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in ttr_b200.h but not exported: {missing}"


def test_binding_table_matches_header(lib):
    from twotowermlretrieval_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == header_symbols()


def test_no_compute_without_gpu_but_version_works(lib):
    assert lib.ttr_version() >= 100
    assert isinstance(lib.ttr_last_error(), bytes)


def test_product_never_imports_oracle():
    pkg = ROOT / "twotowermlretrieval_b200"
    for f in pkg.rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f"{f} imports the oracle"


def test_cpu_tensor_is_rejected_loudly():
    import torch
    from twotowermlretrieval_b200 import TwoTowerModel, synth
    from twotowermlretrieval_b200._lib import TTRError
    cfg = synth.default_config(vocab_size=50, embed_dim=8)
    cfg["HIDDEN_DIM"] = 8
    m = TwoTowerModel(cfg)
    with pytest.raises(TTRError):
        m.encode_query(torch.tensor([[1, 2, 3]]))


def test_binding_arity_and_types_match_header():
    """Every ctypes signature has the same number of arguments, and pointer/integer/float kinds
    in the same positions, as the prototype in include/ttr_b200.h."""
    import ctypes
    from twotowermlretrieval_b200 import _lib
    text = (ROOT / "include" / "ttr_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = dict(re.findall(r"\b(ttr_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text))
    kinds = {ctypes.c_void_p: "p", ctypes.c_int: "i", ctypes.c_int64: "l", ctypes.c_uint64: "u", ctypes.c_float: "f",
             ctypes.c_double: "d", ctypes.c_uint32: "w"}
    for name, sig in _lib._SIGNATURES.items():
        params = [p.strip() for p in protos[name].split(",") if p.strip() and p.strip() != "void"]
        want = []
        for p in params:
            if "*" in p:
                want.append("p")
            elif p.startswith("int64_t"):
                want.append("l")
            elif p.startswith("uint64_t"):
                want.append("u")
            elif p.startswith("uint32_t"):
                want.append("w")
            elif p.startswith("int"):
                want.append("i")
            elif p.startswith("float"):
                want.append("f")
            elif p.startswith("double"):
                want.append("d")
            else:
                raise AssertionError(f"{name}: cannot classify parameter {p!r}")
        got = [kinds[a] for a in sig]
        assert got == want, f"{name}: binding {got} vs header {want}"
