"""CPU: the C-ABI library builds, loads and exports every symbol include/cns_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cns_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cns_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from cns_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 20
    raw = ctypes.CDLL(os.path.join(ROOT, "cns_b200", "libcns_sm100.so"))
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, "declared in include/cns_b200.h but not exported: %s" % missing


def test_binding_covers_header(lib):
    from cns_b200 import _lib
    assert set(_declared()) <= set(_lib.EXPORTED_SYMBOLS) | {"cns_forward_save_bytes"}


def test_batch_struct_matches_header():
    """ctypes mirror of `struct cns_batch`: same field names in the same order as include/cns_b200.h."""
    from cns_b200 import _lib
    src = open(os.path.join(ROOT, "include", "cns_b200.h")).read()
    body = src[src.index("typedef struct cns_batch {"):src.index("} cns_batch;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"([A-Za-z0-9_]+)\s*;", body)
    assert names == [n for n, _ in _lib.cns_batch._fields_]
    assert ctypes.sizeof(_lib.cns_batch) == 8 * len(names)      # int64 / pointer fields only


def test_param_table_matches_checkpoint(lib, state_dict):
    from cns_b200 import _lib
    names = _lib.param_names()
    assert names == list(state_dict.keys())
    for k, n in enumerate(names):
        assert lib.cns_param_numel(k) == state_dict[n].numel()


def test_no_cpu_fallback():
    """The product path must fail loudly without a GPU instead of computing on the CPU."""
    import torch
    from cns_b200 import GraphVS, _lib
    from cns_b200 import synth
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    net = GraphVS(2, 2, 128)
    import numpy as np
    data = synth.Scene(np.random.default_rng(0), 12).frame()
    with pytest.raises(_lib.CnsError):
        net(data)
    with pytest.raises(_lib.CnsError):
        _lib.Handle(torch.device("cuda:0"))


def test_product_does_not_import_oracle():
    """Nothing under cns_b200/ may import oracle/ (voids every parity claim otherwise)."""
    pkg = os.path.join(ROOT, "cns_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "cns_oracle" not in src and "pyg_shim" not in src and "ref_import" not in src, f
