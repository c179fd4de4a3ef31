"""CPU: the C-ABI shared library loads and exports every symbol include/mcb.h declares; the ctypes table in
magmaclust_b200/_lib.py covers exactly that set; compute entry points fail loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mcb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcb_[a-zA-Z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("mcb_create", "mcb_destroy", "mcb_set_data", "mcb_e_step_VEM", "mcb_eval_indiv", "mcb_eval_mean",
                 "mcb_m_step", "mcb_vem_iteration", "mcb_posterior_mu_k", "mcb_predict", "mcb_comm_init"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from magmaclust_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/mcb.h but not exported"


def test_ctypes_table_matches_header():
    from magmaclust_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback():
    """without a CUDA device the product refuses to run (it never routes through the oracle)"""
    from magmaclust_b200 import _lib
    lib = _lib.load()
    assert lib.mcb_version() >= 100
    if lib.mcb_device_count() > 0:
        pytest.skip("a GPU is present")
    h = _lib.H()
    rc = lib.mcb_create(ctypes.byref(h), 0)
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.mcb_last_error(None)
    from magmaclust_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "magmaclust_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dirpath, f)
