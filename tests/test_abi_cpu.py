"""The C-ABI shared library builds for sm_100a, loads without a GPU and exports
every symbol include/soap_b200.h declares (no compute calls here)."""

import ctypes as C
import os
import re

from soap_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "soap_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(soap_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = C.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/soap_b200.h but not exported"
    # and the ctypes table binds exactly the header's functions
    assert sorted(_lib.SYMBOLS) == syms


def test_lib_loads_and_reports_version():
    L = _lib.lib()
    assert L.soap_abi_version() == _lib.ABI_VERSION


def test_layout_is_pure_host_logic():
    from soap_b200.halo_tasks import HaloPropConfig, result_layout

    cfg = HaloPropConfig(boxsize=10.0, G=1.0, critical_density=1.0, mean_density=0.3,
                         so=[("crit", 200.0), ("mean", 200.0)],
                         apertures=[(0.05, 0.03, 0), (0.05, 0.03, 1)], property_flags=1 | 4 | 8)
    ncol, cols = result_layout(cfg.to_c())
    assert cols["InputHalos/status"] == (0, 1)
    assert "SO/1/r" in cols and "Aperture/1/HalfMassRadiusStar" in cols
    assert ncol == sum(w for _, w in cols.values())
    # halo_tasks.py:306-317: the lowest threshold sets the target density
    assert cfg.target_density() == 200.0 * 0.3
    # property_flags bit 4: the iterative tensor pair next to the non-iterative one, per block
    # (BoundSubhalo, 2 SO, 2 apertures) and per projection axis
    cfg_it = HaloPropConfig(boxsize=10.0, G=1.0, critical_density=1.0, mean_density=0.3,
                            so=[("crit", 200.0), ("mean", 200.0)],
                            apertures=[(0.05, 0.03, 0), (0.05, 0.03, 1)], projected=[(0.03, 0.03)],
                            property_flags=1 | 4 | 8 | 16)
    cfg_no = HaloPropConfig(boxsize=10.0, G=1.0, critical_density=1.0, mean_density=0.3,
                            so=[("crit", 200.0), ("mean", 200.0)],
                            apertures=[(0.05, 0.03, 0), (0.05, 0.03, 1)], projected=[(0.03, 0.03)],
                            property_flags=1 | 4 | 8)
    n_it, c_it = result_layout(cfg_it.to_c())
    n_no, c_no = result_layout(cfg_no.to_c())
    assert n_it - n_no == 5 * 12 + 3 * 6
    assert c_it["BoundSubhalo/TotalInertiaTensorReduced"][1] == 6
    assert c_it["Aperture/0/StellarInertiaTensor"][1] == 6
    assert c_it["ProjectedAperture/0/projy/ProjectedTotalInertiaTensor"][1] == 3
    assert "SO/0/TotalInertiaTensor" not in c_no


def test_no_product_import_of_oracle():
    """the product package must never import the oracle (tier rule 3)"""
    pkg = os.path.join(ROOT, "soap_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
