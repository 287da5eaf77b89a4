"""The C-ABI library loads on a CPU-only box, exports every symbol include/dmdb200.h declares, and FAILS LOUDLY
(no CPU fallback) when asked to compute without a CUDA device."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as g
from parallel_dmd_for_biomolecules_b200 import dmd, tables

ROOT = os.path.dirname(os.path.abspath(g.__file__))


@pytest.fixture(scope="module")
def lib():
    g.build()
    return dmd.load_library()


def test_header_and_binding_agree(lib):
    hdr = open(os.path.join(ROOT, "include", "dmdb200.h")).read()
    declared = set(re.findall(r"\b(dmdb_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(dmd.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layouts_match_header():
    assert C.sizeof(tables.Tables) == 8 * (12 + 400 * 3 + 28 + 120 + 100)
    assert C.sizeof(tables.Event) == 24
    assert C.sizeof(tables.Params) == 56
    assert C.sizeof(tables.Stats) == 8 * (2 + 32 + 5) + 8 + 8
    assert C.sizeof(tables.Energy) == 56


def test_no_cpu_fallback(lib, tab):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    topo = tables.Topology([tables.Species.from_sequence("KLVFFAE", 2)])
    with pytest.raises(dmd.DMDError) as e:
        dmd.DMD(tables.make_params(n_replicas=1), topo, tab)
    assert e.value.code == 2  # DMDB_ERR_NO_DEVICE
    assert "CUDA device" in str(e.value)


def test_missing_library_is_an_error(tmp_path, tab):
    topo = tables.Topology([tables.Species.from_sequence("KLVFFAE", 2)])
    with pytest.raises(dmd.DMDError):
        dmd.DMD(tables.make_params(), topo, tab, lib_path=str(tmp_path / "nope.so"))


def test_argument_validation_happens_before_device_use(lib):
    h = C.c_void_p()
    assert lib.dmdb_create(None, None, None, C.byref(h)) == 1  # DMDB_ERR_ARG
    assert b"null" in lib.dmdb_last_error(None)
