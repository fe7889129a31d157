"""The C-ABI library loads without a GPU and exports every symbol include/ftb200.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

from forwardtacotron_b200 import _lib

HEADER = Path(__file__).resolve().parent.parent / 'include' / 'ftb200.h'


def declared_functions():
    text = re.sub(r'/\*.*?\*/', '', HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r'\b(ftb_[a-z0-9_]+)\s*\(', text)))


def test_library_builds_and_loads():
    lib = _lib.lib()
    assert lib.ftb_abi_version() == 1


def test_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(_lib.lib_path()))
    names = declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f'declared in ftb200.h but not exported: {missing}'
    unbound = [n for n in names if n not in _lib.SIGNATURES]
    assert not unbound, f'declared in ftb200.h but not bound in _lib.SIGNATURES: {unbound}'


def test_struct_layouts_match():
    lib = _lib.lib()
    for sid, cls in _lib.STRUCT_IDS.items():
        assert lib.ftb_struct_size(sid) == ctypes.sizeof(cls), cls.__name__


def test_errors_are_reported_not_swallowed():
    lib = _lib.lib()
    assert lib.ftb_length_plan(None, None, None, 0, 0, None) == -1
    assert b'ftb_length_plan' in lib.ftb_last_error()
    with pytest.raises(_lib.FtbError):
        _lib.check(lib.ftb_length_plan(None, None, None, 0, 0, None))
