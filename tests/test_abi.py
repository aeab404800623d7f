"""CPU: the C-ABI library loads and exports every symbol include/elem_b200.h declares; no compute without a GPU."""
import ctypes
import os
import re

import pytest

from elementary_b200 import Runtime, runtime

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "elem_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(elem_b200_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(runtime.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/elem_b200.h but not exported"


def test_no_cpu_fallback_without_device():
    lib = runtime.load_library()
    if lib.elem_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError):
        Runtime(48000.0, 512, 4, device=0)          # no device -> creation fails loudly
    rt = Runtime(48000.0, 512, 4, device=-1)        # plan-only: host logic yes, rendering no
    assert rt.apply_instructions([[0, 1, "root"], [0, 2, "const"], [2, 1, 2, 0], [3, 1, "channel", 0], [4, [1]], [5]]) == 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rt.process(None, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rt.process_voices(None, 1)


def test_product_does_not_link_or_import_the_oracle():
    """The product path must never route through oracle/: neither the Python package nor the library mention it."""
    pkg = os.path.join(ROOT, "elementary_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libelem_oracle" not in src and "libelem_ref" not in src and "import oracle" not in src \
                    and "from oracle" not in src, f"{f} references the oracle"


def test_return_code_descriptions():
    assert runtime.describe_return_code(0) == "Ok"
    assert runtime.describe_return_code(1) == "Node type not recognized"
    assert runtime.describe_return_code(8) == "Invalid instruction format"
