"""The C-ABI library: loads, and exports every symbol include/density_b200.h declares (no compute calls, CPU only)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "density_b200.h")).read()
    names = re.findall(r"^DENSITY_B200_API\s+[^;(]*?\b([a-z_0-9]+)\s*\(", hdr, flags=re.M)
    assert len(names) >= 20
    return names


def test_header_declares_reference_ffi_surface():
    names = set(declared_symbols())
    # chameleon.rs:70-83, cheetah.rs:105-118, lion.rs:193-206
    for a in ("chameleon", "cheetah", "lion"):
        for op in ("encode", "decode", "safe_encode_buffer_size"):
            assert f"{a}_{op}" in names


def test_library_builds_and_exports_every_declared_symbol():
    from density_b200 import build as b
    so = b.build()
    lib = ctypes.CDLL(so)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in density_b200.h but not exported"


def test_python_binding_covers_header():
    from density_b200 import _lib
    assert set(declared_symbols()) == set(_lib.EXPORTED_SYMBOLS)
    _lib.load()


@pytest.mark.parametrize("alg,block,sig", [("chameleon", 256, 8), ("cheetah", 128, 8), ("lion", 64, 6)])
def test_safe_encode_buffer_size_matches_reference_formula(alg, block, sig):
    # pure host arithmetic (codec.rs:18-21): callable without a GPU
    from density_b200 import CODECS
    import oracle
    C = CODECS[alg]
    assert (C.block_size(), C.signature_significant_bytes()) == (block, sig)
    for n in (0, 1, block - 1, block, block + 1, 10192446, 1 << 30, (1 << 32) + 5):
        assert C.safe_encode_buffer_size(n) == oracle.safe_encode_buffer_size(alg, n)


def test_no_gpu_means_loud_failure_not_fallback():
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from density_b200 import Chameleon, EncodeError
    data = np.frombuffer(b"test" * 64, dtype=np.uint8)
    out = np.zeros(Chameleon.safe_encode_buffer_size(data.size), dtype=np.uint8)
    with pytest.raises(EncodeError):
        Chameleon.encode(data, out)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "density_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "oracle/" not in src.replace("oracle/density_oracle.c", ""), f
