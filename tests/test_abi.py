"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/vello_b200.h declares (no compute calls: there is no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    out = set()
    inc = os.path.join(ROOT, "include")
    for h in sorted(os.listdir(inc)):  # vello_b200.h (renderer) and vello_b200_scene.h (scene front end)
        if h.endswith(".h"):
            src = re.sub(r"/\*.*?\*/", "", open(os.path.join(inc, h)).read(), flags=re.S)
            out |= set(re.findall(r"\b(vb_[a-z_0-9]+)\s*\(", src))
    return sorted(out)


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from vello_b200.renderer import load_library, EXPORTED_SYMBOLS
    lib = load_library()
    declared = _declared_symbols()
    assert declared, "header parse failed"
    for s in declared:
        assert hasattr(lib, s), f"{s} declared under include/ but not exported"
    from vello_b200.scene_native import SCENE_SYMBOLS
    assert set(EXPORTED_SYMBOLS) | set(SCENE_SYMBOLS) <= set(declared)


def test_no_gpu_means_loud_failure():
    """The product path must fail loudly without a CUDA device (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vello_b200.renderer import Renderer, VelloB200Error
    with pytest.raises(VelloB200Error):
        Renderer()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vello_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "libvbo" not in txt, f
