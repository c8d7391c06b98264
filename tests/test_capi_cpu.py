"""CPU: the C-ABI library loads and exports every symbol include/softpool_b200.h declares; the
product path has no CPU fallback; the host modules mirror the reference's parameter layout."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "softpool_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:int|size_t|const char\*)\s+(\w+)\s*\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    from softpool_b200 import _lib
    names = header_functions()
    assert len(names) >= 12
    h = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(h, n), "libsoftpool_b200.so does not export %s" % n
    assert sorted(_lib.EXPORTS) == names                       # ctypes table == header
    assert _lib.lib().spk_abi_version() == _lib.ABI_VERSION


def test_library_is_sm100a_with_bulk_copies():
    """The shipped cubin is sm_100a and the gather kernels really use TMA bulk copies (UBLKCP)."""
    import shutil
    import subprocess
    from softpool_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([cuobjdump, "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass
    # the Chamfer forward really runs on the 5th-gen tensor cores: tcgen05.mma, TMEM loads with 16-bit packing,
    # and the packed 3-input 16-bit minimum (DPX) of its epilogue
    for mnemonic in ("UTCHMMA", "LDTM.x16.PACK16BIT", "VIMNMX3.U16x2"):
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback():
    import softpool_b200 as spb
    from softpool_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.softpool_topk(torch.randn(1, 2, 8), 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.chamfer_forward(torch.rand(1, 4, 3), torch.rand(1, 5, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        spb.chamferDist()(torch.rand(1, 4, 3), torch.rand(1, 5, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "softpool_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "oracle/" not in text or f.endswith((".cu", ".cuh")), f


@pytest.mark.skipif(torch.cuda.is_available(), reason="parameter layout is checked on CPU construction")
def test_module_parameter_layout_matches_reference():
    """Names/shapes of reference softpool.py:107-125 (checkpoints of the reference must load)."""
    import softpool_b200 as spb
    C, R, cab = 16, 4, 8
    m = spb.SoftPool(regions=R, cabins=cab, sp_ratio=4, size_feat=C)
    got = {n: tuple(p.shape) for n, p in m.state_dict().items()}
    assert got == {
        "conv2d_1.weight": (C, C, 1, 3), "conv2d_1.bias": (C,),
        "conv2d_2.weight": (C, C, 1, 3), "conv2d_2.bias": (C,),
        "conv2d_3.weight": (C, C, 1, cab - 4), "conv2d_3.bias": (C,),
        "conv2d_5.weight": (C, C, R, 1), "conv2d_5.bias": (C,),
        "sorter.conv1d.weight": (R, C, 1), "sorter.conv1d.bias": (R,),
    }
    f = spb.SoftPoolFeat(num_points=2048, regions=8, sp_points=2048, sp_ratio=8)
    keys = set(f.state_dict())
    for pre in ("conv1", "conv2", "conv3", "bn1", "bn2", "bn3"):
        assert pre + ".weight" in keys and pre + ".bias" in keys
    assert "softpool.sorter.conv1d.weight" in keys and f.sp_points == 256
    with pytest.raises(ValueError):
        spb.SoftPool(cabins=4)


def test_dropin_import_names():
    """`import softpool as sp`, `import dist_chamfer as cd`, `from extensions.chamfer_dist import ...`
    (reference model.py:12, train.py:18-19, GRNet test.py:19) resolve to this package via dropin/."""
    import importlib
    import sys
    d = os.path.join(ROOT, "dropin")
    sys.path.insert(0, d)
    try:
        for name in ("softpool", "dist_chamfer", "extensions.chamfer_dist"):
            sys.modules.pop(name, None)
        sp = importlib.import_module("softpool")
        cd = importlib.import_module("dist_chamfer")
        ext = importlib.import_module("extensions.chamfer_dist")
        assert sp.__file__.startswith(d)
        for n in ("SoftPool", "SoftPoolFeat", "Sorter", "train2cabins", "Periodics"):
            assert hasattr(sp, n)
        assert hasattr(cd, "chamferDist") and hasattr(cd, "chamferFunction")
        assert hasattr(ext, "ChamferDistance") and hasattr(ext, "ChamferFunction")
    finally:
        sys.path.remove(d)
        for name in ("softpool", "dist_chamfer", "extensions.chamfer_dist", "extensions"):
            sys.modules.pop(name, None)
