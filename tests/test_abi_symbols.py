"""CPU: the C-ABI library loads and exports every symbol include/vilgod_b200.h declares.
No compute call is made (there is no GPU here and no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from vilgod_b200 import build
    return build.build_all()[0]


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vilgod_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vg_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_documented_surface():
    names = declared_symbols()
    for must in ("vg_create", "vg_destroy", "vg_load_vit_weights", "vg_set_text_features",
                 "vg_workspace_bytes", "vg_project", "vg_encode_score", "vg_vote", "vg_classify",
                 "vg_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_ctypes_binding_covers_the_header(lib_path):
    from vilgod_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared_symbols()
    for dt, code in (("f16", 1), ("bf16", 0)):
        lib = _lib.load(dt)
        assert lib.vg_abi_version() == _lib.VG_ABI_VERSION and lib.vg_operand_dtype() == code


def test_struct_layouts_match_the_header():
    from vilgod_b200 import _lib
    # VgConfig: 9 int32 (+4 padding), 3 double, 16*9 + 9 floats (+4 tail padding to 8)
    assert ctypes.sizeof(_lib.VgConfig) == 40 + 24 + (16 * 9 + 9) * 4 + 4
    assert _lib.VgConfig.obj_ratio.offset == 40 and _lib.VgConfig.div_mode.offset == 24
    assert ctypes.sizeof(_lib.VgVitLayerWeights) == 12 * 8
    assert ctypes.sizeof(_lib.VgVitWeights) == (5 + 12 * 12 + 3) * 8


def test_default_library_is_the_fp16_operand_build(lib_path):
    """The default (benchmarked) build computes in the reference's own GPU dtype, fp16 operands with
    fp32 accumulation (third_party/CLIP/clip/model.py:375-396); bf16 is the alternative build."""
    from vilgod_b200 import _lib, build
    assert lib_path == build.LIB_PATH and _lib.LIB_PATHS["f16"] == build.LIB_PATH
    assert ctypes.CDLL(lib_path).vg_operand_dtype() == 1


def test_production_library_holds_no_first_generation_kernels(lib_path):
    """Only the tcgen05 / TMA kernels ship: no mma.sync (HMMA) attention or single-CTA GEMM."""
    import subprocess
    r = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "UTCHMMA" in r.stdout and "UTMALDG" in r.stdout
    assert " HMMA" not in r.stdout


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vilgod_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(num_views=4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vilgod_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle;", ""), \
                    f"{f} mentions the oracle package"
