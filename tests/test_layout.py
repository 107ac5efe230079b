"""Byte layouts and the C-ABI surface (CPU only; no compute calls)."""
import ctypes as C
import os
import re

from lumen_b200 import _ctypes_types as T
from lumen_b200 import integrator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_struct_sizes_match_reference_layouts():
    # SURVEY.md appendix B / src/shaders/commons.h:180-340, path_commons.h:3-14
    assert C.sizeof(T.Vertex) == 32 and C.sizeof(T.Light) == 128 and C.sizeof(T.Material) == 104
    assert C.sizeof(T.PrimMeshInfo) == 48 and C.sizeof(T.PCPath) == 52 and C.sizeof(T.SceneUBO) == 492
    assert T.Material.bsdf_type.offset == 28 and T.Material.texture_id.offset == 48 and T.Material.thin.offset == 100
    assert T.Light.pos.offset == 64 and T.Light.light_flags.offset == 108 and T.Light.world_radius.offset == 124
    assert T.PCPath.frame_num.offset == 12 and T.PCPath.max_depth.offset == 32 and T.PCPath.direct_lighting.offset == 48
    # PCBDPT (bdpt_commons.h:5-16) = PCPath without its last field
    assert C.sizeof(T.PCBdpt) == 48 and [f[0] for f in T.PCBdpt._fields_] == [f[0] for f in T.PCPath._fields_][:-1]
    assert T.PCBdpt.time.offset == T.PCPath.time.offset == 28 and T.PCBdpt.dir_light_idx.offset == 44
    assert T.SceneUBO.inv_view.offset == 192 and T.SceneUBO.inv_projection.offset == 256


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lmb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = integrator.lib()  # loads liblumen_b200.so; no CUDA call is made
    for header in ("lumen_b200.h", "lumen_b200_testhooks.h"):
        names = _declared(header)
        assert names, header
        for n in names:
            assert hasattr(lib, n), f"{n} declared in {header} but not exported"
    assert set(_declared("lumen_b200.h")) == set(integrator.EXPORTS)
    assert set(_declared("lumen_b200_testhooks.h")) == set(integrator.TESTHOOK_EXPORTS)


def test_host_library_exports():
    from lumen_b200 import host
    text = open(os.path.join(ROOT, "include", "lumen_host.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for n in set(re.findall(r"\b(lmh_[a-z0-9_]+)\s*\(", text)):
        assert hasattr(host.lib(), n), n


def test_no_cpu_fallback_without_device():
    """Without a GPU the product path must fail loudly, not fall back (the parity claim depends on it)."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        integrator.Device(0)
    from conftest import scene_path
    from lumen_b200 import host
    sc = host.Scene(scene_path("cornell"), 16, 16)
    for cls in (integrator.PathB200, integrator.BDPTB200):  # both integrator mirrors own a Device: same rule
        with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
            cls(sc)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under lumen_b200/ or include/ may reference it."""
    bad = []
    for base in ("lumen_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                    src = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|liboracle|pyoracle|orc_[a-z]", src):
                        # comments that merely mention the oracle by name are fine; code references are not
                        code = re.sub(r"//.*|/\*.*?\*/|#.*|\"\"\".*?\"\"\"", "", src, flags=re.S)
                        if re.search(r"(from|import)\s+oracle|liboracle|pyoracle|orc_[a-z]", code):
                            bad.append(os.path.join(dirpath, f))
    assert not bad, bad
