"""The C-ABI library loads and exports every symbol include/nvblox_b200.h declares.
No compute calls: this runs without a GPU."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "nvblox_b200.h")).read()
    return sorted(set(re.findall(r"NVB_API[^;(]*?\b(nvb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    syms = header_symbols()
    for needed in ("nvb_view_raycast", "nvb_mapper_integrate_depth", "nvb_mapper_update_esdf",
                   "nvb_esdf_integrate_blocks", "nvb_layer_get_blocks"):
        assert needed in syms


def test_library_exports_every_declared_symbol(built):
    from isaac_ros_nvblox_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "libnvblox_b200.so does not export %s" % s
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms, "python binding list and header disagree"


def test_no_torch_types_in_signatures():
    text = open(os.path.join(ROOT, "include", "nvblox_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # declarations only, comments cite the reference's C++
    assert "torch" not in code and "at::" not in code and "std::" not in code and "Eigen" not in code


def test_block_bytes_and_defaults(built):
    from isaac_ros_nvblox_b200 import _lib
    L = _lib.load()
    assert L.nvb_layer_block_bytes(0) == 4096 and L.nvb_layer_block_bytes(1) == 10240
    p = _lib.NvbTsdfParams()
    L.nvb_default_tsdf_params(ctypes.byref(p))
    # projective_integrator_params.h:24-63, view_calculator_params.h:22-25
    assert (p.truncation_distance_vox, p.max_integration_distance_m, p.max_weight) == (4.0, 7.0, 5.0)
    assert p.invalid_depth_decay_factor == -1.0 and p.weighting_type == 2 and p.raycast_subsampling == 4
    e = _lib.NvbEsdfParams()
    L.nvb_default_esdf_params(ctypes.byref(e))
    assert e.max_esdf_distance_m == 2.0 and e.max_site_distance_vox == 1.0 and abs(e.min_weight - 1e-4) < 1e-9


def test_fails_loudly_without_a_gpu(built):
    """No CPU fallback: creating a mapper without a CUDA device is an error, not a slow path."""
    from isaac_ros_nvblox_b200 import _lib
    import isaac_ros_nvblox_b200 as nvb
    import pytest
    if _lib.load().nvb_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(_lib.NvbError) as ei:
        nvb.Mapper(0.05)
    assert ei.value.code == -5
    assert "no CPU fallback" in str(ei.value)


def _compile_cpp_dropin(tmp_path, name="test_mapper_dropin"):
    import subprocess
    from isaac_ros_nvblox_b200 import _lib
    exe = str(tmp_path / name)
    src = os.path.join(ROOT, "tests", "cpp", name + ".cpp")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
           "-L", libdir, "-lnvblox_b200", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return exe


def test_cpp_mirror_headers_compile_and_link(built, tmp_path):
    """include/nvblox/*.h (the source-compatible subset of the reference's headers) builds with plain g++
    against the C-ABI library; without a GPU the program reports that and exits 77."""
    import subprocess
    from isaac_ros_nvblox_b200 import _lib
    for name in ("test_mapper_dropin", "test_mirror_surface", "test_multi_mapper_dropin", "test_mesh_dropin", "test_streamer_dropin"):  # the second covers Plane, planar slices and EsdfSlicer; the third MultiMapper + the reference's Mapper constructor
        exe = _compile_cpp_dropin(tmp_path, name)
        if _lib.load().nvb_device_count() == 0:
            assert subprocess.call([exe]) == 77


def test_header_is_strict_c99(tmp_path):
    """include/nvblox_b200.h is the FFI boundary (cgo / ctypes / JNI read it as C): it must compile as plain C99."""
    import subprocess
    src = tmp_path / "use_header.c"
    src.write_text('#include "nvblox_b200.h"\n'
                   'int use(void) { NvbColorParams c; NvbTsdfParams t; nvb_default_color_params(&c); nvb_default_tsdf_params(&t);\n'
                   '  return (int)sizeof(NvbEsdfVoxel) + (int)sizeof(NvbColorVoxel) + (int)sizeof(NvbFreespaceVoxel); }\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                           str(src)])
