"""CPU: the C-ABI library loads, exports exactly what include/gloc3d.h declares, and
every compute entry point fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gloc3d_b200 as g
from gloc3d_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "gloc3d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gloc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 25
    lib = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gloc3d.h but not exported"
    # and the python binding covers the same set
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_string():
    lib = _lib.lib()
    assert lib.gloc_version() >= 100
    assert isinstance(lib.gloc_last_error(), bytes)


def test_product_never_touches_the_oracle():
    # the product package and its native sources must not reference oracle/
    pkg = os.path.join(ROOT, "gloc3d_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in txt and "gloc_oracle" not in txt and "libgloc_oracle" not in txt, f


def _no_gpu():
    return _lib.lib().gloc_device_count() == 0


def test_compute_fails_loudly_without_gpu():
    if not _no_gpu():
        pytest.skip("a B200 is present")
    with pytest.raises(g.GlocError) as e:
        g.KnnIndex(512, 0)
    assert e.value.code == _lib.GLOC_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(g.GlocError) as e:
        g.CsmStore(0)
    assert e.value.code == _lib.GLOC_ERR_CUDA
    with pytest.raises(g.GlocError):
        g.InvKeyTree(512, np.zeros((4, 512), np.float32))
    with pytest.raises(g.GlocError) as e:
        g.BevProjector(0)
    assert e.value.code == _lib.GLOC_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_argument_validation_without_gpu():
    lib = _lib.lib()
    assert lib.gloc_knn_create(None, 512, 0) == _lib.GLOC_ERR_INVALID
    h = C.c_void_p()
    assert lib.gloc_knn_create(C.byref(h), 0, 0) == _lib.GLOC_ERR_INVALID
    assert lib.gloc_knn_size(None) == 0
    assert lib.gloc_knn_query(None, None, 1, 1, None, None) == _lib.GLOC_ERR_INVALID
    assert lib.gloc_csm_num_grids(None) == 0
    assert lib.gloc_bev_create(None, 0, 0.2, 100.0) == _lib.GLOC_ERR_INVALID
    assert lib.gloc_bev_create(C.byref(h), 0, 0.0, 100.0) == _lib.GLOC_ERR_RANGE       # resolution > 0
    assert lib.gloc_bev_create(C.byref(h), 0, 0.001, 100.0) == _lib.GLOC_ERR_RANGE     # dense arrays bounded
    assert lib.gloc_bev_project(None, None, 0, 4, None) == _lib.GLOC_ERR_INVALID
    assert lib.gloc_bev_get_image(None, None, 0) == _lib.GLOC_ERR_NOT_BUILT
    assert lib.gloc_bev_kernel_launches(None) == 0
    assert lib.gloc_csm_add_grid_from_bev(None, None, None) == _lib.GLOC_ERR_INVALID
    assert lib.gloc_csm_add_grid_from_bev_aligned(None, None, None) == _lib.GLOC_ERR_INVALID
    # host-side helpers of the boundary work without a device
    pts = np.array([[3, 4, 0], [-60, 45, 1]], np.float32)
    nl, na, st = g.search_parameters(3.0, 3.0, pts, 0.2)
    assert nl == 15 and na == int(np.ceil(3.0 / st))
    with pytest.raises(AssertionError):
        g.InvKeyTree(512, np.zeros((0, 512), np.float32))


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/gloc3d.h compiles as C99 and links against the library."""
    import subprocess

    src = tmp_path / "abi.c"
    src.write_text('#include "gloc3d.h"\n'
                   'int main(void) { gloc_grid_info i; i.nx = 0; return (gloc_version() < 0) + i.nx; }\n')
    exe = tmp_path / "abi"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT}/include", str(src),
                        f"-L{ROOT}/gloc3d_b200", "-lgloc3d", f"-Wl,-rpath,{ROOT}/gloc3d_b200", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert subprocess.run([str(exe)]).returncode == 0


def test_microbench_validates_arguments_and_needs_a_gpu():
    import ctypes as C

    L = _lib.lib()
    assert L.gloc_bench_smem_gather(0, None) == _lib.GLOC_ERR_INVALID
    v = C.c_double(-1.0)
    rc = L.gloc_bench_smem_gather(0, C.byref(v))
    if L.gloc_device_count() == 0:
        assert rc == _lib.GLOC_ERR_CUDA and v.value == 0.0
