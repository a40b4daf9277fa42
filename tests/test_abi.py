"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/pba_b200.h declares, its structs match the ctypes mirror, and without a GPU it
fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT
from photobundle_b200 import capi

HEADER = os.path.join(ROOT, "include", "pba_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pba_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    names = _declared_functions()
    assert len(names) >= 18
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (pba_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in pba_b200.h but not exported: {missing}"
    assert sorted(capi.EXPORTED_SYMBOLS) == names  # the Python mirror tracks the header
    L = capi.lib()
    assert b"sm_100a" in L.pba_version()


def test_struct_layouts_match_header(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include "pba_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                    "sizeof(pba_config),sizeof(pba_solver_options),sizeof(pba_iteration_summary),"
                    "sizeof(pba_summary),sizeof(pba_eval_out));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(capi.Config), C.sizeof(capi.SolverOptions), C.sizeof(capi.IterationSummary),
                     C.sizeof(capi.Summary), C.sizeof(capi.EvalOut)]


def test_default_solver_options_are_the_reference_ones():
    """src/photobundle.cc:738-761 + Ceres defaults."""
    o = capi.SolverOptions()
    capi.lib().pba_default_solver_options(C.byref(o))
    assert o.max_num_iterations == 500
    assert o.function_tolerance == o.gradient_tolerance == o.parameter_tolerance == 1e-6
    assert o.initial_trust_region_radius == 1e4 and o.max_trust_region_radius == 1e16
    assert o.min_trust_region_radius == 1e-32 and o.min_relative_decrease == 1e-3
    assert o.min_lm_diagonal == 1e-6 and o.max_lm_diagonal == 1e32
    assert o.max_num_consecutive_invalid_steps == 5 and o.jacobi_scaling == 1


def test_default_solver_options_against_the_reference_source_text():
    """The values GetSolverOptions sets (src/photobundle.cc:738-761), read from the reference's own source text when the
    tree is present: max_num_iterations and the three tolerances are what pba_default_solver_options returns; the solver /
    strategy it names (SPARSE_SCHUR, TRUST_REGION, LEVENBERG_MARQUARDT) are what K_B implements; the loss threshold the
    residual blocks get comes from Options::robustThreshold (:797)."""
    import re
    src_path = "/root/reference/src/photobundle.cc"
    if not os.path.exists(src_path):
        pytest.skip("reference tree not present (GPU box)")
    src = open(src_path).read()
    body = src[src.index("GetSolverOptions(int num_threads"):src.index("void PhotometricBundleAdjustment::optimize(Result* result)")]
    o = capi.SolverOptions()
    capi.lib().pba_default_solver_options(C.byref(o))
    assert int(re.search(r"options\.max_num_iterations\s*=\s*(\d+);", body).group(1)) == o.max_num_iterations
    tol = float(re.search(r"double tol = ([0-9.e+-]+)\)", body).group(1))
    for name in ("function_tolerance", "gradient_tolerance", "parameter_tolerance"):
        assert re.search(r"options\.%s\s*=\s*tol;" % name, body) and getattr(o, name) == tol
    for token in ("ceres::SPARSE_SCHUR", "ceres::TRUST_REGION", "ceres::LEVENBERG_MARQUARDT"):
        assert token in body
    assert "huber_t = _options.robustThreshold" in src and "huber_t > 0.0 ? new ceres::HuberLoss(huber_t) : nullptr" in src


def test_argument_errors_are_reported():
    L = capi.lib()
    h = C.c_void_p()
    bad = capi.Config(rows=376, cols=1241, n_channels=1, patch_radius=9, max_frames=8, max_points=10,
                      max_observations=80, device=-1, fx=700, fy=700, cx=600, cy=180, huber=0.05)
    assert L.pba_create(C.byref(bad), C.byref(h)) == -1  # PBA_ERR_ARGUMENT
    assert b"patch_radius" in L.pba_last_error()
    bad.patch_radius = 2
    bad.max_frames = 99
    assert L.pba_create(C.byref(bad), C.byref(h)) == -1
    assert L.pba_solve(None, None, None) != 0


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.PbaError, match="no CUDA device|CUDA"):
        capi.Handle(376, 1241, 718.856, 718.856, 607.19, 185.21)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under photobundle_b200/ may reference it."""
    pkg = os.path.join(ROOT, "photobundle_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "pba_oracle" not in text and "oracle.binding" not in text and "from oracle" not in text, os.path.join(dp, f)


def test_descriptor_channel_counts():
    """pba_descriptor_channels is host arithmetic (DescriptorFrame::Create's channel counts, src/photobundle.cc:225-245)."""
    L = capi.lib()
    assert [L.pba_descriptor_channels(t) for t in (0, 1, 2)] == [1, 3, 8]
    assert L.pba_descriptor_channels(3) == -1 and L.pba_descriptor_channels(-1) == -1
