"""The drop-in boundary, checked by the compiler: the reference's own application (apps/run_kitti.cc:32-55,
`PhotometricBundleAdjustment photoba(calibration, imageSize, {cf}); photoba.addFrame(I, Z, T, &result);
writePosesKittiFormat(fn, result.poses)`) must type-check, unmodified and from where it lies, against this repo's
host/photobundle.h.  The reference tree exists only in the build container, so the test skips on the GPU box."""
import os
import subprocess

import pytest

HOST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "photobundle_b200", "host")
REF_APP = "/root/reference/apps/run_kitti.cc"


@pytest.mark.skipif(not os.path.exists(REF_APP), reason="reference tree not present (GPU box)")
def test_reference_application_type_checks_against_host_headers():
    out = subprocess.run(["make", "-C", HOST, "boundary-proof"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "type-checks against host/photobundle.h" in out.stdout


@pytest.mark.skipif(not os.path.exists(REF_APP), reason="reference tree not present (GPU box)")
def test_boundary_proof_rejects_a_broken_interface(tmp_path):
    """The proof has teeth: with addFrame's signature changed the same command fails."""
    broken = tmp_path / "photobundle.h"
    src = open(os.path.join(HOST, "photobundle.h")).read()
    assert "void addFrame(const uint8_t* image, const float* depth_map, const Mat44& T, Result* = nullptr);" in src
    broken.write_text(src.replace("void addFrame(const uint8_t* image, const float* depth_map, const Mat44& T, Result* = nullptr);",
                                  "void addFrame(const uint8_t* image, const Mat44& T);"))
    out = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", str(tmp_path), "-I", os.path.join(HOST, "boundary_proof"),
                          "-I", HOST, REF_APP], capture_output=True, text=True, timeout=120)
    assert out.returncode != 0
