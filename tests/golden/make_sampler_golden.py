"""Generates tests/golden/sampler_ref.npz from the REFERENCE's own sampler
(/root/reference/src/sample_eigen.h compiled into oracle/_ref/libref_sampler.so by
oracle/Makefile).  Run in the build container (needs /root/reference):

    python tests/golden/make_sampler_golden.py

The GPU box has no /root/reference; tests there compare against this fixture.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding  # noqa: E402


def main():
    binding.build()
    ref = binding.ref_lib()
    assert ref is not None, "oracle/_ref/libref_sampler.so missing (needs /root/reference)"
    rng = np.random.default_rng(20161201)
    rows, cols = 23, 37
    I = rng.integers(0, 256, size=(rows, cols)).astype(np.float32)
    gx = np.zeros_like(I)
    gy = np.zeros_like(I)
    binding.lib().oracle_imgradient(C.c_void_p(I.ctypes.data), rows, cols, C.c_void_p(gx.ctypes.data),
                                    C.c_void_p(gy.ctypes.data))
    n = 4096
    # interior, border, outside, exact integers, the (-1,0) extrapolation band, huge / NaN
    xs = rng.uniform(-3.0, cols + 2.0, size=n).astype(np.float32)
    ys = rng.uniform(-3.0, rows + 2.0, size=n).astype(np.float32)
    xs[:64] = np.round(xs[:64])
    ys[32:96] = np.round(ys[32:96])
    xs[96:128] = rng.uniform(-1.0, 0.0, size=32).astype(np.float32)
    ys[128:160] = rng.uniform(-1.0, 0.0, size=32).astype(np.float32)
    xs[160:164] = [cols - 1, cols - 2, cols - 1.5, cols - 2.0000002]
    ys[164:168] = [rows - 1, rows - 2, rows - 1.5, rows - 2.0000002]
    xs[168:172] = [3e9, -3e9, np.nan, 1e20]
    ys[172:176] = [3e9, -3e9, np.nan, -1e20]
    out = np.zeros((n, 3), dtype=np.float32)
    for i in range(n):
        ref.ref_sample_linear(C.c_void_p(I.ctypes.data), C.c_void_p(gx.ctypes.data), C.c_void_p(gy.ctypes.data),
                              rows, cols, C.c_float(ys[i]), C.c_float(xs[i]), C.c_void_p(out[i].ctypes.data))
    # Jet chain rule (src/jet_extras.h:86-111) on a few hundred points
    m = 256
    xv = rng.normal(size=(m, 9))
    yv = rng.normal(size=(m, 9))
    xa = rng.uniform(1.0, cols - 2.0, size=m)
    ya = rng.uniform(1.0, rows - 2.0, size=m)
    ja = np.zeros(m)
    jv = np.zeros((m, 9))
    for i in range(m):
        a = C.c_double()
        ref.ref_sample_with_derivative_jet9(
            C.c_void_p(I.ctypes.data), C.c_void_p(gx.ctypes.data), C.c_void_p(gy.ctypes.data), rows, cols,
            C.c_double(xa[i]), C.c_void_p(xv[i].ctypes.data), C.c_double(ya[i]), C.c_void_p(yv[i].ctypes.data),
            C.byref(a), C.c_void_p(jv[i].ctypes.data))
        ja[i] = a.value
    path = os.path.join(ROOT, "tests", "golden", "sampler_ref.npz")
    np.savez_compressed(path, I=I, gx=gx, gy=gy, xs=xs, ys=ys, out=out, xa=xa, ya=ya, xv=xv, yv=yv, ja=ja, jv=jv)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
