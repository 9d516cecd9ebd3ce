"""Builds tests/native/ffi_harness.cpp: integration/xla_ffi/eincm_xla_ffi.cc compiled unchanged against the stand-in FFI header."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, 'native', 'ffi_harness.cpp')
SO = os.path.join(HERE, 'native', '_ffi_harness.so')
DEPS = [SRC, os.path.join(ROOT, 'integration', 'xla_ffi', 'eincm_xla_ffi.cc'), os.path.join(ROOT, 'include', 'eincm.h'),
        os.path.join(HERE, 'native', 'mock_xla_ffi', 'xla', 'ffi', 'api', 'ffi.h')]
LIBDIR = os.path.join(ROOT, 'edge-informed-contrast-maximization_b200', 'lib')


def build():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in DEPS):
        subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-I', os.path.join(HERE, 'native', 'mock_xla_ffi'), '-I', os.path.join(ROOT, 'include'),
                        '-I', '/usr/local/cuda/include', SRC, '-o', SO, '-L', LIBDIR, '-leincm_b200', '-L', '/usr/local/cuda/lib64', '-lcudart',
                        '-Wl,-rpath,' + LIBDIR], check=True)
    return SO


def load():
    lib = C.CDLL(build())
    vp, i32, dbl = C.c_void_p, C.c_int, C.c_double
    lib.ffi_set_window.restype = i32
    lib.ffi_set_window.argtypes = [vp, i32, vp, vp, vp, C.c_longlong, vp, i32, i32, i32, vp, i32, vp, C.c_char_p, i32]
    lib.ffi_value_and_grad.restype = i32
    lib.ffi_value_and_grad.argtypes = [vp, i32, vp, i32, i32, vp, dbl, dbl, dbl, dbl, i32, i32, i32, vp, vp, C.c_char_p, i32]
    lib.ffi_binders_ok.restype = i32
    return lib
