"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/eincm.h declares.
No compute call is made here (no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'eincm.h')


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build_cuda()
    from eincm_b200 import plan
    return plan.load_library()


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(eincm_[a-z_]+)\s*\(', src)))


def test_header_functions_all_exported(lib):
    from eincm_b200 import plan
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/eincm.h but not exported'
    assert set(names) == set(plan.EXPORTED_SYMBOLS)


def test_abi_version_and_error_string(lib):
    assert lib.eincm_abi_version() == 1
    assert lib.eincm_last_error(None) is not None


def test_header_is_plain_c():
    # the boundary must be consumable from C (cgo / JNI / ctypes / XLA-FFI shims): compile it with gcc -std=c99
    prog = '#include "eincm.h"\nint main(void){ eincm_hparams hp; (void)hp; return EINCM_ABI_VERSION == 1 ? 0 : 1; }\n'
    r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), '-x', 'c', '-', '-fsyntax-only'],
                       input=prog.encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()


def test_library_is_sm100a_only():
    from eincm_b200 import plan
    out = subprocess.run(['cuobjdump', '-lelf', plan.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    assert not re.search(r'sm_(?!100a)\d+', out)


def test_no_cpu_fallback_fails_loudly(lib):
    """Without a CUDA device the product must raise, not silently compute elsewhere."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from eincm_b200 import losses, plan
    import numpy as np
    with pytest.raises(plan.EincmError):
        plan.Plan((48, 64), max_events=100)
    h = ctypes.c_void_p()
    rc = lib.eincm_plan_create(ctypes.byref(h), 0, 48, 64, 100, 3, 0)
    assert rc == plan.EINCM_ECUDA and not h.value
    assert b'no CUDA device' in lib.eincm_last_error(None)
    with pytest.raises(plan.EincmError):
        losses.loss_func(np.zeros((1, 1, 2)), np.zeros(1, np.int16), np.zeros(1, np.int16), np.zeros(1), np.zeros((1, 48, 64)),
                         np.zeros(1), 1, 1, 0, 0, 0, 5, (48, 64), 'bilinear')


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'edge-informed-contrast-maximization_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.cc')):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), f'{f} imports the oracle'
