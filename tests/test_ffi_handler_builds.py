"""The XLA FFI handler source (integration/xla_ffi/eincm_xla_ffi.cc) compiles unchanged against the stand-in for jaxlib's FFI header
and links against the C-ABI library: every entry point it calls exists with the signature include/eincm.h declares.  No GPU needed
(nothing is executed beyond the binder type check)."""
import subprocess

from tests import _ffi_harness


def test_handler_source_compiles_and_links_against_the_c_abi():
    so = _ffi_harness.build()
    syms = subprocess.run(['nm', '-D', so], capture_output=True, text=True, check=True).stdout
    for name in ('ffi_set_window', 'ffi_value_and_grad', 'EincmSetWindow_mock_bound', 'EincmValueAndGrad_mock_bound'):
        assert f' T {name}' in syms
    for name in ('eincm_plan_set_window_device_ts', 'eincm_value_and_grad', 'eincm_plan_create'):
        assert f' U {name}' in syms                      # resolved by libeincm_b200.so


def test_handler_takes_no_window_identity_at_trace_time():
    """`window_id` was an attribute (a trace-time constant of the cached executable) and `edge_ts` a host attribute.  The
    objective call now has no attribute or operand that changes from window to window except the token BUFFER."""
    import os
    import re
    src = open(os.path.join(_ffi_harness.ROOT, 'integration', 'xla_ffi', 'eincm_xla_ffi.cc')).read()
    vg = src[src.index('XLA_FFI_DEFINE_HANDLER_SYMBOL(EincmValueAndGrad'):]
    attrs = re.findall(r'\.Attr<[^>]+>\("(\w+)"\)', vg)
    assert attrs == ['alpha', 'beta', 'gamma', 'delta', 'cur_pyr_lvl', 'n_pyr_lvls', 'slot']
    sw = src[src.index('XLA_FFI_DEFINE_HANDLER_SYMBOL(EincmSetWindow'):src.index('XLA_FFI_DEFINE_HANDLER_SYMBOL(EincmValueAndGrad')]
    assert re.findall(r'\.Attr<[^>]+>\("(\w+)"\)', sw) == ['slot'] and sw.count('.Arg<') == 5      # edge_ts is a device operand
    py = open(os.path.join(_ffi_harness.ROOT, 'integration', 'xla_ffi', 'eincm_jax.py')).read()
    assert re.search(r'(?<!j)np\.asarray\(edge_ts', py) is None and '_window_id' not in py      # no host conversion of a tracer
