"""CPU: the C-ABI library loads and exports every symbol include/sg2b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, 'include', 'sg2b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sg2_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported():
    lib = ctypes.CDLL(os.path.join(ROOT, 'animeface_b200', 'libsg2b200.so'))
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/sg2b200.h but not exported'


def test_binding_matches_header():
    from animeface_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.sg2_version() >= 100
    assert lib.sg2_launch_count() == 0 or lib.sg2_launch_count() > 0


def test_argument_errors_without_gpu():
    """Validation happens before any CUDA call, so bad arguments are reported even without a device."""
    from animeface_b200 import _lib
    lib = _lib.load()
    s4 = (ctypes.c_int64 * 4)(1, 1, 1, 1)
    rc = lib.sg2_upfirdn2d(1, 1, 1, 0, 1, 1, 4, 4, s4, 4, 4, s4, 1, 1, 0, 1, 1, 1, 0, 0, 0, 0, 0, 1.0, None)
    assert rc == -1 and b'upsampling factor' in lib.sg2_last_error()
    rc = lib.sg2_bias_act(1, None, None, None, None, 1, 0, 16, 0, 1, 0, 42, 0.0, 1.0, -1.0, None)
    assert rc == -1 and b'activation' in lib.sg2_last_error()
    assert lib.sg2_conv2d_packed_size(8, 8, 5, 0) == -1


def test_product_path_has_no_oracle_import():
    """The shipped package must not import the CPU oracle (or /root/reference) anywhere."""
    pkg = os.path.join(ROOT, 'animeface_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
                assert '/root/reference' not in src, f
