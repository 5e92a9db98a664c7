"""
CPU tier: the C-ABI shared library (built by nvcc for sm_100a) loads without a GPU
and exports every symbol include/psfmc_b200.h declares; argument validation and
error reporting work without touching a device; the ctypes structures match the
header's layout. No compute calls here.
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, 'include', 'psfmc_b200.h')


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(psfmc_[a-z0-9_]+)\s*\(', text)))


def test_header_and_binding_agree():
    from psfmc_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_functions()


def test_library_exports_every_declared_symbol(cuda_library):
    lib = ctypes.CDLL(cuda_library)
    for name in _declared_functions():
        assert hasattr(lib, name), name
    # and they are C symbols (unmangled) in the dynamic table
    out = subprocess.run(['nm', '-D', '--defined-only', cuda_library],
                         stdout=subprocess.PIPE, universal_newlines=True).stdout
    exported = set(re.findall(r' T (psfmc_\w+)', out))
    assert set(_declared_functions()) <= exported


def test_library_carries_sm100a_code_only(cuda_library):
    out = subprocess.run(['cuobjdump', '-lelf', cuda_library], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, universal_newlines=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, out


def test_struct_layout_matches_header(cuda_library, tmp_path):
    """sizeof/offsetof of the ABI structs as the C compiler sees them."""
    from psfmc_b200 import _lib
    src = tmp_path / 'layout.c'
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "psfmc_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(psfmc_slot),'
        'sizeof(psfmc_component), sizeof(psfmc_desc), sizeof(psfmc_info),'
        'offsetof(psfmc_desc, psf_index), offsetof(psfmc_desc, devices),'
        'offsetof(psfmc_info, flops_per_eval), offsetof(psfmc_info, launches_total));'
        'printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(psfmc_prior_column),'
        'offsetof(psfmc_prior_column, log_shape), sizeof(psfmc_prior_plan),'
        'offsetof(psfmc_prior_plan, other_columns), sizeof(psfmc_ensemble),'
        'offsetof(psfmc_ensemble, mt_pos), offsetof(psfmc_ensemble, n_accepted));'
        'return 0;}\n')
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)],
                   check=True)
    got = [int(v) for v in subprocess.run([str(exe)], stdout=subprocess.PIPE,
                                          universal_newlines=True).stdout.split()]
    want = [ctypes.sizeof(_lib.Slot), ctypes.sizeof(_lib.Component),
            ctypes.sizeof(_lib.Desc), ctypes.sizeof(_lib.Info),
            _lib.Desc.psf_index.offset, _lib.Desc.devices.offset,
            _lib.Info.flops_per_eval.offset, _lib.Info.launches_total.offset,
            ctypes.sizeof(_lib.PriorColumn), _lib.PriorColumn.log_shape.offset,
            ctypes.sizeof(_lib.PriorPlan), _lib.PriorPlan.other_columns.offset,
            ctypes.sizeof(_lib.Ensemble), _lib.Ensemble.mt_pos.offset,
            _lib.Ensemble.n_accepted.offset]
    assert got == want


def test_invalid_arguments_are_reported_not_thrown(cuda_library):
    from psfmc_b200 import _lib
    lib = _lib.load(cuda_library)
    assert lib.psfmc_abi_version() == _lib.ABI_VERSION
    handle = ctypes.c_void_p()
    assert lib.psfmc_engine_create(None, ctypes.byref(handle)) == 1
    assert b'null' in lib.psfmc_last_error()
    desc = _lib.Desc()
    desc.abi_version = 999
    assert lib.psfmc_engine_create(ctypes.byref(desc), ctypes.byref(handle)) == 1
    assert b'abi_version' in lib.psfmc_last_error()
    desc.abi_version = _lib.ABI_VERSION
    desc.height, desc.width = 100, 101          # odd width: the reference rejects it too
    assert lib.psfmc_engine_create(ctypes.byref(desc), ctypes.byref(handle)) == 2
    assert b'odd' in lib.psfmc_last_error()
    assert handle.value is None
    # too large for the padded transform frame (1000 + 64 - 1 > 1024)
    px = np.zeros(1000 * 1000)
    stamp = np.zeros(64 * 64)
    bad = np.zeros(1000 * 1000, dtype=np.uint8)
    dbl = ctypes.POINTER(ctypes.c_double)
    desc.height, desc.width = 1000, 1000
    desc.obs_data = desc.obs_var = px.ctypes.data_as(dbl)
    desc.bad_px = bad.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
    desc.n_psf, desc.psf_height, desc.psf_width = 1, 64, 64
    desc.psf = desc.psf_var = stamp.ctypes.data_as(dbl)
    assert lib.psfmc_engine_create(ctypes.byref(desc), ctypes.byref(handle)) == 2
    assert b'too large' in lib.psfmc_last_error()
    assert handle.value is None
    # a slot index below -1 (-1 = constant) is rejected before any device is touched
    desc.height, desc.width = 128, 128
    comps = (_lib.Component * 1)()
    comps[0].kind = _lib.SKY
    for sidx in range(_lib.NSLOTS):
        comps[0].slot[sidx].theta_index = -1
    comps[0].slot[_lib.P_ADU].theta_index = -2
    desc.n_components, desc.components = 1, comps
    desc.psf_index.theta_index = -1
    assert lib.psfmc_engine_create(ctypes.byref(desc), ctypes.byref(handle)) == 1
    assert b'theta_index' in lib.psfmc_last_error()
    comps[0].slot[_lib.P_ADU].theta_index = 0
    desc.psf_index.theta_index = -7
    assert lib.psfmc_engine_create(ctypes.byref(desc), ctypes.byref(handle)) == 1
    assert b'psf_index' in lib.psfmc_last_error()
    assert handle.value is None
    out = np.zeros(1)
    dbl_p = ctypes.POINTER(ctypes.c_double)
    assert lib.psfmc_lnlike_batch(None, out.ctypes.data_as(dbl_p), 1, 1,
                                  out.ctypes.data_as(dbl_p)) == 1
    lib.psfmc_engine_destroy(None)              # no-op, must not crash


def test_every_engine_entry_point_rejects_a_null_engine(cuda_library):
    """No entry point dereferences a null handle: each one that takes an engine reports
    PSFMC_ERR_INVALID_ARG (1) with a message -- without a GPU, before any CUDA call."""
    from psfmc_b200 import _lib
    lib = _lib.load(cuda_library)
    checked = []
    for name in _lib.EXPORTED_SYMBOLS:
        func = getattr(lib, name)
        argtypes = func.argtypes
        if not argtypes or argtypes[0] is not ctypes.c_void_p or \
                name == 'psfmc_engine_destroy':
            continue
        args = [None]
        for kind in argtypes[1:]:
            args.append(0 if kind in (ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32)
                        else None)
        assert func(*args) == 1, name
        assert lib.psfmc_last_error(), name
        checked.append(name)
    assert len(checked) >= 15, checked
    for must in ('psfmc_lnpost_batch_sharded', 'psfmc_ensemble_run',
                 'psfmc_lnlike_batch_exchange', 'psfmc_peer_connect'):
        assert must in checked


def test_product_fails_loudly_without_the_library(tmp_path, monkeypatch):
    from psfmc_b200 import _lib
    monkeypatch.setenv('PSFMC_B200_LIB', str(tmp_path / 'nope.so'))
    with pytest.raises(ImportError, match='no CPU fallback'):
        _lib.load()


def test_no_gpu_means_an_error_not_a_fallback(cuda_library):
    """On a box without a GPU, creating an engine reports PSFMC_ERR_NO_DEVICE /
    PSFMC_ERR_CUDA; the product never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from conftest import model_from_file
    from psfmc_b200._lib import EngineError
    with pytest.raises(EngineError) as err:
        model_from_file('j0005/model_c1.py', 'fp32', library=cuda_library)
    assert err.value.code in (3, 4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'psfmc_b200')
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, name)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, re.M), name
                assert 'psfmc_oracle' not in text, name
