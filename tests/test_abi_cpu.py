"""The C-ABI library loads and exports every symbol include/*.h declares; argument errors are reported, not aborted."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT
from trajsde_b200 import _lib


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'trajsde_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(trajsde_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    syms = header_symbols()
    assert 'trajsde_euler_fwd' in syms and 'trajsde_euler_bwd' in syms and 'trajsde_philox_dw' in syms
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/trajsde_b200.h but not exported"
    assert L.trajsde_abi_version() == _lib.ABI_VERSION


def test_struct_sizes_match_header(tmp_path):
    """ctypes mirrors must have the C sizes (compiled with gcc from the header)."""
    import subprocess
    src = tmp_path / 'sz.c'
    src.write_text('#include "trajsde_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(TrajsdeEulerFwdArgs),sizeof(TrajsdeEulerBwdArgs),sizeof(TrajsdeSchedule),sizeof(TrajsdeNoise),'
                   'sizeof(TrajsdeMlp),sizeof(TrajsdeEncFwdArgs),sizeof(TrajsdeGru),sizeof(TrajsdeEncBwdArgs),'
                   'sizeof(TrajsdeGruGrad),sizeof(TrajsdeGruArgs),sizeof(TrajsdeHead),sizeof(TrajsdeHeadsArgs),sizeof(TrajsdeHeadsBwdArgs),'
                   'sizeof(TrajsdeAggrArgs),sizeof(TrajsdePiArgs),sizeof(TrajsdeL2Args),sizeof(TrajsdeBceArgs));return 0;}\n')
    exe = tmp_path / 'sz'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(_lib.EulerFwdArgs), C.sizeof(_lib.EulerBwdArgs), C.sizeof(_lib.Schedule),
                     C.sizeof(_lib.Noise), C.sizeof(_lib.Mlp), C.sizeof(_lib.EncFwdArgs), C.sizeof(_lib.Gru),
                     C.sizeof(_lib.EncBwdArgs), C.sizeof(_lib.Gru), C.sizeof(_lib.GruArgs), C.sizeof(_lib.Head),
                     C.sizeof(_lib.HeadsArgs), C.sizeof(_lib.HeadsBwdArgs), C.sizeof(_lib.AggrArgs), C.sizeof(_lib.PiArgs),
                     C.sizeof(_lib.L2Args), C.sizeof(_lib.BceArgs)]


def test_invalid_arguments_return_status_and_message():
    L = _lib.lib()
    assert L.trajsde_euler_fwd(None, None) == -1
    a = _lib.EulerFwdArgs()
    a.struct_bytes = 8
    assert L.trajsde_euler_fwd(C.byref(a), None) == -1 and b'ABI mismatch' in L.trajsde_last_error_string()
    a.struct_bytes = C.sizeof(a)
    a.dim = 32
    assert L.trajsde_euler_fwd(C.byref(a), None) == -2 and b'dim 32' in L.trajsde_last_error_string()
    a.dim = 64
    assert L.trajsde_euler_fwd(C.byref(a), None) == -1 and b'schedule' in L.trajsde_last_error_string()
    b = _lib.EulerBwdArgs()
    assert L.trajsde_euler_bwd(C.byref(b), None) == -1
    e = _lib.EncBwdArgs()
    assert L.trajsde_enc_bwd(C.byref(e), None) == -1 and b'ABI mismatch' in L.trajsde_last_error_string()
    e.struct_bytes, e.dim, e.mode = C.sizeof(e), 64, _lib.MODE_EXACT_F32
    assert L.trajsde_enc_bwd(C.byref(e), None) == -2 and b'TC_F16' in L.trajsde_last_error_string()
    assert L.trajsde_enc_bwd_workspace_bytes(_lib.MODE_TC_F16, 1000, 21, 1) > 0
    assert L.trajsde_euler_fwd_workspace_bytes(99, 10, 10, 0) == -2
    h = _lib.HeadsArgs()
    assert L.trajsde_heads_fwd(C.byref(h), None) == -1 and b'ABI mismatch' in L.trajsde_last_error_string()
    h.struct_bytes, h.dim, h.mode, h.n_heads = C.sizeof(h), 64, _lib.MODE_TC_F16, 3
    assert L.trajsde_heads_fwd(C.byref(h), None) == -1 and b'n_heads' in L.trajsde_last_error_string()
    assert L.trajsde_heads_workspace_bytes(_lib.MODE_TC_F16) > 0 and L.trajsde_heads_workspace_bytes(_lib.MODE_EXACT_F32) == -2
    with pytest.raises(_lib.TrajsdeError):
        _lib.check(-2, "x")


def test_no_cpu_fallback():
    """Product path must fail loudly off-GPU (north star: no CPU fallback)."""
    import trajsde_b200 as tb
    from helpers import DecoderSDE
    sde = DecoderSDE()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tb.sdeint(sde, torch.zeros(4, 64), torch.linspace(0, 1, 11), dt=0.1, method='euler')


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'trajsde_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), f"{f} imports oracle/"
