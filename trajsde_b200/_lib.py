"""ctypes binding of include/trajsde_b200.h.  There is NO fallback: a missing library raises at first use."""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('TRAJSDE_LIB_PATH') or os.path.join(_HERE, 'lib', 'libtrajsde_b200.so')   # override: instrumented debug builds

ABI_VERSION = 9
MODE_EXACT_F32 = 0
MODE_TC_F16 = 1
MODES = {'exact': MODE_EXACT_F32, 'tc_f16': MODE_TC_F16}
STATUS_ADJOINT_RANGE = 1
STATUS_SWEEP_TIMEOUT = 2
HEADS_FLAG_CAT4 = 1

EXPORTED_SYMBOLS = (
    'trajsde_abi_version', 'trajsde_last_error_string', 'trajsde_device_sm_count',
    'trajsde_euler_fwd_workspace_bytes', 'trajsde_euler_fwd',
    'trajsde_euler_bwd_workspace_bytes', 'trajsde_euler_bwd',
    'trajsde_philox_dw', 'trajsde_enc_fwd_workspace_bytes', 'trajsde_enc_fwd',
    'trajsde_enc_bwd_workspace_bytes', 'trajsde_enc_bwd',
    'trajsde_gru_workspace_bytes', 'trajsde_gru_fwd', 'trajsde_gru_bwd',
    'trajsde_heads_workspace_bytes', 'trajsde_heads_fwd', 'trajsde_heads_bwd_workspace_bytes', 'trajsde_heads_bwd',
    'trajsde_aggr_embed_workspace_bytes', 'trajsde_aggr_embed_fwd', 'trajsde_aggr_embed_bwd', 'trajsde_pi_head_fwd',
    'trajsde_l2_loss_workspace_bytes', 'trajsde_l2_loss_fwd', 'trajsde_l2_loss_bwd', 'trajsde_diff_bce_workspace_bytes', 'trajsde_diff_bce',
)

_fp = C.c_void_p  # device pointers travel as integers


class Mlp(C.Structure):
    _fields_ = [(n, _fp) for n in ('w1', 'b1', 'w2', 'b2', 'w3', 'b3')]


class Schedule(C.Structure):
    _fields_ = [('n_steps', C.c_int32), ('n_outputs', C.c_int32), ('step_tab', _fp), ('out_begin', _fp), ('out_w', _fp)]


class Noise(C.Structure):
    _fields_ = [('dw', _fp), ('seed', C.c_uint64), ('row_offset', C.c_uint64), ('step_offset', C.c_uint32),
                ('reserved', C.c_uint32), ('seed_dev', _fp)]


class EulerFwdArgs(C.Structure):
    _fields_ = [('struct_bytes', C.c_uint32), ('mode', C.c_int32), ('rows', C.c_int64), ('dim', C.c_int32),
                ('flags', C.c_int32), ('sched', Schedule), ('drift', Mlp), ('diffusion', Mlp), ('diffusion_alt', Mlp),
                ('alt_mask', _fp), ('noise', Noise), ('y0', _fp), ('y0_row_stride', C.c_int64), ('ys', _fp),
                ('ys_t_stride', C.c_int64), ('ys_row_stride', C.c_int64), ('g_last', _fp), ('states', _fp),
                ('workspace', _fp), ('workspace_bytes', C.c_int64)]


class EulerBwdArgs(C.Structure):
    _fields_ = [('struct_bytes', C.c_uint32), ('mode', C.c_int32), ('rows', C.c_int64), ('dim', C.c_int32),
                ('flags', C.c_int32), ('sched', Schedule), ('drift', Mlp), ('diffusion', Mlp), ('diffusion_alt', Mlp),
                ('alt_mask', _fp), ('noise', Noise), ('states', _fp), ('grad_ys', _fp), ('grad_ys_t_stride', C.c_int64),
                ('grad_ys_row_stride', C.c_int64), ('grad_g_last', _fp), ('grad_y0', _fp), ('grad_drift', Mlp),
                ('grad_diffusion', Mlp), ('grad_diffusion_alt', Mlp), ('status', _fp), ('grad_amax', _fp), ('row_flags', _fp),
                ('workspace', _fp), ('workspace_bytes', C.c_int64)]


class Gru(C.Structure):
    """GRU_Unit parameters (models/utils/ode_utils.py:111-134): three 2-layer nets, nn.Linear layout."""
    _fields_ = [(n, _fp) for n in ('u1', 'ub1', 'u2', 'ub2', 'r1', 'rb1', 'r2', 'rb2', 'n1', 'nb1', 'n2', 'nb2')]


class EncFwdArgs(C.Structure):
    _fields_ = [('struct_bytes', C.c_uint32), ('mode', C.c_int32), ('rows', C.c_int64), ('dim', C.c_int32),
                ('flags', C.c_int32), ('sched', Schedule), ('drift', Mlp), ('diffusion', Mlp), ('diffusion_alt', Mlp),
                ('alt_mask', _fp), ('gru', Gru), ('noise', Noise), ('h0', _fp), ('h0_row_stride', C.c_int64),
                ('aa_out', _fp), ('slot', _fp), ('obs_mask', _fp), ('obs_mask_row_stride', C.c_int64),
                ('latent', _fp), ('g_out', _fp), ('y1_out', _fp), ('workspace', _fp), ('workspace_bytes', C.c_int64)]


class EncBwdArgs(C.Structure):
    _fields_ = [('struct_bytes', C.c_uint32), ('mode', C.c_int32), ('rows', C.c_int64), ('dim', C.c_int32),
                ('flags', C.c_int32), ('sched', Schedule), ('drift', Mlp), ('diffusion', Mlp), ('diffusion_alt', Mlp),
                ('alt_mask', _fp), ('gru', Gru), ('noise', Noise), ('h0', _fp), ('aa_out', _fp), ('n_slots', C.c_int32),
                ('reserved', C.c_int32), ('slot', _fp), ('obs_mask', _fp), ('obs_mask_row_stride', C.c_int64),
                ('latent', _fp), ('y1', _fp), ('grad_latent', _fp), ('grad_g', _fp), ('grad_h0', _fp), ('grad_aa_out', _fp),
                ('grad_drift', Mlp), ('grad_diffusion', Mlp), ('grad_diffusion_alt', Mlp), ('grad_gru', Gru),
                ('status', _fp), ('workspace', _fp), ('workspace_bytes', C.c_int64)]


class GruArgs(C.Structure):
    _fields_ = [('struct_bytes', C.c_uint32), ('mode', C.c_int32), ('rows', C.c_int64), ('dim', C.c_int32),
                ('flags', C.c_int32), ('gru', Gru), ('h_cur', _fp), ('x', _fp), ('mask', _fp), ('h_next', _fp),
                ('grad_h_next', _fp), ('grad_h_cur', _fp), ('grad_x', _fp), ('grad_gru', Gru), ('status', _fp),
                ('workspace', _fp), ('workspace_bytes', C.c_int64)]


class Head(C.Structure):
    """One decoder head (nn.Sequential(Linear, LayerNorm, ReLU, Linear(64, 2)), dec_hivt_nusargo_sde.py:50-61)."""
    _fields_ = [(n, _fp) for n in ('w1', 'b1', 'ln_g', 'ln_b', 'w2', 'b2')]


class HeadsArgs(C.Structure):
    _fields_ = [('struct_bytes', C.c_uint32), ('mode', C.c_int32), ('rows', C.c_int64), ('dim', C.c_int32),
                ('flags', C.c_int32), ('n_t', C.c_int32), ('n_heads', C.c_int32), ('head', Head * 2), ('ln_eps', C.c_float),
                ('min_scale', C.c_float), ('x', _fp), ('x_row_stride', C.c_int64), ('x_t_stride', C.c_int64), ('out', _fp * 2),
                ('workspace', _fp), ('workspace_bytes', C.c_int64)]


class HeadsBwdArgs(C.Structure):
    _fields_ = [('struct_bytes', C.c_uint32), ('mode', C.c_int32), ('rows', C.c_int64), ('dim', C.c_int32),
                ('flags', C.c_int32), ('n_t', C.c_int32), ('n_heads', C.c_int32), ('head', Head * 2), ('ln_eps', C.c_float),
                ('min_scale', C.c_float), ('x', _fp), ('x_row_stride', C.c_int64), ('x_t_stride', C.c_int64), ('grad_out', _fp * 2),
                ('grad_x', _fp), ('gx_row_stride', C.c_int64), ('gx_t_stride', C.c_int64), ('grad_head', Head * 2),
                ('grad_amax', _fp), ('row_flags', _fp), ('workspace', _fp), ('workspace_bytes', C.c_int64)]


class AggrArgs(C.Structure):
    _fields_ = [('struct_bytes', C.c_uint32), ('n_modes', C.c_int32), ('n_actors', C.c_int64), ('global_embed', _fp), ('local_embed', _fp),
                ('w', _fp), ('b', _fp), ('ln_g', _fp), ('ln_b', _fp), ('ln_eps', C.c_float), ('reserved', C.c_float), ('out', _fp),
                ('grad_out', _fp), ('grad_global', _fp), ('grad_local', _fp), ('grad_w', _fp), ('grad_b', _fp), ('grad_ln_g', _fp),
                ('grad_ln_b', _fp), ('workspace', _fp), ('workspace_bytes', C.c_int64)]


class PiArgs(C.Structure):
    _fields_ = [('struct_bytes', C.c_uint32), ('n_modes', C.c_int32), ('n_actors', C.c_int64), ('global_embed', _fp), ('local_embed', _fp),
                ('w1', _fp), ('b1', _fp), ('ln_g', _fp), ('ln_b', _fp), ('w2', _fp), ('b2', _fp), ('ln_eps', C.c_float),
                ('reserved', C.c_float), ('out', _fp)]


class L2Args(C.Structure):
    _fields_ = [('struct_bytes', C.c_uint32), ('n_modes', C.c_int32), ('n_actors', C.c_int64), ('n_t', C.c_int32), ('reserved', C.c_int32),
                ('loc', _fp), ('loc_stride', C.c_int64), ('target', _fp), ('reg_mask', _fp), ('loss', _fp), ('count', _fp), ('best_mode', _fp),
                ('grad_loss', _fp), ('grad_loc', _fp), ('grad_loc_stride', C.c_int64), ('workspace', _fp), ('workspace_bytes', C.c_int64)]


class BceArgs(C.Structure):
    _fields_ = [('struct_bytes', C.c_uint32), ('reserved', C.c_int32), ('n_in', C.c_int64), ('n_out', C.c_int64), ('diff_in', _fp),
                ('diff_out', _fp), ('loss', _fp), ('grad_in', _fp), ('grad_out', _fp), ('workspace', _fp), ('workspace_bytes', C.c_int64)]


_lock = threading.Lock()
_lib = None


class TrajsdeError(RuntimeError):
    pass


def lib():
    """Load libtrajsde_b200.so (built by `python -m trajsde_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise TrajsdeError(f"{LIB_PATH} is missing: build it with `python -m trajsde_b200.build`. "
                               "trajsde_b200 has no CPU / PyTorch fallback by design.")
        L = C.CDLL(LIB_PATH)
        L.trajsde_abi_version.restype = C.c_int
        L.trajsde_last_error_string.restype = C.c_char_p
        L.trajsde_device_sm_count.restype = C.c_int
        for name in ('trajsde_euler_fwd_workspace_bytes', 'trajsde_euler_bwd_workspace_bytes',
                     'trajsde_enc_fwd_workspace_bytes', 'trajsde_enc_bwd_workspace_bytes'):
            if hasattr(L, name):
                getattr(L, name).restype = C.c_int64
                getattr(L, name).argtypes = [C.c_int32, C.c_int64, C.c_int32, C.c_int32]
        L.trajsde_euler_fwd.restype = C.c_int
        L.trajsde_euler_fwd.argtypes = [C.POINTER(EulerFwdArgs), C.c_void_p]
        L.trajsde_euler_bwd.restype = C.c_int
        L.trajsde_euler_bwd.argtypes = [C.POINTER(EulerBwdArgs), C.c_void_p]
        L.trajsde_philox_dw.restype = C.c_int
        L.trajsde_philox_dw.argtypes = [C.POINTER(Schedule), C.POINTER(Noise), C.c_int64, C.c_void_p, C.c_void_p]
        if hasattr(L, 'trajsde_enc_fwd'):
            L.trajsde_enc_fwd.restype = C.c_int
            L.trajsde_enc_fwd.argtypes = [C.POINTER(EncFwdArgs), C.c_void_p]
        if hasattr(L, 'trajsde_enc_bwd'):
            L.trajsde_enc_bwd.restype = C.c_int
            L.trajsde_enc_bwd.argtypes = [C.POINTER(EncBwdArgs), C.c_void_p]
        if hasattr(L, 'trajsde_gru_fwd'):
            L.trajsde_gru_workspace_bytes.restype = C.c_int64
            L.trajsde_gru_workspace_bytes.argtypes = [C.c_int32, C.c_int64]
            for f in (L.trajsde_gru_fwd, L.trajsde_gru_bwd):
                f.restype = C.c_int
                f.argtypes = [C.POINTER(GruArgs), C.c_void_p]
        L.trajsde_heads_workspace_bytes.restype = C.c_int64
        L.trajsde_heads_workspace_bytes.argtypes = [C.c_int32]
        L.trajsde_heads_fwd.restype = C.c_int
        L.trajsde_heads_fwd.argtypes = [C.POINTER(HeadsArgs), C.c_void_p]
        L.trajsde_heads_bwd_workspace_bytes.restype = C.c_int64
        L.trajsde_heads_bwd_workspace_bytes.argtypes = [C.c_int32]
        L.trajsde_heads_bwd.restype = C.c_int
        L.trajsde_heads_bwd.argtypes = [C.POINTER(HeadsBwdArgs), C.c_void_p]
        L.trajsde_aggr_embed_workspace_bytes.restype = C.c_int64
        L.trajsde_aggr_embed_workspace_bytes.argtypes = [C.c_int64, C.c_int64]
        for f in (L.trajsde_aggr_embed_fwd, L.trajsde_aggr_embed_bwd):
            f.restype = C.c_int
            f.argtypes = [C.POINTER(AggrArgs), C.c_void_p]
        L.trajsde_pi_head_fwd.restype = C.c_int
        L.trajsde_pi_head_fwd.argtypes = [C.POINTER(PiArgs), C.c_void_p]
        L.trajsde_l2_loss_workspace_bytes.restype = C.c_int64
        L.trajsde_l2_loss_workspace_bytes.argtypes = [C.c_int64]
        for f in (L.trajsde_l2_loss_fwd, L.trajsde_l2_loss_bwd):
            f.restype = C.c_int
            f.argtypes = [C.POINTER(L2Args), C.c_void_p]
        L.trajsde_diff_bce_workspace_bytes.restype = C.c_int64
        L.trajsde_diff_bce_workspace_bytes.argtypes = []
        L.trajsde_diff_bce.restype = C.c_int
        L.trajsde_diff_bce.argtypes = [C.POINTER(BceArgs), C.c_void_p]
        v = L.trajsde_abi_version()
        if v != ABI_VERSION:
            raise TrajsdeError(f"ABI version mismatch: library {v}, binding {ABI_VERSION}")
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc < 0:
        msg = lib().trajsde_last_error_string()
        raise TrajsdeError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")
    return rc
