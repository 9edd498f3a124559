"""ctypes binding of libsigb200.so (C ABI: include/sigb200.h).  Fails loudly when the library is
missing -- there is no CPU fallback behind it."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libsigb200.so')

SIGB_OK = 0
SIGB_EINVAL, SIGB_ESHAPE, SIGB_EINDEX, SIGB_ECRIT, SIGB_EUNSUPPORTED, SIGB_ECUDA, SIGB_ENOMEM, SIGB_ESTATE = \
    -1, -2, -3, -4, -5, -6, -7, -8

(NODE_ZERO, NODE_FIXED, NODE_OSC, NODE_GAIN, NODE_MIX, NODE_RINGMOD, NODE_AMP, NODE_FILTER, NODE_MERGE,
 NODE_GROUPSUM, NODE_PANSUM, NODE_BUFFER, NODE_TAP) = range(13)
WAVE_SINE, WAVE_SQUARE, WAVE_SAWTOOTH, WAVE_TRIANGLE = range(4)
FILT_LOWPASS, FILT_HIGHPASS = range(2)


class SigbNode(ctypes.Structure):
    """struct sigb_node (include/sigb200.h)."""
    _fields_ = [
        ('kind', ctypes.c_int32),
        ('subtype', ctypes.c_int32),
        ('channels', ctypes.c_int32),
        ('inputs', ctypes.c_int32 * 3),
        ('order', ctypes.c_int32),
        ('context', ctypes.c_int32),
        ('rows', ctypes.c_int32),
        ('reserved', ctypes.c_int32),
        ('data_off', ctypes.c_int64),
    ]


class LibraryMissing(RuntimeError):
    pass


_lib = None


def build_hint() -> str:
    return ('libsigb200.so not found at %s -- build it with `sh signals_b200/csrc/build.sh` '
            '(or `python -c "import __graft_entry__ as g; g.build()"`); there is no CPU fallback' % LIB_PATH)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(build_hint())
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    L.sigb_plan_create.argtypes = [ctypes.POINTER(SigbNode), i32, i32, ctypes.POINTER(ctypes.c_double), i64, i32, i32,
                                   ctypes.POINTER(vp)]
    L.sigb_plan_create.restype = ctypes.c_int
    L.sigb_plan_bind_buffer.argtypes = [vp, i32, vp, i64]
    L.sigb_plan_bind_buffer.restype = ctypes.c_int
    L.sigb_plan_bind_buffer_window.argtypes = [vp, i32, vp, i64, i64]
    L.sigb_plan_bind_buffer_window.restype = ctypes.c_int
    L.sigb_render.argtypes = [vp, i64, i32, vp, i64, vp]
    L.sigb_render.restype = ctypes.c_int
    L.sigb_render_host.argtypes = [vp, i64, i32, vp, i64, vp]
    L.sigb_render_host.restype = ctypes.c_int
    L.sigb_render_block.argtypes = [vp, i64, i32, vp, i64]
    L.sigb_render_block.restype = ctypes.c_int
    L.sigb_plan_graph_launches.argtypes = [vp]
    L.sigb_plan_graph_launches.restype = i64
    L.sigb_plan_tap_count.argtypes = [vp]
    L.sigb_plan_tap_count.restype = ctypes.c_int
    L.sigb_plan_read_tap.argtypes = [vp, i32, vp, i64, ctypes.POINTER(i32)]
    L.sigb_plan_read_tap.restype = ctypes.c_int
    L.sigb_state_reset.argtypes = [vp]
    L.sigb_state_reset.restype = ctypes.c_int
    L.sigb_plan_destroy.argtypes = [vp]
    L.sigb_plan_destroy.restype = ctypes.c_int
    L.sigb_plan_describe.argtypes = [vp, ctypes.c_char_p, i64]
    L.sigb_plan_describe.restype = i64
    L.sigb_plan_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    L.sigb_plan_set_option.restype = ctypes.c_int
    L.sigb_set_default_option.argtypes = [ctypes.c_char_p, i64]
    L.sigb_set_default_option.restype = ctypes.c_int
    L.sigb_plan_launch_count.argtypes = [vp]
    L.sigb_plan_launch_count.restype = i64
    L.sigb_plan_last_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    L.sigb_plan_last_kernel_ms.restype = ctypes.c_int
    L.sigb_design_butter.argtypes = [i32, i32, ctypes.c_double, ctypes.POINTER(ctypes.c_double), i32]
    L.sigb_design_butter.restype = ctypes.c_int
    L.sigb_host_alloc.argtypes = [ctypes.POINTER(vp), i64]
    L.sigb_host_alloc.restype = ctypes.c_int
    L.sigb_host_free.argtypes = [vp]
    L.sigb_host_free.restype = ctypes.c_int
    L.sigb_strerror.argtypes = [ctypes.c_int]
    L.sigb_strerror.restype = ctypes.c_char_p
    L.sigb_last_error.argtypes = []
    L.sigb_last_error.restype = ctypes.c_char_p
    L.sigb_abi_version.restype = ctypes.c_int
    L.sigb_device_count.restype = ctypes.c_int
    # test hooks (not part of the public header)
    L.sigb_probe_sin.argtypes = [vp, i32, vp, i32, vp]
    L.sigb_probe_sin.restype = ctypes.c_int
    L.sigb_probe_ratio_q64.argtypes = [ctypes.c_double, i32]
    L.sigb_probe_ratio_q64.restype = ctypes.c_uint64
    L.sigb_probe_frac_q64.argtypes = [ctypes.c_double]
    L.sigb_probe_frac_q64.restype = ctypes.c_uint64
    _lib = L
    return L


def last_error() -> str:
    return lib().sigb_last_error().decode('utf-8', 'replace')


def strerror(status: int) -> str:
    return lib().sigb_strerror(status).decode()
