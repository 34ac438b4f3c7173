"""
ctypes binding of liblime_b200.so (C ABI: include/lime_b200.h).

There is NO fallback: if the shared library is missing, or no CUDA device is visible when
a compute entry point is called, the call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'liblime_b200.so')

c_int, c_ll, c_dbl, c_vp = C.c_int, C.c_longlong, C.c_double, C.c_void_p
P = C.POINTER

# name -> (restype, [argtypes]);  must list every symbol include/lime_b200.h declares
SIGNATURES = {
    'limeb200_version': (c_int, []),
    'limeb200_last_error': (C.c_char_p, []),
    'limeb200_device_info': (c_int, [c_int, P(c_int), P(c_int), P(c_int), P(c_ll), P(c_ll)]),
    'limeb200_qme_create': (c_int, [P(c_vp), c_int, c_int]),
    'limeb200_qme_destroy': (c_int, [c_vp]),
    'limeb200_qme_set_generator_dense': (c_int, [c_vp, c_vp, c_int]),
    'limeb200_qme_set_generator_csr': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int]),
    'limeb200_qme_add_sandwich_dense': (c_int, [c_vp, c_vp, c_vp, c_int]),
    'limeb200_qme_add_sandwich_csr': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int]),
    'limeb200_qme_set_right_generator_dense': (c_int, [c_vp, c_vp, c_int]),
    'limeb200_qme_add_drive_dense': (c_int, [c_vp, c_vp, c_vp]),
    'limeb200_qme_set_observables': (c_int, [c_vp, c_vp, c_int]),
    'limeb200_qme_set_step_values': (c_int, [c_vp, c_int]),
    'limeb200_qme_set_path': (c_int, [c_vp, c_int]),
    'limeb200_qme_finalize': (c_int, [c_vp]),
    'limeb200_qme_get_path': (c_int, [c_vp]),
    'limeb200_qme_get_info': (c_int, [c_vp, c_vp, c_vp]),
    'limeb200_qme_run': (c_int, [c_vp, c_vp, c_int, c_dbl, c_int, c_vp, c_vp, c_vp, c_int, c_vp]),
    'limeb200_qme_rhs': (c_int, [c_vp, c_vp, c_vp, c_int, c_vp]),
    'limeb200_qme_last_launches': (c_ll, [c_vp]),
    'limeb200_zgemm': (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_vp]),
    'limeb200_rkf45_stage': (c_int, [c_int, c_ll, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_dbl, c_vp, c_vp]),
    'limeb200_rkf45_error': (c_int, [c_ll, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_vp, c_vp, c_vp]),
    'limeb200_rkf45_hinit': (c_int, [c_ll, c_vp, c_vp, c_dbl, c_dbl, c_dbl, c_vp, c_vp]),
    'limeb200_rkf45_axpy': (c_int, [c_ll, c_dbl, c_vp, c_vp, c_vp]),
    'limeb200_liouville_rk4_csr': (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_int,
                                           c_dbl, c_int, c_vp]),
    'limeb200_heom_count_states': (c_ll, [c_vp, c_int, c_int]),
    'limeb200_heom_build_tables': (c_int, [c_vp, c_int, c_int, c_ll, c_vp, c_vp, c_vp]),
    'limeb200_heom_create': (c_int, [P(c_vp), c_int, c_int, c_int, c_int, c_ll, c_vp, c_vp, c_vp, c_vp, c_vp,
                                     c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, c_ll]),
    'limeb200_heom_create_batched': (c_int, [P(c_vp), c_int, c_int, c_int, c_int, c_ll, c_vp, c_vp, c_vp, c_vp, c_vp,
                                             c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, c_ll]),
    'limeb200_heom_destroy': (c_int, [c_vp]),
    'limeb200_heom_set_path': (c_int, [c_vp, c_int]),
    'limeb200_heom_get_path': (c_int, [c_vp]),
    'limeb200_heom_run': (c_int, [c_vp, c_vp, c_int, c_dbl, c_int, c_vp, c_int, c_vp, c_vp, c_int, c_vp]),
    'limeb200_heom_rhs': (c_int, [c_vp, c_vp, c_vp, c_int, c_vp]),
    'limeb200_heom_stage': (c_int, [c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_dbl, c_vp]),
    'limeb200_heom_last_launches': (c_ll, [c_vp]),
    'limeb200_memcpy_d2d': (c_int, [c_vp, c_vp, c_ll, c_vp]),
    'limeb200_peer_alloc': (c_int, [c_int, c_ll, P(c_vp), c_vp]),
    'limeb200_peer_open': (c_int, [c_int, c_vp, P(c_vp)]),
    'limeb200_peer_close': (c_int, [c_int, c_vp]),
    'limeb200_peer_free': (c_int, [c_int, c_vp]),
    'limeb200_heom_persist_grid': (c_int, [c_vp, c_int]),
    'limeb200_heom_run_sharded': (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_dbl, c_int, C.c_uint, c_vp]),
    'limeb200_heom_flow_supported': (c_int, [c_vp]),
    'limeb200_heom_flow_pack': (c_int, [c_vp, c_vp, c_vp, C.c_ulonglong, c_vp]),
    'limeb200_heom_flow_unpack': (c_int, [c_vp, c_vp, C.c_ulonglong, c_vp, c_vp]),
    'limeb200_heom_flow_run_sharded': (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_dbl, c_int, C.c_ulonglong, c_vp]),
    'limeb200_heom_run_sharded_halo': (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_dbl, c_int, C.c_ulonglong, c_vp]),
    'limeb200_heom_sharded_error': (c_int, [c_vp, c_vp]),
    'limeb200_heom_dl_euler': (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_int, c_dbl, c_int, c_vp, c_vp]),
    'limeb200_sos_factor': (c_int, [c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp]),
    'limeb200_sos_factor_time': (c_int, [c_vp, c_int, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp]),
    'limeb200_sos_outer': (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_dbl, c_int, c_vp, c_vp]),
    'limeb200_etpa_reduce': (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_int, c_vp, c_vp]),
    'limeb200_sos_tpa2d': (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_int, c_vp, c_int, c_vp, c_int,
                                   c_int, c_vp, c_vp]),
}

_lib = None


class LimeB200Error(RuntimeError):
    pass


def lib():
    """load liblime_b200.so (once); raise if it has not been built"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LimeB200Error(
                'liblime_b200.so not found at %s -- build it with `python -c "import __graft_entry__ as g; g.build()"` '
                'or `make -C lime_b200/csrc`; there is no CPU fallback' % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc is not None and rc < 0:
        raise LimeB200Error('liblime_b200: %s (code %d)' % (lib().limeb200_last_error().decode(), rc))
    return rc


def hptr(a):
    """host pointer of a C-contiguous numpy array (or None)"""
    if a is None:
        return None
    assert a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(c_vp)
