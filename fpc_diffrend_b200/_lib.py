"""ctypes loader for libfpc_b200.so — the C-ABI declared in include/fpc_b200.h.

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# FPC_B200_LIB selects another build of the SAME library (kernel-variant experiments, scripts/exp_variants.py)
LIB_PATH = os.environ.get('FPC_B200_LIB') or os.path.join(_HERE, 'libfpc_b200.so')
HEADER_PATH = os.path.join(_HERE, '..', 'include', 'fpc_b200.h')

ABI_VERSION = 4        # FPC_B200_ABI_VERSION of include/fpc_b200.h these signatures were written against

_lib = None

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_F = _c.c_float
_Z = _c.c_size_t
_L = _c.c_longlong



class AdamFusedArgs(_c.Structure):
    """fpc_adam_fused_args of include/fpc_b200.h (the optimiser step that rides in fpc_geometry_bwd)."""
    _fields_ = [('params', _P), ('m', _P), ('v', _P), ('step_count', _P), ('optimize_pose', _I), ('quat_mode', _I),
                ('lr_w', _F), ('lr_t', _F), ('lr_q', _F), ('b1', _F), ('b2', _F), ('eps', _F), ('lr_ramp', _F), ('max_iter', _F)]


# name -> (restype, argtypes); kept in the order of include/fpc_b200.h
SIGNATURES = {
    'fpc_abi_version': (_I, []),
    'fpc_last_error': (_c.c_char_p, []),
    'fpc_check_device': (_I, []),
    'fpc_rasterize_scratch_bytes': (_Z, [_I, _I, _I, _I]),
    'fpc_rasterize_fwd': (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    'fpc_rasterize_bwd': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    'fpc_interpolate_fwd': (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P, _P]),
    'fpc_interpolate_bwd': (_I, [_P, _I, _I, _I, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    'fpc_texture_linear_fwd': (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _P, _P]),
    'fpc_texture_linear_bwd': (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _P, _P, _P]),
    'fpc_rasterize_bwd_db': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    'fpc_interpolate_da_fwd': (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    'fpc_interpolate_da_bwd': (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    'fpc_texture_mip_levels': (_I, [_I, _I, _I]),
    'fpc_texture_mip_floats': (_Z, [_I, _I, _I, _I, _I]),
    'fpc_texture_mip_build': (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    'fpc_texture_mip_fwd': (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    'fpc_texture_mip_bwd': (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P]),
    'fpc_topology_scratch_bytes': (_Z, [_I]),
    'fpc_topology_build': (_I, [_P, _I, _I, _P, _P, _Z, _P]),
    'fpc_antialias_fwd': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    'fpc_antialias_bwd': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    'fpc_blend_fwd': (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    'fpc_blend_fwd_ex': (_I, [_P, _P, _P, _I, _I, _I, _F, _I, _P, _P]),
    'fpc_basis_code_fwd': (_I, [_P, _P, _P, _I, _I, _P, _P, _P]),
    'fpc_basis_grad': (_I, [_P, _P, _I, _I, _I, _F, _P, _P]),
    'fpc_basis_code_bwd': (_I, [_P, _P, _P, _P, _I, _I, _F, _P, _P, _P]),
    'fpc_l2_reg_scratch_bytes': (_Z, [_L]),
    'fpc_l2_reg_fwd_bwd': (_I, [_P, _I, _L, _F, _P, _P, _P, _F, _P, _P, _Z, _P]),
    'fpc_blend_bwd_scratch_bytes': (_Z, [_I, _I, _I]),
    'fpc_blend_bwd': (_I, [_P, _P, _I, _I, _I, _P, _P, _Z, _P]),
    'fpc_blend_tc_supported': (_I, [_I, _I, _I]),
    'fpc_blend_fwd_tc': (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    'fpc_blend_bwd_tc_scratch_bytes': (_Z, [_I, _I, _I]),
    'fpc_blend_bwd_tc': (_I, [_P, _P, _I, _I, _I, _P, _P, _Z, _P]),
    'fpc_pose_mvp_fwd': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _P, _P]),
    'fpc_pose_mvp_bwd': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P]),
    'fpc_pose_cam_bwd': (_I, [_P] * 7 + [_I, _I, _P, _P, _P]),
    'fpc_project_fwd': (_I, [_P, _P, _I, _I, _I, _P, _P]),
    'fpc_project_bwd_scratch_bytes': (_Z, [_I, _I, _I]),
    'fpc_project_bwd': (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _Z, _P]),
    'fpc_geometry_fused_supported': (_I, [_I, _I, _I, _I]),
    'fpc_geometry_fwd': (_I, [_P] * 9 + [_I] * 4 + [_P] * 4),
    'fpc_geometry_bwd_scratch_bytes': (_Z, [_I, _I, _I, _I]),
    'fpc_geometry_bwd_counter_bytes': (_Z, [_I, _I]),
    'fpc_geometry_bwd': (_I, [_P] * 11 + [_I] * 4 + [_P] * 8 + [_Z, _P]),
    'fpc_image_loss_scratch_bytes': (_Z, [_I, _I, _I, _I]),
    'fpc_image_loss_fwd_bwd': (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _F, _I, _P, _P, _P, _P, _Z, _P]),
    'fpc_render_loss_fused_scratch_bytes': (_Z, [_I, _I, _I, _I]),
    'fpc_render_loss_fused': (_I, [_P, _P, _P, _P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'fpc_render_loss_fused_aa': (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'fpc_blend_fwd_bcast': (_I, [_P, _P, _P, _I, _I, ctypes.c_longlong, _P, _I, _P]),
    'fpc_peer_store_rows': (_I, [_P, _P, _I, _I, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _P]),
    'fpc_peer_sum_rows': (_I, [_P, _I, _I, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _P, _P]),
    'fpc_peer_sum': (_I, [_P, _I, ctypes.c_longlong, _P, _P]),
    'fpc_vertex_adjacency_scratch_bytes': (_Z, [_I]),
    'fpc_vertex_adjacency_build': (_I, [_P, _I, _I, _P, _P, _P, _Z, _P]),
    'fpc_raster_bin_px': (_I, []),
    'fpc_rasterize_clip_pieces': (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    'fpc_render_loss_fused_band': (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _F, _F, _I, _I, _I, _I,
                                        _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'fpc_mesh_reg_scratch_bytes': (_Z, [_I, _I, _I]),
    'fpc_mesh_reg_fwd_bwd': (_I, [_P, _I, _I, _P, _P, _I, _P, _I, _F, _F, _F, _F, _P, _P, _P, _I, _P, _Z, _P]),
    'fpc_adam_step': (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _F, _P, _P]),
    'fpc_adam_step_from': (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _F, _P, _F, _P]),
    'fpc_adam_advance': (_I, [_P, _P]),
    'fpc_quat_renorm': (_I, [_P, _I, _I, _P]),
    'fpc_adam_fused': (_I, [_P] * 4 + [_I] * 3 + [_F] * 8 + [_I, _P, _P]),
}


def header_symbols():
    """Function names declared in include/fpc_b200.h (used by the CPU tests to check the exports)."""
    with open(HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(fpc_[a-z0-9_]+)\s*\(', src)))


def load():
    """Load the shared library (no GPU needed for loading); raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'libfpc_b200.so not found at %s — build it with `python -m fpc_diffrend_b200.build`; '
                'there is no CPU or PyTorch fallback for the fit hot path' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.fpc_abi_version() != ABI_VERSION:
            raise RuntimeError('libfpc_b200.so at %s has ABI version %d, this package expects %d — rebuild it with '
                               '`python -m fpc_diffrend_b200.build --force`' % (LIB_PATH, lib.fpc_abi_version(), ABI_VERSION))
        _lib = lib
    return _lib


def check(status):
    if status != 0:
        raise RuntimeError('fpc_b200: %s' % load().fpc_last_error().decode())


def call(name, *args):
    check(getattr(load(), name)(*args))


def vertex_adjacency(tri, n_vertices):
    """vadj_off [V+1], vadj_item [3T] (int32, on tri's device): the vertex -> (triangle, corner) adjacency the fused kernels'
    gradient gather needs (fpc_vertex_adjacency_build; once per mesh).  tri: contiguous int32 CUDA tensor [T,3]."""
    import torch
    T, V = int(tri.shape[0]), int(n_vertices)
    off = torch.empty(V + 1, dtype=torch.int32, device=tri.device)
    item = torch.empty(3 * T, dtype=torch.int32, device=tri.device)
    sc = torch.empty(int(load().fpc_vertex_adjacency_scratch_bytes(V)), dtype=torch.uint8, device=tri.device)
    with torch.cuda.device(tri.device):
        call('fpc_vertex_adjacency_build', ctypes.c_void_p(tri.data_ptr()), T, V, ctypes.c_void_p(off.data_ptr()), ctypes.c_void_p(item.data_ptr()),
             ctypes.c_void_p(sc.data_ptr()), sc.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.current_stream().synchronize()      # sc is freed on return
    return off, item
