"""nvdiffrast-compatible front end of the B200 kernels — the drop-in for `import nvdiffrast.torch as dr`.

The reference calls exactly these:
    dr.RasterizeGLContext(device='cuda')                                  fit.py:484
    dr.rasterize(glctx, pos_clip, pos_idx, resolution=(H, W))             fit.py:151
    dr.interpolate(attr[None], rast_out, idx)                             fit.py:157
    dr.texture(tex[None], texc, filter_mode='linear')                     fit.py:158
    dr.antialias(colour, rast_out, pos_clip, pos_idx)                     fit.py:160
Signatures, defaults, tensor layouts (rast = (u, v, z/w, tri_id+1), row 0 = bottom) and the error
behaviour (RuntimeError naming the offending argument) follow upstream nvdiffrast v0.3.x.  Every op is a
torch.autograd.Function whose forward/backward call the C-ABI of include/fpc_b200.h on the current CUDA
stream.  There is no CPU or PyTorch fallback: tensors must live on a CUDA device of compute capability 10.x.
"""
import ctypes

import torch

from . import _lib

_checked_devices = set()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _check_tensor(name, t, dtype, ndim=None):
    _require(isinstance(t, torch.Tensor), '%s must be a torch.Tensor' % name)
    _require(t.is_cuda, '%s must reside on a CUDA device (fpc_diffrend_b200 has no CPU path)' % name)
    _require(t.dtype == dtype, '%s must have dtype %s (got %s)' % (name, dtype, t.dtype))
    if ndim is not None:
        dims = ndim if isinstance(ndim, (tuple, list)) else (ndim,)
        _require(t.dim() in dims, '%s must have %s dimensions (got shape %s)' % (name, ' or '.join(map(str, dims)), tuple(t.shape)))


def _check_device(t):
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx not in _checked_devices:
        with torch.cuda.device(idx):
            _lib.call('fpc_check_device')
        _checked_devices.add(idx)


class _Scratch:
    """Grow-only device scratch buffer (per context / per device)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        return self.buf


# ---------------------------------------------------------------------------------------------------------
# contexts
# ---------------------------------------------------------------------------------------------------------

class RasterizeCudaContext:
    """Rasterizer state (scratch for the bin lists).  Not to be shared across concurrently running streams."""

    def __init__(self, device=None):
        if device is None:
            self.device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else None
        else:
            self.device = torch.device(device)
            _require(self.device.type == 'cuda', 'RasterizeCudaContext: device must be a CUDA device')
        self.output_db = True
        self._scratch = _Scratch()


class RasterizeGLContext(RasterizeCudaContext):
    """Alias kept so that `dr.RasterizeGLContext(device='cuda')` (fit.py:484) runs unchanged; the rasterizer
    behind it is the CUDA one (no OpenGL on a headless B200 box)."""

    def __init__(self, output_db=True, mode='automatic', device=None):
        _require(mode in ('automatic', 'manual'), "RasterizeGLContext: mode must be 'automatic' or 'manual'")
        super().__init__(device=device)
        self.output_db = bool(output_db)

    def set_context(self):
        pass

    def release_context(self):
        pass


# ---------------------------------------------------------------------------------------------------------
# rasterize
# ---------------------------------------------------------------------------------------------------------

class _rasterize_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, glctx, pos, tri, resolution, want_db):
        N, V, _ = pos.shape
        T = tri.shape[0]
        H, W = resolution
        rast = torch.empty((N, H, W, 4), dtype=torch.float32, device=pos.device)
        rast_db = torch.empty((N, H, W, 4), dtype=torch.float32, device=pos.device) if want_db else None
        nbytes = _lib.load().fpc_rasterize_scratch_bytes(N, T, H, W)
        scratch = glctx._scratch.get(nbytes, pos.device)
        with torch.cuda.device(pos.device):
            _lib.call('fpc_rasterize_fwd', _ptr(pos), _ptr(tri), N, V, T, H, W, _ptr(rast), _ptr(rast_db),
                      _ptr(scratch), scratch.numel(), _stream())
        ctx.save_for_backward(pos, tri, rast)
        if rast_db is None:
            rast_db = torch.empty((N, H, W, 0), dtype=torch.float32, device=pos.device)
        ctx.mark_non_differentiable(rast_db)
        return rast, rast_db

    @staticmethod
    def backward(ctx, dy, ddb):
        pos, tri, rast = ctx.saved_tensors
        N, V, _ = pos.shape
        _, H, W, _ = rast.shape
        g_pos = torch.empty_like(pos)
        dy = dy.contiguous()
        with torch.cuda.device(pos.device):
            _lib.call('fpc_rasterize_bwd', _ptr(pos), _ptr(tri), _ptr(rast), _ptr(dy), N, V, tri.shape[0], H, W,
                      _ptr(g_pos), _stream())
        return None, g_pos, None, None, None


def rasterize(glctx, pos, tri, resolution, ranges=None, grad_db=True):
    """pos [N,V,4] clip space, tri [T,3] int32, resolution (H, W) -> (rast [N,H,W,4], rast_db [N,H,W,4])."""
    _require(isinstance(glctx, RasterizeCudaContext), 'glctx must be a RasterizeCudaContext / RasterizeGLContext')
    _require(ranges is None, 'rasterize: range mode (ranges != None) is not supported; use instanced mode pos [N,V,4]')
    _check_tensor('pos', pos, torch.float32, 3)
    _check_tensor('tri', tri, torch.int32, 2)
    _require(pos.shape[2] == 4 and pos.shape[0] > 0 and pos.shape[1] > 0, 'pos must have shape [>0, >0, 4]')
    _require(tri.shape[1] == 3 and tri.shape[0] > 0, 'tri must have shape [>0, 3]')
    _require(len(resolution) == 2 and int(resolution[0]) > 0 and int(resolution[1]) > 0, 'resolution must be [>0, >0]')
    _require(pos.device == tri.device, 'pos and tri must reside on the same device')
    _check_device(pos)
    # grad_db only controls whether gradients flow into rast_db upstream; rast_db is not differentiable here
    # (mip path is out of scope for this round).
    return _rasterize_func.apply(glctx, pos.contiguous(), tri.contiguous(), (int(resolution[0]), int(resolution[1])),
                                 glctx.output_db)


# ---------------------------------------------------------------------------------------------------------
# interpolate
# ---------------------------------------------------------------------------------------------------------

class _interpolate_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attr, rast, tri):
        Na, Vt, A = attr.shape
        N, H, W, _ = rast.shape
        out = torch.empty((N, H, W, A), dtype=torch.float32, device=rast.device)
        with torch.cuda.device(rast.device):
            _lib.call('fpc_interpolate_fwd', _ptr(attr), Na, Vt, A, _ptr(rast), _ptr(tri), N, tri.shape[0], H, W,
                      _ptr(out), _stream())
        ctx.save_for_backward(attr, rast, tri)
        return out

    @staticmethod
    def backward(ctx, dy):
        attr, rast, tri = ctx.saved_tensors
        Na, Vt, A = attr.shape
        N, H, W, _ = rast.shape
        g_attr = torch.empty_like(attr)
        g_rast = torch.empty_like(rast)
        dy = dy.contiguous()
        with torch.cuda.device(rast.device):
            _lib.call('fpc_interpolate_bwd', _ptr(attr), Na, Vt, A, _ptr(rast), _ptr(tri), _ptr(dy), N, tri.shape[0],
                      H, W, _ptr(g_attr), _ptr(g_rast), _stream())
        return g_attr, g_rast, None


def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
    """attr [1|N,V,A], rast [N,H,W,4], tri [T,3] -> (out [N,H,W,A], out_da [N,H,W,0])."""
    _require(diff_attrs is None or (isinstance(diff_attrs, (list, tuple)) and len(diff_attrs) == 0),
             'interpolate: attribute pixel differentials (diff_attrs) are not supported (mip path, SURVEY §8(f) rank 4)')
    _check_tensor('attr', attr, torch.float32, 3)
    _check_tensor('rast', rast, torch.float32, 4)
    _check_tensor('tri', tri, torch.int32, 2)
    _require(rast.shape[3] == 4 and min(rast.shape) > 0, 'rast must have shape [>0, >0, >0, 4]')
    _require(tri.shape[1] == 3 and tri.shape[0] > 0, 'tri must have shape [>0, 3]')
    _require(attr.shape[0] in (1, rast.shape[0]) and attr.shape[1] > 0 and attr.shape[2] > 0,
             'attr must have shape [1 or minibatch, >0, >0]')
    _require(attr.device == rast.device == tri.device, 'attr, rast and tri must reside on the same device')
    _check_device(rast)
    out = _interpolate_func.apply(attr.contiguous(), rast.contiguous(), tri.contiguous())
    out_da = torch.empty(tuple(out.shape[:3]) + (0,), dtype=torch.float32, device=out.device)
    return out, out_da


# ---------------------------------------------------------------------------------------------------------
# texture
# ---------------------------------------------------------------------------------------------------------

class _texture_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tex, uv):
        Nt, Ht, Wt, C = tex.shape
        N, H, W, _ = uv.shape
        out = torch.empty((N, H, W, C), dtype=torch.float32, device=uv.device)
        with torch.cuda.device(uv.device):
            _lib.call('fpc_texture_linear_fwd', _ptr(tex), Nt, Ht, Wt, C, _ptr(uv), N, H, W, _ptr(out), _stream())
        ctx.save_for_backward(tex, uv)
        return out

    @staticmethod
    def backward(ctx, dy):
        tex, uv = ctx.saved_tensors
        Nt, Ht, Wt, C = tex.shape
        N, H, W, _ = uv.shape
        g_tex = torch.empty_like(tex) if ctx.needs_input_grad[0] else None
        g_uv = torch.empty_like(uv)
        dy = dy.contiguous()
        with torch.cuda.device(uv.device):
            _lib.call('fpc_texture_linear_bwd', _ptr(tex), Nt, Ht, Wt, C, _ptr(uv), _ptr(dy), N, H, W,
                      _ptr(g_tex), _ptr(g_uv), _stream())
        return g_tex, g_uv


def texture(tex, uv, uv_da=None, mip_level_bias=None, mip=None, filter_mode='auto', boundary_mode='wrap',
            max_mip_level=None):
    """tex [1|N,Ht,Wt,C], uv [N,H,W,2] -> [N,H,W,C]; filter_mode 'linear' (or 'auto' without uv_da), boundary 'wrap'."""
    if filter_mode == 'auto':
        filter_mode = 'linear-mipmap-linear' if (uv_da is not None or mip_level_bias is not None) else 'linear'
    _require(filter_mode == 'linear',
             "texture: only filter_mode='linear' is supported (got %r; mip modes are SURVEY §8(f) rank 4)" % (filter_mode,))
    _require(boundary_mode == 'wrap', "texture: only boundary_mode='wrap' is supported (got %r)" % (boundary_mode,))
    _require(uv_da is None and mip_level_bias is None and mip is None, 'texture: mip inputs are not supported with filter_mode=linear')
    _check_tensor('tex', tex, torch.float32, 4)
    _check_tensor('uv', uv, torch.float32, 4)
    _require(uv.shape[3] == 2 and min(uv.shape) > 0, 'uv must have shape [>0, >0, >0, 2]')
    _require(min(tex.shape) > 0 and tex.shape[0] in (1, uv.shape[0]), 'tex must have shape [1 or minibatch, >0, >0, >0]')
    _require(tex.device == uv.device, 'tex and uv must reside on the same device')
    _check_device(uv)
    return _texture_func.apply(tex.contiguous(), uv.contiguous())


# ---------------------------------------------------------------------------------------------------------
# antialias
# ---------------------------------------------------------------------------------------------------------

class TopologyHashWrapper:
    """Per-topology adjacency table (tri_opp [T,3]) — the role of upstream's TopologyHashWrapper."""

    def __init__(self, tri_opp):
        self.tri_opp = tri_opp


_topology_cache = {}


def antialias_construct_topology_hash(tri):
    _check_tensor('tri', tri, torch.int32, 2)
    _require(tri.shape[1] == 3 and tri.shape[0] > 0, 'tri must have shape [>0, 3]')
    _check_device(tri)
    tri = tri.contiguous()
    T = tri.shape[0]
    tri_opp = torch.empty((T, 3), dtype=torch.int32, device=tri.device)
    nbytes = _lib.load().fpc_topology_scratch_bytes(T)
    scratch = torch.empty(int(nbytes), dtype=torch.uint8, device=tri.device)
    with torch.cuda.device(tri.device):
        _lib.call('fpc_topology_build', _ptr(tri), T, int(0), _ptr(tri_opp), _ptr(scratch), scratch.numel(), _stream())
    return TopologyHashWrapper(tri_opp)


def _topology_for(tri):
    key = (tri.data_ptr(), tri._version, tuple(tri.shape), tri.device)
    hit = _topology_cache.get(key)
    if hit is None:
        if len(_topology_cache) > 16:
            _topology_cache.clear()
        hit = antialias_construct_topology_hash(tri)
        _topology_cache[key] = hit
    return hit


class _antialias_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color, rast, pos, tri, tri_opp, pos_gradient_boost):
        N, H, W, C = color.shape
        V = pos.shape[1]
        out = torch.empty_like(color)
        with torch.cuda.device(color.device):
            _lib.call('fpc_antialias_fwd', _ptr(color), _ptr(rast), _ptr(pos), _ptr(tri), _ptr(tri_opp), N, V,
                      tri.shape[0], H, W, C, _ptr(out), _stream())
        ctx.save_for_backward(color, rast, pos, tri, tri_opp)
        ctx.pos_gradient_boost = pos_gradient_boost
        return out

    @staticmethod
    def backward(ctx, dy):
        color, rast, pos, tri, tri_opp = ctx.saved_tensors
        N, H, W, C = color.shape
        V = pos.shape[1]
        g_color = torch.empty_like(color)
        g_pos = torch.empty_like(pos)
        dy = dy.contiguous()
        with torch.cuda.device(color.device):
            _lib.call('fpc_antialias_bwd', _ptr(color), _ptr(rast), _ptr(pos), _ptr(tri), _ptr(tri_opp), _ptr(dy), N, V,
                      tri.shape[0], H, W, C, _ptr(g_color), _ptr(g_pos), _stream())
        if ctx.pos_gradient_boost != 1.0:
            g_pos = g_pos * ctx.pos_gradient_boost
        return g_color, None, g_pos, None, None, None


def antialias(color, rast, pos, tri, topology_hash=None, pos_gradient_boost=1.0):
    """color [N,H,W,C], rast [N,H,W,4], pos [N,V,4], tri [T,3] -> [N,H,W,C]."""
    _check_tensor('color', color, torch.float32, 4)
    _check_tensor('rast', rast, torch.float32, 4)
    _check_tensor('pos', pos, torch.float32, 3)
    _check_tensor('tri', tri, torch.int32, 2)
    _require(min(color.shape) > 0, 'color must have shape [>0, >0, >0, >0]')
    _require(rast.shape[3] == 4 and tuple(rast.shape[:3]) == tuple(color.shape[:3]),
             'rast must have shape [N, H, W, 4] matching color')
    _require(pos.shape[2] == 4 and pos.shape[0] == color.shape[0], 'pos must have shape [N, >0, 4] (instanced mode)')
    _require(tri.shape[1] == 3 and tri.shape[0] > 0, 'tri must have shape [>0, 3]')
    _require(color.device == rast.device == pos.device == tri.device, 'all inputs must reside on the same device')
    _check_device(color)
    tri = tri.contiguous()
    if topology_hash is None:
        topology_hash = _topology_for(tri)
    _require(isinstance(topology_hash, TopologyHashWrapper), 'topology_hash must come from antialias_construct_topology_hash')
    return _antialias_func.apply(color.contiguous(), rast.contiguous(), pos.contiguous(), tri, topology_hash.tri_opp,
                                 float(pos_gradient_boost))


# upstream module-level helpers kept for API compatibility
_log_level = 1


def get_log_level():
    return _log_level


def set_log_level(level):
    global _log_level
    _log_level = int(level)
