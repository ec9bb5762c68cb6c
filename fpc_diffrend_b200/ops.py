"""nvdiffrast-compatible front end of the B200 kernels — the drop-in for `import nvdiffrast.torch as dr`.

The reference calls exactly these:
    dr.RasterizeGLContext(device='cuda')                                  fit.py:484
    dr.rasterize(glctx, pos_clip, pos_idx, resolution=(H, W))             fit.py:151
    dr.interpolate(attr[None], rast_out, idx)                             fit.py:157
    dr.texture(tex[None], texc, filter_mode='linear')                     fit.py:158
    dr.antialias(colour, rast_out, pos_clip, pos_idx)                     fit.py:160
and, in the enable_mip branch of its render() (fit.py:153-155; off in the shipped configuration):
    dr.interpolate(uv[None], rast_out, uv_idx, rast_db=rast_out_db, diff_attrs='all')
    dr.texture(tex[None], texc, texd, filter_mode='linear-mipmap-linear', max_mip_level=k)
Signatures, defaults, tensor layouts (rast = (u, v, z/w, tri_id+1), row 0 = bottom) and the error
behaviour (RuntimeError naming the offending argument) follow upstream nvdiffrast v0.3.x.  Every op is a
torch.autograd.Function whose forward/backward call the C-ABI of include/fpc_b200.h on the current CUDA
stream.  There is no CPU or PyTorch fallback: tensors must live on a CUDA device of compute capability 10.x.
"""
import ctypes
import weakref

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
    def forward(ctx, glctx, pos, tri, resolution, want_db, grad_db):
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
        ctx.grad_db = bool(grad_db and rast_db is not None)
        if rast_db is None:
            rast_db = torch.empty((N, H, W, 0), dtype=torch.float32, device=pos.device)
        if not ctx.grad_db:
            ctx.mark_non_differentiable(rast_db)
        return rast, rast_db

    @staticmethod
    def backward(ctx, dy, ddb):
        pos, tri, rast = ctx.saved_tensors
        N, V, _ = pos.shape
        _, H, W, _ = rast.shape
        g_pos = torch.empty_like(pos)
        dy = dy.contiguous() if dy is not None else torch.zeros_like(rast)
        with torch.cuda.device(pos.device):
            if ctx.grad_db and ddb is not None:
                # upstream's rasterize_grad_db: the gradient also flows through the barycentric pixel differentials
                _lib.call('fpc_rasterize_bwd_db', _ptr(pos), _ptr(tri), _ptr(rast), _ptr(dy), _ptr(ddb.contiguous()), N, V,
                          tri.shape[0], H, W, _ptr(g_pos), _stream())
            else:
                _lib.call('fpc_rasterize_bwd', _ptr(pos), _ptr(tri), _ptr(rast), _ptr(dy), N, V, tri.shape[0], H, W,
                          _ptr(g_pos), _stream())
        return None, g_pos, None, None, None, None


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
    # grad_db: whether gradients flow back through rast_db (mip-mapped texturing path, fit.py:153-155)
    return _rasterize_func.apply(glctx, pos.contiguous(), tri.contiguous(), (int(resolution[0]), int(resolution[1])),
                                 glctx.output_db, bool(grad_db))


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


class _interpolate_da_func(torch.autograd.Function):
    """interpolate with attribute pixel differentials (upstream's interpolate_fwd_da / interpolate_grad_da)."""

    @staticmethod
    def forward(ctx, attr, rast, tri, rast_db, diff_list):
        Na, Vt, A = attr.shape
        N, H, W, _ = rast.shape
        K = A if diff_list is None else len(diff_list)
        sel = None if diff_list is None else (ctypes.c_int32 * K)(*diff_list)
        out = torch.empty((N, H, W, A), dtype=torch.float32, device=rast.device)
        out_da = torch.empty((N, H, W, 2 * K), dtype=torch.float32, device=rast.device)
        with torch.cuda.device(rast.device):
            _lib.call('fpc_interpolate_da_fwd', _ptr(attr), Na, Vt, A, _ptr(rast), _ptr(rast_db), _ptr(tri), sel, K, N, tri.shape[0],
                      H, W, _ptr(out), _ptr(out_da), _stream())
        ctx.save_for_backward(attr, rast, tri, rast_db)
        ctx.sel, ctx.K = sel, K
        return out, out_da

    @staticmethod
    def backward(ctx, dy, dda):
        attr, rast, tri, rast_db = ctx.saved_tensors
        Na, Vt, A = attr.shape
        N, H, W, _ = rast.shape
        g_attr, g_rast, g_db = torch.empty_like(attr), torch.empty_like(rast), torch.empty_like(rast_db)
        dy = dy.contiguous() if dy is not None else None
        dda = dda.contiguous() if (dda is not None and ctx.K > 0) else None
        if dy is None and dda is None:
            return torch.zeros_like(attr), torch.zeros_like(rast), None, torch.zeros_like(rast_db), None
        with torch.cuda.device(rast.device):
            _lib.call('fpc_interpolate_da_bwd', _ptr(attr), Na, Vt, A, _ptr(rast), _ptr(rast_db), _ptr(tri), ctx.sel, ctx.K, _ptr(dy),
                      _ptr(dda), N, tri.shape[0], H, W, _ptr(g_attr), _ptr(g_rast), _ptr(g_db), _stream())
        return g_attr, g_rast, None, g_db, None


def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
    """attr [1|N,V,A], rast [N,H,W,4], tri [T,3] -> (out [N,H,W,A], out_da [N,H,W,2k]); k = 0 unless rast_db and diff_attrs
    ('all' or a list of attribute indices) are given (fit.py:154)."""
    _check_tensor('attr', attr, torch.float32, 3)
    _check_tensor('rast', rast, torch.float32, 4)
    _check_tensor('tri', tri, torch.int32, 2)
    _require(rast.shape[3] == 4 and min(rast.shape) > 0, 'rast must have shape [>0, >0, >0, 4]')
    _require(tri.shape[1] == 3 and tri.shape[0] > 0, 'tri must have shape [>0, 3]')
    _require(attr.shape[0] in (1, rast.shape[0]) and attr.shape[1] > 0 and attr.shape[2] > 0,
             'attr must have shape [1 or minibatch, >0, >0]')
    _require(attr.device == rast.device == tri.device, 'attr, rast and tri must reside on the same device')
    _check_device(rast)
    diff_list = None
    want_da = diff_attrs is not None and not (isinstance(diff_attrs, (list, tuple)) and len(diff_attrs) == 0)
    if want_da:
        _require(rast_db is not None, 'interpolate: diff_attrs needs rast_db')
        _check_tensor('rast_db', rast_db, torch.float32, 4)
        _require(tuple(rast_db.shape) == tuple(rast.shape), 'rast_db must have the shape of rast [N, H, W, 4]')
        _require(rast_db.device == rast.device, 'rast_db must reside on the same device as rast')
        if diff_attrs != 'all':
            _require(isinstance(diff_attrs, (list, tuple)), "diff_attrs must be 'all' or a list of attribute indices")
            diff_list = [int(i) for i in diff_attrs]
            _require(all(0 <= i < attr.shape[2] for i in diff_list), 'diff_attrs indices out of range')
        _require((attr.shape[2] if diff_list is None else len(diff_list)) <= 32, 'interpolate: at most 32 differentiated attributes')
        return _interpolate_da_func.apply(attr.contiguous(), rast.contiguous(), tri.contiguous(), rast_db.contiguous(), diff_list)
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


class TextureMipWrapper:
    """Pre-built mip stack (levels 1..L back to back) — the role of upstream's TextureMipWrapper."""

    def __init__(self, mip, levels, tex_shape):
        self.mip, self.levels, self.tex_shape = mip, levels, tuple(tex_shape)


def _mip_levels(tex, max_mip_level):
    L = _lib.load().fpc_texture_mip_levels(int(tex.shape[1]), int(tex.shape[2]), -1 if max_mip_level is None else int(max_mip_level))
    if max_mip_level is None:
        # upstream builds the chain down to 1x1 and refuses extents that stop being even on the way
        _require((tex.shape[1] >> L) == 1 or (tex.shape[2] >> L) == 1,
                 'texture: extents %dx%d do not halve evenly down to 1; pass max_mip_level' % (tex.shape[2], tex.shape[1]))
    else:
        _require(L == int(max_mip_level), 'texture: max_mip_level=%d needs extents divisible by %d (got %dx%d)'
                 % (max_mip_level, 1 << int(max_mip_level), tex.shape[2], tex.shape[1]))
    return L


def _build_mip(tex, L):
    Nt, Ht, Wt, C = tex.shape
    mip = torch.empty(int(_lib.load().fpc_texture_mip_floats(Nt, Ht, Wt, C, L)), dtype=torch.float32, device=tex.device)
    with torch.cuda.device(tex.device):
        _lib.call('fpc_texture_mip_build', _ptr(tex), Nt, Ht, Wt, C, L, _ptr(mip), _stream())
    return mip


def texture_construct_mip(tex, max_mip_level=None, cube_mode=False):
    """Build the mip stack of tex [1|N,Ht,Wt,C] once, for reuse across texture() calls with a constant texture."""
    _require(not cube_mode, 'texture_construct_mip: cube maps are not supported')
    _check_tensor('tex', tex, torch.float32, 4)
    _check_device(tex)
    tex = tex.contiguous()
    L = _mip_levels(tex, max_mip_level)
    return TextureMipWrapper(_build_mip(tex.detach(), L), L, tex.shape)


class _texture_mip_func(torch.autograd.Function):
    """upstream's texture_fwd_mip / texture_grad_linear_mipmap_{nearest,linear}"""

    @staticmethod
    def forward(ctx, tex, uv, uv_da, bias, mip, L, nearest, mip_const):
        Nt, Ht, Wt, C = tex.shape
        N, H, W, _ = uv.shape
        if mip is None:
            mip = _build_mip(tex, L)
        out = torch.empty((N, H, W, C), dtype=torch.float32, device=uv.device)
        with torch.cuda.device(uv.device):
            _lib.call('fpc_texture_mip_fwd', _ptr(tex), _ptr(mip), Nt, Ht, Wt, C, L, _ptr(uv), _ptr(uv_da), _ptr(bias), nearest, N, H, W,
                      _ptr(out), _stream())
        ctx.save_for_backward(tex, uv, uv_da, bias, mip)
        ctx.L, ctx.nearest, ctx.mip_const = L, nearest, mip_const
        return out

    @staticmethod
    def backward(ctx, dy):
        tex, uv, uv_da, bias, mip = ctx.saved_tensors
        Nt, Ht, Wt, C = tex.shape
        N, H, W, _ = uv.shape
        g_tex = torch.empty_like(tex) if ctx.needs_input_grad[0] else None
        g_mip = torch.empty_like(mip) if g_tex is not None else None
        g_uv = torch.empty_like(uv)
        g_da = torch.empty_like(uv_da) if uv_da is not None else None
        g_bias = torch.empty_like(bias) if bias is not None else None
        with torch.cuda.device(uv.device):
            _lib.call('fpc_texture_mip_bwd', _ptr(tex), _ptr(mip), Nt, Ht, Wt, C, ctx.L, _ptr(uv), _ptr(uv_da), _ptr(bias), ctx.nearest,
                      _ptr(dy.contiguous()), N, H, W, _ptr(g_tex), _ptr(g_mip), 1 if ctx.mip_const else 0, _ptr(g_uv), _ptr(g_da),
                      _ptr(g_bias), _stream())
        return g_tex, g_uv, g_da, g_bias, None, None, None, None


def texture(tex, uv, uv_da=None, mip_level_bias=None, mip=None, filter_mode='auto', boundary_mode='wrap',
            max_mip_level=None):
    """tex [1|N,Ht,Wt,C], uv [N,H,W,2] -> [N,H,W,C].  filter_mode 'linear', 'linear-mipmap-nearest' or 'linear-mipmap-linear'
    ('auto' = the latter when uv_da / mip_level_bias is given, else 'linear'); boundary_mode 'wrap' (fit.py:155,158)."""
    if filter_mode == 'auto':
        filter_mode = 'linear-mipmap-linear' if (uv_da is not None or mip_level_bias is not None) else 'linear'
    _require(filter_mode in ('linear', 'linear-mipmap-nearest', 'linear-mipmap-linear'),
             "texture: filter_mode must be 'linear', 'linear-mipmap-nearest' or 'linear-mipmap-linear' (got %r)" % (filter_mode,))
    _require(boundary_mode == 'wrap', "texture: only boundary_mode='wrap' is supported (got %r)" % (boundary_mode,))
    _check_tensor('tex', tex, torch.float32, 4)
    _check_tensor('uv', uv, torch.float32, 4)
    _require(uv.shape[3] == 2 and min(uv.shape) > 0, 'uv must have shape [>0, >0, >0, 2]')
    _require(min(tex.shape) > 0 and tex.shape[0] in (1, uv.shape[0]), 'tex must have shape [1 or minibatch, >0, >0, >0]')
    _require(tex.device == uv.device, 'tex and uv must reside on the same device')
    _check_device(uv)
    if filter_mode == 'linear':
        # upstream ignores the mip inputs in the non-mip modes
        return _texture_func.apply(tex.contiguous(), uv.contiguous())
    _require(uv_da is not None or mip_level_bias is not None, 'texture: mip filter modes need uv_da or mip_level_bias')
    if uv_da is not None:
        _check_tensor('uv_da', uv_da, torch.float32, 4)
        _require(tuple(uv_da.shape) == tuple(uv.shape[:3]) + (4,), 'uv_da must have shape [N, H, W, 4]')
        _require(uv_da.device == uv.device, 'uv_da must reside on the same device as uv')
        uv_da = uv_da.contiguous()
    if mip_level_bias is not None:
        _check_tensor('mip_level_bias', mip_level_bias, torch.float32, 3)
        _require(tuple(mip_level_bias.shape) == tuple(uv.shape[:3]), 'mip_level_bias must have shape [N, H, W]')
        _require(mip_level_bias.device == uv.device, 'mip_level_bias must reside on the same device as uv')
        mip_level_bias = mip_level_bias.contiguous()
    tex = tex.contiguous()
    if mip is not None:
        _require(isinstance(mip, TextureMipWrapper) and mip.tex_shape == tuple(tex.shape), 'mip must come from texture_construct_mip(tex)')
        L, mip_buf = mip.levels, mip.mip
        if max_mip_level is not None:
            L = min(L, int(max_mip_level))
    else:
        L, mip_buf = _mip_levels(tex, max_mip_level), None
    return _texture_mip_func.apply(tex, uv.contiguous(), uv_da, mip_level_bias, mip_buf, L, 1 if filter_mode == 'linear-mipmap-nearest' else 0,
                                   mip is not None)


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
    """Adjacency table of `tri`, cached per tensor OBJECT (upstream rebuilds its edge hash on every call unless topology_hash is
    passed; the reference never passes it, fit.py:160, but hands over the same pos_idx tensor every iteration).  The entry holds a
    weak reference to the tensor it was built from: a different tensor that merely reuses a freed tensor's address, or an in-place
    modification (version counter), rebuilds."""
    key = tri.data_ptr()
    hit = _topology_cache.get(key)
    if hit is not None:
        ref, version, wrapper = hit
        if ref() is tri and version == tri._version:
            return wrapper
    if len(_topology_cache) > 16:
        _topology_cache.clear()
    wrapper = antialias_construct_topology_hash(tri)
    _topology_cache[key] = (weakref.ref(tri), tri._version, wrapper)
    return wrapper


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


class DepthPeeler:
    """Upstream's depth-peeling context manager.  The reference never uses it; constructing one fails loudly
    instead of rendering something else."""

    def __init__(self, glctx, pos, tri, resolution, ranges=None, grad_db=True):
        raise RuntimeError('DepthPeeler is not supported by fpc_diffrend_b200.ops (the fit path renders the first surface only)')
