"""Forward renderers on the drop-in ops: the reference's `render()` (fit.py:134-162, including its
enable_mip branch :153-155) and the result-sequence renderer of render_multicam.py:112-167 / render_result.py (all cameras of
every fitted frame, saved pose re-applied), without the viewer / mp4 plumbing.

Everything here calls fpc_diffrend_b200.ops (the nvdiffrast-compatible front end of the CUDA kernels); there is no CPU path.
"""
import codecs
import json
import os

import numpy as np
import torch

from . import camera as cam
from . import dataio
from . import ops as dr

BG = 45.0 / 255.0   # fit.py:161


def transform_clip(mtx, pos):
    """camera.py:11-23: [V,3] object-space positions -> [1,V,4] clip space, pos_clip = [pos 1] mtx^T."""
    posw = torch.cat([pos, torch.ones((pos.shape[0], 1), dtype=pos.dtype, device=pos.device)], dim=1)
    return torch.matmul(posw, mtx.t())[None, ...].contiguous()


def render(glctx, mtx, pos, pos_idx, uv, uv_idx, tex, resolution, enable_mip=False, max_mip_level=None):
    """The reference's render() (fit.py:134-162) on the drop-in: rasterize -> interpolate -> texture (bilinear, or trilinear
    mip-mapped when enable_mip) -> antialias -> background 45/255.  Returns [H, W, C] (row 0 = bottom).  Differentiable."""
    pos_clip = transform_clip(mtx, pos)
    rast_out, rast_out_db = dr.rasterize(glctx, pos_clip, pos_idx, resolution=(resolution[0], resolution[1]))
    if enable_mip:
        texc, texd = dr.interpolate(uv[None, ...], rast_out, uv_idx, rast_db=rast_out_db, diff_attrs='all')
        colour = dr.texture(tex[None, ...], texc, texd, filter_mode='linear-mipmap-linear', max_mip_level=max_mip_level)
    else:
        texc, _ = dr.interpolate(uv[None, ...], rast_out, uv_idx)
        colour = dr.texture(tex[None, ...], texc, filter_mode='linear')
    colour = dr.antialias(colour, rast_out, pos_clip, pos_idx)
    colour = torch.where(rast_out[..., 3:] > 0, colour, torch.tensor(BG, device=colour.device))
    return colour[0]


def camera_mvp(calib, t_frame=None, q_frame=None, y_offset=0.0, device='cuda'):
    """MVP of one calibration entry as render_multicam.py:131-145 builds it: P (R_frame|t_frame) MV T(0, y_offset, 0).
    The fit loop uses y_offset = 170 (fit.py:545); the result renderers 0 (the saved vertices already sit in place)."""
    P = cam.intrinsic_to_projection(np.asarray(calib['intrinsic'], dtype=np.float32)).astype(np.float64)
    MV = cam.extrinsic_to_modelview(np.asarray(calib['rotation'], dtype=np.float32), np.asarray(calib['translation'], dtype=np.float32))
    m = MV.astype(np.float64) @ cam.translate(0.0, y_offset, 0.0).astype(np.float64)
    if t_frame is not None:
        m = cam.rigid(t_frame, cam.unitquat_to_rotmat(np.asarray(q_frame, dtype=np.float64))) @ m
    return torch.tensor((P @ m).astype(np.float32), device=device)


def render_result(result_dir, calibpath, cam_names, resolution, frames=None, reproduce_pose=True, texpath=None, enable_mip=False,
                  max_mip_level=None, out_dir=None, y_offset=0.0):
    """Render a saved fit (the `result/` directory written by dataio.save_results / the reference's save(), fit.py:235-286)
    from every camera in cam_names: for each frame <i>.obj, the texture and — when reproduce_pose — the per-frame head pose
    of pose.json.  Returns uint8 images [F, C, H, W, Ch] in display orientation (row 0 = top); with out_dir also writes
    frame<i>_<cam>.png.  Role of render_multicam.py:112-167 and render_result.py."""
    from PIL import Image
    dev = torch.device('cuda', torch.cuda.current_device())
    calibs = cam.load_calibration(calibpath)
    objs = sorted((f for f in os.listdir(result_dir) if f.endswith('.obj') and f[:-4].isdigit()), key=lambda f: int(f[:-4]))
    if frames is not None:
        objs = ['%d.obj' % i for i in frames]
    if not objs:
        raise ValueError('no <i>.obj files in %s' % result_dir)
    mesh = dataio.MeshData(os.path.join(result_dir, objs[0]))
    pos_idx = torch.tensor(mesh.faces, dtype=torch.int32, device=dev)
    uv = torch.tensor(mesh.uv, dtype=torch.float32, device=dev)
    uv_idx = torch.tensor(mesh.fuv, dtype=torch.int32, device=dev)
    tex = np.array(Image.open(texpath or os.path.join(result_dir, 'texture.png'))).astype(np.float32) / 255.0
    if tex.ndim == 2:
        tex = tex[..., None]
    tex = torch.tensor(np.ascontiguousarray(np.flip(tex, 0)), dtype=torch.float32, device=dev)
    pose = None
    if reproduce_pose:
        with codecs.open(os.path.join(result_dir, 'pose.json'), 'r', encoding='utf-8') as f:
            pose = json.load(f)
    glctx = dr.RasterizeGLContext(device=dev)
    out = []
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
    for obj in objs:
        i = int(obj[:-4])
        verts = torch.tensor(dataio.read_obj_vertices(os.path.join(result_dir, obj)), dtype=torch.float32, device=dev).reshape(-1, 3)
        views = []
        for name in cam_names:
            calib = dataio.calibration_for(calibs, name) if name not in calibs else calibs[name]
            t, q = (pose['translation'][i], pose['rotation'][i]) if pose is not None else (None, None)
            mvp = camera_mvp(calib, t, q, y_offset=y_offset, device=dev)
            with torch.no_grad():
                img = render(glctx, mvp, verts, pos_idx, uv, uv_idx, tex, resolution, enable_mip, max_mip_level) * 255.0
            img8 = torch.flip(img, dims=[0]).round().clamp(0, 255).to(torch.uint8).cpu().numpy()
            views.append(img8)
            if out_dir is not None:
                Image.fromarray(img8[..., 0] if img8.shape[2] == 1 else img8).save(os.path.join(out_dir, 'frame%d_%s.png' % (i, name)))
        out.append(np.stack(views))
    return np.stack(out)
