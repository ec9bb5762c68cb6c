"""Fit a whole take from disk (role of the reference's fitTake / main.py, fit.py:327-657, under the north-star
parameterisation: per-frame activations and head pose, all cameras of a frame in every iteration).

    results = fit_take(basemeshpath, localblpath, imdir, calibpath, out_dir, iters_per_frame=200)

What differs from the reference's loop by design: the frames of a take are decoded once (dataio.load_reference_frames)
and stay on the device as uint8 instead of one PIL decode + H2D copy per iteration (fit.py:529-533); every iteration
renders ALL cameras of every frame of the batch instead of one random (camera, frame) pair (fit.py:525-526); frame
batches are sharded over the ranks of torch.distributed when it is initialised (no exchange step).
"""
from dataclasses import replace
from types import SimpleNamespace

import numpy as np
import torch

from . import camera as cam
from . import dataio, shard
from .fit import FitConfig, FitSession


def load_take_rig(basemeshpath, localblpath, calibpath, cams, texpath=None, texshape=(1024, 1024, 1), blend_order='listdir', seed=0):
    """Everything constant over a take, as a rig object FitSession understands (fit.py:418-439, 183-230, 514-521)."""
    mesh = dataio.MeshData(basemeshpath)
    D, names = dataio.load_blendshape_dir(localblpath, mesh.vertices, order=blend_order)
    calibs = cam.load_calibration(calibpath)
    P, A = cam.camera_constants([dataio.calibration_for(calibs, c) for c in cams])
    if texpath:
        from PIL import Image
        tex = np.array(Image.open(texpath)).astype(np.float32) / 255.0          # fit.py:434-436
        if tex.ndim == 2:
            tex = tex[..., None]
        tex = np.ascontiguousarray(np.flip(tex, 0))
    else:
        tex = np.random.default_rng(seed).uniform(0.0, 1.0, size=texshape).astype(np.float32)   # fit.py:438
    V = mesh.vertices.shape[0] // 3
    return SimpleNamespace(v_base=mesh.vertices, pos_idx=mesh.faces, uv=mesh.uv, uv_idx=mesh.fuv, D=D, tex=tex.astype(np.float32),
                           vcol=np.full((V, 3), 0.5, np.float32), P=P, A=A, shape_names=names, B=D.shape[1], V=V, T=mesh.faces.shape[0])


def fit_take(basemeshpath, localblpath, imdir, calibpath, out_dir=None, iters_per_frame=200, frame_batch=16, frames=None,
             cams=None, texpath=None, config=None, blend_order='listdir', use_graph=True, log=None, max_iter=None):
    """Fit every frame of a take; returns dict(vertices [F,3V], w [F,B], t [F,3], q [F,4], loss [F_batches]) on rank 0
    (None elsewhere) and, when out_dir is given, writes result/<i>.obj, pose.json, texture.png and config.txt there.

    Every batch of frames runs exactly `iters_per_frame` optimiser steps from fresh parameters.  The schedule the reference
    spans over its whole run — LambdaLR lr_ramp ** (i / max_iter), correctives unlocked after max_iter / 2 (fit.py:503-505,
    603-608) — spans one batch here: its length is `max_iter` (default: iters_per_frame; `config.max_iter` is not used, and
    the caller's config object is not modified)."""
    cams = list(cams) if cams is not None else dataio.list_cameras(imdir)
    n_frames, digits = dataio.assert_num_frames(cams, imdir)
    frames = list(range(n_frames)) if frames is None else list(frames)
    rig = load_take_rig(basemeshpath, localblpath, calibpath, cams, texpath=texpath, blend_order=blend_order)
    first = dataio.read_frame(dataio.frame_path(imdir, cams[0], frames[0], digits))
    H, W = first.shape[:2]
    cfg = replace(config or FitConfig(shading='texture', antialias=True), resolution=(H, W), ref_dtype='u8',
                  max_iter=int(max_iter or iters_per_frame), reorder_vertices=True)    # (results come back in the rig's order)
    f0, f1 = shard.frame_shard(len(frames))
    mine = frames[f0:f1]
    # parameters shared by all frames (texture, per-camera pose corrections, the learned basis of the free / combined modes)
    # live in ONE session: such takes are fitted as a single batch per rank (their gradients are all-reduced over ranks)
    shared = cfg.optimize_texture or cfg.optimize_cam_pose or cfg.mode != 'prior'
    if shared:
        frame_batch = max(frame_batch, len(mine))
        if cfg.mode != 'prior':
            cfg = replace(cfg, n_frames_total=len(frames))
    verts, ws, ts, qs, losses = [], [], [], [], []
    sessions = {}
    for a in range(0, len(mine), frame_batch):
        batch = mine[a:a + frame_batch]
        ref = dataio.load_reference_frames(imdir, cams, batch, digits)
        s = sessions.get(len(batch))
        if s is None:
            s = sessions[len(batch)] = FitSession(rig, len(batch), cfg, frame_ids=list(range(f0 + a, f0 + a + len(batch))))
        s.set_reference(torch.from_numpy(ref))
        if use_graph and s.graph is None:
            s.capture(keep_state=True)                # the eager warm-up iteration is not one of the iters_per_frame steps
        s.reset_state()                               # fresh parameters and optimiser state for this batch of frames
        for _ in range(iters_per_frame):
            s.replay() if use_graph else s.iteration()
        torch.cuda.synchronize()
        losses.append(float(s.loss))
        if log:
            log('frames %s: loss %.4f after %d iterations' % (batch, losses[-1], iters_per_frame))
        verts.append(s.result_vertices()); ws.append(s.w.clone()); ts.append(s.t.clone()); qs.append(s.q.clone())
    dev = torch.device('cuda', torch.cuda.current_device())
    cat = lambda xs, shape: torch.cat(xs) if xs else torch.empty(shape, device=dev)
    local = dict(vertices=cat(verts, (0, rig.V * 3)), w=cat(ws, (0, rig.B)), t=cat(ts, (0, 3)), q=cat(qs, (0, 4)))
    out = {k: shard.gather_frames(v, len(frames)) for k, v in local.items()}
    if out['vertices'] is None:
        return None
    out = {k: v.cpu().numpy() for k, v in out.items()}
    out['loss'] = losses
    out['frames'] = frames
    if cfg.optimize_texture and sessions:
        rig.tex = next(iter(sessions.values())).tex[0].cpu().numpy()          # tex_opt is what fit.py:655 saves
    if out_dir is not None:
        dataio.save_results(out['vertices'], rig.uv, rig.tex, out['t'], out['q'], out_dir,
                            faces_lines=dataio.faces_lines_for(rig.pos_idx, rig.uv_idx))
        dataio.write_config(out_dir, dict(basemeshpath=basemeshpath, localblpath=localblpath, imdir=imdir, calibpath=calibpath,
                                          iters_per_frame=iters_per_frame, frame_batch=frame_batch, cams=cams, resolution=(W, H),
                                          **{k: getattr(cfg, k) for k in ('shading', 'antialias', 'lr_base', 'lr_t', 'lr_q', 'lr_ramp', 'max_iter',
                                                                          'weight_laplacian', 'weight_meshedge', 'weight_normalconsistency')}))
    return out
