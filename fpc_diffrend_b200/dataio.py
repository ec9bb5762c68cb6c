"""On-disk formats either side of the fit hot path, host-side numpy / PIL only.

Inputs, as the reference reads them:
  base mesh            Wavefront OBJ, `v` / `vt` / `f v/vt v/vt v/vt`, 1-based, triangles only      (data.py:7-39)
  blendshape directory one OBJ per shape, only `v ` lines are read; shape order = os.listdir order      (fit.py:200-216)
  reference frames     <imdir>/<cam>/<cam>_<frame:0{digits}d>.tif, 8-bit grey, clipped to [0,140] and flipped
                       vertically before use                                                            (fit.py:29-43,514-533)
  calibration          calibration.json keyed by the part of the camera directory name after '_'        (fit.py:514-521)
Outputs, as the reference writes them:
  <out>/result/<i>.obj (v lines, vt lines, the verbatim lines of result/faces.txt if present), texture.png (flipped, x255),
  pose.json {'rotation', 'translation'} (sorted keys, indent 4), <out>/config.txt ("key: 'value'" lines)  (fit.py:235-286,655-657)

The reference decodes ONE tif from disk inside every iteration (fit.py:529-533); here a take is decoded once into a
uint8 array [F, C, H, W, 1] that stays resident on the device (or is streamed through FitSession.fit_stream).
"""
import codecs
import json
import os

import numpy as np


class MeshData:
    """OBJ reader with the reference's conventions (data.py:7-39): `vertices` [3V] f32 (x,y,z,x,...), `uv` [Vt,2] f32,
    `faces` [T,3] i32 and `fuv` [T,3] i32 (0-based)."""

    def __init__(self, obj):
        vertices, uv, faces, fuv = [], [], [], []
        with open(obj, 'r') as f:
            for line in f:
                if line.startswith('v '):
                    vertices.extend(float(x) for x in line.split()[1:4])
                elif line.startswith('vt '):
                    uv.append([float(x) for x in line.split()[1:3]])
                elif line.startswith('f '):
                    idxs = [tok.split('/') for tok in line.split()[1:]]
                    if len(idxs) != 3:
                        raise ValueError('%s: only triangles are supported (data.py:29), got a face with %d corners' % (obj, len(idxs)))
                    faces.append([int(x[0]) - 1 for x in idxs])
                    fuv.append([int(x[1]) - 1 if len(x) > 1 and x[1] else int(x[0]) - 1 for x in idxs])
        self.vertices = np.asarray(vertices, dtype=np.float32)
        self.uv = np.asarray(uv, dtype=np.float32).reshape(-1, 2)
        self.faces = np.asarray(faces, dtype=np.int32).reshape(-1, 3)
        self.fuv = np.asarray(fuv, dtype=np.int32).reshape(-1, 3)


def read_obj_vertices(path):
    """Only the `v ` lines of an OBJ -> [3V] f32 (the fast path the reference uses for blendshapes, fit.py:208-214)."""
    out = []
    with open(path, 'r') as f:
        for line in f:
            if line.startswith('v '):
                out.extend(float(x) for x in line.split()[1:4])
    return np.asarray(out, dtype=np.float32)


def load_blendshape_dir(path, v_base, order='listdir'):
    """D [3V, B] f32 = (blendshape vertices - base) transposed, and the shape names (fit.py:200-219).
    order='listdir' reproduces the reference (os.listdir order, i.e. file-system dependent); 'sorted' is reproducible."""
    names = os.listdir(path)
    if order == 'sorted':
        names = sorted(names)
    elif order != 'listdir':
        raise ValueError("order must be 'listdir' or 'sorted'")
    names = [n for n in names if n.lower().endswith('.obj')]
    v_base = np.asarray(v_base, dtype=np.float32)
    D = np.empty((len(names), v_base.shape[0]), dtype=np.float32)
    for i, n in enumerate(names):
        v = read_obj_vertices(os.path.join(path, n))
        if v.shape != v_base.shape:
            raise ValueError('%s has %d coordinates, the base mesh %d' % (n, v.shape[0], v_base.shape[0]))
        D[i] = v - v_base
    return np.ascontiguousarray(D.T), names


def list_cameras(imdir):
    """Camera directories of a take, in os.listdir order like fit.py:415."""
    return [d for d in os.listdir(imdir) if os.path.isdir(os.path.join(imdir, d))]


def assert_num_frames(cams, imdir):
    """(n_frames, digits) of a take; every camera must hold the same number of frames (fit.py:29-43)."""
    n = [len(os.listdir(os.path.join(imdir, c))) for c in cams]
    if any(x != n[0] for x in n):
        raise AssertionError('All cameras do not have the same number of frames!')
    return (n[0], 2) if n[0] < 100 else (n[0], 3)


def calibration_for(calibs, cam):
    """The calibration entry of camera directory `cam` ('<prefix>_<key>...': key = second '_' field, fit.py:515)."""
    return calibs[cam.split('_')[1]]


def frame_path(imdir, cam, frame_idx, digits):
    return os.path.join(imdir, cam, '%s_%0*d.tif' % (cam, digits, frame_idx))


def read_frame(path):
    """One reference frame as the loop sees it (fit.py:529-533): decode, clip to [0,140], flip vertically so that row 0 is
    the bottom row (the rasterizer's convention) -> [H, W, 1] uint8."""
    from PIL import Image
    img = np.array(Image.open(path))
    if img.ndim == 3:
        img = img[..., 0]
    img = np.clip(img, 0, 140)
    return np.ascontiguousarray(np.flip(img, 0)).astype(np.uint8)[..., None]


def load_reference_frames(imdir, cams, frames, digits=None, out=None):
    """[F, C, H, W, 1] uint8 for the given frame indices and camera directories (decoded once per take)."""
    if digits is None:
        digits = assert_num_frames(cams, imdir)[1]
    first = read_frame(frame_path(imdir, cams[0], frames[0], digits))
    H, W, _ = first.shape
    if out is None:
        out = np.empty((len(frames), len(cams), H, W, 1), dtype=np.uint8)
    for fi, f in enumerate(frames):
        for ci, c in enumerate(cams):
            img = first if (fi == 0 and ci == 0) else read_frame(frame_path(imdir, c, f, digits))
            if img.shape != (H, W, 1):
                raise ValueError('frame %d of %s is %s, expected %s' % (f, c, img.shape[:2], (H, W)))
            out[fi, ci] = img
    return out


def write_frame(path, img_bottom_up):
    """Inverse of read_frame for synthetic takes: [H, W(,1)] grey levels with row 0 = bottom -> 8-bit tif on disk."""
    from PIL import Image
    a = np.asarray(img_bottom_up)
    if a.ndim == 3:
        a = a[..., 0]
    Image.fromarray(np.ascontiguousarray(np.flip(np.clip(np.rint(a), 0, 255).astype(np.uint8), 0))).save(path, format='TIFF')


def save_results(meshes, uv, texture, translation, rotation, out_dir, faces_lines=None):
    """The reference's `save` (fit.py:235-286): result/<i>.obj per frame, texture.png, pose.json."""
    directory = os.path.join(out_dir, 'result')
    os.makedirs(directory, exist_ok=True)
    if faces_lines is None:
        try:
            with open(os.path.join(directory, 'faces.txt')) as f:
                faces_lines = f.readlines()
        except OSError:
            faces_lines = []
    meshes = np.asarray(meshes, dtype=np.float32)
    for i, mesh in enumerate(meshes):
        with open(os.path.join(directory, '%d.obj' % i), 'w') as f:
            for x, y, z in mesh.reshape(-1, 3):
                f.write('v %r %r %r\n' % (float(x), float(y), float(z)))
            for u in np.asarray(uv):
                f.write('vt %r %r\n' % (float(u[0]), float(u[1])))
            f.writelines(faces_lines)
    if texture is not None:
        from PIL import Image
        t = np.asarray(texture, dtype=np.float32)
        t8 = (np.flip(t, 0) * 255).astype(np.uint8)
        Image.fromarray(t8[..., 0] if t8.ndim == 3 and t8.shape[2] == 1 else t8).save(os.path.join(directory, 'texture.png'), format='PNG')
    pose = {'translation': np.asarray(translation, dtype=np.float32).tolist(), 'rotation': np.asarray(rotation, dtype=np.float32).tolist()}
    with codecs.open(os.path.join(directory, 'pose.json'), 'w', encoding='utf-8') as f:
        json.dump(pose, f, separators=(',', ':'), sort_keys=True, indent=4)
    return directory


def faces_lines_for(pos_idx, uv_idx):
    """`f v/vt v/vt v/vt` lines (what the reference expects in result/faces.txt)."""
    return ['f %d/%d %d/%d %d/%d\n' % (a[0] + 1, b[0] + 1, a[1] + 1, b[1] + 1, a[2] + 1, b[2] + 1) for a, b in zip(pos_idx, uv_idx)]


def write_config(out_dir, args):
    """config.txt with one "key: 'value'" line per setting (fit.py:655-657)."""
    with open(os.path.join(out_dir, 'config.txt'), 'w') as f:
        for k, v in args.items():
            f.write("%s: '%s'\n" % (k, v))
