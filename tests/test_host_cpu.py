"""CPU tests: host-side mirror of the reference's camera code, rig generator, C-ABI exports."""
import ctypes
import json
import os

import numpy as np
import pytest

from fpc_diffrend_b200 import _lib, camera, rig as rigmod

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def kat():
    with open(os.path.join(HERE, 'golden', 'camera_kat.json')) as f:
        return json.load(f)


def test_camera_matches_reference_kat(kat):
    """camera.py mirror == the reference's own camera.py on the real calibration (bit-exact fp32)."""
    for name, c in kat['cameras'].items():
        P = camera.intrinsic_to_projection(np.asarray(c['intrinsic'], np.float32))
        MV = camera.extrinsic_to_modelview(np.asarray(c['rotation'], np.float32), np.asarray(c['translation'], np.float32))
        assert np.array_equal(P, np.asarray(c['P'], np.float32)), name
        assert np.array_equal(MV, np.asarray(c['MV'], np.float32)), name
        Pc, Ac = camera.camera_constants([c])
        assert np.array_equal(Ac[0], np.asarray(c['A'], np.float32)), name
        mvp = (Pc[0] @ Ac[0]).astype(np.float32)
        assert np.array_equal(mvp, np.asarray(c['MVP'], np.float32)), name


def test_survey_appendix_c3_values(kat):
    c = kat['cameras']['pod2primary']
    assert abs(c['P'][0][0] - 12.083149909973145) < 1e-6
    assert abs(c['P'][1][1] - 8.954282760620117) < 1e-6
    np.testing.assert_allclose(c['clip'][0], [-9.326126, -56.269230, 171.465439, 171.468292], rtol=1e-6)


def test_unitquat_to_rotmat_is_rotation():
    rng = np.random.default_rng(0)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    R = camera.unitquat_to_rotmat(q)
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-12)
    assert abs(np.linalg.det(R) - 1.0) < 1e-12
    # identity quaternion (0,0,0,1) -> identity (XYZW order, fit.py:446-447)
    np.testing.assert_allclose(camera.unitquat_to_rotmat([0, 0, 0, 1]), np.eye(3))
    # 90 degrees about z
    s = np.sqrt(0.5)
    np.testing.assert_allclose(camera.unitquat_to_rotmat([0, 0, s, s]) @ [1, 0, 0], [0, 1, 0], atol=1e-12)


@pytest.mark.parametrize('nv', [1000, 2500])
def test_rig_counts(nv):
    r = rigmod.make_rig(n_vertices=nv, n_shapes=4, n_cams=2, width=64, height=64, tex_size=16)
    assert r.V == nv and r.T == 2 * nv - 4                     # closed genus-0: T = 2V - 4
    assert r.uv.shape[0] > nv                                   # seam duplicates: Vt > V
    assert (r.uv_idx != r.pos_idx).any()
    assert r.uv.min() > 0 and r.uv.max() < 1
    assert r.D.shape == (3 * nv, 4) and r.D.dtype == np.float32
    assert r.pos_idx.min() == 0 and r.pos_idx.max() == nv - 1
    # every edge is shared by exactly two triangles (closed manifold)
    e = np.sort(np.concatenate([r.pos_idx[:, [0, 1]], r.pos_idx[:, [1, 2]], r.pos_idx[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    assert (cnt == 2).all()
    assert list(r.calib.keys()) == ['pod1primary', 'pod1secondary']
    assert set(r.calib['pod1primary'].keys()) == {'distortion', 'intrinsic', 'rotation', 'translation'}


def test_obj_roundtrip(tmp_path, tiny_rig):
    """write_obj follows the reference's OBJ conventions (data.py:17-39): parse it back the way MeshData does."""
    p = tmp_path / 'base.obj'
    rigmod.write_obj(str(p), tiny_rig.v_base, tiny_rig.uv, tiny_rig.pos_idx, tiny_rig.uv_idx)
    verts, uv, faces, fuv = [], [], [], []
    for line in open(p):
        if line.startswith('v '):
            verts.extend(float(x) for x in line.strip().split(' ')[1:])
        elif line.startswith('vt '):
            uv.append([float(x) for x in line.strip().split(' ')[1:]])
        elif line.startswith('f '):
            idx = [l.split('/') for l in line.strip().split(' ')[1:]]
            assert len(idx) == 3
            faces.append([int(x[0]) - 1 for x in idx])
            fuv.append([int(x[1]) - 1 for x in idx])
    assert np.array_equal(np.asarray(verts, np.float32), tiny_rig.v_base)
    assert np.array_equal(np.asarray(faces, np.int32), tiny_rig.pos_idx)
    assert np.array_equal(np.asarray(fuv, np.int32), tiny_rig.uv_idx)
    assert np.array_equal(np.asarray(uv, np.float32), tiny_rig.uv)


# ---- C-ABI ------------------------------------------------------------------------------------------------

def test_library_exports_every_header_symbol():
    lib = _lib.load()
    names = _lib.header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), 'libfpc_b200.so does not export %s declared in include/fpc_b200.h' % n
    assert set(names) == set(_lib.SIGNATURES.keys())
    assert lib.fpc_abi_version() == 4


def test_argument_errors_do_not_throw_and_set_message():
    """Error convention: int status + fpc_last_error(), validated before any CUDA work."""
    lib = _lib.load()
    st = lib.fpc_rasterize_fwd(None, None, 1, 1, 1, 8, 8, None, None, None, 0, None)
    assert st == 1
    assert b'rasterize_fwd' in lib.fpc_last_error()
    st = lib.fpc_interpolate_fwd(ctypes.c_void_p(8), 3, 4, 2, ctypes.c_void_p(8), ctypes.c_void_p(8), 2, 1, 4, 4, ctypes.c_void_p(8), None)
    assert st == 1 and b'attr batch must be 1 or N' in lib.fpc_last_error()
    st = lib.fpc_texture_linear_fwd(ctypes.c_void_p(8), 1, 0, 4, 1, ctypes.c_void_p(8), 1, 4, 4, ctypes.c_void_p(8), None)
    assert st == 1
    st = lib.fpc_quat_renorm(ctypes.c_void_p(8), 4, 7, None)
    assert st == 1 and b'mode' in lib.fpc_last_error()
    with pytest.raises(RuntimeError, match='fpc_b200'):
        _lib.call('fpc_blend_fwd', None, None, None, 1, 1, 1, None, None)


def test_scratch_size_queries():
    lib = _lib.load()
    small = lib.fpc_rasterize_scratch_bytes(1, 2000, 128, 128)
    big = lib.fpc_rasterize_scratch_bytes(9, 39996, 1024, 1024)
    assert 0 < small < big < 64 * 2 ** 20
    assert lib.fpc_blend_bwd_scratch_bytes(60000, 200, 1) > 0
    assert lib.fpc_topology_scratch_bytes(39996) >= 6 * 39996 * 16


def test_ops_refuse_cpu_tensors():
    import torch
    import fpc_diffrend_b200.ops as dr
    ctx = dr.RasterizeGLContext(device='cuda')
    with pytest.raises(RuntimeError, match='CUDA'):
        dr.rasterize(ctx, torch.zeros(1, 3, 4), torch.zeros(1, 3, dtype=torch.int32), resolution=(8, 8))
    with pytest.raises(RuntimeError, match='CUDA'):
        dr.interpolate(torch.zeros(1, 3, 2), torch.zeros(1, 4, 4, 4), torch.zeros(1, 3, dtype=torch.int32))
    with pytest.raises(RuntimeError, match='range mode'):
        dr.rasterize(ctx, torch.zeros(3, 4), torch.zeros(1, 3, dtype=torch.int32), resolution=(8, 8), ranges=torch.zeros(1, 2))
    with pytest.raises(RuntimeError, match='linear'):
        dr.texture(torch.zeros(1, 4, 4, 1), torch.zeros(1, 4, 4, 2), filter_mode='nearest')


# ---- on-disk formats ------------------------------------------------------------------

def _write_take(tmp_path, rig, frames_u8, cams):
    """A synthetic take laid out like the reference's data: base.obj, blendshapes/*.obj, calibration.json, <cam>/<cam>_NN.tif."""
    import json
    from fpc_diffrend_b200 import dataio
    rigmod.write_obj(str(tmp_path / 'base.obj'), rig.v_base, rig.uv, rig.pos_idx, rig.uv_idx)
    bdir = tmp_path / 'blendshapes'
    bdir.mkdir()
    for b in range(rig.B):
        rigmod.write_obj(str(bdir / ('shape_%03d.obj' % b)), rig.v_base + rig.D[:, b], rig.uv, rig.pos_idx, rig.uv_idx)
    with open(tmp_path / 'calibration.json', 'w') as f:
        json.dump(rig.calib, f)
    imdir = tmp_path / 'frames'
    imdir.mkdir()
    F = frames_u8.shape[0]
    digits = 2 if F < 100 else 3
    for ci, c in enumerate(cams):
        (imdir / c).mkdir()
        for fi in range(F):
            dataio.write_frame(str(imdir / c / ('%s_%0*d.tif' % (c, digits, fi))), frames_u8[fi, ci])
    return str(tmp_path / 'base.obj'), str(bdir), str(imdir), str(tmp_path / 'calibration.json')


def test_take_formats_roundtrip(tmp_path):
    import json
    from fpc_diffrend_b200 import dataio
    rig = rigmod.make_rig(n_vertices=200, n_shapes=5, n_cams=2, width=40, height=24, tex_size=8, seed=1)
    cams = ['take_%s' % k for k in rig.calib.keys()]
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, size=(3, 2, 24, 40, 1), dtype=np.uint8)
    base, bdir, imdir, calib = _write_take(tmp_path, rig, frames, cams)
    m = dataio.MeshData(base)
    assert np.array_equal(m.vertices, rig.v_base) and np.array_equal(m.faces, rig.pos_idx)
    assert np.array_equal(m.fuv, rig.uv_idx) and np.array_equal(m.uv, rig.uv)
    D, names = dataio.load_blendshape_dir(bdir, m.vertices, order='sorted')
    assert names == ['shape_%03d.obj' % b for b in range(rig.B)]
    assert D.shape == rig.D.shape and np.abs(D - rig.D).max() <= 2e-6          # (base + delta) - base in float32
    got_cams = dataio.list_cameras(imdir)
    assert sorted(got_cams) == sorted(cams)
    assert dataio.assert_num_frames(cams, imdir) == (3, 2)
    ref = dataio.load_reference_frames(imdir, cams, [0, 1, 2])
    assert ref.dtype == np.uint8 and ref.shape == (3, 2, 24, 40, 1)
    assert np.array_equal(ref, np.clip(frames, 0, 140))                         # clip [0,140]; flip undone by write_frame
    calibs = json.load(open(calib))
    assert dataio.calibration_for(calibs, cams[0]) == rig.calib[cams[0].split('_')[1]]
    # writers
    out = tmp_path / 'out'
    out.mkdir()
    meshes = np.stack([rig.v_base, rig.v_base + 1.0])
    t = rng.normal(size=(2, 3)).astype(np.float32)
    q = np.tile(np.array([0, 0, 0, 1], np.float32), (2, 1))
    d = dataio.save_results(meshes, rig.uv, rig.tex, t, q, str(out), faces_lines=dataio.faces_lines_for(rig.pos_idx, rig.uv_idx))
    back = dataio.MeshData(os.path.join(d, '1.obj'))
    assert np.array_equal(back.vertices, meshes[1]) and np.array_equal(back.faces, rig.pos_idx) and np.array_equal(back.fuv, rig.uv_idx)
    pose = json.load(open(os.path.join(d, 'pose.json')))
    assert list(pose.keys()) == ['rotation', 'translation'] and np.allclose(pose['translation'], t)
    from PIL import Image
    png = np.array(Image.open(os.path.join(d, 'texture.png')))
    assert np.array_equal(png, (np.flip(rig.tex, 0) * 255).astype(np.uint8)[..., 0])
    dataio.write_config(str(out), {'max_iter': 5, 'mode': 'prior'})
    assert open(out / 'config.txt').read() == "max_iter: '5'\nmode: 'prior'\n"


def test_algorithmic_byte_model_matches_survey():
    """bench.py's per-op algorithmic bytes are the formulas of SURVEY.md §8(d): per-pixel terms of the op-boundary chain and the
    fused-path model (56 + 20 C) B/px + geometry the bench line's roofline is quoted on."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(HERE, '..', 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    wl = bench.WORKLOADS['config2']
    b = bench.algorithmic_bytes(wl, 1, wl['V'])
    px = 9 * 1024 * 1024
    geo = 9 * 16 * wl['V'] + 12 * (2 * wl['V'] - 4)
    # rasterize fwd 16 px + geo, bwd 32 px + geo + 16 N V; interpolate (A = 3) fwd 28 px, bwd 44 px (+ index / attribute terms)
    assert b['rasterize_fwd'] == 16 * px + geo and b['rasterize_bwd'] == 32 * px + geo + 16 * 9 * wl['V']
    assert b['interpolate_fwd'] // px == 28 and b['interpolate_bwd'] // px == 44
    assert b['render_loss_fused'] == (56 + 20 * 3) * px + geo + 12 * (2 * wl['V'] - 4) + 4 * 3 * wl['V']
    assert abs(b['render_loss_fused'] / 1e9 - 1.0988) < 1e-3           # the figure the bench line's roofline is quoted on
    # the single-frame GEMV streams D once per direction
    assert b['geometry_fwd'] > 4 * 3 * wl['V'] * wl['B'] and b['geometry_fwd'] < 1.1 * 4 * 3 * wl['V'] * wl['B']


def test_reorder_rig_is_a_pure_renumbering(tiny_rig):
    """fit.reorder_rig (FitConfig.reorder_vertices): vertices renumbered along a Morton curve — a permutation; every triangle
    keeps its position in the list and its corner positions, D / v_base / vcol rows move with their vertex, the uv set is
    untouched; neighbouring corners end up close in index; the result is cached on the rig."""
    from fpc_diffrend_b200.fit import morton_order, reorder_rig
    rig = tiny_rig
    V = rig.V
    r2, perm = reorder_rig(rig)
    assert sorted(perm.tolist()) == list(range(V))
    assert np.array_equal(r2.v_base.reshape(V, 3), rig.v_base.reshape(V, 3)[perm])
    assert np.array_equal(r2.D.reshape(V, 3, -1), rig.D.reshape(V, 3, -1)[perm])
    assert np.array_equal(r2.vcol, rig.vcol[perm])
    assert np.array_equal(r2.v_base.reshape(V, 3)[r2.pos_idx], rig.v_base.reshape(V, 3)[rig.pos_idx])      # same triangles, same order
    assert r2.uv is rig.uv and r2.uv_idx is rig.uv_idx and r2.pos_idx.dtype == np.int32
    spread = lambda idx: np.abs(idx[:, 0].astype(np.int64) - idx[:, 1]).mean()
    assert spread(r2.pos_idx) < 0.25 * spread(rig.pos_idx)
    assert reorder_rig(rig)[0] is r2
    # the curve itself: points on a line come back in line order, whatever the input order
    pts = np.stack([np.linspace(0, 1, 50)] * 3, axis=1)
    shuffle = np.random.default_rng(0).permutation(50)
    assert np.array_equal(shuffle[morton_order(pts[shuffle])], np.arange(50))
