"""Host-side camera math (numpy) — mirror of the reference's camera helpers.

Names and argument meaning follow /root/reference/src/torch/camera.py so a user of the
reference finds the same functions:

* ``intrinsic_to_projection``  — camera.py:27-41  (OpenGL projection from a 3x3 intrinsic;
  the principal point is used only as the half-extent, quirk kept: SURVEY App. B)
* ``extrinsic_to_modelview``   — camera.py:46-66  ([R t; 0 1] with the y and z rows negated,
  OpenCV -> OpenGL axes)
* ``translate``                — camera.py:108-112
* ``unitquat_to_rotmat``       — roma.unitquat_to_rotmat as called at fit.py:548,550 (XYZW order,
  no normalisation)
* ``rigid``                    — camera.py:128-132 (``rigid_grad``)
* ``camera_constants``         — the per-camera constant part of the MVP chain at fit.py:541-546:
  ``P`` and ``A = MV @ translate(0, 170, 0)``

Everything here runs once at set-up time on the host; the per-iteration chain
``mvp = P @ T_frame @ T_cam @ A`` (fit.py:551-553) is evaluated on the device by
``fpc_pose_mvp_fwd`` (csrc/project.cu).
"""
import json

import numpy as np

MODEL_Y_OFFSET = 170.0  # fit.py:545


def intrinsic_to_projection(intr, zn=0.01, zf=200.0):
    intr = np.asarray(intr, dtype=np.float64)
    proj = np.zeros((4, 4), dtype=np.float64)
    proj[0, 0] = intr[0, 0] / intr[0, 2]
    proj[1, 1] = intr[1, 1] / intr[1, 2]
    proj[2, 2] = -(zf + zn) / (zf - zn)
    proj[2, 3] = -(2.0 * zf * zn) / (zf - zn)
    proj[3, 2] = -1.0
    return proj.astype(np.float32)


def extrinsic_to_modelview(rmat, tvec):
    mdv = np.eye(4, dtype=np.float32)
    mdv[:3, :3] = np.asarray(rmat, dtype=np.float32)
    mdv[:3, 3] = np.asarray(tvec, dtype=np.float32).reshape(3)
    mdv[1:3, :] *= -1.0  # flip camera y and z: OpenCV looks down +z, OpenGL down -z
    return mdv


def translate(x, y, z):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = (x, y, z)
    return m


def unitquat_to_rotmat(q):
    """XYZW unit quaternion -> 3x3 rotation (the formula roma/SciPy use; q is NOT normalised here)."""
    x, y, z, w = (np.asarray(q, dtype=np.float64)[i] for i in range(4))
    return np.array([
        [x * x - y * y - z * z + w * w, 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), -x * x + y * y - z * z + w * w, 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), -x * x - y * y + z * z + w * w],
    ], dtype=np.float64)


def rigid(tvec, rotmat):
    m = np.eye(4, dtype=np.float64)
    m[:3, :3] = rotmat
    m[:3, 3] = np.asarray(tvec, dtype=np.float64).reshape(3)
    return m


def load_calibration(path):
    """Read a calibration.json (schema written by the reference's calibrate.py:71-72)."""
    with open(path) as f:
        return json.load(f)


def camera_constants(calib_entries):
    """Per-camera constants of the MVP chain.

    ``calib_entries``: list of dicts with 'intrinsic' [3,3], 'rotation' [3,3], 'translation' [3,1]
    (one entry of calibration.json each, fit.py:514-521).  Returns ``P [C,4,4]`` and
    ``A [C,4,4] = MV @ translate(0,170,0)`` as float32 — the products are formed in float32 like the
    reference does (numpy float32 matrices multiplied with torch.matmul, fit.py:546).
    """
    P, A = [], []
    t170 = translate(0.0, MODEL_Y_OFFSET, 0.0)
    for c in calib_entries:
        intr = np.asarray(c['intrinsic'], dtype=np.float32)
        rot = np.asarray(c['rotation'], dtype=np.float32)
        tr = np.asarray(c['translation'], dtype=np.float32)
        P.append(intrinsic_to_projection(intr))
        A.append((extrinsic_to_modelview(rot, tr) @ t170).astype(np.float32))
    return np.stack(P), np.stack(A)
