"""Generates tests/golden/camera_kat.json by importing the REFERENCE's own numpy camera code
(/root/reference/src/torch/camera.py: intrinsic_to_projection :27-41, extrinsic_to_modelview :46-66,
translate :108-112) on the real calibration (/root/reference/calibration/calibration.json).

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_camera_kat.py
"""
import importlib.util
import json
import os

import numpy as np

REF = '/root/reference'
spec = importlib.util.spec_from_file_location('ref_camera', os.path.join(REF, 'src/torch/camera.py'))
ref_camera = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_camera)

with open(os.path.join(REF, 'calibration/calibration.json')) as f:
    calibs = json.load(f)

points = np.array([[0, 0, 0], [5, -3, 8], [-7.5, 10, 2.25]], dtype=np.float32)
out = {'source': 'reference camera.py run on calibration/calibration.json', 'points': points.tolist(), 'cameras': {}}
for name, c in calibs.items():
    intr = np.asarray(c['intrinsic'], dtype=np.float32)
    rot = np.asarray(c['rotation'], dtype=np.float32)
    tr = np.asarray(c['translation'], dtype=np.float32)
    P = ref_camera.intrinsic_to_projection(intr)
    MV = ref_camera.extrinsic_to_modelview(rot, tr)
    T170 = ref_camera.translate(0.0, 170.0, 0.0)
    A = (MV @ T170).astype(np.float32)
    MVP = (P @ A).astype(np.float32)
    clip = (np.concatenate([points, np.ones((3, 1), np.float32)], axis=1) @ MVP.T).astype(np.float32)
    out['cameras'][name] = {
        'intrinsic': c['intrinsic'], 'rotation': c['rotation'], 'translation': c['translation'],
        'P': P.tolist(), 'MV': MV.tolist(), 'A': A.tolist(), 'MVP': MVP.tolist(), 'clip': clip.tolist(),
    }
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'camera_kat.json')
with open(dst, 'w') as f:
    json.dump(out, f, indent=1)
print('wrote', dst, len(out['cameras']), 'cameras')
