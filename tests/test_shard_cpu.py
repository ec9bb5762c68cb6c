"""Multi-GPU host logic on CPU: world_size-2 gloo runs of the two sharding modes.

The per-rank compute is the CPU oracle here (the CUDA path needs a GPU); what is under test is the sharding
arithmetic, the packed-gradient all-reduce of the camera-split mode and the frame gather of the frame-sharded mode."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fpc_diffrend_b200 import shard


def test_split_range_properties():
    for n in (0, 1, 7, 9, 64, 512):
        for world in (1, 2, 3, 4, 8):
            parts = [shard.split_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert [shard.camera_shard(9, r, 2) for r in range(2)] == [(0, 5), (5, 9)]
    assert [b - a for a, b in (shard.camera_shard(9, r, 4) for r in range(4))] == [3, 2, 2, 2]
    assert [b - a for a, b in (shard.camera_shard(9, r, 8) for r in range(8))] == [2, 1, 1, 1, 1, 1, 1, 1]
    assert shard.frame_shard(512, 3, 8) == (192, 256)
    with pytest.raises(ValueError):
        shard.camera_shard(1, 1, 2)
    with pytest.raises(ValueError):
        shard.split_range(4, 2, 2)


def test_view_band_shard_covers_every_bin_row_once():
    """Camera split at bin-row granularity: over all ranks every (view, bin row) is rendered exactly once and the ranks'
    shares differ by at most one row."""
    for C, H, world in ((9, 2048, 8), (9, 2048, 2), (9, 1024, 4), (3, 152, 2), (3, 152, 5), (2, 64, 4), (9, 1600, 16)):
        R = -(-H // 32)
        seen = np.zeros((C, R), int)
        sizes = []
        for r in range(world):
            (c0, c1), (lo, hi) = shard.view_band_shard(C, H, r, world)
            assert 0 <= c0 < c1 <= C and 0 <= lo < R and 0 < hi <= R
            n = 0
            for c in range(c0, c1):
                a = lo if c == c0 else 0
                b = hi if c == c1 - 1 else R
                assert a < b
                seen[c, a:b] += 1
                n += b - a
            sizes.append(n)
        assert (seen == 1).all(), (C, H, world)
        assert max(sizes) - min(sizes) <= 1
    assert shard.view_band_shard(9, 2048, 0, 8) == ((0, 2), (0, 8))          # 72 of 576 rows: view 0 and an eighth of view 1
    with pytest.raises(ValueError):
        shard.view_band_shard(1, 32, 1, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_grads(rig, w, t, q, ref, cams, H, W):
    """Packed [d_w | d_t | d_q] of sum_{c in cams} loss_c / C_total for one frame through the CPU oracle."""
    from oracle import golden as G
    w, t, q = (torch.tensor(x, requires_grad=True) for x in (w, t, q))
    C = rig.P.shape[0]
    verts = G.blend(torch.tensor(rig.v_base), torch.tensor(rig.D), w).reshape(-1, 3)
    total = torch.zeros(())
    for c in cams:
        mvp = G.mvp_chain(torch.tensor(rig.P[c]), torch.tensor(rig.A[c]), t, q)
        img = G.render(mvp, verts, torch.tensor(rig.pos_idx), (H, W), vcol=torch.tensor(rig.vcol), use_antialias=False)
        total = total + G.image_loss(ref[c], img) / C
    total.backward()
    return torch.cat([w.grad.reshape(-1), t.grad.reshape(-1), q.grad.reshape(-1)]), float(total.detach())


def _make_case():
    from fpc_diffrend_b200 import rig as rigmod
    from oracle import golden as G
    H, W = 48, 64
    rig = rigmod.make_rig(n_vertices=300, n_shapes=6, n_cams=3, width=W, height=H, tex_size=16, seed=5)
    w_true, t_true, q_true = rigmod.make_targets(4, rig.B, seed=2)
    refs = []
    for f in range(4):
        verts = G.blend(torch.tensor(rig.v_base), torch.tensor(rig.D), torch.tensor(w_true[f])).reshape(-1, 3)
        views = []
        for c in range(3):
            mvp = G.mvp_chain(torch.tensor(rig.P[c]), torch.tensor(rig.A[c]), torch.tensor(0.2 * t_true[f]), torch.tensor(q_true[f]))
            img = G.render(mvp, verts, torch.tensor(rig.pos_idx), (H, W), vcol=torch.tensor(rig.vcol), use_antialias=False)
            views.append(torch.clamp(img * 255, 0, 140))
        refs.append(views)
    rng = np.random.default_rng(1)
    w0 = (0.05 * rng.random((4, rig.B))).astype(np.float32)
    t0 = (0.1 * rng.normal(size=(4, 3))).astype(np.float32)
    q0 = np.tile(np.array([0, 0, 0, 1], np.float32), (4, 1))
    return rig, refs, w0, t0, q0, H, W


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        rig, refs, w0, t0, q0, H, W = _make_case()
        # ---- camera split: each rank renders its views of frame 0, all-reduce of the packed gradient ----
        c0, c1 = shard.camera_shard(rig.P.shape[0])
        g, part = _oracle_grads(rig, w0[0], t0[0], q0[0], refs[0], range(c0, c1), H, W)
        shard.allreduce_gradients(g)
        # ---- frame shard: each rank fits its own frames (here: one gradient evaluation per frame), gather on rank 0 ----
        f0, f1 = shard.frame_shard(4)
        local = torch.stack([_oracle_grads(rig, w0[f], t0[f], q0[f], refs[f], range(rig.P.shape[0]), H, W)[0] for f in range(f0, f1)])
        allf = shard.gather_frames(local, 4)
        if rank == 0:
            torch.save({'cam_split': g, 'frames': allf, 'slices': (c0, c1, f0, f1)}, out)
        else:
            assert allf is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_camera_split_and_frame_shard(tmp_path):
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    rig, refs, w0, t0, q0, H, W = _make_case()
    assert got['slices'] == (0, 2, 0, 2)
    full = [_oracle_grads(rig, w0[f], t0[f], q0[f], refs[f], range(3), H, W)[0] for f in range(4)]
    # camera split: the all-reduced partial gradients equal the gradient over all views (sum order differs -> tolerance)
    assert torch.allclose(got['cam_split'], full[0], rtol=1e-5, atol=1e-7)
    # frame shard: rank 0 holds every frame's result in frame order, bit-identical to the single-process run
    assert torch.equal(got['frames'], torch.stack(full))
