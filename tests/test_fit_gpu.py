"""GPU parity of the whole fit iteration (FitSession: pose -> blend -> project -> render -> loss -> backward ->
Adam) against the oracle pipeline (torch CPU stages + golden ops with autograd), BASELINE config 1."""
import numpy as np
import pytest
import torch

from oracle import golden as G

pytestmark = pytest.mark.gpu


def oracle_iteration(rig, params, ref, shading, use_aa, H, W, opp):
    """loss and gradients (d_w [F,B], d_t [F,3], d_q [F,4]) of the batch loss (sum over frames of the mean over
    cameras of fit.py:579's first term) through the oracle."""
    w, t, q = (p.clone().requires_grad_(True) for p in params)
    F, C = w.shape[0], rig.P.shape[0]
    total = 0.0
    for f in range(F):
        verts = G.blend(torch.tensor(rig.v_base), torch.tensor(rig.D), w[f]).reshape(-1, 3)
        for c in range(C):
            mvp = G.mvp_chain(torch.tensor(rig.P[c]), torch.tensor(rig.A[c]), t[f], q[f])
            if shading == 'vcol':
                img = G.render(mvp, verts, torch.tensor(rig.pos_idx), (H, W), vcol=torch.tensor(rig.vcol), tri_opp=opp,
                               use_antialias=use_aa)
            else:
                img = G.render(mvp, verts, torch.tensor(rig.pos_idx), (H, W), uv=torch.tensor(rig.uv),
                               uv_idx=torch.tensor(rig.uv_idx), tex=torch.tensor(rig.tex), tri_opp=opp, use_antialias=use_aa)
            total = total + G.image_loss(ref[f, c], img) / C
    total.backward()
    return float(total.detach()), w.grad, t.grad, q.grad


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def oracle_render_loss(rig, pos_clip, ref, shading, use_aa, H, W, opp, n_cams_total):
    """Golden render chain of fit.py:151-161 + loss for ONE view, starting from clip-space positions."""
    rast, _ = G.rasterize(pos_clip, torch.tensor(rig.pos_idx), (H, W))
    if shading == 'vcol':
        col = G.interpolate(torch.tensor(rig.vcol)[None], rast, torch.tensor(rig.pos_idx))
    else:
        texc = G.interpolate(torch.tensor(rig.uv)[None], rast, torch.tensor(rig.uv_idx))
        col = G.texture(torch.tensor(rig.tex)[None], texc)
    if use_aa:
        col = G.antialias(col, rast, pos_clip, torch.tensor(rig.pos_idx), opp)
    img = torch.where(rast[..., 3:] > 0, col, torch.tensor(G.BG))[0]
    return G.image_loss(ref, img) / n_cams_total


@pytest.mark.parametrize('shading,use_aa,fused,geom', [('vcol', False, True, True), ('vcol', False, True, False),
                                                       ('vcol', False, False, True), ('texture', False, True, True),
                                                       ('texture', True, False, False), ('vcol', True, False, True),
                                                       ('texture', True, True, True), ('vcol', True, True, False)])
def test_iteration_gradients(small_rig3, shading, use_aa, fused, geom):
    """Every link of one fit iteration against the oracle ON IDENTICAL INPUT BITS.

    The chain is not continuous in its inputs (a 1-ulp change of a clip-space coordinate can move a snapped
    vertex across a 1/16-px boundary and flip a silhouette pixel, which alone changes d_w by ~1 %), so the
    comparison is staged: (A) blend / MVP / clip transform vs the torch CPU stages; (B) render + loss and
    d loss / d pos_clip with the oracle fed the GPU's pos_clip; (C) the transpose chain (project bwd, D^T,
    pose bwd) vs torch autograd fed the GPU's d pos_clip."""
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rig, H, W, F = small_rig3, 152, 200, 2
    C = rig.P.shape[0]
    # geom: pose+blend+project fused into one kernel per direction (csrc/geometry.cu) vs the separate kernels
    cfg = FitConfig(resolution=(H, W), shading=shading, antialias=use_aa, fused=fused, fused_geometry=geom)
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=1)
    ref = synthesize_reference(rig, w_true, t_true * 0.2, q_true, cfg)
    s = FitSession(rig, F, cfg)
    assert s.use_fused == fused and s.use_geom_fused == geom
    s.set_reference(ref)
    rng = np.random.default_rng(0)
    w0 = (0.05 * rng.random((F, rig.B))).astype(np.float32)
    t0 = (0.1 * rng.normal(size=(F, 3))).astype(np.float32)
    q0 = rng.normal(size=(F, 4)).astype(np.float32) * 0.01 + np.array([0, 0, 0, 1], np.float32)
    q0 /= np.linalg.norm(q0, axis=1, keepdims=True)
    s.set_parameters(w=w0, t=t0, q=q0)
    s.forward()
    s.backward()
    torch.cuda.synchronize()
    opp = torch.tensor(G.topology_build(rig.pos_idx))
    ref_cpu = ref.cpu()

    # (A) forward torch stages
    w, t, q = (torch.tensor(x, requires_grad=True) for x in (w0, t0, q0))
    pcs = []
    for f in range(F):
        verts = G.blend(torch.tensor(rig.v_base), torch.tensor(rig.D), w[f]).reshape(-1, 3)
        assert rel(s.verts[f].cpu().reshape(-1, 3), verts.detach()) < 1e-6
        for c in range(C):
            mvp = G.mvp_chain(torch.tensor(rig.P[c]), torch.tensor(rig.A[c]), t[f], q[f])
            assert rel(s.mvp[f * C + c].cpu().reshape(4, 4), mvp.detach()) < 1e-6
            pc = G.transform_clip(mvp, verts)
            assert rel(s.pos_clip[f * C + c].cpu(), pc[0].detach()) < 1e-6
            pcs.append(pc)

    # (B) render + loss from the GPU's own pos_clip bits
    total, g_pos_ref = 0.0, []
    for n in range(F * C):
        pc = s.pos_clip[n:n + 1].cpu().clone().requires_grad_(True)
        loss = oracle_render_loss(rig, pc, ref_cpu[n // C, n % C], shading, use_aa, H, W, opp, C)
        loss.backward()
        total += float(loss.detach())
        g_pos_ref.append(pc.grad[0])
    g_pos_ref = torch.stack(g_pos_ref)
    assert abs(float(s.loss) - total) / total < 1e-5
    assert rel(s.g_pos.cpu(), g_pos_ref) < 1e-4

    # (C) transpose chain from the GPU's d pos_clip
    g_pos_gpu = s.g_pos.cpu()
    sum((pcs[n][0] * g_pos_gpu[n]).sum() for n in range(F * C)).backward()
    assert rel(s.d_w.cpu(), w.grad) < 1e-4
    assert rel(s.d_t.cpu(), t.grad) < 1e-4
    assert rel(s.d_q.cpu(), q.grad) < 1e-4


@pytest.mark.parametrize('geom', [True, False])
def test_iteration_with_mesh_regularisers(small_rig3, geom):
    """Whole-iteration gradients with the shipped mesh terms on (fit.py:578-582): the GPU's d_w must equal the image-only
    d_w plus D^T (d reg / d V) from the oracle's torch restatement, and the loss the sum of both."""
    from fpc_diffrend_b200 import rig as rigmod, topology
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rig, H, W, F = small_rig3, 152, 200, 2
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=1)
    kw = dict(resolution=(H, W), shading='vcol', fused_geometry=geom)
    ref = synthesize_reference(rig, w_true, t_true * 0.2, q_true, FitConfig(**kw))
    rng = np.random.default_rng(0)
    w0 = (0.3 * rng.random((F, rig.B))).astype(np.float32)
    out = {}
    for name, reg in (('img', {}), ('reg', dict(weight_laplacian=5000.0, weight_meshedge=70.0, meshedge_target=0.05,
                                                  weight_normalconsistency=400.0))):
        s = FitSession(rig, F, FitConfig(**kw, **reg))
        assert s.use_geom_fused == geom
        s.set_reference(ref)
        s.set_parameters(w=w0)
        s.forward(); s.backward()
        torch.cuda.synchronize()
        out[name] = (float(s.loss), s.d_w.cpu().clone(), s.d_t.cpu().clone(), s.verts.cpu().clone())
    tp = topology.build_topology(rig.pos_idx, rig.V)
    verts = out['img'][3].reshape(F, -1, 3).clone().requires_grad_(True)
    tot = sum(G.mesh_regularisers(verts[f], torch.tensor(tp.edges).long(), torch.tensor(tp.edge_quads).long(), 5000.0, 70.0, 0.05, 400.0)[0]
              for f in range(F))
    tot.backward()
    d_w_reg = verts.grad.reshape(F, -1) @ torch.tensor(rig.D)            # D^T d V per frame
    assert abs(out['reg'][0] - out['img'][0] - float(tot.detach())) <= 1e-5 * abs(out['reg'][0])
    assert rel(out['reg'][1] - out['img'][1], d_w_reg) < 1e-4
    assert rel(out['reg'][2], out['img'][2]) < 1e-5                        # the mesh terms do not see the pose (float REDs: not bitwise)


@pytest.mark.parametrize('geom', [True, False])
def test_camera_pose_correction_gradients(small_rig3, geom):
    """Per-camera pose corrections t_opt / q_opt (fit.py:443-448): forward through the MVP chain and the gradient summed over
    the frames, against torch autograd through the oracle's mvp_chain fed the GPU's d mvp."""
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rig, H, W, F = small_rig3, 152, 200, 2
    C = rig.P.shape[0]
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=1)
    cfg = FitConfig(resolution=(H, W), shading='vcol', fused_geometry=geom, optimize_cam_pose=True)
    ref = synthesize_reference(rig, w_true, t_true * 0.2, q_true, FitConfig(resolution=(H, W), shading='vcol'))
    s = FitSession(rig, F, cfg)
    s.set_reference(ref)
    rng = np.random.default_rng(3)
    tc = (0.2 * rng.normal(size=(C, 3))).astype(np.float32)
    qc = rng.normal(size=(C, 4)).astype(np.float32) * 0.02 + np.array([0, 0, 0, 1], np.float32)
    qc /= np.linalg.norm(qc, axis=1, keepdims=True)
    t0 = (0.1 * rng.normal(size=(F, 3))).astype(np.float32)
    s.set_parameters(w=(0.05 * rng.random((F, rig.B))).astype(np.float32), t=t0)
    s.t_cam.copy_(torch.tensor(tc)); s.q_cam.copy_(torch.tensor(qc))
    s.forward(); s.backward()
    torch.cuda.synchronize()
    tct, qct = torch.tensor(tc, requires_grad=True), torch.tensor(qc, requires_grad=True)
    tf, qf = s.t.cpu().clone().requires_grad_(True), s.q.cpu().clone().requires_grad_(True)
    d_mvp = s.d_mvp.cpu().reshape(F, C, 4, 4)
    tot = 0.0
    for f in range(F):
        for c in range(C):
            mvp = G.mvp_chain(torch.tensor(rig.P[c]), torch.tensor(rig.A[c]), tf[f], qf[f], tct[c], qct[c])
            assert rel(s.mvp[f * C + c].cpu().reshape(4, 4), mvp.detach()) < 1e-6
            tot = tot + (mvp * d_mvp[f, c]).sum()
    tot.backward()
    assert rel(s.d_t_cam.cpu(), tct.grad) < 1e-4 and rel(s.d_q_cam.cpu(), qct.grad) < 1e-4
    assert rel(s.d_t.cpu(), tf.grad) < 1e-4 and rel(s.d_q.cpu(), qf.grad) < 1e-4
    before = s.cam_params.clone()
    s.optimizer_step()
    torch.cuda.synchronize()
    assert (s.cam_params != before).any() and torch.allclose(s.q_cam.norm(dim=1), torch.ones(C, device='cuda'), atol=1e-6)


@pytest.mark.parametrize('use_aa,fused', [(False, True), (True, True), (True, False)])
def test_texture_optimisation(small_rig3, use_aa, fused):
    """tex_opt (fit.py:439,502): the texture as a shared parameter.  d loss / d tex of one iteration against the oracle fed
    the GPU's pos_clip bits, then a fixed number of texture-only Adam steps against torch.optim.Adam driven by the same
    gradients' oracle, and the loss must fall."""
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rig, H, W, F = small_rig3, 152, 200, 2
    C = rig.P.shape[0]
    cfg = FitConfig(resolution=(H, W), shading='texture', antialias=use_aa, fused=fused, optimize_texture=True, optimize_pose=False,
                    lr_base=2e-2, lr_tex_coef=0.5, max_iter=100)
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=1)
    ref = synthesize_reference(rig, w_true, t_true * 0.2, q_true, cfg)
    s = FitSession(rig, F, cfg)
    s.set_reference(ref)
    s.set_parameters(w=w_true, t=t_true * 0.2, q=q_true)
    tex0 = (0.7 * torch.tensor(rig.tex) + 0.1).reshape(s.tex.shape)
    s.tex.copy_(tex0)
    s.forward()
    s.backward()
    torch.cuda.synchronize()
    opp = torch.tensor(G.topology_build(rig.pos_idx))
    ref_cpu = ref.cpu()

    def oracle_tex_grad(tex):
        tex = tex.clone().requires_grad_(True)
        total = 0.0
        for n in range(F * C):
            pc = s.pos_clip[n:n + 1].cpu()
            rast, _ = G.rasterize(pc, torch.tensor(rig.pos_idx), (H, W))
            texc = G.interpolate(torch.tensor(rig.uv)[None], rast, torch.tensor(rig.uv_idx))
            col = G.texture(tex, texc)
            if use_aa:
                col = G.antialias(col, rast, pc, torch.tensor(rig.pos_idx), opp)
            img = torch.where(rast[..., 3:] > 0, col, torch.tensor(G.BG))[0]
            total = total + G.image_loss(ref_cpu[n // C, n % C], img) / C
        total.backward()
        return float(total.detach()), tex.grad

    loss_ref, g_ref = oracle_tex_grad(tex0)
    assert abs(float(s.loss) - loss_ref) / loss_ref < 1e-5
    assert g_ref.abs().max() > 0
    assert rel(s.d_tex.cpu(), g_ref) < 1e-4

    # texture-only fit: the activations are restored before every step (the oracle below optimises the texture alone)
    iters = 6
    tex_o = tex0.clone().requires_grad_(True)
    opt = torch.optim.Adam([{'params': tex_o, 'lr': s.cfg.lr_base * s.cfg.lr_tex_coef}])
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda x: cfg.lr_ramp ** (float(x) / float(cfg.max_iter)))
    losses = []
    for _ in range(iters):
        s.set_parameters(w=w_true, t=t_true * 0.2, q=q_true)
        s.iteration()
        losses.append(float(s.loss))
        _, g = oracle_tex_grad(tex_o.detach())
        opt.zero_grad()
        tex_o.grad = g
        opt.step()
        sched.step()
    assert losses[-1] < losses[0]
    travel = s.cfg.lr_base * s.cfg.lr_tex_coef * iters
    d = (s.tex.cpu() - tex_o.detach()).abs()
    # texels whose gradient is rounding noise move by +-lr per step either way (Adam divides by sqrt(v)): tight on the texels
    # with a real signal, loose elsewhere (same criterion as test_fitted_parameters_after_fixed_iterations)
    well = g_ref.abs() > 1e-2 * g_ref.abs().max()
    assert well.sum() > 50
    assert d[well].max() < 2e-3 * travel + 1e-6, float(d[well].max())
    assert d.max() <= 2.01 * travel


def test_fitted_parameters_after_fixed_iterations(tiny_rig):
    """North-star: fitted activations after a fixed iteration count must match the reference path within tolerance.
    Deterministic all-frames schedule, 12 iterations, config 1 (1 camera 128x128, 1 frame)."""
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rig, H, W, F, iters = tiny_rig, 128, 128, 1, 12
    cfg = FitConfig(resolution=(H, W), shading='vcol', antialias=True, lr_base=1e-2, lr_t=1e-3, lr_q=1e-4, max_iter=100)
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=2)
    ref = synthesize_reference(rig, w_true, t_true * 0.2, q_true, cfg)
    s = FitSession(rig, F, cfg)
    s.set_reference(ref)
    losses = []
    for _ in range(iters):
        s.iteration()
        losses.append(float(s.loss))
    torch.cuda.synchronize()

    # oracle: torch.optim.Adam + LambdaLR exactly as fit.py:493-505,610-618 (per-row quaternion renorm)
    w = torch.zeros(F, rig.B, requires_grad=True)
    t = torch.zeros(F, 3, requires_grad=True)
    q = torch.tensor([[0., 0, 0, 1]] * F, requires_grad=True)
    opt = torch.optim.Adam([{'params': w, 'lr': cfg.lr_base}, {'params': t, 'lr': cfg.lr_t}, {'params': q, 'lr': cfg.lr_q}])
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda x: cfg.lr_ramp ** (float(x) / float(cfg.max_iter)))
    opp = torch.tensor(G.topology_build(rig.pos_idx))
    ref_cpu = ref.cpu()
    ref_losses, grad_log = [], []
    for _ in range(iters):
        loss, gw, gt, gq = oracle_iteration(rig, (w.detach(), t.detach(), q.detach()), ref_cpu, 'vcol', True, H, W, opp)
        ref_losses.append(loss)
        grad_log.append(gw[0].numpy().copy())
        opt.zero_grad()
        w.grad, t.grad, q.grad = gw, gt, gq
        opt.step()
        sched.step()
        with torch.no_grad():
            q /= q.norm(dim=1, keepdim=True)
    assert losses[-1] < losses[0]
    # the loss trajectory is chaotic at the 1e-3 level (single silhouette-pixel flips, see test_iteration_gradients)
    np.testing.assert_allclose(losses, ref_losses, rtol=5e-3)
    # Adam divides by sqrt(v): a component whose gradient is ~0 (a blendshape the camera barely sees) moves by
    # +-lr per step on rounding noise alone, so parity is asserted tightly (1e-4 of the distance travelled + 1e-5)
    # on the well-conditioned components and loosely (5 % of the maximum travel) on the rest.
    gsig = np.stack(grad_log)                                  # [iters, B] oracle activation gradients
    well = (np.abs(gsig) > 1e-2 * np.abs(gsig).max()).all(axis=0)
    assert well.sum() >= 3
    dw = np.abs(s.w.cpu().numpy() - w.detach().numpy())[0]
    travel = cfg.lr_base * iters
    assert dw[well].max() < 1e-4 * travel + 1e-5, dw
    assert dw.max() < 0.05 * travel, dw
    assert np.abs(s.t.cpu().numpy() - t.detach().numpy()).max() < 0.05 * cfg.lr_t * iters
    assert np.abs(s.q.cpu().numpy() - q.detach().numpy()).max() < 0.05 * cfg.lr_q * iters


def test_graph_replay_matches_eager(tiny_rig):
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rig, F = tiny_rig, 2
    cfg = FitConfig(resolution=(128, 128), shading='texture', antialias=True, lr_base=1e-2, max_iter=100)
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=3)
    ref = synthesize_reference(rig, w_true, t_true * 0.2, q_true, cfg)
    a, b = FitSession(rig, F, cfg), FitSession(rig, F, cfg)
    a.set_reference(ref)
    b.set_reference(ref)
    for _ in range(5):
        a.iteration()
    b.capture()                      # one eager warm-up iteration; the capture itself does not execute
    for _ in range(4):
        b.replay()
    torch.cuda.synchronize()
    assert float(b.step_count) == 5.0 and float(a.step_count) == 5.0
    # no atomics on the gradient path (per-bin slots + per-vertex gather, fixed-order reductions everywhere else): two runs of
    # the same fit agree to the last bit, whether enqueued eagerly or replayed from a graph
    assert torch.equal(a.w, b.w) and torch.equal(a.t, b.t) and torch.equal(a.q, b.q)


@pytest.mark.parametrize('shading,use_aa,size', [('vcol', False, 'small'), ('texture', True, 'small'), ('texture', True, 'config2'), ('vcol', False, 'config2')])
def test_gradients_are_bit_reproducible(small_rig3, shading, use_aa, size):
    """North-star item 4 ("atomics-free gradient scatter"): d loss / d pos_clip, the packed parameter gradient and the fitted
    parameters are BIT-identical between repeated runs — at the small rig and at BASELINE config 2 size (20k vertices, 9 views
    1024 x 1024), with and without antialias.  (No float atomic is left between the image and the activations: per-(view,
    triangle, bin) gradient slots written once, a per-vertex gather in adjacency order, fixed-order reductions in the geometry
    backward and the loss.)"""
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    if size == 'small':
        rig, H, W, F = small_rig3, 152, 200, 2
    else:
        H = W = 1024
        F = 1
        rig = rigmod.make_rig(n_vertices=20000, n_shapes=200, n_cams=9, width=W, height=H, tex_size=256, seed=0)
    cfg = FitConfig(resolution=(H, W), shading=shading, antialias=use_aa, lr_base=1e-2)
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=1)
    ref = synthesize_reference(rig, w_true, 0.2 * t_true, q_true, cfg)
    runs = []
    for rep in range(3):
        s = FitSession(rig, F, cfg)
        s.set_reference(ref)
        s.forward(); s.backward()
        torch.cuda.synchronize()
        g_pos, grads = s.g_pos.clone(), s.grads.clone()
        assert float(g_pos.abs().max()) > 0
        for _ in range(4):
            s.iteration()
        torch.cuda.synchronize()
        runs.append((g_pos, grads, s.params.clone(), float(s.loss)))
        del s
    for r in runs[1:]:
        assert torch.equal(r[0], runs[0][0]), 'd loss / d pos_clip differs between runs'
        assert torch.equal(r[1], runs[0][1]), 'packed gradient differs between runs'
        assert torch.equal(r[2], runs[0][2]), 'fitted parameters differ between runs'
        assert r[3] == runs[0][3]


@pytest.mark.parametrize('shading,use_aa,reg', [('vcol', False, False), ('texture', True, True)])
def test_vertex_reordering_changes_nothing_visible(small_rig3, shading, use_aa, reg):
    """FitConfig.reorder_vertices (what fit_take and bench.py run with): the session renumbers the vertices along a Morton curve.
    Triangle ids do not change, so the loss of the first iteration is the same to the last bit; the per-vertex gradient is the
    unordered session's, permuted (bit-equal: same slots, same adjacency order per vertex); the packed gradient differs only by
    the summation order of D^T over the vertices; result_vertices() comes back in the rig's order."""
    from dataclasses import replace
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rig, H, W, F = small_rig3, 152, 200, 2
    cfg = FitConfig(resolution=(H, W), shading=shading, antialias=use_aa, lr_base=1e-2,
                    weight_laplacian=5.0 if reg else 0.0, weight_meshedge=0.5 if reg else 0.0)
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=1)
    ref = synthesize_reference(rig, w_true, 0.2 * t_true, q_true, cfg)
    a, b = FitSession(rig, F, cfg), FitSession(rig, F, replace(cfg, reorder_vertices=True))
    assert b.vertex_order is not None and not torch.equal(b.vertex_order, torch.arange(rig.V, device='cuda'))
    for s in (a, b):
        s.set_reference(ref)
        s.forward(); s.backward()
    torch.cuda.synchronize()
    assert float(a.loss) == float(b.loss)
    assert torch.equal(a.g_pos[:, b.vertex_order], b.g_pos)
    rel = lambda x, y: float((x - y).abs().max() / y.abs().max().clamp_min(1e-30))
    assert rel(b.grads, a.grads) < 2e-5, rel(b.grads, a.grads)
    assert torch.equal(a.result_vertices(), b.result_vertices())
    for s in (a, b):
        for _ in range(5):
            s.iteration()
    torch.cuda.synchronize()
    assert abs(float(a.loss) - float(b.loss)) <= 1e-4 * abs(float(a.loss))
    assert rel(b.result_vertices(), a.result_vertices()) < 1e-4


def test_fit_stream_matches_resident(tiny_rig):
    """Pipelined host-buffer API (double-buffered uploads, one graph per buffer) == resident-frame iterations."""
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rig, F = tiny_rig, 2
    cfg = FitConfig(resolution=(128, 128), shading='vcol', antialias=False, lr_base=1e-2, max_iter=100, ref_dtype='u8')
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=3)
    ref = synthesize_reference(rig, w_true, t_true * 0.2, q_true, cfg).round().clamp(0, 255).to(torch.uint8)
    a, b = FitSession(rig, F, cfg), FitSession(rig, F, cfg)
    a.set_reference(ref)
    la = []
    for _ in range(6):
        a.iteration()
        la.append(float(a.loss))
    host = ref.cpu().pin_memory()
    lb = list(b.fit_stream((host for _ in range(6))))
    assert len(lb) == 6 and float(b.step_count) == 6.0
    np.testing.assert_allclose(lb, la, rtol=1e-4)
    assert torch.allclose(a.w, b.w, atol=1e-5)


@pytest.mark.parametrize('use_aa', [False, True])
def test_view_band_split_adds_up(small_rig3, use_aa):
    """Camera split at bin-row granularity (shard.view_band_shard, fpc_render_loss_fused_band): the partial losses and packed
    gradients of the shares of 2, 3 and 5 ranks — rendered here one after the other on one GPU — add up to the unsplit
    iteration (every pixel and every antialias pair is owned by exactly one bin; only the summation order differs)."""
    from fpc_diffrend_b200 import rig as rigmod, shard
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rig, H, W, F = small_rig3, 152, 200, 2
    C = rig.P.shape[0]
    base = dict(resolution=(H, W), shading='texture', antialias=use_aa, optimize_texture=True)
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=1)
    ref = synthesize_reference(rig, w_true, 0.2 * t_true, q_true, FitConfig(**base))
    rng = np.random.default_rng(0)
    w0 = (0.05 * rng.random((F, rig.B))).astype(np.float32)
    full = FitSession(rig, F, FitConfig(**base))
    full.set_reference(ref)
    full.set_parameters(w=w0)
    full.forward(); full.backward()
    torch.cuda.synchronize()
    for world in (2, 3, 5):
        loss, grads, d_tex = 0.0, torch.zeros_like(full.grads), torch.zeros_like(full.d_tex)
        for r in range(world):
            sl, band = shard.view_band_shard(C, H, r, world)
            s = FitSession(rig, F, FitConfig(cam_slice=sl, cam_band=band, **base))
            s.set_reference(ref[:, sl[0]:sl[1]])
            s.set_parameters(w=w0)
            s.forward(); s.backward()
            torch.cuda.synchronize()
            loss += float(s.loss)
            grads += s.grads
            d_tex += s.d_tex
        assert abs(loss - float(full.loss)) / float(full.loss) < 1e-5, world
        assert rel(grads.cpu(), full.grads.cpu()) < 1e-5, world
        assert rel(d_tex.cpu(), full.d_tex.cpu()) < 1e-5, world


def _cam_split_worker(rank, world, port, out, peer=None):
    import os
    import torch.distributed as dist
    from fpc_diffrend_b200 import rig as rigmod, shard
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        rig = rigmod.make_rig(n_vertices=600, n_shapes=8, n_cams=3, width=200, height=152, tex_size=32, seed=3)
        F, H, W = 2, 152, 200
        w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=1)
        # rank 0 renders view 0 and the lower half of view 1, rank 1 the rest (bin-row granularity)
        sl, band = shard.view_band_shard(3, H)
        cfg = FitConfig(resolution=(H, W), shading='texture', antialias=True, cam_slice=sl, cam_band=band, lr_base=1e-2, peer_exchange=peer)
        ref = synthesize_reference(rig, w_true, 0.2 * t_true, q_true, cfg)
        s = FitSession(rig, F, cfg)
        assert s.row_shard is not None and (s.peer is not None) == bool(peer)
        s.set_reference(ref)
        s.iteration()
        torch.cuda.synchronize()
        g1 = s.g_total.cpu().clone()
        s.capture()                      # the exchanges (NCCL collectives or peer-memory kernels + barriers) are captured too
        s.replay()
        torch.cuda.synchronize()
        verts = s.result_vertices()          # a collective in the row-sharded mode: every rank calls it
        torch.cuda.synchronize()
        if rank == 0:
            torch.save({'grads': g1, 'params': s.params.cpu(), 'verts': verts.cpu()}, out)
        # captured graphs hold NCCL work: drop them before the process group goes away
        s.invalidate_graphs()
        del s
        torch.cuda.synchronize()
        dist.barrier()
    finally:
        import threading
        t = threading.Timer(20.0, os._exit, (0 if os.path.exists(out) or rank != 0 else 1,))     # never let a stuck teardown hang the test
        t.daemon = True
        t.start()
        dist.destroy_process_group()
        t.cancel()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
@pytest.mark.parametrize('peer', [False, True])
def test_camera_split_two_gpus(tmp_path, peer):
    """Camera-split mode on 2 GPUs, rows of D sharded, exchanges over NCCL (peer=False) or over NVLink peer memory (peer=True:
    blend + all-gather in one kernel, peer sums, device-side barriers): the summed packed gradient of 2 ranks (1.5 + 1.5 views,
    cut at a bin row) equals the single-GPU gradient over all 3 views (summation order differs -> tolerance), the replicated
    Adam steps and the blended vertices follow — through three iterations, two of them replayed from a captured graph."""
    import socket
    import torch.multiprocessing as mp
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    sock = socket.socket(); sock.bind(('127.0.0.1', 0)); port = sock.getsockname()[1]; sock.close()
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_cam_split_worker, args=(2, port, out, peer), nprocs=2, join=True)
    got = torch.load(out)
    rig = rigmod.make_rig(n_vertices=600, n_shapes=8, n_cams=3, width=200, height=152, tex_size=32, seed=3)
    F, H, W = 2, 152, 200
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=1)
    cfg = FitConfig(resolution=(H, W), shading='texture', antialias=True, lr_base=1e-2)
    ref = synthesize_reference(rig, w_true, 0.2 * t_true, q_true, cfg)
    s = FitSession(rig, F, cfg)
    s.set_reference(ref)
    s.iteration()
    torch.cuda.synchronize()
    assert rel(got['grads'], s.grads.cpu()) < 1e-5                 # first iteration: identical parameters on both sides
    for _ in range(2):
        s.iteration()
    torch.cuda.synchronize()
    # three Adam steps later (sign-like steps amplify last-bit differences of small gradient components)
    assert rel(got['params'], s.params.cpu()) < 2e-2
    assert rel(got['verts'], s.result_vertices().cpu()) < 1e-3


def test_fit_take_from_disk(tmp_path):
    """The callers / formats either side of the hot path: a synthetic take written in the reference's
    on-disk layout (OBJ base + blendshape directory, calibration.json, <cam>/<cam>_NN.tif) is fitted by fit_take and the
    results come back in the reference's output layout; the image loss of every frame batch must decrease."""
    import json
    import os
    from fpc_diffrend_b200 import dataio, rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, synthesize_reference
    from fpc_diffrend_b200.take import fit_take
    from test_host_cpu import _write_take
    H, W, F = 96, 128, 3
    rig = rigmod.make_rig(n_vertices=400, n_shapes=6, n_cams=2, width=W, height=H, tex_size=32, seed=2)
    cams = ['take_%s' % k for k in rig.calib.keys()]
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=3)
    w_true = np.full((F, rig.B), 0.8, np.float32) * (1.0 + 0.25 * np.arange(F, dtype=np.float32)[:, None])
    cfg = FitConfig(resolution=(H, W), shading='texture', antialias=True)
    ref = synthesize_reference(rig, w_true, 0.0 * t_true, q_true * 0 + np.array([0, 0, 0, 1], np.float32), cfg, out_dtype=torch.uint8)
    base, bdir, imdir, calib = _write_take(tmp_path, rig, ref.cpu().numpy(), cams)
    from PIL import Image
    texpath = str(tmp_path / 'tex.png')
    Image.fromarray((np.flip(rig.tex, 0) * 255).astype(np.uint8)[..., 0]).save(texpath)
    out_dir = tmp_path / 'out'
    out_dir.mkdir()
    fit_cfg = FitConfig(shading='texture', antialias=True, lr_base=5e-3, optimize_pose=False, lr_ramp=1.0, max_iter=400, weight_laplacian=1.0)
    res0 = fit_take(base, bdir, imdir, calib, None, iters_per_frame=0, frame_batch=2, cams=cams, texpath=texpath, config=fit_cfg,
                    blend_order='sorted', use_graph=False)
    res1 = fit_take(base, bdir, imdir, calib, None, iters_per_frame=1, frame_batch=2, cams=cams, texpath=texpath, config=fit_cfg,
                    blend_order='sorted', use_graph=False)
    res = fit_take(base, bdir, imdir, calib, str(out_dir), iters_per_frame=150, frame_batch=2, cams=cams, texpath=texpath, config=fit_cfg,
                   blend_order='sorted')
    assert all(b < a for a, b in zip(res1['loss'], res['loss'])), (res1['loss'], res['loss'])
    assert res['vertices'].shape == (F, rig.V * 3) and res['w'].shape == (F, rig.B) and len(res['loss']) == 2
    # zero iterations return the base mesh; the fit moves the vertices (the image loss above is what it minimises — with two
    # views the 3-D shape itself is not uniquely determined, so no claim about the distance to the ground truth is made)
    assert np.allclose(res0['vertices'], rig.v_base[None], atol=1e-5)
    assert np.abs(res['vertices'] - rig.v_base[None]).max() > 1e-2 and np.abs(res['w']).max() > 1e-2
    d = out_dir / 'result'
    assert sorted(os.listdir(d)) == ['0.obj', '1.obj', '2.obj', 'pose.json', 'texture.png']
    back = dataio.MeshData(str(d / '2.obj'))
    assert np.array_equal(back.vertices, res['vertices'][2]) and np.array_equal(back.faces, rig.pos_idx)
    pose = json.load(open(d / 'pose.json'))
    assert np.allclose(pose['rotation'], res['q']) and np.allclose(pose['translation'], res['t'])
    assert "iters_per_frame: '150'" in open(out_dir / 'config.txt').read()

    # forward-only result renderer on the drop-in (role of render_multicam.py / render_result.py): the saved meshes, texture and
    # pose rendered from every camera reproduce what the fit session itself renders from the fitted parameters
    from fpc_diffrend_b200.fit import FitSession
    from fpc_diffrend_b200.render import render_result
    imgs = render_result(str(d), calib, cams, (H, W), reproduce_pose=True, y_offset=170.0, out_dir=str(tmp_path / 'rendered'))
    assert imgs.shape == (F, len(cams), H, W, 1) and imgs.dtype == np.uint8
    s = FitSession(rig, F, FitConfig(resolution=(H, W), shading='texture', antialias=True))
    s.tex.copy_(torch.tensor((rig.tex * 255).astype(np.uint8) / 255.0).reshape(s.tex.shape))   # texture.png: 8-bit, truncated (fit.py:270)
    s.set_parameters(w=res['w'], t=res['t'], q=res['q'])
    own = torch.flip(s.forward(with_loss=False).reshape(F, len(cams), H, W, 1), dims=[2]) * 255.0
    diff = np.abs(own.cpu().numpy() - imgs.astype(np.float32))
    assert diff.max() <= 2.0 and (diff > 0.51).mean() < 1e-3, (diff.max(), (diff > 0.51).mean())
    assert len(os.listdir(tmp_path / 'rendered')) == F * len(cams)


def test_basis_kernels_match_reference_kat():
    """fpc_basis_code_fwd + fpc_blend_fwd_ex and their backward (fpc_blend_bwd, fpc_basis_grad, fpc_basis_code_bwd) on the
    vectors the reference's own blend_free / blend_combined produced (tests/golden/blend_kat.json), all frames as one batch."""
    import ctypes
    import json
    import os
    from fpc_diffrend_b200 import _lib
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'blend_kat.json')) as f:
        k = json.load(f)
    R, B, Fn = k['R'], k['B'], k['F']
    cu = lambda name: torch.tensor(k[name], dtype=torch.float32).cuda().contiguous()
    v_base, D, M1, M2, m1, m2, m3, dy = (cu(n) for n in ('v_base', 'D', 'M1', 'M2', 'm1', 'm2', 'm3', 'dy'))
    ids = torch.tensor([3, 0, 4, 1, 2], dtype=torch.int32).cuda()         # a shuffled batch of all frames
    Fb = ids.numel()
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    x1, x2 = torch.empty(Fb, Fn, device='cuda'), torch.empty(Fb, Fn, device='cuda')
    _lib.call('fpc_basis_code_fwd', P(m1), P(m2), P(ids), Fn, Fb, P(x1), P(x2), st)
    scratch = torch.empty(int(_lib.load().fpc_blend_bwd_scratch_bytes(R, max(B, Fn), Fb)), dtype=torch.uint8, device='cuda')
    d_verts = dy[None].repeat(Fb, 1).contiguous()
    for mode, coef in (('free', 1.0), ('combined', 0.5)):
        verts = torch.full((Fb, R), float('nan'), device='cuda')
        if mode == 'free':
            _lib.call('fpc_blend_fwd_ex', P(m3), P(v_base), P(x2), R, Fn, Fb, 1.0, 0, P(verts), st)
        else:
            w = (M2 @ M1[:, ids.long()]).t().contiguous()                  # [Fb,B] activations of the prior half
            _lib.call('fpc_blend_fwd', P(D), P(v_base), P(w), R, B, Fb, P(verts), st)
            _lib.call('fpc_blend_fwd_ex', P(m3), None, P(x2), R, Fn, Fb, coef, 1, P(verts), st)
        d_x2 = torch.empty(Fb, Fn, device='cuda')
        d_m1, d_m2, d_m3 = torch.full((Fn, Fn), 7.0, device='cuda'), torch.full((Fn, Fn), 7.0, device='cuda'), torch.full((R, Fn), 7.0, device='cuda')
        _lib.call('fpc_blend_bwd', P(m3), P(d_verts), R, Fn, Fb, P(d_x2), P(scratch), scratch.numel(), st)
        _lib.call('fpc_basis_grad', P(d_verts), P(x2), R, Fn, Fb, coef, P(d_m3), st)
        _lib.call('fpc_basis_code_bwd', P(m2), P(x1), P(d_x2), P(ids), Fn, Fb, coef, P(d_m1), P(d_m2), st)
        torch.cuda.synchronize()
        g1, g2, g3 = torch.zeros(Fn, Fn), torch.zeros(Fn, Fn), torch.zeros(R, Fn)
        for b, f in enumerate(ids.tolist()):
            rec = k['frames'][str(f)][mode]
            ref = torch.tensor(rec['vtx_pos'])
            assert (verts[b].cpu() - ref).abs().max() <= 1e-5 * ref.abs().max(), (mode, f)
            g1 += torch.tensor(rec['grad']['m1']); g2 += torch.tensor(rec['grad']['m2']); g3 += torch.tensor(rec['grad']['m3'])
        # the batch gradient is the sum of the reference's single-frame gradients
        assert rel(d_m1.cpu(), g1) < 1e-5 and rel(d_m2.cpu(), g2) < 1e-5 and rel(d_m3.cpu(), g3) < 1e-5, mode


def _oracle_mode_iteration(rig, s, mode, ids, ref, H, W, opp, reg_corr=False, reg_prior=False):
    """Loss and parameter gradients of one batch iteration in the given mode through the oracle (torch CPU stages + golden
    ops with autograd), the blend written exactly as the reference's blend / blend_free / blend_combined."""
    F, C, Fn = s.F, rig.P.shape[0], (s.Fn if mode != 'prior' else s.F)
    w = s.w.cpu().clone().requires_grad_(True)
    t = s.t.cpu().clone().requires_grad_(True)
    q = s.q.cpu().clone().requires_grad_(True)
    ms = [x.cpu().clone().requires_grad_(True) for x in (s.m1, s.m2, s.m3)] if mode != 'prior' else None
    vb, D = torch.tensor(rig.v_base), torch.tensor(rig.D)
    total = 0.0
    for f in range(F):
        if mode == 'prior':
            verts = G.blend(vb, D, w[f])
        else:
            e = torch.zeros(Fn)
            e[int(ids[f])] = 1.0
            corr = ms[2] @ (ms[1] @ (ms[0] @ e))
            verts = G.blend_free(vb, ms[0], ms[1], ms[2], e) if mode == 'free' else G.blend(vb, D, w[f]) + s.basis_coef * corr
            if reg_corr:
                total = total + torch.mean(corr ** 2)
        if reg_prior:
            total = total + torch.mean(w[f] ** 2)
        verts = verts.reshape(-1, 3)
        for c in range(C):
            mvp = G.mvp_chain(torch.tensor(rig.P[c]), torch.tensor(rig.A[c]), t[f], q[f])
            img = G.render(mvp, verts, torch.tensor(rig.pos_idx), (H, W), vcol=torch.tensor(rig.vcol), tri_opp=opp, use_antialias=False)
            total = total + G.image_loss(ref[f, c], img) / C
    total.backward()
    return float(total.detach()), w.grad, t.grad, q.grad, ([m.grad for m in ms] if ms else None)


@pytest.mark.parametrize('mode,reg', [('free', False), ('combined', False), ('combined', True), ('prior', True)])
def test_free_and_combined_modes(tiny_rig, mode, reg):
    """the 'free' (blend_free) and 'combined' (blend_combined) optimisation modes and the optional L2 terms.
    One iteration's loss and gradients against the oracle from a state where every matrix is non-trivial, then the Adam
    schedule of the learned basis (locked until corrective_start, its own bias-correction clock afterwards)."""
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rig, H, W, F, Fn = tiny_rig, 128, 128, 2, 5
    ids = [3, 1]
    cfg = FitConfig(resolution=(H, W), shading='vcol', antialias=False, mode=mode, n_frames_total=Fn, lr_base=1e-2, lr_t=1e-3, lr_q=1e-4,
                    max_iter=8, regularize_correctives=reg and mode == 'combined', regularize_prior=reg and mode == 'prior')
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=2)
    ref = synthesize_reference(rig, w_true, t_true * 0.2, q_true, cfg)
    s = FitSession(rig, F, cfg, frame_ids=ids if mode != 'prior' else None)
    s.set_reference(ref)
    g = torch.Generator().manual_seed(5)
    s.set_parameters(w=0.2 * torch.rand(F, rig.B, generator=g))
    if mode != 'prior':
        assert s.basis_start == (0 if mode == 'free' else cfg.max_iter // 2 + 2) and not s.use_geom_fused
        s.m1.add_((0.1 * torch.randn(Fn, Fn, generator=g)).cuda())
        s.m2.add_((0.1 * torch.randn(Fn, Fn, generator=g)).cuda())
        s.m3.copy_((0.05 * torch.randn(3 * rig.V, Fn, generator=g)).cuda())
    s.forward()
    s.backward()
    torch.cuda.synchronize()
    opp = torch.tensor(G.topology_build(rig.pos_idx))
    loss_o, gw, gt, gq, gm = _oracle_mode_iteration(rig, s, mode, ids, ref.cpu(), H, W, opp, reg_corr=cfg.regularize_correctives,
                                                    reg_prior=cfg.regularize_prior)
    # the render chain is discontinuous in its inputs (test_iteration_gradients): whole-chain comparisons carry ~1 % noise from
    # single silhouette-pixel flips, the staged test above pins every link tightly; here the new links are what is checked
    assert abs(float(s.loss) - loss_o) / loss_o < 2e-3
    if mode != 'free':
        assert rel(s.d_w.cpu(), gw) < 3e-2
    else:
        assert float(s.d_w.abs().max()) == 0.0
    if mode != 'prior':
        # exact link check: the basis gradients are linear maps of the GPU's own d_verts
        dv = s.d_verts.cpu()
        x1 = s.m1.cpu()[:, ids].t()
        x2 = x1 @ s.m2.cpu().t()
        c = s.basis_coef
        dvc = c * dv + (2.0 * (x2 @ s.m3.cpu().t()) / (3 * rig.V) if cfg.regularize_correctives else 0.0)
        d_x2 = dvc @ s.m3.cpu()
        assert rel(s.d_m3.cpu(), dvc.t() @ x2) < 1e-5
        assert rel(s.d_m2.cpu(), d_x2.t() @ x1) < 1e-5
        d_m1 = torch.zeros(Fn, Fn)
        d_m1[:, ids] = (d_x2 @ s.m2.cpu()).t()
        assert rel(s.d_m1.cpu(), d_m1) < 1e-5
        for a, b in zip((s.d_m1, s.d_m2, s.d_m3), gm):
            assert rel(a.cpu(), b) < 3e-2
    if reg:
        vb = torch.tensor(rig.v_base)
        if mode == 'prior':
            term = sum(float(torch.mean(s.w[f].cpu() ** 2)) for f in range(F))
        else:
            term = sum(float(torch.mean((x2[f] @ s.m3.cpu().t()) ** 2)) for f in range(F))
        assert abs(float(s.reg_l2_term) - term) <= 1e-5 * term

    if mode == 'prior':
        return
    # Adam schedule of the basis group: torch.optim.Adam on the same gradient stream; the group is skipped while locked
    s2 = FitSession(rig, F, cfg, frame_ids=ids)
    s2.set_reference(ref)
    basis_o = s2.basis.cpu().clone().requires_grad_(True)
    opt = torch.optim.Adam([{'params': basis_o, 'lr': s2.basis_lr}])
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda x: cfg.lr_ramp ** (float(x) / float(cfg.max_iter)))
    for it in range(cfg.max_iter):
        s2.forward()
        s2.backward()
        grads = s2.basis_grads.cpu().clone()
        before = s2.basis.clone()
        s2.optimizer_step()
        if it >= s2.basis_start:
            opt.zero_grad()
            basis_o.grad = grads
            opt.step()
        else:
            assert torch.equal(before, s2.basis), 'the learned basis must stay locked before corrective_start'
        sched.step()
    torch.cuda.synchronize()
    assert float((s2.basis.cpu() - torch.cat([torch.eye(Fn).flatten(), torch.eye(Fn).flatten(), torch.zeros(3 * rig.V * Fn)])).abs().max()) > 0
    d = (s2.basis.cpu() - basis_o.detach()).abs().max()
    assert float(d) <= 1e-6 + 1e-3 * s2.basis_lr * cfg.max_iter, float(d)


def test_every_kernel_family_runs_on_a_ragged_resolution():
    """scripts/sanitize_target.py: three iterations in every configuration family (fused with / without antialias, texture and
    camera-correction gradients, op-level path, free / combined modes, L2 terms, mesh regularisers, tensor-core blend, band
    split, mip chain) at 72 x 104 — no CUDA error, finite parameters and losses."""
    import os
    import runpy
    runpy.run_path(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'scripts', 'sanitize_target.py'), run_name='__main__')


def test_camera_and_prior_blend_kernels_match_reference_kat():
    """The geometry kernels on the vectors the REFERENCE's own code produced: (a) fpc_pose_mvp_fwd + fpc_project_fwd on the real
    calibration against camera.py's P, MV, MVP and clip-space points (tests/golden/camera_kat.json); (b) fpc_blend_fwd / fpc_blend_bwd
    against fit.py's blend() and its autograd gradient of the intermediate mapping (tests/golden/blend_kat.json)."""
    import ctypes
    import json
    import os
    from fpc_diffrend_b200 import _lib, camera
    here = os.path.dirname(os.path.abspath(__file__))
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    with open(os.path.join(here, 'golden', 'camera_kat.json')) as f:
        kat = json.load(f)
    names = sorted(kat['cameras'])
    entries = [kat['cameras'][n] for n in names]
    Pc, Ac = camera.camera_constants(entries)
    C = len(names)
    dP, dA = torch.tensor(Pc).reshape(C, 16).cuda(), torch.tensor(Ac).reshape(C, 16).cuda()
    t0, q0 = torch.zeros(1, 3, device='cuda'), torch.tensor([[0., 0, 0, 1]], device='cuda')
    mvp = torch.empty(C, 16, device='cuda')
    _lib.call('fpc_pose_mvp_fwd', P(dP), P(dA), P(t0), P(q0), None, None, 1, C, P(mvp), st)
    mvp_ref = np.stack([np.asarray(e['MVP'], np.float32) for e in entries]).reshape(C, 16)
    assert rel(mvp.cpu().numpy(), mvp_ref) < 1e-6
    pts = torch.tensor(kat['points'], dtype=torch.float32).cuda().contiguous()          # [3,3]
    clip = torch.empty(C, pts.shape[0], 4, device='cuda')
    _lib.call('fpc_project_fwd', P(pts), P(mvp), 1, C, pts.shape[0], P(clip), st)
    clip_ref = np.stack([np.asarray(e['clip'], np.float32) for e in entries])
    assert np.abs(clip.cpu().numpy() - clip_ref).max() <= 1e-5 * np.abs(clip_ref).max()

    with open(os.path.join(here, 'golden', 'blend_kat.json')) as f:
        k = json.load(f)
    R, B, Fn = k['R'], k['B'], k['F']
    T = lambda name: torch.tensor(k[name], dtype=torch.float32)
    M1, M2 = T('M1'), T('M2')
    w = (M2 @ M1).t().contiguous()                                    # [Fn,B]: w_f = M2 M1 e_f for every frame
    dD, dbase, dw = T('D').cuda().contiguous(), T('v_base').cuda(), w.cuda()
    verts = torch.empty(Fn, R, device='cuda')
    _lib.call('fpc_blend_fwd', P(dD), P(dbase), P(dw), R, B, Fn, P(verts), st)
    d_verts = T('dy')[None].repeat(Fn, 1).cuda().contiguous()
    scratch = torch.empty(int(_lib.load().fpc_blend_bwd_scratch_bytes(R, B, Fn)), dtype=torch.uint8, device='cuda')
    d_w = torch.empty(Fn, B, device='cuda')
    _lib.call('fpc_blend_bwd', P(dD), P(d_verts), R, B, Fn, P(d_w), P(scratch), scratch.numel(), st)
    torch.cuda.synchronize()
    for f in range(Fn):
        rec = k['frames'][str(f)]['prior']
        ref = torch.tensor(rec['vtx_pos'])
        assert (verts[f].cpu() - ref).abs().max() <= 1e-5 * ref.abs().max()
        # d loss / d M2 = d_w (x) (M1 e_f): the reference's autograd gradient of the intermediate mapping
        g_m2 = torch.outer(d_w[f].cpu(), M1[:, f])
        assert rel(g_m2, torch.tensor(rec['grad']['M2'])) < 1e-5


@pytest.mark.parametrize('fused,use_aa', [(True, False), (True, True), (False, True)])
def test_l1_image_loss(small_rig3, fused, use_aa):
    """loss='l1' (north-star: L1/L2 loss): mean(|ref - 255 colour|) and its gradient (sign of the residual) through the fused
    kernels and the op-level image-loss kernel, against the oracle fed the GPU's pos_clip bits."""
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rig, H, W, F = small_rig3, 152, 200, 2
    C = rig.P.shape[0]
    cfg = FitConfig(resolution=(H, W), shading='texture', antialias=use_aa, fused=fused, loss='l1')
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=1)
    ref = synthesize_reference(rig, w_true, t_true * 0.2, q_true, cfg)
    s = FitSession(rig, F, cfg)
    s.set_reference(ref)
    s.set_parameters(w=(0.05 * np.random.default_rng(0).random((F, rig.B))).astype(np.float32))
    s.forward()
    s.backward()
    torch.cuda.synchronize()
    opp = torch.tensor(G.topology_build(rig.pos_idx))
    total, g_ref = 0.0, []
    for n in range(F * C):
        pc = s.pos_clip[n:n + 1].cpu().clone().requires_grad_(True)
        rast, _ = G.rasterize(pc, torch.tensor(rig.pos_idx), (H, W))
        col = G.texture(torch.tensor(rig.tex)[None], G.interpolate(torch.tensor(rig.uv)[None], rast, torch.tensor(rig.uv_idx)))
        if use_aa:
            col = G.antialias(col, rast, pc, torch.tensor(rig.pos_idx), opp)
        img = torch.where(rast[..., 3:] > 0, col, torch.tensor(G.BG))[0]
        loss = G.image_loss(ref.cpu()[n // C, n % C], img, 'l1') / C
        loss.backward()
        total += float(loss.detach())
        g_ref.append(pc.grad[0])
    assert abs(float(s.loss) - total) / total < 1e-5
    assert rel(s.g_pos.cpu(), torch.stack(g_ref)) < 1e-4
    # and a few Adam steps reduce it
    l0 = float(s.loss)
    s.cfg.lr_base = 5e-3
    for _ in range(15):
        s.iteration()
    assert float(s.loss) < l0


def test_fit_recovers_ground_truth_activations():
    """End-to-end sanity of the analysis-by-synthesis loop: starting from the neutral face, 400 graph-replayed iterations against
    frames rendered from known activations (3 cameras, vertex colours + antialias, the reference's decaying learning-rate
    schedule) bring the image loss down several-fold and move activations and blended vertices most of the way to the ground
    truth.  Two properties of rasterization-based gradients bound what to expect (tests/tools/convergence_probe.py): the loss is
    only piecewise smooth — single pixels popping across silhouettes are 0.06 each here, and Adam turns that noise into steps of
    the size of the learning rate, so the schedule has to decay for the loss to settle —, and the silhouette term of the gradient
    comes from the antialias op alone, which sees an outline pixel only when the triangle covering it owns the silhouette edge
    (rim triangles of a smooth mesh are sub-pixel slivers): along the line to the ground truth the analytic slope has the right
    sign and a quarter of the long-baseline finite-difference slope.  The fused and the op-level paths agree on it."""
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    H, W, F, iters = 192, 192, 2, 400
    rig = rigmod.make_rig(n_vertices=1500, n_shapes=12, n_cams=3, width=W, height=H, tex_size=128, seed=5)
    cfg = FitConfig(resolution=(H, W), shading='vcol', antialias=True, lr_base=2e-2, lr_ramp=0.005, max_iter=iters, optimize_pose=False)
    _, t_true, q_true = rigmod.make_targets(F, rig.B, seed=7)
    w_true = np.random.default_rng(7).uniform(0.3, 0.9, size=(F, rig.B)).astype(np.float32)       # every shape clearly active
    ref = synthesize_reference(rig, w_true, 0.0 * t_true, q_true * 0 + np.array([0, 0, 0, 1], np.float32), cfg)
    s = FitSession(rig, F, cfg)
    s.set_reference(ref)
    v_true = torch.tensor(rig.v_base)[None] + torch.tensor(w_true) @ torch.tensor(rig.D).t()
    err0 = float((s.result_vertices().cpu() - v_true).abs().mean())
    s.iteration()
    torch.cuda.synchronize()
    loss0 = float(s.loss)
    s.capture()
    for _ in range(iters - 1):
        s.replay()
    torch.cuda.synchronize()
    loss1 = float(s.loss)
    err1 = float((s.result_vertices().cpu() - v_true).abs().mean())
    werr0, werr1 = float(np.abs(w_true).mean()), float((s.w.cpu() - torch.tensor(w_true)).abs().mean())
    print('loss %.4f -> %.4f, vertex error %.4f -> %.4f, activation error %.3f -> %.3f' % (loss0, loss1, err0, err1, werr0, werr1))
    assert loss1 < 0.5 * loss0, (loss0, loss1)
    assert err1 < 0.4 * err0, (err0, err1)
    assert werr1 < 0.6 * werr0, (werr0, werr1)


@pytest.mark.parametrize('opt_tex', [False, True])
def test_fit_session_mip_path(small_rig3, opt_tex):
    """FitConfig.enable_mip (the mip branch of the reference's render(), fit.py:153-155, inside the accelerated loop): loss and
    gradients of the C-ABI chain against the SAME chain written with the drop-in ops under torch autograd (render.render with
    enable_mip=True), which the op-level tests pin against the float64 oracle."""
    from fpc_diffrend_b200 import rig as rigmod, render as R
    import fpc_diffrend_b200.ops as dr
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rig, H, W, F, maxl = small_rig3, 152, 200, 2, 3
    C = rig.P.shape[0]
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=1)
    ref = synthesize_reference(rig, w_true, t_true * 0.2, q_true, FitConfig(resolution=(H, W), shading='texture', antialias=True))
    cfg = FitConfig(resolution=(H, W), shading='texture', antialias=True, fused=False, enable_mip=True, max_mip_level=maxl, optimize_texture=opt_tex)
    s = FitSession(rig, F, cfg)
    s.set_reference(ref)
    s.set_parameters(w=(0.05 * np.random.default_rng(0).random((F, rig.B))).astype(np.float32))
    s.forward()
    s.backward()
    torch.cuda.synchronize()
    glctx = dr.RasterizeGLContext(device='cuda')
    tex = s.tex[0].clone().requires_grad_(True)
    uv, uvi, tri = s.attr[0], s.attr_idx, s.pos_idx
    total, g_pos = 0.0, []
    for n in range(F * C):
        pc = s.pos_clip[n].clone().requires_grad_(True)                      # [V,4]: feed the session's own clip positions
        rast, rast_db = dr.rasterize(glctx, pc[None], tri, resolution=(H, W))
        texc, texd = dr.interpolate(uv[None], rast, uvi, rast_db=rast_db, diff_attrs='all')
        col = dr.texture(tex[None], texc, texd, filter_mode='linear-mipmap-linear', max_mip_level=maxl)
        col = dr.antialias(col, rast, pc[None], tri)
        img = torch.where(rast[..., 3:] > 0, col, torch.tensor(R.BG, device='cuda'))[0]
        loss = torch.mean((ref[n // C, n % C] - 255.0 * img) ** 2) / C
        loss.backward()
        total += float(loss.detach())
        g_pos.append(pc.grad)
    assert abs(float(s.loss) - total) / total < 1e-5
    assert rel(s.g_pos.cpu(), torch.stack(g_pos).cpu()) < 1e-4
    if opt_tex:
        assert rel(s.d_tex.cpu(), tex.grad[None].cpu()) < 1e-4
    l0 = float(s.loss)
    s.cfg.lr_base = 5e-3
    for _ in range(10):
        s.iteration()
    assert float(s.loss) < l0


def test_batched_fit_with_regularisers_matches_op_level(small_rig3):
    """A frame batch large enough for the tensor-core blend (F = 9) WITH mesh regularisers: the fused configuration against the
    op-level configuration with separate fp32 kernels.  (tests/tools/session_sweep.py found d_w 5e-4 off here when the transpose
    D^T d_verts ran on the 3xTF32 tensor-core kernel: the Laplacian gradient makes that contraction ill-conditioned; the fp32
    kernel is used for it under regularisers.)"""
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rig, H, W, F = small_rig3, 152, 200, 9
    base = dict(resolution=(H, W), shading='vcol', antialias=False, loss='l1', weight_laplacian=50.0, weight_meshedge=1.0)
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=3)
    ref = synthesize_reference(rig, w_true, 0.3 * t_true, q_true, FitConfig(**base))
    rng = np.random.default_rng(5)
    w0 = (0.3 * rng.random((F, rig.B))).astype(np.float32)
    got = []
    for kw in (dict(fused=True), dict(fused=False, fused_geometry=False, tc_blend=False)):
        s = FitSession(rig, F, FitConfig(**base, **kw))
        s.set_reference(ref)
        s.set_parameters(w=w0)
        s.forward(); s.backward()
        torch.cuda.synchronize()
        got.append((float(s.loss), s.grads.clone(), s.use_tc_blend))
    assert got[0][2] and not got[1][2]                          # the forward blend of the first session does run on tcgen05
    assert abs(got[0][0] - got[1][0]) / got[1][0] < 1e-5
    assert rel(got[0][1].cpu(), got[1][1].cpu()) < 1e-4


@pytest.mark.parametrize('tool,n,seed', [('parity_sweep.py', 16, 4), ('session_sweep.py', 16, 4)])
def test_randomised_sweeps(tool, n, seed):
    """tests/tools/parity_sweep.py (rasterizer vs oracle, fused kernels vs op-level chain, op-level chain vs oracle chain over random
    rigs, poses — partly off screen —, ragged resolutions, shading modes) and tests/tools/session_sweep.py (fully fused FitSession
    vs op-level FitSession over random batch sizes, losses, regularisers, shared parameters): a seeded slice of each."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'tests', 'tools', tool), str(n), str(seed)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and ('%d / %d cases clean' % (n, n)) in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_two_devices_in_one_process(tiny_rig):
    """One process driving two GPUs: the kernels that need more than 48 KB of dynamic shared memory (k_fused_aa, the geometry
    backward, the tcgen05 GEMM, k_fill) are configured per DEVICE — the second device must work like the first."""
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rig, H, W, F = tiny_rig, 128, 128, 9
    cfg = FitConfig(resolution=(H, W), shading='texture', antialias=True, lr_base=1e-2)
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=2)
    losses = []
    for dev in (0, 1):
        with torch.cuda.device(dev):
            ref = synthesize_reference(rig, w_true, 0.2 * t_true, q_true, cfg, device='cuda:%d' % dev)
            s = FitSession(rig, F, cfg, device='cuda:%d' % dev)
            s.set_reference(ref)
            for _ in range(3):
                s.iteration()
            torch.cuda.synchronize(dev)
            losses.append(float(s.loss))
            assert s.use_tc_blend
    assert abs(losses[0] - losses[1]) <= 1e-3 * abs(losses[0])


def test_quaternion_renorm_frobenius_quirk(tiny_rig):
    """fit.py:616-618 divides the WHOLE [n,4] quaternion tensor by its Frobenius norm (`q /= torch.sum(q ** 2) ** 0.5`, SURVEY
    App. B) — with F frames every row ends up with norm 1/sqrt(F), and roma's un-normalised quat -> R then SCALES the rig by
    1/F.  FitConfig(quat_norm='frobenius') reproduces exactly that; checked (a) on the kernel alone, (b) on the fused Adam step
    against torch.optim.Adam followed by the reference's two lines, over several iterations."""
    import ctypes
    from fpc_diffrend_b200 import _lib
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    # (a) fpc_quat_renorm, mode 1 = whole tensor, mode 0 = per row
    g = torch.Generator().manual_seed(0)
    q0 = torch.randn(5, 4, generator=g)
    for mode in (1, 0):
        qd = q0.cuda().contiguous()
        _lib.call('fpc_quat_renorm', ctypes.c_void_p(qd.data_ptr()), 5, mode, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        want = q0 / torch.sum(q0 ** 2) ** 0.5 if mode == 1 else q0 / q0.norm(dim=1, keepdim=True)
        np.testing.assert_allclose(qd.cpu().numpy(), want.numpy(), rtol=2e-6, atol=1e-7)
    # (b) the packed Adam step of a 3-frame session: same gradients fed to torch.optim.Adam + the reference's renorm lines
    rig, F, iters = tiny_rig, 3, 4
    cfg = FitConfig(resolution=(128, 128), shading='vcol', antialias=False, lr_base=1e-2, lr_t=1e-3, lr_q=1e-3, max_iter=50,
                    quat_norm='frobenius')
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=2)
    ref = synthesize_reference(rig, w_true, 0.2 * t_true, q_true, replace_cfg(cfg, quat_norm='row'))
    s = FitSession(rig, F, cfg)
    s.set_reference(ref)
    w = torch.zeros(F, rig.B, requires_grad=True)
    t = torch.zeros(F, 3, requires_grad=True)
    q = torch.tensor([[0., 0, 0, 1]] * F, requires_grad=True)
    opt = torch.optim.Adam([{'params': w, 'lr': cfg.lr_base}, {'params': t, 'lr': cfg.lr_t}, {'params': q, 'lr': cfg.lr_q}])
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda x: cfg.lr_ramp ** (float(x) / float(cfg.max_iter)))
    for it in range(iters):
        s.set_parameters(w.detach(), t.detach(), q.detach())          # identical state on both sides before every step
        s.iteration()
        torch.cuda.synchronize()
        opt.zero_grad()
        w.grad, t.grad, q.grad = s.d_w.cpu().clone(), s.d_t.cpu().clone(), s.d_q.cpu().clone()
        opt.step()
        sched.step()
        with torch.no_grad():
            q /= torch.sum(q ** 2) ** 0.5                              # fit.py:616-618, verbatim semantics
        np.testing.assert_allclose(s.q.cpu().numpy(), q.detach().numpy(), rtol=1e-5, atol=1e-7, err_msg='iteration %d' % it)
        np.testing.assert_allclose(s.w.cpu().numpy(), w.detach().numpy(), rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(s.t.cpu().numpy(), t.detach().numpy(), rtol=1e-4, atol=1e-8)
    # the quirk itself: the whole tensor has unit Frobenius norm, so every row ends up near 1/sqrt(F), not 1
    np.testing.assert_allclose(s.q.norm(dim=1).cpu().numpy(), np.full(F, 1.0 / np.sqrt(F)), rtol=2e-2)
    assert abs(float(torch.sum(s.q ** 2)) - 1.0) < 1e-5


def replace_cfg(cfg, **kw):
    from dataclasses import replace
    return replace(cfg, **kw)


@pytest.mark.parametrize('band', [False, True])
def test_camera_split_with_regularisers_adds_up(small_rig3, band):
    """Camera split + terms that do not depend on the views (mesh regularisers, regularize_prior): the ranks' gradients are
    SUMMED, so exactly one rank may evaluate them.  The partial losses / gradients of 2 and 3 ranks (rendered one after the
    other on one GPU) must add up to the unsplit session — with the regularisers counted once."""
    from fpc_diffrend_b200 import rig as rigmod, shard
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    rig, H, W, F = small_rig3, 152, 200, 2
    C = rig.P.shape[0]
    base = dict(resolution=(H, W), shading='texture', antialias=True, weight_laplacian=5000.0, weight_meshedge=70.0, regularize_prior=True)
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=1)
    ref = synthesize_reference(rig, w_true, 0.2 * t_true, q_true, FitConfig(**base))
    rng = np.random.default_rng(0)
    w0 = (0.05 * rng.random((F, rig.B))).astype(np.float32)
    full = FitSession(rig, F, FitConfig(**base))
    full.set_reference(ref)
    full.set_parameters(w=w0)
    full.forward(); full.backward()
    torch.cuda.synchronize()
    plain = FitSession(rig, F, FitConfig(**dict(base, weight_laplacian=0.0, weight_meshedge=0.0, regularize_prior=False)))
    plain.set_reference(ref)
    plain.set_parameters(w=w0)
    plain.forward(); plain.backward()
    torch.cuda.synchronize()
    reg_effect = rel(plain.grads.cpu(), full.grads.cpu())
    assert reg_effect > 1e-4, reg_effect                            # the regularisers matter in this set-up (1e-5 is the test's tolerance)
    for world in (2, 3):
        loss, grads, owners = 0.0, torch.zeros_like(full.grads), 0
        for r in range(world):
            if band:
                sl, bd = shard.view_band_shard(C, H, r, world)
            else:
                sl, bd = shard.camera_shard(C, r, world), None
            s = FitSession(rig, F, FitConfig(cam_slice=sl, cam_band=bd, **base))
            owners += int(s.reg_owner)
            s.set_reference(ref[:, sl[0]:sl[1]])
            s.set_parameters(w=w0)
            s.forward(); s.backward()
            torch.cuda.synchronize()
            loss += float(s.loss)
            grads += s.grads
        assert owners == 1
        assert abs(loss - float(full.loss)) / float(full.loss) < 1e-5, (world, loss, float(full.loss))
        assert rel(grads.cpu(), full.grads.cpu()) < 1e-5, (world, rel(grads.cpu(), full.grads.cpu()))
    with pytest.raises(ValueError):
        FitSession(rig, F, FitConfig(cam_slice=(0, 2), cam_band=(0, 3), optimize_cam_pose=True, **base))


def test_fitted_activations_config2_size():
    """North-star: "fitted activations after a fixed iteration count" at the size the metric is quoted on (BASELINE config 2:
    20k vertices / 40k triangles, 200 blendshapes, 9 cameras 1024 x 1024, vertex colours, the reference's learning rates,
    main.py:14-18) against the CPU oracle driven by torch.optim.Adam + LambdaLR exactly as fit.py:493-505,610-618.

    What CAN be pinned to the north-star's tolerances at this size, and is asserted at every iteration:
      * the free-running LOSS trajectories of the GPU fit and of the oracle fit agree to 1e-5 relative (measured 1e-7);
      * on the GPU's own pos_clip bits: loss 1e-5; d loss / d pos_clip within 1e-4 of the largest gradient on >= 99.95 % of
        the vertices; the fused kernels and the op-level kernels (two independent fp32 formulations) agree to 1e-5;
      * Adam + LambdaLR + renorm fed the GPU's gradients reproduce the GPU's parameters to 1e-5.
    What cannot, for ANY pair of fp32 implementations (tests/tools/fit_traj_probe.py, debug_sliver_grad.py):
    the largest position gradient of the whole mesh sits on a sliver triangle seen edge-on at the silhouette, where
    1 / (a0 + a1 + a2) cancels five digits — the float64-accumulating oracle changes its own d pos by 2e-3 .. 2e-2 of the
    maximum when pos_clip moves by ONE ulp, and fp32 evaluation is 8e-3 off on that one vertex (fused and op-level kernels
    alike, 9e-7 from each other).  Through D^T such a spike moves d_w by up to 100 % between two runs that differ in the last
    bit of pos_clip, so free-running activations drift apart by a few per cent of their range (Adam's eps raised above the
    gradient noise floor or not): that figure is bounded loosely below and reported, not presented as parity."""
    from fpc_diffrend_b200 import rig as rigmod
    from fpc_diffrend_b200.fit import FitConfig, FitSession, synthesize_reference
    H = W = 1024
    F, iters = 1, 3
    rig = rigmod.make_rig(n_vertices=20000, n_shapes=200, n_cams=9, width=W, height=H, tex_size=64, seed=0)
    cfg = FitConfig(resolution=(H, W), shading='vcol', antialias=False)
    w_true, t_true, q_true = rigmod.make_targets(F, rig.B, seed=1)
    ref = synthesize_reference(rig, w_true, t_true, q_true, cfg)
    ref_cpu = ref.cpu()
    tri = torch.tensor(rig.pos_idx)
    base, D, vcol = torch.tensor(rig.v_base), torch.tensor(rig.D), torch.tensor(rig.vcol)
    Ps, As = torch.tensor(rig.P), torch.tensor(rig.A)

    def oracle_render_loss(pcs):
        rast, _ = G.rasterize(pcs, tri, (H, W))                 # all views in one call (OpenMP over views)
        col = G.interpolate(vcol[None], rast, tri)
        img = torch.where(rast[..., 3:] > 0, col, torch.tensor(G.BG))
        return sum(G.image_loss(ref_cpu[0, c], img[c]) for c in range(9)) / 9

    def oracle_loss(w, t, q):
        verts = G.blend(base, D, w[0]).reshape(-1, 3)
        return oracle_render_loss(torch.cat([G.transform_clip(G.mvp_chain(Ps[c], As[c], t[0], q[0]), verts) for c in range(9)]))

    def adam(params):
        opt = torch.optim.Adam([{'params': params[0], 'lr': cfg.lr_base}, {'params': params[1], 'lr': cfg.lr_t}, {'params': params[2], 'lr': cfg.lr_q}])
        return opt, torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda x: cfg.lr_ramp ** (float(x) / float(cfg.max_iter)))

    s = FitSession(rig, F, cfg)                                 # fused kernels (the product path)
    s.set_reference(ref)
    o = FitSession(rig, F, replace_cfg(cfg, fused=False))       # op-level kernels on the same parameters
    o.set_reference(ref)
    free = [torch.zeros(F, rig.B, requires_grad=True), torch.zeros(F, 3, requires_grad=True), torch.tensor([[0., 0, 0, 1]] * F, requires_grad=True)]
    fed = [p.detach().clone().requires_grad_(True) for p in free]
    opt_free, sched_free = adam(free)
    opt_fed, sched_fed = adam(fed)
    for it in range(iters):
        o.params.copy_(s.params)
        o.forward(); o.backward()
        s.iteration()
        torch.cuda.synchronize()
        # (1) free-running oracle fit: same loss trajectory
        lf = oracle_loss(*free)
        assert abs(float(s.loss) - float(lf.detach())) <= 1e-5 * float(lf.detach()), (it, float(s.loss), float(lf.detach()))
        opt_free.zero_grad(); lf.backward(); opt_free.step(); sched_free.step()
        with torch.no_grad():
            free[2] /= free[2].norm(dim=1, keepdim=True)
        # (2) the oracle on the GPU's own pos_clip bits
        pc = s.pos_clip.cpu().clone().requires_grad_(True)
        ls = oracle_render_loss(pc)
        ls.backward()
        assert abs(float(s.loss) - float(ls.detach())) <= 1e-5 * float(ls.detach())
        go, gf, gp = pc.grad.numpy(), s.g_pos.cpu().numpy(), o.g_pos.cpu().numpy()
        mx = np.abs(go).max()
        err = np.abs(gf - go).max(axis=-1) / mx
        assert (err <= 1e-4).mean() >= 0.9995, (it, float((err <= 1e-4).mean()))
        assert err.max() <= 5e-2, (it, float(err.max()))               # the sliver corners (see the docstring); measured 8e-3
        assert np.abs(gf - gp).max() <= 1e-5 * mx, (it, float(np.abs(gf - gp).max() / mx))
        # (3) the optimiser fed the GPU's gradients
        opt_fed.zero_grad()
        fed[0].grad, fed[1].grad, fed[2].grad = s.d_w.cpu().clone(), s.d_t.cpu().clone(), s.d_q.cpu().clone()
        opt_fed.step(); sched_fed.step()
        with torch.no_grad():
            fed[2] /= fed[2].norm(dim=1, keepdim=True)
        assert rel(s.w.cpu().numpy(), fed[0].detach().numpy()) <= 1e-5
        assert rel(s.t.cpu().numpy(), fed[1].detach().numpy()) <= 1e-5
        assert np.abs(s.q.cpu().numpy() - fed[2].detach().numpy()).max() <= 1e-6
    wo = free[0].detach().numpy()
    assert np.abs(wo).max() > 0.5 * cfg.lr_base * iters             # the fit moved
    drift = rel(s.w.cpu().numpy(), wo)
    assert drift <= 2.0, drift      # Adam's steps are sign-like: two fits that disagree on a noise-level gradient component differ
                                    # by up to 2 lr per step there (see the docstring); the loss trajectories above are what agrees
