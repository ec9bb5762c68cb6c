/*
 * fpc_b200.h — C-ABI of the B200-native fit hot path (libfpc_b200.so).
 *
 * This is the drop-in boundary for the per-iteration analysis-by-synthesis path of fpc-diffrend
 * (reference: /root/reference/src/torch/fit.py:524-642).  Each entry point states the reference
 * interface it replaces.  For the four rendering ops that interface is the nvdiffrast plugin the
 * reference binds through `import nvdiffrast.torch as dr` (fit.py:13): upstream plugin functions
 * `rasterize_fwd_cuda / rasterize_grad / rasterize_grad_db / interpolate_fwd / interpolate_fwd_da /
 * interpolate_grad / interpolate_grad_da / texture_construct_mip / texture_fwd / texture_fwd_mip /
 * texture_grad_linear / texture_grad_linear_mipmap_{nearest,linear} / antialias_construct_topology_hash /
 * antialias_fwd / antialias_grad`.  The binding a maintainer adds on the reference side is
 * shown in INTEGRATION.md.
 *
 * Groups, in file order: status / device check; the four rendering ops of the drop-in (+ their mip-mapped variants);
 * blendshape combination (SIMT, tcgen05, learned basis of the free / combined modes); pose -> MVP and clip transform;
 * fused geometry stages; image loss (L2 or L1); fused render + loss + gradient (plain, antialiased, band-split for the
 * camera-split mode); mesh regularisers; Adam.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers on the current CUDA device unless
 *     a parameter says "host"; float = IEEE fp32, indices = int32; tensors are dense, row-major,
 *     contiguous with the shapes given in brackets;
 *   - every function returns 0 (FPC_OK) or an fpc_status; it never throws and never synchronises the
 *     device; work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *     there is no host read-back on any path, so a whole fit iteration can be captured in a CUDA graph;
 *   - `fpc_last_error()` returns a thread-local, human-readable description of the last failure;
 *   - scratch memory is caller-owned: ask `*_scratch_bytes()` and pass a device buffer of at least that size.
 *     The same scratch buffer must not be used by two concurrently running streams;
 *   - outputs documented as "overwritten" need no initialisation by the caller.
 *   - image row 0 is the BOTTOM row (NDC y = -1), as in nvdiffrast; rast = (u, v, z/w, float(tri_id+1)),
 *     all four zero on background pixels.
 */
#ifndef FPC_B200_H
#define FPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum fpc_status {
    FPC_OK = 0,
    FPC_ERR_INVALID_ARGUMENT = 1,
    FPC_ERR_CUDA = 2,
    FPC_ERR_UNSUPPORTED = 3
} fpc_status;

typedef void* fpc_stream_t; /* cudaStream_t */

#define FPC_B200_ABI_VERSION 4   /* 2: loss_kind / grad_tex arguments of the loss and fused entries, band-split entry;
                                    3: vadj_off / vadj_item arguments of the fused entries (atomics-free gradient gather);
                                    4: counters / adam arguments of fpc_geometry_bwd (reduction, pose backward and Adam in its tail) */

/* ---- library ------------------------------------------------------------------------------------- */
int fpc_abi_version(void);
const char* fpc_last_error(void);
/* Fails (FPC_ERR_UNSUPPORTED) unless the current device is compute capability 10.x: there is no fallback path. */
int fpc_check_device(void);

/* ---- rasterize   (replaces dr.rasterize, fit.py:151; plugin rasterize_fwd_cuda / rasterize_grad) --- */
/* Bytes of scratch needed by fpc_rasterize_fwd for N instances of T triangles at HxW. */
size_t fpc_rasterize_scratch_bytes(int N, int T, int H, int W);
/* pos [N,V,4] clip space, tri [T,3]  ->  rast [N,H,W,4], rast_db [N,H,W,4] (may be NULL).  Overwritten. */
int fpc_rasterize_fwd(const float* pos, const int32_t* tri, int N, int V, int T, int H, int W,
                      float* rast, float* rast_db, void* scratch, size_t scratch_bytes, fpc_stream_t stream);
/* dy [N,H,W,4] (only d u, d v are used; d(z/w), d(id) ignored as upstream) -> grad_pos [N,V,4], overwritten
 * (x, y, w components; z = 0).  Gradients w.r.t. rast_db are not supported (mip path). */
int fpc_rasterize_bwd(const float* pos, const int32_t* tri, const float* rast, const float* dy,
                      int N, int V, int T, int H, int W, float* grad_pos, fpc_stream_t stream);

/* ---- interpolate (replaces dr.interpolate, fit.py:157; plugin interpolate_fwd / interpolate_grad) -- */
/* attr [Na,Vt,A] with Na in {1,N} (1 = broadcast over instances), rast [N,H,W,4], tri [T,3] -> out [N,H,W,A] */
int fpc_interpolate_fwd(const float* attr, int Na, int Vt, int A, const float* rast, const int32_t* tri,
                        int N, int T, int H, int W, float* out, fpc_stream_t stream);
/* dy [N,H,W,A] -> grad_attr [Na,Vt,A] (overwritten; summed over instances when Na == 1),
 *                 grad_rast [N,H,W,4] = (d u, d v, 0, 0) (overwritten) */
int fpc_interpolate_bwd(const float* attr, int Na, int Vt, int A, const float* rast, const int32_t* tri,
                        const float* dy, int N, int T, int H, int W, float* grad_attr, float* grad_rast,
                        fpc_stream_t stream);

/* ---- texture, filter_mode='linear', boundary_mode='wrap'
 *      (replaces dr.texture(..., filter_mode='linear'), fit.py:158; plugin texture_fwd / texture_grad_linear) */
/* tex [Nt,Ht,Wt,C] with Nt in {1,N}, uv [N,H,W,2] -> out [N,H,W,C] */
int fpc_texture_linear_fwd(const float* tex, int Nt, int Ht, int Wt, int C, const float* uv,
                           int N, int H, int W, float* out, fpc_stream_t stream);
/* dy [N,H,W,C] -> grad_tex [Nt,Ht,Wt,C] (overwritten; may be NULL to skip it), grad_uv [N,H,W,2] (overwritten) */
int fpc_texture_linear_bwd(const float* tex, int Nt, int Ht, int Wt, int C, const float* uv, const float* dy,
                           int N, int H, int W, float* grad_tex, float* grad_uv, fpc_stream_t stream);

/* ---- antialias   (replaces dr.antialias, fit.py:160; plugin antialias_construct_topology_hash /
 *      antialias_fwd / antialias_grad).  The per-call edge hash of upstream is replaced by a per-topology
 *      adjacency table tri_opp [T,3]: opposite vertex across edge e (e0=(v1,v2), e1=(v2,v0), e2=(v0,v1)), -1 = none. */
size_t fpc_topology_scratch_bytes(int T);
int fpc_topology_build(const int32_t* tri, int T, int V, int32_t* tri_opp, void* scratch, size_t scratch_bytes,
                       fpc_stream_t stream);
/* color [N,H,W,C], rast [N,H,W,4], pos [N,V,4], tri [T,3], tri_opp [T,3] -> out [N,H,W,C] */
int fpc_antialias_fwd(const float* color, const float* rast, const float* pos, const int32_t* tri,
                      const int32_t* tri_opp, int N, int V, int T, int H, int W, int C, float* out,
                      fpc_stream_t stream);
/* dy [N,H,W,C] -> grad_color [N,H,W,C], grad_pos [N,V,4] (both overwritten) */
int fpc_antialias_bwd(const float* color, const float* rast, const float* pos, const int32_t* tri,
                      const int32_t* tri_opp, const float* dy, int N, int V, int T, int H, int W, int C,
                      float* grad_color, float* grad_pos, fpc_stream_t stream);

/* ---- blendshape combination (replaces fit.blend, fit.py:103-129, north-star form V = base + D w) ---- */
/* D [R,B] (R = 3V rows, xyz interleaved), base [R], w [F,B]  ->  verts [F,R] */
int fpc_blend_fwd(const float* D, const float* base, const float* w, int R, int B, int F, float* verts,
                  fpc_stream_t stream);
/* General form: verts[f,r] = init + coef * sum_b D[r,b] w[f,b] with init = verts[f,r] (accumulate != 0), base[r], or 0
 * (base == NULL).  Used for the learned basis of the free / combined modes (D := m3, w := x2, below). */
int fpc_blend_fwd_ex(const float* D, const float* base, const float* w, int R, int B, int F, float coef, int accumulate,
                     float* verts, fpc_stream_t stream);
/* transpose gradient: d_verts [F,R] -> d_w [F,B] (overwritten) = D^T d_verts.  Deterministic two-stage reduction. */
size_t fpc_blend_bwd_scratch_bytes(int R, int B, int F);
int fpc_blend_bwd(const float* D, const float* d_verts, int R, int B, int F, float* d_w,
                  void* scratch, size_t scratch_bytes, fpc_stream_t stream);

/* ---- learned vertex basis of the "free" and "combined" modes (replaces blend_free fit.py:47-62 and the learned half of
 *      blend_combined fit.py:66-99 with setup_dataset_free fit.py:166-179) -----------------------
 * m1, m2 [Fn,Fn], m3 [R,Fn] are shared by the Fn frames of the take; frame_ids [Fb] = take-wide ids of the batch's frames
 * (distinct).  x1[b,:] = m1[:, id_b] (m1 e_f), x2[b,:] = m2 x1[b,:];  the vertices follow from
 * fpc_blend_fwd_ex(m3, base-or-accumulate, x2, R, Fn, Fb, coef, ...).  Backward: d_x2 [Fb,Fn] = fpc_blend_bwd(m3, d_verts),
 *   fpc_basis_grad:      d_m3 [R,Fn]  = coef * d_verts^T x2                         (overwritten)
 *   fpc_basis_code_bwd:  d_m2 [Fn,Fn] = coef * d_x2^T x1,  d_m1[:, id_b] = coef * m2^T d_x2[b,:], 0 elsewhere (overwritten)
 * Fixed summation order, no atomics. */
int fpc_basis_code_fwd(const float* m1, const float* m2, const int32_t* frame_ids, int Fn, int Fb, float* x1, float* x2,
                       fpc_stream_t stream);
int fpc_basis_grad(const float* d_verts, const float* x2, int R, int Fn, int Fb, float coef, float* d_m3, fpc_stream_t stream);
int fpc_basis_code_bwd(const float* m2, const float* x1, const float* d_x2, const int32_t* frame_ids, int Fn, int Fb, float coef,
                       float* d_m1, float* d_m2, fpc_stream_t stream);
/* The optional L2 terms of the loop: regularize_prior, loss += mean(activations^2) (fit.py:591-595), and
 * regularize_correctives, loss += mean(deformations^2) (fit.py:584-589).  x [F,n]:
 *   term [1] (nullable) = weight * sum_f mean_n x^2;  loss_accum [1] (nullable) += term;
 *   g_out [F,n] (nullable; may alias g_in) = g_scale * g_in (0 if g_in == NULL) + (2 weight / n) x. */
size_t fpc_l2_reg_scratch_bytes(long long total);
int fpc_l2_reg_fwd_bwd(const float* x, int F, long long n, float weight, float* loss_accum, float* term,
                       const float* g_in, float g_scale, float* g_out, void* scratch, size_t scratch_bytes, fpc_stream_t stream);

/* Tensor-core path for frame batches (north-star item 1): the same mathematics as fpc_blend_fwd / fpc_blend_bwd as one
 * TMA + tcgen05 (kind::tf32, 3xTF32 hi/lo split: fp32-level accuracy) GEMM kernel with the accumulator in TMEM.
 * fpc_blend_tc_supported() != 0 requires B % 4 == 0 and R % 4 == 0 (16-byte row pitches for TMA).
 * The backward takes DT [B,R] = D^T (a transposed copy made once at set-up) so that both operands are K-major;
 * it is a split-K GEMM over one wave of CTAs with a fixed-order second-stage sum (deterministic). */
int fpc_blend_tc_supported(int R, int B, int F);
int fpc_blend_fwd_tc(const float* D, const float* base, const float* w, int R, int B, int F, float* verts, fpc_stream_t stream);
size_t fpc_blend_bwd_tc_scratch_bytes(int R, int B, int F);
int fpc_blend_bwd_tc(const float* DT, const float* d_verts, int R, int B, int F, float* d_w,
                     void* scratch, size_t scratch_bytes, fpc_stream_t stream);

/* ---- pose + projection (replaces the MVP chain fit.py:546-553 with camera.rigid_grad camera.py:128-132 and
 *      roma.unitquat_to_rotmat, and camera.transform_clip camera.py:11-23) --------------------------------- */
/* P [C,16], A [C,16] (= MV @ translate(0,170,0)), row-major 4x4; t [F,3], q [F,4] XYZW (not normalised);
 * t_cam [C,3] / q_cam [C,4] per-camera corrections (fit.py:443-448) or NULL for identity
 * -> mvp [F*C,16], instance n = f*C + c:   mvp = P_c @ Rigid(t_f,q_f) @ Rigid(t_cam_c,q_cam_c) @ A_c */
int fpc_pose_mvp_fwd(const float* P, const float* A, const float* t, const float* q,
                     const float* t_cam, const float* q_cam, int F, int C, float* mvp, fpc_stream_t stream);
/* d_mvp [F*C,16] -> d_t [F,3], d_q [F,4] (overwritten) */
int fpc_pose_mvp_bwd(const float* P, const float* A, const float* t, const float* q,
                     const float* t_cam, const float* q_cam, const float* d_mvp, int F, int C,
                     float* d_t, float* d_q, fpc_stream_t stream);
/* Gradient of the per-camera pose corrections (t_opt / q_opt of fit.py:443-448, optimiser groups fit.py:498-499):
 * d_mvp [F*C,16] -> d_t_cam [C,3], d_q_cam [C,4] (overwritten), summed over the F frames in index order. */
int fpc_pose_cam_bwd(const float* P, const float* A, const float* t, const float* q, const float* t_cam, const float* q_cam,
                     const float* d_mvp, int F, int C, float* d_t_cam, float* d_q_cam, fpc_stream_t stream);
/* verts [F,V,3], mvp [F*C,16] -> pos_clip [F*C,V,4] */
int fpc_project_fwd(const float* verts, const float* mvp, int F, int C, int V, float* pos_clip, fpc_stream_t stream);
/* d_pos_clip [F*C,V,4] -> d_verts [F,V,3] (overwritten), d_mvp [F*C,16] (overwritten) */
size_t fpc_project_bwd_scratch_bytes(int F, int C, int V);
int fpc_project_bwd(const float* verts, const float* mvp, const float* d_pos_clip, int F, int C, int V,
                    float* d_verts, float* d_mvp, void* scratch, size_t scratch_bytes, fpc_stream_t stream);

/* ---- fused geometry stages (small frame batches): one pass over D per direction -----------------------------
 * Same mathematics as fpc_pose_mvp_fwd + fpc_blend_fwd + fpc_project_fwd (forward) and fpc_project_bwd +
 * fpc_blend_bwd + fpc_pose_mvp_bwd (backward), i.e. reference fit.py:546-553, fit.py:103-129, camera.py:11-23 and
 * their part of loss.backward() (fit.py:611), fused so that D [3V,B] is streamed exactly once per direction and
 * every frame's C cameras are handled by the lanes of the warp that owns the vertex.  Deterministic.
 * Requirements (fpc_geometry_fused_supported() != 0): B % 4 == 0, B <= 1024, C <= 32, F <= 65535.  D is re-read for
 * every frame, so for large frame batches the GEMM path (fpc_blend_fwd / fpc_blend_bwd) is the better choice. */
int fpc_geometry_fused_supported(int V, int B, int F, int C);
/* -> mvp [F*C,16], verts [F,V,3], pos_clip [F*C,V,4] (all overwritten) */
int fpc_geometry_fwd(const float* P, const float* A, const float* t, const float* q, const float* t_cam, const float* q_cam,
                     const float* D, const float* base, const float* w, int V, int B, int F, int C,
                     float* mvp, float* verts, float* pos_clip, fpc_stream_t stream);
/* The optimiser step of fpc_adam_fused as a rider of fpc_geometry_bwd (same meaning member by member). */
typedef struct fpc_adam_fused_args {
    float* params; float* m; float* v; float* step_count;
    int optimize_pose, quat_mode;
    float lr_w, lr_t, lr_q, b1, b2, eps, lr_ramp, max_iter;
} fpc_adam_fused_args;
/* g_pos [F*C,V,4] = d loss / d pos_clip; d_verts_add [F,V,3] or NULL: extra gradient on the blended vertices
 * (mesh regularisers, fit.py:580-582) added before the D^T contraction.
 * -> d_w [F,B], d_t [F,3], d_q [F,4] (overwritten); optional outputs d_verts [F,V,3], d_mvp [F*C,16] (NULL to skip).
 * ONE launch: the CTAs that arrive last combine the per-CTA partials in a fixed order (bit-reproducible), run the pose
 * backward and — when `adam` is non-NULL — the optimiser step on the packed parameters (fit.py:610-618); that needs
 * d_w, d_t, d_q to be the packed vector [d_w | d_t | d_q].  counters: fpc_geometry_bwd_counter_bytes() bytes of device
 * memory owned by the caller, ZERO before the first call; every call leaves them zero (calls sharing a counter buffer must not
 * overlap in time). */
size_t fpc_geometry_bwd_scratch_bytes(int V, int B, int F, int C);
size_t fpc_geometry_bwd_counter_bytes(int V, int F);
int fpc_geometry_bwd(const float* P, const float* A, const float* t, const float* q, const float* t_cam, const float* q_cam,
                     const float* D, const float* verts, const float* mvp, const float* g_pos, const float* d_verts_add,
                     int V, int B, int F, int C, float* d_w, float* d_t, float* d_q, float* d_verts, float* d_mvp,
                     int32_t* counters, const fpc_adam_fused_args* adam,
                     void* scratch, size_t scratch_bytes, fpc_stream_t stream);

/* ---- mip-mapped texturing path (reference fit.py:153-155 with enable_mip) -------------------------
 * Replaces, in nvdiffrast plugin terms: rasterize_grad_db, interpolate_fwd_da / interpolate_grad_da,
 * texture_construct_mip, texture_fwd_mip, texture_grad_linear_mipmap_{nearest,linear}. */
/* rasterize backward with the gradient of rast_db: d_rast [N,H,W,4] (u, v used), d_rast_db [N,H,W,4] -> grad_pos (overwritten) */
int fpc_rasterize_bwd_db(const float* pos, const int32_t* tri, const float* rast, const float* d_rast, const float* d_rast_db,
                         int N, int V, int T, int H, int W, float* grad_pos, fpc_stream_t stream);
/* interpolate with attribute pixel differentials: diff_attrs = HOST array of K attribute indices (NULL = all A attributes,
 * K ignored; at most 32); out [N,H,W,A], out_da [N,H,W,2K] = (d a_j / dX, d a_j / dY) per selected attribute. */
int fpc_interpolate_da_fwd(const float* attr, int Na, int Vt, int A, const float* rast, const float* rast_db, const int32_t* tri,
                           const int32_t* diff_attrs, int K, int N, int T, int H, int W, float* out, float* out_da,
                           fpc_stream_t stream);
/* dy [N,H,W,A] and/or dda [N,H,W,2K] (either may be NULL) -> grad_attr, grad_rast (overwritten), grad_rast_db (nullable) */
int fpc_interpolate_da_bwd(const float* attr, int Na, int Vt, int A, const float* rast, const float* rast_db, const int32_t* tri,
                           const int32_t* diff_attrs, int K, const float* dy, const float* dda, int N, int T, int H, int W,
                           float* grad_attr, float* grad_rast, float* grad_rast_db, fpc_stream_t stream);
/* Mip chain: level l has extents (Ht >> l, Wt >> l), every texel the mean of its 2x2 parents.  fpc_texture_mip_levels: how
 * many levels exist (extents must stay even; max_mip_level < 0 = no limit; at most 16).  The levels 1..L are stored back to
 * back in `mip` (fpc_texture_mip_floats floats). */
int fpc_texture_mip_levels(int Ht, int Wt, int max_mip_level);
size_t fpc_texture_mip_floats(int Nt, int Ht, int Wt, int C, int L);
int fpc_texture_mip_build(const float* tex, int Nt, int Ht, int Wt, int C, int L, float* mip, fpc_stream_t stream);
/* filter_mode 'linear-mipmap-linear' (nearest_level == 0) / 'linear-mipmap-nearest' (1), boundary wrap.
 * level = 0.5 log2(squared major axis of the pixel's texel-space footprint from uv_da [N,H,W,4] = (du/dX, du/dY, dv/dX, dv/dY))
 *         + mip_level_bias [N,H,W] (either may be NULL, not both), clamped to [0, L]. */
int fpc_texture_mip_fwd(const float* tex, const float* mip, int Nt, int Ht, int Wt, int C, int L, const float* uv,
                        const float* uv_da, const float* mip_level_bias, int nearest_level, int N, int H, int W, float* out,
                        fpc_stream_t stream);
/* dy -> grad_uv (overwritten), grad_uv_da / grad_bias (nullable, overwritten), grad_tex (nullable, overwritten; needs grad_mip,
 * a scratch of fpc_texture_mip_floats floats that receives the coarse levels' gradients, folded down the chain into grad_tex
 * unless mip_is_constant != 0: a caller-supplied mip stack is treated as constant data). */
int fpc_texture_mip_bwd(const float* tex, const float* mip, int Nt, int Ht, int Wt, int C, int L, const float* uv,
                        const float* uv_da, const float* mip_level_bias, int nearest_level, const float* dy, int N, int H, int W,
                        float* grad_tex, float* grad_mip, int mip_is_constant, float* grad_uv, float* grad_uv_da, float* grad_bias,
                        fpc_stream_t stream);

/* ---- background composite + image loss (replaces fit.py:161 and the first term of fit.py:579) --------------
 * colour [N,H,W,C], rast [N,H,W,4], ref [N,H,W,C] (grey levels, 0..255 scale):
 *   comp = rast.w > 0 ? colour : bg;   loss = scale * sum_n mean_{h,w,c} (ref - 255 comp)^2      (loss_kind 0, fit.py:579)
 *                                      loss = scale * sum_n mean_{h,w,c} |ref - 255 comp|        (loss_kind 1: the L1 form the
 *                                      north-star also asks for; d|e|/de = sign(e), 0 at e = 0)
 * -> loss [1] (overwritten), d_colour [N,H,W,C] (overwritten; 0 on background), comp [N,H,W,C] or NULL. */
size_t fpc_image_loss_scratch_bytes(int N, int H, int W, int C);
int fpc_image_loss_fwd_bwd(const float* colour, const float* rast, const float* ref, int N, int H, int W, int C,
                           float bg, float scale, int loss_kind, float* loss, float* d_colour, float* comp,
                           void* scratch, size_t scratch_bytes, fpc_stream_t stream);

/* ---- fused render + loss + gradient (no antialias): rasterize -> interpolate -> [texture] -> background ->
 *      image loss -> backward to clip-space positions, in one kernel per (64x64-px bin, view).  Replaces the chain
 *      fit.py:151-158,161,579 and its part of loss.backward() (fit.py:611) when antialias is off.
 * attr [Va,A] + attr_tri [T,3]: vertex colours (tex == NULL, A == C) or uv (tex [Ht,Wt,C] given, A == 2);
 * ref [N,H,W,C] float32 (ref_is_u8 == 0) or uint8 (ref_is_u8 == 1) grey levels on the 0..255 scale; C in {1,3}.
 *   loss [1]            = scale * sum_n mean_{h,w,c} (ref - 255 comp)^2  (loss_kind 0) or |ref - 255 comp| (loss_kind 1), overwritten
 *   grad_pos [N,V,4]    = d loss / d pos (x, y, w; z = 0), overwritten; NULL = forward only.  Needs vadj_off [V+1] /
 *                         vadj_item [3T], the vertex -> (triangle, corner) adjacency of the mesh (fpc_vertex_adjacency_build,
 *                         once per mesh): the position gradient is assembled WITHOUT atomics — every triangle's contribution is
 *                         written once into a per-(view, triangle, bin) slot and a per-vertex pass gathers the slots around the
 *                         vertex in a fixed order — so it is bit-reproducible from run to run (exception: triangles larger than
 *                         128 px or 2 x 2 bins, or clipped by the near plane, accumulate with float REDs)
 *   grad_tex [Ht,Wt,C]  = d loss / d tex (texture optimisation, tex_opt of fit.py:439,502), overwritten; NULL to skip
 *                         (needs tex and grad_pos; 4C float REDs per covered pixel)
 *   rast_out [N,H,W,4], colour_out [N,H,W,C] (composited image): optional outputs, NULL to skip the HBM writes. */
size_t fpc_render_loss_fused_scratch_bytes(int N, int T, int H, int W);
int fpc_render_loss_fused(const float* pos, const int32_t* tri, const float* attr, const int32_t* attr_tri, int Va, int A,
                          const float* tex, int Ht, int Wt, const void* ref, int ref_is_u8,
                          int N, int V, int T, int H, int W, int C, float bg, float scale, int loss_kind,
                          float* loss, float* grad_pos, float* grad_tex, float* rast_out, float* colour_out,
                          const int32_t* vadj_off, const int32_t* vadj_item,
                          void* scratch, size_t scratch_bytes, fpc_stream_t stream);

/* Vertex -> (triangle, corner) adjacency of a mesh in CSR form, for the gradient gather of the fused kernels: vadj_off [V+1],
 * vadj_item [3T] with item = triangle * 4 + corner, every vertex's items in ascending order (canonical: the gather order, and
 * with it the rounding of grad_pos, does not depend on how the list was built).  Once per mesh (role of the per-call hash
 * upstream's antialias builds; there is no reference counterpart for the rasterizer gradient, which upstream scatters with
 * atomics). */
size_t fpc_vertex_adjacency_scratch_bytes(int V);
int fpc_vertex_adjacency_build(const int32_t* tri, int T, int V, int32_t* vadj_off, int32_t* vadj_item,
                               void* scratch, size_t scratch_bytes, fpc_stream_t stream);

/* The same with dr.antialias (fit.py:160) between shading and the background composite: replaces the chain
 * fit.py:151-161,579 and its part of loss.backward() (fit.py:611).  tri_opp [T,3] from fpc_topology_build.
 * One kernel per (32x32-px bin, view) resolves the bin plus a 2-px halo in shared memory; the antialias forward, the colour
 * gradient AND the silhouette position gradient are atomics-free gathers (the latter joins the per-(view, triangle, bin)
 * gradient slots, see grad_pos above); results equal fpc_antialias_fwd/bwd applied to the op-level chain.
 * Same scratch size as fpc_render_loss_fused; colour_out is the antialiased, composited image. */
int fpc_render_loss_fused_aa(const float* pos, const int32_t* tri, const int32_t* tri_opp, const float* attr,
                             const int32_t* attr_tri, int Va, int A, const float* tex, int Ht, int Wt,
                             const void* ref, int ref_is_u8, int N, int V, int T, int H, int W, int C, float bg, float scale, int loss_kind,
                             float* loss, float* grad_pos, float* grad_tex, float* rast_out, float* colour_out,
                             const int32_t* vadj_off, const int32_t* vadj_item,
                             void* scratch, size_t scratch_bytes, fpc_stream_t stream);

/* Camera split at bin-row granularity (multi-GPU; fpc_diffrend_b200/shard.py: view_band_shard): the same
 * kernels (tri_opp == NULL: without antialias), but of the first view of every frame (view index n with
 * n % views_per_frame == 0) only the rows of fpc_raster_bin_px()-pixel bins >= row_lo are rendered, and of the last view
 * (n % views_per_frame == views_per_frame - 1) only the rows < row_hi; the other bins contribute neither loss nor gradient
 * (another rank renders them) and their rast_out / colour_out pixels are left untouched.  Every pixel, and every antialias
 * pixel pair (owned by its lower / left pixel), belongs to exactly one bin: the ranks' losses and gradients add up to those of
 * the unsplit call. */
int fpc_raster_bin_px(void);
/* Near-plane clipper bookkeeping of the LAST binning that used `scratch` (fpc_rasterize_fwd or a fused entry with the same
 * N, T, H, W): pieces of triangles crossing w <= 0 live in a pool of `capacity` = N*T/32 + 1024 entries; `requested` > `capacity`
 * means pieces were dropped (a camera inside the mesh, or a wildly wrong pose).  A validation call: it synchronises the stream
 * and reads 4 bytes back — not for the steady-state loop. */
int fpc_rasterize_clip_pieces(const void* scratch, int N, int T, int H, int W, int* requested, int* capacity, fpc_stream_t stream);
int fpc_render_loss_fused_band(const float* pos, const int32_t* tri, const int32_t* tri_opp, const float* attr,
                               const int32_t* attr_tri, int Va, int A, const float* tex, int Ht, int Wt,
                               const void* ref, int ref_is_u8, int N, int V, int T, int H, int W, int C, float bg, float scale, int loss_kind,
                               int views_per_frame, int row_lo, int row_hi,
                               float* loss, float* grad_pos, float* grad_tex, float* rast_out, float* colour_out,
                               const int32_t* vadj_off, const int32_t* vadj_item,
                               void* scratch, size_t scratch_bytes, fpc_stream_t stream);

/* ---- camera-split exchanges over NVLink peer memory (multi-GPU part of the north-star; fit.py is single-process:
 *      no reference counterpart).  `peers` = HOST array of `world` device pointers, entry p = rank p's copy of the buffer as
 *      mapped into this process (symmetric memory; the caller's own buffer included at its rank).  The kernels only move
 *      data; a device-side barrier between producer and consumer kernels is the caller's job.
 *   fpc_blend_fwd_bcast : verts[row0 + r] = base[r] + sum_b D[r,b] w[b] for this rank's R_local rows (single frame), stored
 *                         by the GEMV epilogue into the vertex buffer [R_total] of EVERY rank: blend + all-gather in one kernel
 *   fpc_peer_store_rows : src [F, rows_local] -> peers[p][f * rows_total + row0 + r] for every p (frame batches)
 *   fpc_peer_sum_rows   : out [F, rows_local] = sum_p peers[p][f * rows_total + row0 + r], rank order (reduce-scatter)
 *   fpc_peer_sum        : out [n] = sum_p peers[p][i], rank order: the same bits on every rank (all-reduce) */
int fpc_blend_fwd_bcast(const float* D, const float* base, const float* w, int R_local, int B, long long row0,
                        void* const* peer_verts, int world, fpc_stream_t stream);
int fpc_peer_store_rows(const float* src, void* const* peers, int world, int F, long long rows_local, long long rows_total,
                        long long row0, fpc_stream_t stream);
int fpc_peer_sum_rows(void* const* peers, int world, int F, long long rows_local, long long rows_total, long long row0,
                      float* out, fpc_stream_t stream);
int fpc_peer_sum(void* const* peers, int world, long long n, float* out, fpc_stream_t stream);

/* ---- mesh regularisers (replaces the pytorch3d terms of fit.py:578-582: weight_laplacian * laplacian(mesh)^2 +
 *      weight_meshedge * mesh_edge_loss(mesh, target) + weight_normalconsistency * mesh_normal_consistency(mesh)) -------
 * verts [F,V,3]; static topology (fpc_diffrend_b200/topology.py): neighbour CSR nbr_off [V+1] / nbr_idx [2E] over the E
 * unique edges and edge_quads [E2,4] = (v0, v1, a, b) for every pair of faces sharing an edge (NULL allowed if w_nc == 0).
 *   loss_accum [1] (nullable) += sum_f (w_lap lap_f^2 + w_edge edge_f + w_nc nc_f)
 *   terms [F,3] (nullable)     = raw (lap_f, edge_f, nc_f)
 *   d_verts [F,V,3]            = (accumulate ? += : =) gradient of that sum w.r.t. the vertices.
 * The Laplacian and edge gradients are gathers over the CSR (deterministic); only the normal-consistency term (weight 0 in
 * the shipped configuration, main.py:40) scatters with float REDs. */
size_t fpc_mesh_reg_scratch_bytes(int F, int V, int E2);
int fpc_mesh_reg_fwd_bwd(const float* verts, int F, int V, const int32_t* nbr_off, const int32_t* nbr_idx, int E,
                         const int32_t* edge_quads, int E2, float w_lap, float w_edge, float edge_target, float w_nc,
                         float* loss_accum, float* terms, float* d_verts, int accumulate,
                         void* scratch, size_t scratch_bytes, fpc_stream_t stream);

/* ---- Adam (replaces torch.optim.Adam + LambdaLR + quaternion renorm, fit.py:493-505,610-618) ---------------
 * p, g, m, v [n]; step_count [1]: device float holding the number of optimiser steps taken so far
 * (advanced by fpc_adam_advance after all parameter groups of an iteration have been stepped).
 * lr_eff = lr * lr_ramp^(step_count/max_iter)  (LambdaLR semantics), Adam's t = step_count + 1.
 * torch.optim.Adam update rule with bias correction, betas (b1,b2), eps, no weight decay / amsgrad. */
int fpc_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                  float lr_ramp, float max_iter, const float* step_count, fpc_stream_t stream);
/* The same for a parameter group that is unlocked at optimiser step `start_step` (combined mode: the learned basis gets
 * requires_grad after max_iter/2, fit.py:603-608): no-op while step_count < start_step (torch skips parameters without a
 * gradient, their state stays untouched); afterwards Adam's t = step_count - start_step + 1, the LambdaLR factor still
 * follows the global step_count. */
int fpc_adam_step_from(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                       float lr_ramp, float max_iter, const float* step_count, float start_step, fpc_stream_t stream);
/* step_count [1] += 1 (device side, so the iteration stays graph-capturable) */
int fpc_adam_advance(float* step_count, fpc_stream_t stream);
/* q [n,4] /= norm.  mode 0: per-row norm (default of this build);  mode 1: Frobenius norm of the whole tensor
 * (the reference's quirk, fit.py:616-618). */
int fpc_quat_renorm(float* q, int n, int mode, fpc_stream_t stream);
/* One launch for the packed parameter vector params = [w (F*B) | t (F*3) | q (F*4)] (grads, m, v alike):
 * fpc_adam_step for the three groups (learning rates lr_w, lr_t, lr_q; pose groups skipped when optimize_pose == 0),
 * fpc_quat_renorm (quat_mode as there) and fpc_adam_advance.  step_count [1] is read and then incremented. */
int fpc_adam_fused(float* params, const float* grads, float* m, float* v, int B, int F, int optimize_pose,
                   float lr_w, float lr_t, float lr_q, float b1, float b2, float eps, float lr_ramp, float max_iter,
                   int quat_mode, float* step_count, fpc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FPC_B200_H */
