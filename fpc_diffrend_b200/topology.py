"""Static mesh topology for the regularisers of the loss (reference fit.py:578-582), built once on the host.

The reference lets pytorch3d rebuild `Meshes(...).edges_packed()` and the sparse Laplacian on every iteration
(fit.py:578) and has its own O(T) neighbour builder (data.py:44-66, `vertex_neighbours`); here the same information is
computed once with numpy and handed to the kernels of csrc/meshreg.cu as flat arrays.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class MeshTopology:
    edges: np.ndarray       # [E,2] i32 unique undirected edges, v0 < v1, sorted (pytorch3d edges_packed order)
    nbr_off: np.ndarray     # [V+1] i32 CSR offsets: neighbours of vertex i are nbr_idx[nbr_off[i]:nbr_off[i+1]]
    nbr_idx: np.ndarray     # [2E] i32 sorted neighbour lists (role of data.py:44-66 without the pad-to-8)
    edge_quads: np.ndarray  # [E2,4] i32 (v0, v1, a, b): the two faces (v0,v1,a) and (v0,v1,b) sharing edge (v0,v1);
                            #   edges with k > 2 faces contribute all k(k-1)/2 pairs (mesh_normal_consistency semantics)

    @property
    def E(self):
        return self.edges.shape[0]

    @property
    def E2(self):
        return self.edge_quads.shape[0]


def build_topology(tri, n_vertices):
    tri = np.asarray(tri, dtype=np.int64)
    T = tri.shape[0]
    e = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]])          # [3T,2] directed half-edges
    opp = np.concatenate([tri[:, 2], tri[:, 0], tri[:, 1]])                        # opposite vertex of each half-edge
    lo, hi = e.min(axis=1), e.max(axis=1)
    key = lo * n_vertices + hi
    order = np.argsort(key, kind='stable')
    key_s, opp_s = key[order], opp[order]
    uniq, first, counts = np.unique(key_s, return_index=True, return_counts=True)
    edges = np.stack([uniq // n_vertices, uniq % n_vertices], axis=1).astype(np.int32)
    # neighbour CSR (both directions of every unique edge)
    src = np.concatenate([edges[:, 0], edges[:, 1]]).astype(np.int64)
    dst = np.concatenate([edges[:, 1], edges[:, 0]]).astype(np.int64)
    o2 = np.lexsort((dst, src))
    nbr_idx = dst[o2].astype(np.int32)
    nbr_off = np.zeros(n_vertices + 1, dtype=np.int32)
    np.cumsum(np.bincount(src, minlength=n_vertices), out=nbr_off[1:])
    # face pairs across edges
    quads = []
    two = counts == 2
    if two.any():
        f2 = first[two]
        quads.append(np.stack([edges[two, 0], edges[two, 1], opp_s[f2], opp_s[f2 + 1]], axis=1))
    for i in np.nonzero(counts > 2)[0]:                                            # non-manifold edges: all pairs
        os_ = opp_s[first[i]:first[i] + counts[i]]
        for a in range(len(os_)):
            for b in range(a + 1, len(os_)):
                quads.append(np.array([[edges[i, 0], edges[i, 1], os_[a], os_[b]]]))
    edge_quads = np.concatenate(quads).astype(np.int32) if quads else np.zeros((0, 4), np.int32)
    return MeshTopology(edges=edges, nbr_off=nbr_off, nbr_idx=nbr_idx, edge_quads=np.ascontiguousarray(edge_quads))
