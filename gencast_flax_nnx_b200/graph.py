"""Static graph construction for the GenCast denoiser (host side, runs once).

This is the host-side precompute of SURVEY.md §8 row a13.  It builds the three
static graphs the denoiser works on and everything the CUDA kernels need to
walk them:

* the refined icosahedral mesh, relabelled to a banded ordering
  (reference: common/icosahedral_mesh.py:60-284, gencast/denoiser.py:849-867),
* grid->mesh edges from a radius query
  (reference: common/grid_mesh_connectivity.py:40-86, gencast/denoiser.py:443-510),
* mesh->grid edges, three per grid point, from the containing mesh triangle
  (reference: common/grid_mesh_connectivity.py:89-133, gencast/denoiser.py:551-600),
* structural node / edge features
  (reference: common/model_utils.py:364-591),
* the k-hop attention neighbourhoods of the mesh
  (reference: gencast/transformer.py:21-47, gencast/sparse_transformer.py:86-96,555),
* receiver-sorted CSR tables for the deterministic segment sums.

Everything here is numpy/scipy; nothing is on the per-forward path.
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Tuple

import numpy as np
from scipy import sparse
from scipy.sparse import csgraph
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation


# --------------------------------------------------------------------------
# Icosphere
# --------------------------------------------------------------------------

@dataclasses.dataclass(frozen=True)
class TriMesh:
    """Triangular mesh on the unit sphere: vertices [V,3] f32, faces [F,3] i32."""
    vertices: np.ndarray
    faces: np.ndarray


# Face table of the regular icosahedron for the vertex enumeration used below;
# every face is counter-clockwise seen from outside
# (reference: common/icosahedral_mesh.py:131-151; pinned by
# common/icosahedral_mesh_test.py:106-127).
_ICO_FACES = np.array(
    [(0, 1, 2), (0, 6, 1), (8, 0, 2), (8, 4, 0), (3, 8, 2), (3, 2, 7), (7, 2, 1),
     (0, 4, 6), (4, 11, 6), (6, 11, 5), (1, 5, 7), (4, 10, 11), (4, 8, 10),
     (10, 8, 3), (10, 3, 9), (11, 10, 9), (11, 9, 5), (5, 9, 7), (9, 3, 7),
     (1, 6, 5)], dtype=np.int32)


def icosahedron() -> TriMesh:
    """Regular icosahedron with two faces parallel to the xy plane.

    Reference: common/icosahedral_mesh.py:103-181.  Vertex enumeration, the f32
    rounding points and the final rotation about y follow the reference so that
    vertex coordinates agree bit for bit (tests/test_graph.py checks this against
    fixtures generated from the reference module).
    """
    golden = (1 + np.sqrt(5)) / 2
    verts = []
    for s1 in (1.0, -1.0):
        for s2 in (golden, -golden):
            verts += [(s1, s2, 0.0), (0.0, s1, s2), (s2, 0.0, s1)]
    verts = np.array(verts, dtype=np.float32)
    verts /= np.linalg.norm([1.0, golden])
    dihedral = 2 * np.arcsin(golden / np.sqrt(3))
    rot = Rotation.from_euler(seq="y", angles=(np.pi - dihedral) / 2).as_matrix()
    verts = np.dot(verts, rot)
    return TriMesh(vertices=verts.astype(np.float32), faces=_ICO_FACES.copy())


def _split_faces(mesh: TriMesh) -> TriMesh:
    """One 1->4 refinement of every triangle, new vertices pushed to the sphere.

    Reference: common/icosahedral_mesh.py:184-256.  New vertices are numbered in
    order of first use while walking faces in order and, inside a face, edges
    (v0,v1), (v1,v2), (v2,v0); that ordering is load-bearing (it fixes vertex
    ids, hence every index table downstream).
    """
    faces = mesh.faces
    parents = mesh.vertices
    n_parent = parents.shape[0]
    # Edge keys in creation order: per face (01, 12, 20), endpoints sorted.
    ends = np.stack([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=1)
    ends = np.sort(ends.reshape(-1, 2), axis=1).astype(np.int64)
    keys = ends[:, 0] * n_parent + ends[:, 1]
    uniq, first_pos, inverse = np.unique(keys, return_index=True, return_inverse=True)
    creation_rank = np.argsort(np.argsort(first_pos))       # rank of each unique key by first use
    child_of_edge = (n_parent + creation_rank[inverse]).reshape(-1, 3)

    order = np.argsort(first_pos)
    new_ends = np.stack([uniq[order] // n_parent, uniq[order] % n_parent], axis=1)
    children = np.empty((new_ends.shape[0], 3), dtype=parents.dtype)
    for i, (a, b) in enumerate(new_ends):
        # Midpoint then projection, in the vertex dtype (f32), one vertex at a
        # time so rounding matches the reference's per-vertex arithmetic.
        mid = parents[[a, b]].mean(0)
        mid /= np.linalg.norm(mid)
        children[i] = mid
    verts = np.concatenate([parents, children], axis=0)

    v0, v1, v2 = faces[:, 0], faces[:, 1], faces[:, 2]
    m01, m12, m20 = child_of_edge[:, 0], child_of_edge[:, 1], child_of_edge[:, 2]
    new_faces = np.stack([
        np.stack([v0, m01, m20], -1),
        np.stack([m01, v1, m12], -1),
        np.stack([m20, m12, v2], -1),
        np.stack([m01, m12, m20], -1)], axis=1).reshape(-1, 3)
    return TriMesh(vertices=verts, faces=new_faces.astype(np.int32))


def icosphere(splits: int) -> TriMesh:
    """Finest mesh after `splits` refinements (reference: icosahedral_mesh.py:284)."""
    mesh = icosahedron()
    for _ in range(splits):
        mesh = _split_faces(mesh)
    return mesh


def faces_to_edges(faces: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Directed edges 0->1, 1->2, 2->0 of every face, column by column.

    Reference: common/icosahedral_mesh.py:259-281 (ordering pinned by
    common/icosahedral_mesh_test.py:73-92).
    """
    assert faces.ndim == 2 and faces.shape[-1] == 3
    senders = np.concatenate([faces[:, 0], faces[:, 1], faces[:, 2]])
    receivers = np.concatenate([faces[:, 1], faces[:, 2], faces[:, 0]])
    return senders, receivers


def permute_mesh_to_banded(mesh: TriMesh) -> TriMesh:
    """Relabel vertices by reverse Cuthill-McKee (reference: denoiser.py:849-867)."""
    s, r = faces_to_edges(mesh.faces)
    n = mesh.vertices.shape[0]
    adj = sparse.csr_matrix((np.ones(s.shape[0]), (s, r)), shape=(n, n))
    adj.data[:] = 1
    perm = csgraph.reverse_cuthill_mckee(adj, symmetric_mode=True)
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    return TriMesh(vertices=mesh.vertices[perm], faces=inv[mesh.faces])


def max_edge_length(mesh: TriMesh) -> float:
    """Longest mesh edge in R^3 (reference: denoiser.py:840-846)."""
    s, r = faces_to_edges(mesh.faces)
    return np.linalg.norm(mesh.vertices[s] - mesh.vertices[r], axis=-1).max()


# --------------------------------------------------------------------------
# Coordinates
# --------------------------------------------------------------------------

def lat_lon_deg_to_spherical(lat, lon):
    """(phi=lon, theta=colatitude) in radians (reference: model_utils.py:175-180)."""
    return np.deg2rad(lon), np.deg2rad(90 - lat)


def spherical_to_cartesian(phi, theta):
    """Reference: model_utils.py:202-208."""
    return (np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta))


def mesh_lat_lon(mesh: TriMesh) -> Tuple[np.ndarray, np.ndarray]:
    """Mesh vertex lat/lon in degrees, f32 (reference: denoiser.py:419-429)."""
    x, y, z = mesh.vertices[:, 0], mesh.vertices[:, 1], mesh.vertices[:, 2]
    phi = np.arctan2(y, x)
    with np.errstate(invalid="ignore"):
        theta = np.arccos(z)
    lon = np.mod(np.rad2deg(phi), 360)
    lat = 90 - np.rad2deg(theta)
    return lat.astype(np.float32), lon.astype(np.float32)


def grid_positions(grid_lat: np.ndarray, grid_lon: np.ndarray) -> np.ndarray:
    """[n_lat*n_lon, 3] unit vectors, lat-major (reference: grid_mesh_connectivity.py:22-37)."""
    phi, theta = np.meshgrid(np.deg2rad(grid_lon), np.deg2rad(90 - grid_lat))
    return np.stack([np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta),
                     np.cos(theta)], axis=-1).reshape(-1, 3)


# --------------------------------------------------------------------------
# Connectivity
# --------------------------------------------------------------------------

def radius_query_indices(grid_lat, grid_lon, mesh: TriMesh, radius: float):
    """Grid->mesh edges: every (grid point, mesh vertex) closer than `radius` in R^3.

    Reference: common/grid_mesh_connectivity.py:40-86.  Edges come out
    grid-point-major with mesh ids ascending inside a grid point (cKDTree sorts
    multi-point queries).
    """
    pos = grid_positions(grid_lat, grid_lon)
    tree = cKDTree(mesh.vertices)
    hits = tree.query_ball_point(x=pos, r=radius)
    counts = np.fromiter((len(h) for h in hits), dtype=np.int64, count=len(hits))
    grid_idx = np.repeat(np.arange(pos.shape[0]), counts)
    mesh_idx = np.concatenate([np.asarray(h, dtype=np.int64) for h in hits]) if counts.sum() else np.zeros(0, np.int64)
    return grid_idx.astype(np.int64), mesh_idx.astype(np.int64)


def _closest_point_sqdist(p, a, b, c):
    """Squared distance from points p[n,3] to triangles (a,b,c)[n,3] (Ericson, RTCD 5.1.5)."""
    ab, ac, ap = b - a, c - a, p - a
    d1 = np.einsum("ij,ij->i", ab, ap); d2 = np.einsum("ij,ij->i", ac, ap)
    bp = p - b
    d3 = np.einsum("ij,ij->i", ab, bp); d4 = np.einsum("ij,ij->i", ac, bp)
    cp = p - c
    d5 = np.einsum("ij,ij->i", ab, cp); d6 = np.einsum("ij,ij->i", ac, cp)
    vc = d1 * d4 - d3 * d2
    vb = d5 * d2 - d1 * d6
    va = d3 * d6 - d5 * d4
    out = np.empty_like(p)
    done = np.zeros(p.shape[0], dtype=bool)

    def put(mask, val):
        m = mask & ~done
        out[m] = val[m]
        done[m] = True

    with np.errstate(divide="ignore", invalid="ignore"):
        put((d1 <= 0) & (d2 <= 0), a)
        put((d3 >= 0) & (d4 <= d3), b)
        put((d6 >= 0) & (d5 <= d6), c)
        v = d1 / (d1 - d3)
        put((vc <= 0) & (d1 >= 0) & (d3 <= 0), a + v[:, None] * ab)
        w = d2 / (d2 - d6)
        put((vb <= 0) & (d2 >= 0) & (d6 <= 0), a + w[:, None] * ac)
        w2 = (d4 - d3) / ((d4 - d3) + (d5 - d6))
        put((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), b + w2[:, None] * (c - b))
        denom = 1.0 / (va + vb + vc)
        put(np.ones_like(done), a + ab * (vb * denom)[:, None] + ac * (vc * denom)[:, None])
    diff = p - out
    return np.einsum("ij,ij->i", diff, diff)


def in_mesh_triangle_indices(grid_lat, grid_lon, mesh: TriMesh, n_candidates: int = 12):
    """Mesh->grid edges: the 3 vertices of the mesh face closest to each grid point.

    Reference: common/grid_mesh_connectivity.py:89-133, which delegates the
    closest-face query to trimesh.proximity.closest_point (trimesh is not
    available here).  Restated as: candidate faces by nearest face centroids,
    exact point-triangle distance, smallest face id on ties.  DEVIATION: where a
    grid point is equidistant from several faces (it sits over a mesh vertex or
    edge, e.g. the pole rows) trimesh's pick is implementation-defined; ours is
    the lowest face id.  Output ordering (grid-major, 3 consecutive edges per
    grid point, vertex order of the face) follows the reference.
    """
    pos = grid_positions(grid_lat, grid_lon).astype(np.float64)
    verts = mesh.vertices.astype(np.float64)
    tri = verts[mesh.faces]                                   # [F,3,3]
    centroids = tri.mean(axis=1)
    k = min(n_candidates, mesh.faces.shape[0])
    _, cand = cKDTree(centroids).query(pos, k=k)              # [G,k]
    cand = np.sort(cand, axis=1)                              # ascending face id -> argmin picks lowest id on ties
    n = pos.shape[0]
    best = np.full(n, np.inf)
    best_face = np.zeros(n, dtype=np.int64)
    for j in range(k):
        f = cand[:, j]
        d = _closest_point_sqdist(pos, tri[f, 0], tri[f, 1], tri[f, 2])
        better = d < best - 1e-15
        best[better] = d[better]
        best_face[better] = f[better]
    mesh_idx = mesh.faces[best_face].reshape(-1).astype(np.int64)
    grid_idx = np.repeat(np.arange(n, dtype=np.int64), 3)
    return grid_idx, mesh_idx


# --------------------------------------------------------------------------
# Structural features
# --------------------------------------------------------------------------

def _rotation_to_local(phi, theta):
    """Per-node rotation taking the node to lon 0, lat 0 (reference: model_utils.py:326-339)."""
    return Rotation.from_euler("zy", np.stack([-phi, -theta + np.pi / 2], axis=1)).as_matrix()


def _node_features(phi, theta):
    """[cos(colat), cos(lon), sin(lon)] (reference: model_utils.py:445-457)."""
    return np.stack([np.cos(theta), np.cos(phi), np.sin(phi)], axis=-1)


def bipartite_spatial_features(senders_lat, senders_lon, receivers_lat, receivers_lon,
                               senders, receivers):
    """Structural features of a bipartite graph, GenCast's flag set.

    Reference: common/model_utils.py:364-502 and :505-591 with
    add_node_positions=False, add_node_latitude=True, add_node_longitude=True,
    add_relative_positions=True and both local-coordinate flags True
    (gencast/denoiser.py:241-248).  Returns (sender_nodes[Ns,3],
    receiver_nodes[Nr,3], edges[E,4]); edges = [|d|, dx, dy, dz] / max|d| with d
    the sender-minus-receiver offset in the receiver's rotated frame.
    """
    s_phi, s_theta = lat_lon_deg_to_spherical(senders_lat, senders_lon)
    r_phi, r_theta = lat_lon_deg_to_spherical(receivers_lat, receivers_lon)
    s_feat = _node_features(s_phi, s_theta)
    r_feat = _node_features(r_phi, r_theta)
    s_pos = np.stack(spherical_to_cartesian(s_phi, s_theta), axis=-1)
    r_pos = np.stack(spherical_to_cartesian(r_phi, r_theta), axis=-1)
    rot = _rotation_to_local(r_phi, r_theta)[receivers]
    # einsum("bji,bi->bj") == R @ p per edge (model_utils.py:359-361).
    r_local = np.einsum("bji,bi->bj", rot, r_pos[receivers])
    s_local = np.einsum("bji,bi->bj", rot, s_pos[senders])
    rel = s_local - r_local
    dist = np.linalg.norm(rel, axis=-1, keepdims=True)
    scale = dist.max()
    e_feat = np.concatenate([dist / scale, rel / scale], axis=-1)
    return s_feat, r_feat, e_feat


# --------------------------------------------------------------------------
# k-hop attention neighbourhoods
# --------------------------------------------------------------------------

def khop_neighbourhoods(mesh: TriMesh, k_hop: int) -> sparse.csr_matrix:
    """Boolean reachability within k hops, self included, as CSR with sorted columns.

    Reference: gencast/transformer.py:21-47 builds the int32 adjacency with self
    edges and gencast/sparse_transformer.py:555 raises it to the k-th power; the
    non-zero pattern of that power is what masks the attention.  We compute the
    pattern by boolean reachability (no path counts, so no int32 overflow at
    large k — SURVEY.md Appendix A).
    """
    s, r = faces_to_edges(mesh.faces)
    n = mesh.vertices.shape[0]
    adj = sparse.csr_matrix((np.ones(s.shape[0], dtype=np.float32), (s, r)), shape=(n, n))
    adj = adj + sparse.identity(n, dtype=np.float32, format="csr")
    adj.data[:] = 1
    reach = adj.copy()
    for _ in range(k_hop - 1):
        reach = reach @ adj
        reach.data[:] = 1
    reach.sort_indices()
    return reach.tocsr()


def mask_block_size(mask: sparse.spmatrix) -> int:
    """Band half-width (+1) of a sparse matrix = the reference's attention block size.

    Reference: gencast/sparse_transformer.py:86-96.  Only used as a known-answer
    check of the ordering (649 / 1289 / 2569).
    """
    coo = mask.tocoo()
    n = mask.shape[0]
    first_row = np.full(n, n, dtype=np.int64)
    last_row = np.full(n, -1, dtype=np.int64)
    np.minimum.at(first_row, coo.col, coo.row)
    np.maximum.at(last_row, coo.col, coo.row)
    cols = np.arange(n)
    lower = (cols - first_row + 1).max()
    upper = (last_row - cols + 1).max()
    return int(max(lower, upper))


def khop_tiles(khop: sparse.spmatrix, tile: int = 128):
    """Block-sparse form of the k-hop pattern for the tensor-core attention kernel.

    Returns (tile_ptr[nq+1] i32, tile_kv[nt] i32, tile_mask[nt, tile, tile//32] u32): for each
    query tile the non-empty key tiles in ascending order and one bit mask per pair (bit j%32 of
    word j//32 of row r = key kv*tile + j is a k-hop neighbour of query qt*tile + r).
    """
    coo = khop.tocoo()
    n = khop.shape[0]
    nq = -(-n // tile)
    qt, kt = coo.row // tile, coo.col // tile
    key = qt.astype(np.int64) * nq + kt
    uniq, inv = np.unique(key, return_inverse=True)
    tile_q, tile_kv = (uniq // nq).astype(np.int32), (uniq % nq).astype(np.int32)
    tile_ptr = np.zeros(nq + 1, np.int32)
    np.cumsum(np.bincount(tile_q, minlength=nq), out=tile_ptr[1:])
    bits = np.zeros((len(uniq), tile, tile), dtype=bool)
    bits[inv, coo.row % tile, coo.col % tile] = True
    mask = np.packbits(bits, axis=-1, bitorder="little").view(np.uint32).reshape(len(uniq), tile, tile // 32)
    return tile_ptr, tile_kv, np.ascontiguousarray(mask)


def pack_key_ranges(tile_kv: np.ndarray, tile_mask: np.ndarray) -> np.ndarray:
    """Annotates every listed (query tile, key tile) pair with the 32-key sub-blocks no query of the pair
    attends to at either end of the key tile: bits 24-25 of tile_kv = number of leading dead sub-blocks,
    bits 26-27 = trailing ones (0 = the whole tile is used, which is what an unannotated list says).
    The tensor-core attention kernel then forms S and P V over the live key range only (N = 32 .. 128):
    with the hierarchical patch order the range is 2.86 of 4 sub-blocks on average at 1 deg."""
    kv = np.asarray(tile_kv, np.int64)
    if kv.size and kv.max() >= (1 << 24):
        raise ValueError("pack_key_ranges: key tile index does not fit 24 bits")
    m = np.asarray(tile_mask).view(np.uint32).reshape(len(kv), -1, 4)
    live = (m != 0).any(axis=1)                                   # [tiles, 4]
    any_live = live.any(axis=1)
    first = np.where(any_live, live.argmax(axis=1), 0)
    last = np.where(any_live, 3 - live[:, ::-1].argmax(axis=1), 3)
    return (kv | (first.astype(np.int64) << 24) | ((3 - last).astype(np.int64) << 26)).astype(np.int32)


def khop_compact_steps(khop: sparse.spmatrix, tile_q: int = 128, step: int = 64):
    """Per-query-tile compacted key lists for the gather form of the tensor-core attention kernel.

    For each tile of `tile_q` consecutive queries the sorted union of the keys any of them attends to is
    cut into steps of `step` keys (the last step is padded by repeating its last key; padded columns have
    no mask bit).  A 128-query patch of the 1 deg mesh attends to 690 distinct keys on average, which the
    plain (query tile, key tile) list spreads over 11.3 key tiles of 128 (1 446 keys): compacted it is 5.6
    tiles' worth.  The kernel gathers the listed K / V rows into dense operand tiles itself.

    Returns (step_ptr[nq+1] i32, keys[ns*step] i32, mask[ns, tile_q, step//32] u32, work[nq] i32):
    steps step_ptr[t] .. step_ptr[t+1]-1 belong to query tile t; bit j%32 of word j//32 of mask[s, r] says
    that key keys[s*step + j] is a neighbour of query t*tile_q + r; `work` lists the query tiles by
    decreasing step count (longest first, stable) for the launch order.
    """
    csr = khop.tocsr()
    csr.sort_indices()
    n = csr.shape[0]
    nq = -(-n // tile_q)
    words = step // 32
    step_ptr = np.zeros(nq + 1, np.int64)
    keys_all, mask_all = [], []
    for t in range(nq):
        r0, r1 = t * tile_q, min((t + 1) * tile_q, n)
        lo, hi = csr.indptr[r0], csr.indptr[r1]
        cols = csr.indices[lo:hi]
        uniq = np.unique(cols)
        ns = -(-len(uniq) // step)
        step_ptr[t + 1] = step_ptr[t] + ns
        if ns == 0:
            continue
        padded = np.full(ns * step, uniq[-1], np.int32)
        padded[:len(uniq)] = uniq
        keys_all.append(padded)
        rows = np.repeat(np.arange(r0, r1), np.diff(csr.indptr[r0:r1 + 1])) - r0
        pos = np.searchsorted(uniq, cols)
        bits = np.zeros((ns, tile_q, step), dtype=bool)
        bits[pos // step, rows, pos % step] = True
        mask_all.append(np.packbits(bits, axis=-1, bitorder="little").view(np.uint32).reshape(ns, tile_q, words))
    keys = np.concatenate(keys_all) if keys_all else np.zeros(0, np.int32)
    mask = np.concatenate(mask_all) if mask_all else np.zeros((0, tile_q, words), np.uint32)
    counts = np.diff(step_ptr)
    work = np.argsort(-counts, kind="stable").astype(np.int32)
    return step_ptr.astype(np.int32), keys.astype(np.int32), np.ascontiguousarray(mask), work


def patch_order(xyz: np.ndarray, leaf: int = 128, sub_leaf: int = 32) -> np.ndarray:
    """Permutation that groups mesh nodes into spatially compact patches of `leaf` nodes, each of
    which is itself ordered into compact sub-patches of `sub_leaf` nodes.

    Recursive bisection along the principal axis with left halves sized in multiples of `leaf`
    (then `sub_leaf`), so every aligned block of `leaf` (`sub_leaf`) consecutive nodes is one
    patch.  Attention is permutation equivariant and mesh latents never leave the denoiser, so the
    engine is free to relabel mesh nodes; compact patches put the k-hop neighbourhoods of a query
    tile into fewer key tiles than the band ordering the reference needs for its tri-block mask
    (gencast/denoiser.py:849-867), and compact sub-patches leave about half of the 32 x 32
    sub-blocks of those tiles empty, which the attention kernel skips.
    new position i holds old node order[i].
    """
    xyz = np.asarray(xyz, np.float64)
    out = []

    def rec(ids):
        n = len(ids)
        if n <= sub_leaf:
            out.append(ids)
            return
        step = leaf if n > leaf else sub_leaf
        c = xyz[ids] - xyz[ids].mean(0)
        _, _, vt = np.linalg.svd(c, full_matrices=False)
        order = np.argsort(c @ vt[0], kind="stable")
        nl = ((n // 2 + step - 1) // step) * step
        if nl >= n:
            nl = n - step
        rec(ids[order[:nl]])
        rec(ids[order[nl:]])

    rec(np.arange(len(xyz)))
    return np.concatenate(out)


# --------------------------------------------------------------------------
# CSR by receiver
# --------------------------------------------------------------------------

def csr_by_receiver(receivers: np.ndarray, n_receivers: int):
    """Stable receiver sort -> (row_ptr[n+1] i32, edge_perm[E] i32).

    edge_perm[j] is the original edge id of the j-th edge in receiver order; a
    stable sort keeps the summation order fixed, which is what makes the
    segment sum deterministic (no float atomics).  jraph.segment_sum
    (call sites common/typed_graph_net.py:173,182) is order-free in exact
    arithmetic, so any fixed order is a valid restatement.
    """
    perm = np.argsort(receivers, kind="stable").astype(np.int32)
    counts = np.bincount(receivers, minlength=n_receivers)
    row_ptr = np.zeros(n_receivers + 1, dtype=np.int32)
    np.cumsum(counts, out=row_ptr[1:])
    return row_ptr, perm


# --------------------------------------------------------------------------
# The bundle
# --------------------------------------------------------------------------

@dataclasses.dataclass
class DenoiserGraphs:
    """All static tables for one (grid, mesh_size, k_hop) combination."""
    grid_lat: np.ndarray            # [n_lat] f32
    grid_lon: np.ndarray            # [n_lon] f32
    mesh: TriMesh                   # banded ordering
    query_radius: float
    # grid2mesh
    g2m_senders: np.ndarray         # [E1] grid ids
    g2m_receivers: np.ndarray       # [E1] mesh ids
    g2m_grid_feat: np.ndarray       # [G,3] f32
    g2m_mesh_feat: np.ndarray       # [V,3] f32
    g2m_edge_feat: np.ndarray       # [E1,4] f32
    # mesh2grid
    m2g_senders: np.ndarray         # [E2] mesh ids
    m2g_receivers: np.ndarray       # [E2] grid ids (== repeat(arange(G),3))
    m2g_edge_feat: np.ndarray       # [E2,4] f32
    # mesh attention
    khop: sparse.csr_matrix         # [V,V] boolean pattern
    k_hop: int

    @property
    def num_grid_nodes(self) -> int:
        return self.grid_lat.shape[0] * self.grid_lon.shape[0]

    @property
    def num_mesh_nodes(self) -> int:
        return self.mesh.vertices.shape[0]


def build_denoiser_graphs(grid_lat, grid_lon, mesh_size: int, k_hop: int,
                          radius_query_fraction_edge_length: float = 0.6) -> DenoiserGraphs:
    """Build every static table (reference: gencast/denoiser.py:234-301,343-360,419-600)."""
    grid_lat = np.asarray(grid_lat).astype(np.float32)
    grid_lon = np.asarray(grid_lon).astype(np.float32)
    mesh = permute_mesh_to_banded(icosphere(mesh_size))
    radius = max_edge_length(mesh) * radius_query_fraction_edge_length
    m_lat, m_lon = mesh_lat_lon(mesh)
    lon2d, lat2d = np.meshgrid(grid_lon, grid_lat)
    g_lat = lat2d.reshape(-1).astype(np.float32)
    g_lon = lon2d.reshape(-1).astype(np.float32)

    gi, mi = radius_query_indices(grid_lat, grid_lon, mesh, radius)
    g_feat, m_feat, e1_feat = bipartite_spatial_features(g_lat, g_lon, m_lat, m_lon, gi, mi)

    gi2, mi2 = in_mesh_triangle_indices(grid_lat, grid_lon, mesh)
    _, _, e2_feat = bipartite_spatial_features(m_lat, m_lon, g_lat, g_lon, mi2, gi2)

    return DenoiserGraphs(
        grid_lat=grid_lat, grid_lon=grid_lon, mesh=mesh, query_radius=float(radius),
        g2m_senders=gi, g2m_receivers=mi,
        g2m_grid_feat=g_feat.astype(np.float32), g2m_mesh_feat=m_feat.astype(np.float32),
        g2m_edge_feat=e1_feat.astype(np.float32),
        m2g_senders=mi2, m2g_receivers=gi2, m2g_edge_feat=e2_feat.astype(np.float32),
        khop=khop_neighbourhoods(mesh, k_hop), k_hop=k_hop)


def regular_grid(resolution_deg: float) -> Tuple[np.ndarray, np.ndarray]:
    """Equiangular lat/lon grid with poles: n_lon = 2 (n_lat - 1)."""
    lat = np.arange(-90.0, 90.0 + resolution_deg / 2, resolution_deg, dtype=np.float64)
    lon = np.arange(0.0, 360.0, resolution_deg, dtype=np.float64)
    return lat.astype(np.float32), lon.astype(np.float32)
