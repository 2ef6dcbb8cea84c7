#!/usr/bin/env python
"""Generates tests/golden/graph_tiny.npz from the reference's own static-geometry code
(/root/reference: common/icosahedral_mesh.py, common/grid_mesh_connectivity.radius_query_indices,
common/model_utils.get_bipartite_graph_spatial_features, imported under tools/refshim only to
satisfy their `import xarray / trimesh / jax` lines) for the test-size configuration, following
gencast/denoiser.py:234-301, :419-510, :849-867 for how they are combined.

in_mesh_triangle_indices (mesh2grid) needs trimesh, which is not installable here: that table is
checked by geometric properties instead (tests/test_host_logic.py)."""
import os
import sys

import numpy as np
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("GENCAST_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)


def reference_tables(resolution=10.0, mesh_size=2, fraction=0.6):
    sys.path.insert(0, os.path.join(ROOT, "tools", "refshim"))
    sys.path.insert(0, REFERENCE)
    from common import grid_mesh_connectivity, icosahedral_mesh, model_utils
    lat = np.arange(-90.0, 90.0 + resolution / 2, resolution).astype(np.float32)
    lon = np.arange(0.0, 360.0, resolution).astype(np.float32)
    mesh = icosahedral_mesh.get_last_triangular_mesh_for_sphere(splits=mesh_size)
    # gencast/denoiser.py:849-867
    s, r = icosahedral_mesh.faces_to_edges(mesh.faces)
    n = mesh.vertices.shape[0]
    adj = sparse.lil_matrix((n, n))
    adj[s, r] = 1
    perm = sparse.csgraph.reverse_cuthill_mckee(adj.tocsr(), symmetric_mode=True)
    inv = {j: i for i, j in enumerate(perm)}
    mesh = icosahedral_mesh.TriangularMesh(vertices=mesh.vertices[perm], faces=np.vectorize(lambda x: inv[x])(mesh.faces))
    # gencast/denoiser.py:276-279, :840-846
    s, r = icosahedral_mesh.faces_to_edges(mesh.faces)
    radius = np.linalg.norm(mesh.vertices[s] - mesh.vertices[r], axis=-1).max() * fraction
    gi, mi = grid_mesh_connectivity.radius_query_indices(grid_latitude=lat, grid_longitude=lon, mesh=mesh, radius=radius)
    phi, theta = model_utils.cartesian_to_spherical(mesh.vertices[:, 0], mesh.vertices[:, 1], mesh.vertices[:, 2])
    m_lat, m_lon = model_utils.spherical_to_lat_lon(phi=phi, theta=theta)
    lon2d, lat2d = np.meshgrid(lon, lat)
    kw = dict(add_node_positions=False, add_node_latitude=True, add_node_longitude=True, add_relative_positions=True,
              relative_longitude_local_coordinates=True, relative_latitude_local_coordinates=True)   # denoiser.py:281-288
    sf, rf, ef = model_utils.get_bipartite_graph_spatial_features(
        senders_node_lat=lat2d.reshape(-1).astype(np.float32), senders_node_lon=lon2d.reshape(-1).astype(np.float32),
        receivers_node_lat=m_lat.astype(np.float32), receivers_node_lon=m_lon.astype(np.float32),
        senders=gi, receivers=mi, edge_normalization_factor=None, **kw)
    return dict(vertices=mesh.vertices, faces=mesh.faces, radius=np.asarray(radius), g2m_senders=gi, g2m_receivers=mi,
                g2m_grid_feat=sf, g2m_mesh_feat=rf, g2m_edge_feat=ef)


if __name__ == "__main__":
    t = reference_tables()
    dst = os.path.join(ROOT, "tests", "golden", "graph_tiny.npz")
    np.savez_compressed(dst, **t)
    print("wrote", dst, {k: np.asarray(v).shape for k, v in t.items()})
