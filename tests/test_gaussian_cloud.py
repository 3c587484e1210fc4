"""Host logic of the f1 row: PLY codec in the reference's schema and the stored-parameter
container whose getters are the reference's activations (scene/gaussian_model.py:90-117, 255-358)."""
import os
import struct

import numpy as np
import pytest
import torch

from goi_b200.gaussian_cloud import GaussianCloud, read_ply_vertices, write_ply_vertices
from goi_b200.scenes import make_scene


def _cloud(P=257, S=10, seed=3):
    g, _, _ = make_scene(P, 64, 48, S, seed)
    gain = 0.5 + torch.rand(P, generator=torch.Generator().manual_seed(seed))
    return g, GaussianCloud.from_activated(g.get_xyz, g.get_opacity, g.get_scaling, g.get_rotation, g.get_features,
                                           g.get_semantics, rotation_gain=gain)


def test_getters_are_the_reference_activations():
    g, c = _cloud()
    assert torch.allclose(c.get_opacity, g.get_opacity, atol=1e-6)
    assert torch.allclose(c.get_scaling, g.get_scaling, rtol=1e-6)
    assert torch.allclose(c.get_rotation, g.get_rotation, atol=1e-6)
    assert torch.equal(c.get_features, g.get_features)
    a, b = c.get_covariance(1.3), g.get_covariance(1.3)
    assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())
    m = (torch.arange(c.get_xyz.shape[0]) % 2).float()
    c.set_semantic_masks(m)
    assert torch.equal(c.get_semantics, c._semantics * m.unsqueeze(1))
    c.set_semantic_masks(None)
    assert c.get_semantics is c._semantics


def test_ply_schema_and_round_trip(tmp_path):
    _, c = _cloud(P=300, S=10)
    path = str(tmp_path / "point_cloud" / "iteration_1" / "point_cloud.ply")
    c.save_ply(path)
    with open(path, "rb") as f:
        header = f.read(4096).split(b"end_header\n")[0].decode().split("\n")
    props = [l.split()[2] for l in header if l.startswith("property")]
    # construct_list_of_attributes, scene/gaussian_model.py:255-270
    exp = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(45)] \
        + [f"sem_{i}" for i in range(10)] + ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]
    assert props == exp
    assert all(l.split()[1] == "float" for l in header if l.startswith("property"))
    assert os.path.getsize(path) == len("\n".join(header)) + len("end_header\n") + 300 * 4 * len(exp)

    d = GaussianCloud(3, 10).load_ply(path)
    for k, v in c.parameters().items():
        assert torch.equal(d.parameters()[k], v), k
    assert d.active_sh_degree == 3
    # f_dc / f_rest are stored channel-major: f_rest_0..14 = coefficient 1..15 of the RED channel (:277-278, :322-323)
    cols = read_ply_vertices(path)
    assert np.array_equal(cols["f_rest_1"], c._features_rest[:, 1, 0].numpy())
    assert np.array_equal(cols["f_rest_15"], c._features_rest[:, 0, 1].numpy())


def test_load_ply_semantic_dim_mismatch_gives_zeros(tmp_path):
    """load_ply :331-335: sem_* columns are only taken when their count equals semantic_dim."""
    _, c = _cloud(P=50, S=10)
    path = str(tmp_path / "a.ply")
    c.save_ply(path)
    d = GaussianCloud(3, 16).load_ply(path)
    assert d._semantics.shape == (50, 10) and float(d._semantics.abs().max()) == 0.0
    with pytest.raises(ValueError):
        GaussianCloud(2, 10).load_ply(path)             # the reference asserts on the f_rest count (:318)


def test_reader_handles_big_endian_ascii_and_extra_elements(tmp_path):
    cols = [("x", [1.0, 2.0]), ("y", [3.0, 4.0]), ("opacity", [0.25, -0.5])]
    p1 = str(tmp_path / "le.ply")
    write_ply_vertices(p1, cols)
    a = read_ply_vertices(p1)
    assert a["opacity"].tolist() == [0.25, -0.5]
    p2 = str(tmp_path / "be.ply")
    with open(p2, "wb") as f:
        f.write(b"ply\nformat binary_big_endian 1.0\ncomment made by hand\nelement vertex 2\nproperty float x\n"
                b"property double y\nproperty uchar flag\nelement face 0\nproperty list uchar int vertex_indices\n"
                b"end_header\n")
        f.write(struct.pack(">fdB", 1.5, 2.5, 7) + struct.pack(">fdB", -1.5, 1e-3, 9))
    b = read_ply_vertices(p2)
    assert b["x"].tolist() == [1.5, -1.5] and b["y"].tolist() == [2.5, 1e-3] and b["flag"].tolist() == [7, 9]
    p3 = str(tmp_path / "ascii.ply")
    with open(p3, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 2\nproperty float x\nproperty float y\nend_header\n1 2\n3 4.5\n")
    c = read_ply_vertices(p3)
    assert c["y"].tolist() == [2.0, 4.5]
    with open(p2, "r+b") as f:
        f.truncate(os.path.getsize(p2) - 3)
    with pytest.raises(ValueError):
        read_ply_vertices(p2)
