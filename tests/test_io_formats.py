"""f-4: on-disk formats of the reference (gs.ply, as_occ_prior.ply, pt.yaml) restated in gssdf_b200/io.py: header text, property order
and byte layout as NeuralGS::export_gs_to_ply (neural_gaussian.cpp:928-1039) writes them, round trips, and reading a 3DGS-style file whose
properties come in a different order (tinyply looks properties up by name)."""
import math
import os
import struct

import numpy as np
import torch


def _gs(n, K, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(anchors=r(n, 3), offsets=r(n, 3) * 0.01, features_dc=r(n, 1, 3), features_rest=r(n, K - 1, 3), opacity=r(n), scaling=r(n, 3) - 3,
                quaternion=r(n, 4))


def test_gs_ply_layout_and_round_trip(tmp_path):
    from gssdf_b200 import io
    n, deg = 37, 2
    K = (deg + 1) ** 2
    p = _gs(n, K)
    path = str(tmp_path / "model" / "gs.ply")
    assert io.export_gs_to_ply(path, **p) == n
    raw = open(path, "rb").read()
    end = raw.index(b"end_header\n") + 11
    hdr = raw[:end].decode().splitlines()
    assert hdr[:3] == ["ply", "format binary_little_endian 1.0", f"element vertex {n}"]
    props = [ln.split()[2] for ln in hdr if ln.startswith("property")]
    assert all(ln.split()[1] == "float" for ln in hdr if ln.startswith("property"))
    assert props == ["x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(3 * (K - 1))] + \
        ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    assert len(raw) - end == n * len(props) * 4
    # first vertex, field by field
    v0 = struct.unpack("<" + "f" * len(props), raw[end:end + 4 * len(props)])
    xyz = (p["anchors"] + p["offsets"])[0]
    assert np.allclose(v0[:3], xyz.numpy()) and np.allclose(v0[3:6], p["features_dc"][0, 0].numpy())
    rest = p["features_rest"][0].transpose(0, 1).flatten()  # channel-major
    assert np.allclose(v0[6:6 + 3 * (K - 1)], rest.numpy())
    o = 6 + 3 * (K - 1)
    assert np.isclose(v0[o], p["opacity"][0].item()) and np.allclose(v0[o + 1:o + 3], p["scaling"][0, :2].numpy())
    assert np.isclose(v0[o + 3], math.log(1e-6)) and np.allclose(v0[o + 4:o + 8], p["quaternion"][0].numpy())
    # load: anchors = xyz, offsets = 0, scale_2 is whatever the file says
    q = io.load_ply_to_gs(path, deg)
    assert torch.allclose(q["anchors"], p["anchors"] + p["offsets"]) and float(q["offsets"].abs().max()) == 0
    assert torch.equal(q["features_dc"], p["features_dc"]) and torch.equal(q["features_rest"], p["features_rest"])
    assert torch.equal(q["opacity"], p["opacity"]) and torch.equal(q["quaternion"], p["quaternion"])
    assert torch.equal(q["scaling"][:, :2], p["scaling"][:, :2]) and torch.allclose(q["scaling"][:, 2], torch.tensor(math.log(1e-6)))
    # degree 0: no f_rest properties at all, features_rest comes back as [N,0,3]
    p0 = _gs(5, 1)
    io.export_gs_to_ply(str(tmp_path / "g0.ply"), **p0)
    q0 = io.load_ply_to_gs(str(tmp_path / "g0.ply"), 0)
    assert q0["features_rest"].shape == (5, 0, 3) and b"f_rest" not in open(tmp_path / "g0.ply", "rb").read()


def test_reads_files_with_other_property_order(tmp_path):
    """A 3DGS-style writer puts normals in and orders differently; lookup is by name."""
    from gssdf_b200 import io
    n = 4
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    data = np.arange(n * len(names), dtype="<f4").reshape(n, len(names))
    with open(tmp_path / "x.ply", "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\ncomment made elsewhere\nelement vertex %d\n" % n +
                 "".join(f"property float {k}\n" for k in names) + "end_header\n").encode())
        f.write(data.tobytes())
    q = io.load_ply_to_gs(str(tmp_path / "x.ply"), 0)
    assert np.array_equal(q["anchors"].numpy(), data[:, :3]) and np.array_equal(q["opacity"].numpy(), data[:, 9])
    assert np.array_equal(q["quaternion"].numpy(), data[:, 13:17])


def test_points_ply_and_pt_yaml(tmp_path):
    from gssdf_b200 import io
    xyz = torch.randn(100, 3)
    io.write_points_ply(str(tmp_path / "as_occ_prior.ply"), xyz)
    assert torch.equal(io.read_points_ply(str(tmp_path / "as_occ_prior.ply")), xyz)
    io.write_pt_params(str(tmp_path / "pt.yaml"), [0.25, -1.5, 3.0], 7.0, "/opt/gs_sdf")
    txt = open(tmp_path / "pt.yaml").read()
    assert txt.startswith("%YAML:1.0\nmap_origin: !!opencv-matrix\n   rows: 1\n   cols: 3\n   dt: f\n   data: [")
    r = io.read_pt_params(str(tmp_path / "pt.yaml"), leaf_size=0.05)
    assert np.allclose(r["map_origin"], [0.25, -1.5, 3.0]) and r["inner_map_size"] == 7.0 and r["package_path"] == "/opt/gs_sdf"
    # params.cpp:474-477: level = ceil(log2((inner + 2 leaf) / leaf)) = ceil(log2(142)) = 8 ; map_size = 256 * 0.05
    assert r["octree_level"] == 8 and abs(r["map_size"] - 12.8) < 1e-9 and r["x_max"] == 3.5
