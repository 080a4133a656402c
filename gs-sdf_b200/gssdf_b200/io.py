"""f-4: the reference's on-disk formats, so that `view` / `render` of a reference build can load what this path trained and vice versa.

  gs.ply                  NeuralGS::export_gs_to_ply / load_ply_to_gs (include/neural_gaussian/neural_gaussian.cpp:928-1188): binary
                          little-endian PLY, one `vertex` element, float32 properties in this order:
                          x y z | f_dc_0..2 | f_rest_0..3(K-1)-1 (channel-major: features_rest.transpose(1,2).flatten(1)) | opacity (logit) |
                          scale_0 scale_1 scale_2 (log; scale_2 = log(1e-6) "to make it compatible with 3DGS") | rot_0..3 (w,x,y,z, raw)
                          x y z = anchors + offsets; loading puts everything into anchors and zeroes the offsets (:1137-1143).
  as_occ_prior.ply        point cloud the occupancy octree is rebuilt from (neural_mapping.cpp:1366-1374; property names x y z)
  pt.yaml                 write_pt_params / read_pt_params (include/params/params.cpp:443-481): OpenCV-FileStorage YAML with map_origin,
                          inner_map_size, package_path; the octree level / map size are re-derived from leaf_size on load
  local_map_checkpoint.pt torch::save(local_map_ptr) (neural_mapping.cpp:1331-1342): a libtorch module archive with the flat tcnn parameter
                          `encoder_local_map` and `decoder.{0,2,4,..}.{weight,bias}`; written / read through the libtorch shim
                          (gssdf_shim.LocalMapReplay.save / .load), because only libtorch can produce its own archive format.
Host-side file I/O only: numpy + torch tensors, no device code."""
import math
import os
import re

import numpy as np
import torch


def _ply_header(n, props, element="vertex"):
    lines = ["ply", "format binary_little_endian 1.0", f"element {element} {n}"] + [f"property float {p}" for p in props] + ["end_header"]
    return ("\n".join(lines) + "\n").encode("ascii")


def gs_ply_properties(n_sh_bases):
    props = ["x", "y", "z"] + [f"f_dc_{i}" for i in range(3)]
    props += [f"f_rest_{i}" for i in range((n_sh_bases - 1) * 3)]
    return props + ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]


def export_gs_to_ply(path, anchors, offsets, features_dc, features_rest, opacity, scaling, quaternion):
    """All arguments are the RAW NeuralGS parameters ([N,3] [N,3] [N,1,3] [N,K-1,3] [N] [N,3] [N,4])."""
    c = lambda t: t.detach().float().cpu()
    xyz = c(anchors) + c(offsets)
    n = xyz.shape[0]
    f_dc = c(features_dc).transpose(1, 2).flatten(1)
    cols = [xyz, f_dc]
    K = 1
    if features_rest is not None and features_rest.numel() > 0:
        K = 1 + features_rest.shape[1]
        cols.append(c(features_rest).transpose(1, 2).flatten(1))
    sc = c(scaling)
    sc = torch.cat([sc[:, :2], torch.full((n, 1), math.log(1e-6))], -1)
    cols += [c(opacity).view(n, 1), sc, c(quaternion)]
    data = torch.cat(cols, 1).contiguous().numpy().astype("<f4")
    props = gs_ply_properties(K)
    assert data.shape[1] == len(props)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(_ply_header(n, props))
        f.write(data.tobytes())
    return n


def _read_ply(path):
    with open(path, "rb") as f:
        raw = f.read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    header = raw[:end].decode("ascii").splitlines()
    if header[0] != "ply" or "binary_little_endian" not in header[1]:
        raise ValueError(f"{path}: only binary little-endian PLY is supported (the reference writes binary, neural_gaussian.cpp:1030)")
    n, props, in_vertex = 0, [], False
    types = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1", "int": "<i4", "int32": "<i4",
             "short": "<i2", "ushort": "<u2", "uint": "<u4"}
    for ln in header:
        t = ln.split()
        if t[:1] == ["element"]:
            in_vertex = t[1] == "vertex"
            if in_vertex:
                n = int(t[2])
        elif t[:1] == ["property"] and in_vertex:
            if t[1] == "list":
                raise ValueError("list properties are not part of the GS-SDF formats")
            props.append((t[2], types[t[1]]))
    arr = np.frombuffer(raw, dtype=np.dtype(props), count=n, offset=end)
    return arr


def load_ply_to_gs(path, sh_degree, device="cpu"):
    """NeuralGS::load_ply_to_gs: returns dict(anchors, offsets (zeros), features_dc [N,1,3], features_rest [N,K-1,3], opacity, scaling,
    quaternion); properties are looked up BY NAME like tinyply does, so 3DGS-written files load too."""
    a = _read_ply(path)
    K = (sh_degree + 1) ** 2
    col = lambda names: torch.from_numpy(np.stack([a[k].astype(np.float32) for k in names], 1)).to(device)
    anchors = col(["x", "y", "z"])
    out = dict(anchors=anchors, offsets=torch.zeros_like(anchors))
    out["features_dc"] = col([f"f_dc_{i}" for i in range(3)]).view(-1, 3, 1).transpose(1, 2).contiguous()
    if sh_degree > 0:
        out["features_rest"] = col([f"f_rest_{i}" for i in range((K - 1) * 3)]).view(-1, 3, K - 1).transpose(1, 2).contiguous()
    else:
        out["features_rest"] = torch.zeros(anchors.shape[0], 0, 3, device=device)
    out["opacity"] = col(["opacity"]).view(-1)
    out["scaling"] = col(["scale_0", "scale_1", "scale_2"])
    out["quaternion"] = col(["rot_0", "rot_1", "rot_2", "rot_3"])
    return out


def write_points_ply(path, xyz):
    """as_occ_prior.ply: x y z float32."""
    d = xyz.detach().float().cpu().contiguous().numpy().astype("<f4")
    with open(path, "wb") as f:
        f.write(_ply_header(d.shape[0], ["x", "y", "z"]))
        f.write(d.tobytes())


def read_points_ply(path, device="cpu"):
    a = _read_ply(path)
    return torch.from_numpy(np.stack([a["x"], a["y"], a["z"]], 1).astype(np.float32)).to(device)


def write_pt_params(path, map_origin, inner_map_size, package_path=""):
    """write_pt_params (params.cpp:443-453); cv::Mat's operator<< prints a 1x3 float row as [a, b, c]."""
    o = [float(v) for v in np.asarray(map_origin, np.float32).reshape(3)]
    fmt = lambda v: repr(np.float32(v).item()) if v != int(v) else str(int(v))
    with open(path, "w") as f:
        f.write("%YAML:1.0\n")
        f.write("map_origin: !!opencv-matrix\n   rows: 1\n   cols: 3\n   dt: f\n   data: [" + ", ".join(fmt(v) for v in o) + "]\n")
        f.write(f"inner_map_size: {inner_map_size:g}\n")
        f.write(f"package_path: {package_path}\n")


def read_pt_params(path, leaf_size):
    """read_pt_params (params.cpp:455-481): values + the quantities the reference re-derives from them."""
    txt = open(path).read()
    m = re.search(r"map_origin:.*?data:\s*\[([^\]]*)\]", txt, flags=re.S)
    origin = np.array([float(v) for v in m.group(1).split(",")], np.float32)
    inner = float(re.search(r"inner_map_size:\s*([-+0-9.eE]+)", txt).group(1))
    pkg = re.search(r"package_path:\s*(.*)", txt)
    level = int(math.ceil(math.log2((inner + 2 * leaf_size) / leaf_size)))
    res = 2 ** level
    return dict(map_origin=origin, inner_map_size=inner, package_path=pkg.group(1).strip() if pkg else "", x_max=0.5 * inner, x_min=-0.5 * inner,
                octree_level=level, map_resolution=res, map_size=res * leaf_size, map_size_inv=1.0 / (res * leaf_size))
