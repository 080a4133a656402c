"""Host-side mirror of the reference's OctreeAS (submodules/kaolin_wisp_cpp/kaolin_wisp_cpp/octree_as/octree_as.{h,cpp}) and of
NeuralSLAM::sample (include/neural_mapping/neural_mapping.cpp:73-104) over the C ABI: same method names and argument meaning;
data-dependent output sizes are read back ONCE per call here (the exact-shape API, like the reference's .item() calls) while
`RaySampler` keeps everything capacity-sized on the device for the training step."""
import ctypes as C

import numpy as np
import torch

from . import _lib, cabi
from ._lib import check, lib, make_args


def quantize_points(x, level):
    """spc_ops::quantize_points (spc_ops.cpp:6-15): torch float [-1,1] -> int16 [0, 2^level - 1]."""
    res = 2 ** level
    return torch.floor(torch.clamp(res * (x + 1.0) / 2.0, 0, res - 1)).to(torch.int16)


class OctreeAS:
    def __init__(self, octree, exsum, points, pyramid, level, device, origin=(0.0, 0.0, 0.0), map_size=0.0):
        self.max_level_ = level
        self.octree_h, self.exsum_h, self.points_h, self.pyramid_ = octree, exsum, points, pyramid
        self.device = device
        self.octree_ = torch.from_numpy(octree).to(device)
        self.prefix_ = torch.from_numpy(exsum).to(device)
        self.points_ = torch.from_numpy(points).to(device)
        self.origin, self.map_size = tuple(float(v) for v in origin), float(map_size)
        self.n_nodes = len(octree)
        self.ws = cabi.Workspace(device)

    @staticmethod
    def from_quantized_points(qpts, level, device, origin=(0.0, 0.0, 0.0), map_size=0.0):
        """from_quantized_points (octree_as.cpp:27-31): int16 [n,3] (any device) -> acceleration structure on `device`.
        map_size > 0: coordinates given to query / raytrace / RaySampler are WORLD points of a SubMap centred at `origin`."""
        q = np.ascontiguousarray(qpts.detach().cpu().numpy() if hasattr(qpts, "detach") else qpts, np.int16).reshape(-1, 3)
        qp = q.ctypes.data if len(q) else None
        a = make_args("gssdf_octree_build_args", n=len(q), qpoints=qp, level=level)
        check(lib().gssdf_octree_build_host(C.byref(a)))
        nn, npnt = int(a.n_nodes), int(a.n_points)
        octree, exsum = np.zeros(max(nn, 1), np.uint8), np.zeros(nn + 1, np.int32)
        points, pyramid = np.zeros((max(npnt, 1), 3), np.int16), np.zeros((2, level + 2), np.int32)
        a = make_args("gssdf_octree_build_args", n=len(q), qpoints=qp, level=level, node_cap=max(nn, 1), point_cap=max(npnt, 1),
                      octree=octree.ctypes.data, exsum=exsum.ctypes.data, points=points.ctypes.data, pyramid=pyramid.ctypes.data)
        check(lib().gssdf_octree_build_host(C.byref(a)))
        t = OctreeAS(octree, exsum, points, pyramid, level, device, origin, map_size)
        t.n_nodes = nn
        return t

    def tree_struct(self):
        t = _lib.STRUCTS["gssdf_octree"]()
        t.level, t.n_nodes = self.max_level_, self.n_nodes
        t.octree, t.exsum = self.octree_.data_ptr(), self.prefix_.data_ptr()
        t.origin = (C.c_float * 3)(*self.origin)
        t.inv_size = 1.0 / self.map_size if self.map_size > 0 else 0.0
        t.size = self.map_size
        return t

    def query(self, coords, n_live=None, valid_out=None):
        """OctreeAS::query at the leaf level: pidx [n] int32 (-1 = not occupied)."""
        n = coords.shape[0]
        pidx = torch.empty(n, dtype=torch.int32, device=coords.device)
        a = make_args("gssdf_octree_query_args", n=n, coords=coords, n_live=n_live, pidx=pidx, valid=valid_out)
        a.tree = self.tree_struct()
        check(lib().gssdf_octree_query(C.byref(a), cabi._stream()))
        return pidx

    def valid_mask(self, coords, out, n_live=None):
        """SubMap::get_valid_mask into a caller-owned uint8 buffer (no allocation, no sync: the training step's path)."""
        a = make_args("gssdf_octree_query_args", n=coords.shape[0], coords=coords, n_live=n_live, valid=out)
        a.tree = self.tree_struct()
        check(lib().gssdf_octree_query(C.byref(a), cabi._stream()))
        return out

    def raytrace(self, origins, dirs, cap=None):
        """OctreeAS::raytrace(level = max, with_exit = True): (ridx, pidx, depth[k,2])."""
        n = origins.shape[0]
        cap = int(cap or max(64 * n, 1024))
        dev = origins.device
        ridx, pidx = torch.empty(cap, dtype=torch.int32, device=dev), torch.empty(cap, dtype=torch.int32, device=dev)
        depth, cnt = torch.empty(cap, 2, device=dev), torch.zeros(2, dtype=torch.int32, device=dev)
        w = self.ws.get(lib().gssdf_octree_raytrace_workspace_bytes(C.c_int64(n)))
        a = make_args("gssdf_octree_raytrace_args", n_rays=n, origins=origins, dirs=dirs, cap=cap, ridx=ridx, pidx=pidx, depth=depth,
                      n_nuggets=cnt, workspace=w, workspace_bytes=w.numel())
        a.tree = self.tree_struct()
        check(lib().gssdf_octree_raytrace(C.byref(a), cabi._stream()))
        k, ovf = cnt.tolist()
        if ovf:
            raise RuntimeError("gssdf_b200: raytrace capacity exceeded")
        return ridx[:k], pidx[:k], depth[:k]


class RaySampler:
    """NeuralSLAM::sample as one asynchronous call with capacity buffers (no host sync): `sample(...)` fills xyz / ray_sdf (/ direction /
    depth / ridx) rows [0, counts[0]) -- feed `counts` as n_live to the SDF kernels."""

    def __init__(self, tree, n_rays, device, voxel_sample_num=1, n_free=4, n_surface=4, sample_std=0.1, truncated_dis=0.3,
                 xyz_min=(-7.0, -7.0, -7.0), xyz_max=(7.0, 7.0, 7.0), nugget_cap=None, cap=None, keep_aux=False):
        self.tree, self.n, self.ns, self.n_free, self.n_surf = tree, n_rays, voxel_sample_num, n_free, n_surface
        self.std, self.trunc = sample_std, truncated_dis
        f = np.float32
        self.lo = tuple(float(f(v) + f(1e-6)) for v in xyz_min)   # (xyz_min + padding + 1e-6), padding 0 (sub_map.cpp:41-42)
        self.hi = tuple(float(f(v) - f(1e-6)) for v in xyz_max)
        self.nugget_cap = int(nugget_cap or 48 * n_rays)
        self.cap = int(cap or self.nugget_cap * voxel_sample_num + n_rays * (n_free + n_surface + 1))
        e = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device=device)
        self.xyz, self.ray_sdf = e(self.cap, 3), e(self.cap)
        self.direction, self.depth, self.ridx = (e(self.cap, 3), e(self.cap), e(self.cap, dt=torch.int64)) if keep_aux else (None, None, None)
        self.counts = torch.zeros(4, dtype=torch.int32, device=device)
        self.rand_voxel, self.rand_free = e(max(self.nugget_cap * voxel_sample_num, 1)), e(max(n_rays * n_free, 1))
        self.randn_surface = e(max(n_rays * n_surface, 1))
        self.ws = cabi.Workspace(device)
        self.ws.get(lib().gssdf_sdf_sample_rays_workspace_bytes(C.c_int64(n_rays), C.c_int64(self.nugget_cap), voxel_sample_num, n_free, n_surface))

    def draw(self):
        """the reference's torch::rand_like / randn draws (wisp_spc_ops.cpp:92, utils.cpp:341,377)"""
        self.rand_voxel.uniform_()
        self.rand_free.uniform_()
        self.randn_surface.normal_()

    def sample(self, origin, direction, depth, xyz):
        w = self.ws.buf
        a = make_args("gssdf_sdf_sample_rays_args", n_rays=self.n, origin=origin, direction=direction, depth=depth, xyz=xyz,
                      voxel_sample_num=self.ns, n_free=self.n_free, n_surface=self.n_surf, sample_std=self.std, truncated_dis=self.trunc,
                      xyz_min=list(self.lo), xyz_max=list(self.hi), rand_voxel=self.rand_voxel, rand_free=self.rand_free,
                      randn_surface=self.randn_surface, nugget_cap=self.nugget_cap, cap=self.cap, out_xyz=self.xyz, out_ray_sdf=self.ray_sdf,
                      out_direction=self.direction, out_depth=self.depth, out_ridx=self.ridx, counts=self.counts, workspace=w,
                      workspace_bytes=w.numel())
        a.tree = self.tree.tree_struct()
        check(lib().gssdf_sdf_sample_rays(C.byref(a), cabi._stream()))
        return self.counts
