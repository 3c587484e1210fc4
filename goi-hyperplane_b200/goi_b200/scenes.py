"""Seeded synthetic scenes and cameras -- the inputs of bench.py and the tests.

Everything is generated on the CPU with ``torch.Generator(seed)`` (device independent) and then moved.
Camera conventions follow the reference: ``world_view_transform`` and ``full_proj_transform`` are the
TRANSPOSED (row-vector) matrices of scene/cameras.py:45-47, the projection is
utils/graphics_utils.py:51-71 (getProjectionMatrix), world-to-view is :38-49 (getWorld2View2).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> torch.Tensor:
    """utils/graphics_utils.py:51-71."""
    tan_y, tan_x = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = tan_y * znear, tan_x * znear
    bottom, left = -top, -right
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


class SyntheticCamera:
    """Duck-types scene/cameras.py:Camera / MiniCam for gaussian_renderer.render."""

    def __init__(self, width: int, height: int, fovx: float, world_to_view: torch.Tensor | None = None,
                 znear: float = 0.01, zfar: float = 100.0, device="cpu"):
        self.image_width, self.image_height = int(width), int(height)
        self.FoVx = float(fovx)
        tanfovx = math.tan(fovx * 0.5)
        self.FoVy = 2.0 * math.atan(tanfovx * height / width)        # tanfovy = tanfovx * H / W
        self.znear, self.zfar = znear, zfar
        w2v = torch.eye(4) if world_to_view is None else world_to_view.float()
        self.world_view_transform = w2v.transpose(0, 1).contiguous().to(device)
        proj = projection_matrix(znear, zfar, self.FoVx, self.FoVy).transpose(0, 1)
        self.projection_matrix = proj.contiguous().to(device)
        self.full_proj_transform = (self.world_view_transform.unsqueeze(0).bmm(
            self.projection_matrix.unsqueeze(0))).squeeze(0).contiguous()
        self.camera_center = self.world_view_transform.inverse()[3, :3].contiguous()

    def to(self, device):
        import copy
        c = copy.copy(self)
        for k in ("world_view_transform", "projection_matrix", "full_proj_transform", "camera_center"):
            setattr(c, k, getattr(self, k).to(device))
        return c


def look_at_world_to_view(eye, target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0)) -> torch.Tensor:
    """World-to-view (camera looks down +z, COLMAP-style) for a camera at `eye` looking at `target`."""
    eye = torch.tensor(eye, dtype=torch.float64)
    fwd = torch.tensor(target, dtype=torch.float64) - eye
    fwd = fwd / fwd.norm()
    upv = torch.tensor(up, dtype=torch.float64)
    right = torch.linalg.cross(upv, fwd)
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    R = torch.stack([right, down, fwd], dim=0)               # rows = camera axes in world coordinates
    M = torch.eye(4, dtype=torch.float64)
    M[:3, :3] = R
    M[:3, 3] = -R @ eye
    return M.float()


class SyntheticGaussians:
    """Activated per-Gaussian tensors with the getter names of scene/gaussian_model.py:90-117."""

    def __init__(self, xyz, opacity, scaling, rotation, features, semantics, sh_degree=3):
        self._xyz, self._opacity, self._scaling, self._rotation = xyz, opacity, scaling, rotation
        self._features, self._semantics = features, semantics
        self.active_sh_degree = sh_degree
        self.max_sh_degree = 3

    get_xyz = property(lambda s: s._xyz)
    get_opacity = property(lambda s: s._opacity)
    get_scaling = property(lambda s: s._scaling)
    get_rotation = property(lambda s: s._rotation)
    get_features = property(lambda s: s._features)
    get_semantics = property(lambda s: s._semantics)

    def tensors(self):
        return [self._xyz, self._opacity, self._scaling, self._rotation, self._features, self._semantics]

    def to(self, device):
        return SyntheticGaussians(*[t.to(device) if t is not None else None for t in self.tensors()],
                                  sh_degree=self.active_sh_degree)

    def requires_grad_(self, flag=True):
        for t in self.tensors():
            if t is not None:
                t.requires_grad_(flag)
        return self

    def get_covariance(self, scaling_modifier=1.0):
        """scene/gaussian_model.py:16-20 (build_covariance_from_scaling_rotation -> strip_symmetric)."""
        q = self._rotation / self._rotation.norm(dim=1, keepdim=True)
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                         2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                         2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).view(-1, 3, 3)
        L = R * (scaling_modifier * self._scaling).unsqueeze(1)
        cov = L @ L.transpose(1, 2)
        return torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], dim=1)


@dataclass
class PipeFlags:
    """arguments/__init__.py:56-61 PipelineParams."""
    convert_SHs_python: bool = False
    compute_cov3D_python: bool = False
    debug: bool = False


def make_scene(P: int, W: int, H: int, S: int, seed: int, device="cpu", fovx_deg: float = 60.0,
               px_sigma: float = 2.0):
    """Frustum-filling scene for a camera at the origin looking down +z.

    Returns (gaussians, camera, bg).  ~`px_sigma` pixel standard deviation per Gaussian."""
    gen = torch.Generator().manual_seed(seed)
    fovx = math.radians(fovx_deg)
    tanfovx = math.tan(fovx / 2)
    tanfovy = tanfovx * H / W
    fx = W / (2 * tanfovx)
    u = lambda *shape: torch.rand(*shape, generator=gen)
    n = lambda *shape: torch.randn(*shape, generator=gen)
    z = 1.0 + 9.0 * u(P)
    x = z * tanfovx * (u(P) * 2.2 - 1.1)
    y = z * tanfovy * (u(P) * 2.2 - 1.1)
    xyz = torch.stack([x, y, z], dim=1)
    scaling = (z * (px_sigma / fx)).unsqueeze(1) * torch.exp(0.5 * n(P, 3))
    q = n(P, 4)
    rotation = q / q.norm(dim=1, keepdim=True)
    opacity = 0.02 + 0.96 * u(P, 1)
    features = torch.cat([0.5 * n(P, 1, 3), 0.1 * n(P, 15, 3)], dim=1)
    semantics = n(P, S) if S > 0 else None
    cam = SyntheticCamera(W, H, fovx, device=device)
    g = SyntheticGaussians(xyz, opacity, scaling, rotation, features.contiguous(), semantics).to(device)
    bg = torch.zeros(3, device=device)
    return g, cam, bg


def make_orbit_scene(P: int, W: int, H: int, S: int, n_views: int, seed: int, device="cpu", radius: float = 5.0,
                     fovx_deg: float = 60.0, px_sigma: float = 2.0):
    """Object-centric scene: Gaussians in a unit-ish ball, `n_views` cameras on a circle of radius 5
    looking at the origin."""
    gen = torch.Generator().manual_seed(seed)
    fovx = math.radians(fovx_deg)
    fx = W / (2 * math.tan(fovx / 2))
    n = lambda *shape: torch.randn(*shape, generator=gen)
    u = lambda *shape: torch.rand(*shape, generator=gen)
    d = n(P, 3)
    xyz = d / d.norm(dim=1, keepdim=True) * (u(P, 1) ** (1.0 / 3.0)) * 1.6
    scaling = (radius * (px_sigma / fx)) * torch.exp(0.5 * n(P, 3))
    q = n(P, 4)
    rotation = q / q.norm(dim=1, keepdim=True)
    opacity = 0.02 + 0.96 * u(P, 1)
    features = torch.cat([0.5 * n(P, 1, 3), 0.1 * n(P, 15, 3)], dim=1)
    semantics = n(P, S) if S > 0 else None
    cams = []
    for v in range(n_views):
        a = 2 * math.pi * v / n_views
        eye = (radius * math.cos(a), 0.6 * math.sin(2 * a), radius * math.sin(a))
        cams.append(SyntheticCamera(W, H, fovx, look_at_world_to_view(eye), device=device))
    g = SyntheticGaussians(xyz, opacity, scaling, rotation, features.contiguous(), semantics).to(device)
    return g, cams, torch.zeros(3, device=device)


def make_loss_weights(S: int, W: int, H: int, seed: int, device="cpu"):
    """Fixed linear pseudo-loss L = sum(w * out) over all four outputs."""
    gen = torch.Generator().manual_seed(seed + 7919)
    r = lambda *shape: (torch.rand(*shape, generator=gen) * 2 - 1).to(device)
    return {"render": r(3, H, W), "semantics": r(S, H, W), "depth": r(1, H, W), "alpha": r(1, H, W)}


def make_mask_model(S: int, K: int = 300, D: int = 256, seed: int = 0, device="cpu"):
    """Random stand-ins for semantic_MLP.pt / LUT.pt / the text hyperplane (train.py:64-66)."""
    gen = torch.Generator().manual_seed(seed + 104729)
    bound = math.sqrt(6.0 / (S + K))                         # xavier_uniform_, scene/semantic_model.py:41
    mlp_w = ((torch.rand(K, S, generator=gen) * 2 - 1) * bound).to(device)
    mlp_b = ((torch.rand(K, generator=gen) * 2 - 1) / math.sqrt(S)).to(device)
    lut = torch.randn(K, D, generator=gen)
    lut = (lut / lut.norm(dim=1, keepdim=True) * (0.5 + torch.rand(K, 1, generator=gen))).to(device)
    w = torch.randn(1, D, generator=gen)
    w = (w / w.norm() * 3.0).to(device)
    return mlp_w, mlp_b, lut, w
