"""Parameter container + PLY codec for trained GOI scenes.

``GaussianCloud`` holds the STORED parameters of the reference's ``GaussianModel``
(scene/gaussian_model.py:33-51: ``_xyz, _features_dc, _features_rest, _semantics, _scaling, _rotation,
_opacity``) and exposes the same activated getters (:90-117) so ``gaussian_renderer.render`` takes it
unchanged -- with ``fused_activations=True`` the stored tensors go to the rasterizer directly.

``read_ply`` / ``write_ply`` speak the reference's vertex schema (``construct_list_of_attributes`` :255-270,
``save_ply`` :272-289, ``load_ply`` :307-358):

    x y z nx ny nz  f_dc_0..2  f_rest_0..(3((D+1)^2-1)-1)  sem_0..S-1  opacity  scale_0..2  rot_0..3   (all float32)

without the ``plyfile`` dependency (binary_little_endian, binary_big_endian and ascii; extra elements and
non-float properties are skipped).  f_dc / f_rest are stored channel-major ([P,3,K] flattened, :277-278).
"""
from __future__ import annotations

import os

import numpy as np
import torch

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
              "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
              "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def read_ply_vertices(path: str) -> dict:
    """{property name: float64/whatever numpy column} of the 'vertex' element."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements, cur = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: unterminated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                cur = {"name": tok[1], "count": int(tok[2]), "props": []}
                elements.append(cur)
            elif tok[0] == "property":
                if tok[1] == "list":
                    cur["props"].append((tok[4], "list", tok[2], tok[3]))
                else:
                    if tok[1] not in _PLY_TYPES:
                        raise ValueError(f"{path}: unknown PLY type {tok[1]}")
                    cur["props"].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("binary_little_endian", "binary_big_endian", "ascii"):
            raise ValueError(f"{path}: unsupported PLY format {fmt}")
        for el in elements:                       # elements are stored in header order
            has_list = any(p[1] == "list" for p in el["props"])
            if el["name"] != "vertex":
                if has_list or fmt == "ascii":
                    raise ValueError(f"{path}: element '{el['name']}' before 'vertex' cannot be skipped")
                f.seek(el["count"] * sum(np.dtype(p[1]).itemsize for p in el["props"]), os.SEEK_CUR)
                continue
            if has_list:
                raise ValueError(f"{path}: list property in the vertex element")
            if fmt == "ascii":
                rows = np.loadtxt(f, max_rows=el["count"], ndmin=2, dtype=np.float64)
                if rows.shape[0] != el["count"]:
                    raise ValueError(f"{path}: truncated vertex data")
                return {p[0]: rows[:, i] for i, p in enumerate(el["props"])}
            end = "<" if fmt == "binary_little_endian" else ">"
            dt = np.dtype([(p[0], end + p[1]) for p in el["props"]])
            data = np.fromfile(f, dtype=dt, count=el["count"])
            if data.shape[0] != el["count"]:
                raise ValueError(f"{path}: truncated vertex data")
            return {name: data[name] for name in dt.names}
    raise ValueError(f"{path}: no vertex element")


def write_ply_vertices(path: str, columns: list) -> None:
    """columns = [(name, float array [P])...] -> binary_little_endian PLY, all float32 ('f4' like save_ply)."""
    P = len(columns[0][1]) if columns else 0
    dt = np.dtype([(n, "<f4") for n, _ in columns])
    arr = np.empty(P, dtype=dt)
    for n, c in columns:
        arr[n] = np.asarray(c, dtype=np.float32)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\n")
        f.write(f"element vertex {P}\n".encode())
        for n, _ in columns:
            f.write(f"property float {n}\n".encode())
        f.write(b"end_header\n")
        arr.tofile(f)


def _sorted_cols(cols: dict, prefix: str):
    names = sorted((n for n in cols if n.startswith(prefix)), key=lambda n: int(n.split("_")[-1]))
    return names


def inverse_sigmoid(x):
    return torch.log(x / (1 - x))


class GaussianCloud:
    """Stored (pre-activation) parameters with the reference's getter names (scene/gaussian_model.py:90-117)."""

    def __init__(self, sh_degree: int = 3, semantic_dim: int = 10):
        self.active_sh_degree = 0
        self.max_sh_degree = sh_degree
        self.semantic_dim = semantic_dim
        e = torch.empty(0)
        self._xyz = self._features_dc = self._features_rest = self._semantics = e
        self._scaling = self._rotation = self._opacity = e
        self._semantics_masks = None

    # ---- activated views (what render() marshals in the reference call form) ----
    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_scaling(self):
        return torch.exp(self._scaling)

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self._rotation)

    @property
    def get_features(self):
        return torch.cat((self._features_dc, self._features_rest), dim=1)

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    @property
    def get_semantics(self):
        if self._semantics_masks is None:
            return self._semantics
        return self._semantics * self._semantics_masks

    def set_semantic_masks(self, masks=None):
        self._semantics_masks = None if masks is None else masks.unsqueeze(1)

    def get_covariance(self, scaling_modifier=1.0):
        """build_covariance_from_scaling_rotation + strip_symmetric (scene/gaussian_model.py:16-20)."""
        q = self.get_rotation
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                         2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                         2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).view(-1, 3, 3)
        L = R * (scaling_modifier * self.get_scaling).unsqueeze(1)
        cov = L @ L.transpose(1, 2)
        return torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], dim=1)

    def parameters(self):
        return {"xyz": self._xyz, "f_dc": self._features_dc, "f_rest": self._features_rest,
                "semantics": self._semantics, "opacity": self._opacity, "scaling": self._scaling,
                "rotation": self._rotation}

    def to(self, device):
        for k in ("_xyz", "_features_dc", "_features_rest", "_semantics", "_scaling", "_rotation", "_opacity"):
            setattr(self, k, getattr(self, k).detach().to(device))
        return self

    def requires_grad_(self, flag=True):
        for t in self.parameters().values():
            t.requires_grad_(flag)
        return self

    # ---- construction ----
    @classmethod
    def from_activated(cls, xyz, opacity, scaling, rotation, features, semantics, sh_degree=3, rotation_gain=None):
        """Invert the activations of an activated set (e.g. goi_b200.scenes.make_scene).  `rotation_gain`
        ([P] positive) scales the stored quaternions so that normalisation is actually exercised."""
        c = cls(sh_degree, 0 if semantics is None else semantics.shape[1])
        c._xyz = xyz.clone()
        c._opacity = inverse_sigmoid(opacity.clone())
        c._scaling = torch.log(scaling)
        c._rotation = rotation.clone() if rotation_gain is None else rotation * rotation_gain.view(-1, 1)
        c._features_dc = features[:, :1, :].contiguous()
        c._features_rest = features[:, 1:, :].contiguous()
        c._semantics = torch.zeros(xyz.shape[0], 0) if semantics is None else semantics.clone()
        c.active_sh_degree = sh_degree
        return c

    # ---- PLY (reference schema) ----
    def construct_list_of_attributes(self):
        names = ["x", "y", "z", "nx", "ny", "nz"]
        names += [f"f_dc_{i}" for i in range(self._features_dc.shape[1] * self._features_dc.shape[2])]
        names += [f"f_rest_{i}" for i in range(self._features_rest.shape[1] * self._features_rest.shape[2])]
        names += [f"sem_{i}" for i in range(self._semantics.shape[1])]
        names.append("opacity")
        names += [f"scale_{i}" for i in range(self._scaling.shape[1])]
        names += [f"rot_{i}" for i in range(self._rotation.shape[1])]
        return names

    def save_ply(self, path: str) -> None:
        n = lambda t: t.detach().cpu().numpy()
        xyz = n(self._xyz)
        f_dc = n(self._features_dc.detach().transpose(1, 2).flatten(start_dim=1).contiguous())
        f_rest = n(self._features_rest.detach().transpose(1, 2).flatten(start_dim=1).contiguous())
        attributes = np.concatenate((xyz, np.zeros_like(xyz), f_dc, f_rest, n(self._semantics), n(self._opacity),
                                     n(self._scaling), n(self._rotation)), axis=1)
        names = self.construct_list_of_attributes()
        assert attributes.shape[1] == len(names)
        write_ply_vertices(path, [(name, attributes[:, i]) for i, name in enumerate(names)])

    def load_ply(self, path: str, device="cpu"):
        cols = read_ply_vertices(path)
        P = len(cols["x"])
        xyz = np.stack((cols["x"], cols["y"], cols["z"]), axis=1)
        opacities = np.asarray(cols["opacity"])[..., np.newaxis]
        features_dc = np.stack([cols["f_dc_0"], cols["f_dc_1"], cols["f_dc_2"]], axis=1)[:, :, np.newaxis]
        extra = _sorted_cols(cols, "f_rest_")
        if len(extra) != 3 * (self.max_sh_degree + 1) ** 2 - 3:
            raise ValueError(f"{path}: {len(extra)} f_rest_* properties, expected "
                             f"{3 * (self.max_sh_degree + 1) ** 2 - 3} for SH degree {self.max_sh_degree}")
        features_extra = np.stack([cols[nm] for nm in extra], axis=1) if extra else np.zeros((P, 0))
        features_extra = features_extra.reshape((P, 3, (self.max_sh_degree + 1) ** 2 - 1))
        sem_names = _sorted_cols(cols, "sem_")
        sems = np.zeros((P, len(sem_names) or self.semantic_dim))
        if len(sem_names) == self.semantic_dim:                     # load_ply :331-335: otherwise zeros
            for i, nm in enumerate(sem_names):
                sems[:, i] = cols[nm]
        scales = np.stack([cols[nm] for nm in _sorted_cols(cols, "scale_")], axis=1)
        rots = np.stack([cols[nm] for nm in _sorted_cols(cols, "rot_")], axis=1)
        t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float, device=device)
        self._xyz = t(xyz)
        self._features_dc = t(features_dc).transpose(1, 2).contiguous()
        self._features_rest = t(features_extra).transpose(1, 2).contiguous()
        self._opacity, self._scaling, self._rotation = t(opacities), t(scales), t(rots)
        self._semantics = t(sems)
        self.active_sh_degree = self.max_sh_degree
        return self
