"""torch front end of oracle/_ref/libref_S{S}.so -- the REFERENCE's own CUDA rasterizer compiled for
sm_100a (oracle/build.py, oracle/ref_shim.cu).  TEST / BASELINE ONLY (GPU box)."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs: dict[int, C.CDLL] = {}
BUILT_CHANNELS = (1, 4, 8, 10, 16, 32)


def available(S: int) -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", f"libref_S{_build_for(S)}.so"))


def _build_for(S: int) -> int:
    return 1 if S == 0 else S


def lib(S: int) -> C.CDLL:
    Sb = _build_for(S)
    if Sb not in _libs:
        path = os.path.join(_HERE, "_ref", f"libref_S{Sb}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: reference build for S={S} not present (oracle/build.py ref)")
        L = C.CDLL(path)
        L.ref_create.restype = C.c_void_p
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_error.restype = C.c_char_p
        L.ref_error.argtypes = [C.c_void_p]
        assert L.ref_sem_channels() == Sb
        _libs[Sb] = L
    return _libs[Sb]


def _p(t):
    return None if t is None or t.numel() == 0 else C.c_void_p(t.data_ptr())


class RefRasterizer:
    """One reference rasterizer context (owns the three scratch buffers like the reference's glue)."""

    def __init__(self, S: int):
        self.S = S
        self.Sb = _build_for(S)
        self.L = lib(S)
        self.ctx = C.c_void_p(self.L.ref_create())

    def __del__(self):
        try:
            self.L.ref_destroy(self.ctx)
        except Exception:
            pass

    def forward(self, *, means3D, opacities, W, H, viewmatrix, projmatrix, campos, tanfovx, tanfovy, bg,
                shs=None, colors_precomp=None, semantics=None, scales=None, rotations=None, cov3D_precomp=None,
                sh_degree=3, scale_modifier=1.0, prefiltered=False, debug=False):
        dev = means3D.device
        P = means3D.shape[0]
        M = shs.shape[1] if shs is not None else 0
        if self.S == 0:
            semantics = torch.zeros((P, 1), device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        # torch.empty: the shim performs the reference glue's zero fills itself
        out_color = torch.empty((3, H, W), **f32)
        out_sem = torch.empty((self.Sb, H, W), **f32)
        out_depth = torch.empty((1, H, W), **f32)
        out_alpha = torch.empty((1, H, W), **f32)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        self._saved = dict(means3D=means3D, shs=shs, colors_precomp=colors_precomp, semantics=semantics,
                           scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp, viewmatrix=viewmatrix,
                           projmatrix=projmatrix, campos=campos, bg=bg, tanfovx=tanfovx, tanfovy=tanfovy,
                           sh_degree=sh_degree, scale_modifier=scale_modifier, W=W, H=H, M=M, P=P,
                           alpha=out_alpha, radii=radii, debug=debug)
        rc = self.L.ref_forward(self.ctx, C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(bg), C.c_int(W), C.c_int(H),
                                _p(means3D), _p(shs), _p(colors_precomp), _p(semantics), _p(opacities), _p(scales),
                                C.c_float(scale_modifier), _p(rotations), _p(cov3D_precomp), _p(viewmatrix),
                                _p(projmatrix), _p(campos), C.c_float(tanfovx), C.c_float(tanfovy),
                                C.c_int(int(prefiltered)), _p(out_color), _p(out_sem), _p(out_depth), _p(out_alpha),
                                _p(radii), C.c_int(int(debug)))
        if rc < 0:
            raise RuntimeError(f"ref_forward failed: {self.L.ref_error(self.ctx).decode()}")
        self.num_rendered = rc
        if self.S == 0:
            out_sem = out_sem[:0]
        return dict(color=out_color, semantics=out_sem, depth=out_depth, alpha=out_alpha, radii=radii)

    def backward(self, dL_dcolor, dL_dsemantics, dL_ddepth, dL_dalpha):
        s = self._saved
        dev = s["means3D"].device
        P, M, W, H = s["P"], s["M"], s["W"], s["H"]
        f32 = dict(dtype=torch.float32, device=dev)
        if self.S == 0:
            dL_dsemantics = torch.zeros((1, H, W), **f32)
        e = lambda *shape: torch.empty(shape, **f32)
        g = dict(dL_dmeans2D=e(P, 3), dL_dconic=e(P, 4), dL_dopacity=e(P, 1), dL_dcolors=e(P, 3),
                 dL_dsemantics=e(P, self.Sb), dL_ddepths=e(P, 1), dL_dmeans3D=e(P, 3), dL_dcov3D=e(P, 6),
                 dL_dsh=e(P, M, 3), dL_dscales=e(P, 3), dL_drotations=e(P, 4))
        rc = self.L.ref_backward(self.ctx, C.c_int(P), C.c_int(s["sh_degree"]), C.c_int(M), _p(s["bg"]), C.c_int(W),
                                 C.c_int(H), _p(s["means3D"]), _p(s["shs"]), _p(s["colors_precomp"]),
                                 _p(s["semantics"]), _p(s["alpha"]), _p(s["scales"]), C.c_float(s["scale_modifier"]),
                                 _p(s["rotations"]), _p(s["cov3D_precomp"]), _p(s["viewmatrix"]), _p(s["projmatrix"]),
                                 _p(s["campos"]), C.c_float(s["tanfovx"]), C.c_float(s["tanfovy"]), _p(s["radii"]),
                                 _p(dL_dcolor), _p(dL_dsemantics), _p(dL_ddepth), _p(dL_dalpha),
                                 _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolors"]),
                                 _p(g["dL_dsemantics"]), _p(g["dL_ddepths"]), _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]),
                                 _p(g["dL_dsh"]), _p(g["dL_dscales"]), _p(g["dL_drotations"]), C.c_int(int(s["debug"])))
        if rc < 0:
            raise RuntimeError(f"ref_backward failed: {self.L.ref_error(self.ctx).decode()}")
        if self.S == 0:
            g["dL_dsemantics"] = g["dL_dsemantics"][:, :0]
        return g
