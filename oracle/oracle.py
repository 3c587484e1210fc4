"""numpy front end of the CPU oracle (oracle/goi_oracle.c, built by oracle/build.py).  TEST ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            from . import build
            build.build_oracle()
        L = C.CDLL(path)
        L.oracle_forward.restype = C.c_void_p
        L.oracle_trace.restype = C.c_void_p
        L.oracle_num_rendered.restype = C.c_int64
        L.oracle_num_rendered.argtypes = [C.c_void_p]
        L.oracle_free.argtypes = [C.c_void_p]
        for name, ty in (("oracle_means2D", C.c_float), ("oracle_depths", C.c_float),
                         ("oracle_conic_opacity", C.c_float), ("oracle_rgb", C.c_float), ("oracle_cov3D", C.c_float),
                         ("oracle_point_list", C.c_uint32), ("oracle_ranges", C.c_uint32),
                         ("oracle_n_contrib", C.c_uint32), ("oracle_tiles_touched", C.c_uint32)):
            fn = getattr(L, name)
            fn.restype = C.POINTER(ty)
            fn.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _f(a):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.c_void_p)


class OracleResult:
    """Forward outputs + the opaque oracle state (needed by backward)."""

    def __init__(self, handle, P, W, H, S, outputs):
        self.handle, self.P, self.W, self.H, self.S = handle, P, W, H, S
        self.__dict__.update(outputs)

    @property
    def num_rendered(self):
        return int(lib().oracle_num_rendered(self.handle))

    def _view(self, name, shape, dtype):
        ptr = getattr(lib(), name)(self.handle)
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape, dtype=dtype)
        return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape).copy()

    def state(self):
        P, R = self.P, self.num_rendered
        T = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        return dict(means2D=self._view("oracle_means2D", (P, 2), np.float32),
                    depths=self._view("oracle_depths", (P,), np.float32),
                    conic_opacity=self._view("oracle_conic_opacity", (P, 4), np.float32),
                    rgb=self._view("oracle_rgb", (P, 3), np.float32),
                    cov3D=self._view("oracle_cov3D", (P, 6), np.float32),
                    tiles_touched=self._view("oracle_tiles_touched", (P,), np.uint32),
                    point_list=self._view("oracle_point_list", (R,), np.uint32),
                    ranges=self._view("oracle_ranges", (T, 2), np.uint32),
                    n_contrib=self._view("oracle_n_contrib", (self.H, self.W), np.uint32))

    def free(self):
        if self.handle:
            lib().oracle_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def forward(*, means3D, opacities, W, H, viewmatrix, projmatrix, campos, tanfovx, tanfovy, bg,
            shs=None, colors_precomp=None, semantics=None, scales=None, rotations=None, cov3D_precomp=None,
            sh_degree=3, scale_modifier=1.0):
    """CudaRasterizer::Rasterizer::forward restated on the CPU.  Arrays are numpy float32."""
    L = lib()
    means3D, p_means = _f(means3D)
    P = means3D.shape[0]
    shs, p_shs = _f(shs)
    M = shs.shape[1] if shs is not None else 0
    colors_precomp, p_col = _f(colors_precomp)
    semantics, p_sem = _f(semantics)
    S = semantics.shape[1] if semantics is not None else 0
    opacities, p_opa = _f(np.asarray(opacities).reshape(-1))
    scales, p_sc = _f(scales)
    rotations, p_rot = _f(rotations)
    cov3D_precomp, p_cov = _f(cov3D_precomp)
    viewmatrix, p_view = _f(np.asarray(viewmatrix).reshape(-1))
    projmatrix, p_proj = _f(np.asarray(projmatrix).reshape(-1))
    campos, p_cam = _f(np.asarray(campos).reshape(-1))
    bg, p_bg = _f(np.asarray(bg).reshape(-1))
    out_color = np.zeros((3, H, W), np.float32)
    out_sem = np.zeros((S, H, W), np.float32)
    out_depth = np.zeros((1, H, W), np.float32)
    out_alpha = np.zeros((1, H, W), np.float32)
    radii = np.zeros((P,), np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    h = L.oracle_forward(C.c_int(P), C.c_int(sh_degree), C.c_int(M), C.c_int(S), p_bg, C.c_int(W), C.c_int(H),
                         p_means, p_shs, p_col, p_sem, p_opa, p_sc, C.c_float(scale_modifier), p_rot, p_cov,
                         p_view, p_proj, p_cam, C.c_float(tanfovx), C.c_float(tanfovy),
                         vp(out_color), vp(out_sem), vp(out_depth), vp(out_alpha), vp(radii))
    res = OracleResult(h, P, W, H, S, dict(color=out_color, semantics=out_sem, depth=out_depth, alpha=out_alpha,
                                             radii=radii))
    res._inputs = dict(means3D=means3D, shs=shs, colors_precomp=colors_precomp, semantics=semantics,
                       scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp, viewmatrix=viewmatrix,
                       projmatrix=projmatrix, campos=campos, bg=bg, tanfovx=tanfovx, tanfovy=tanfovy,
                       sh_degree=sh_degree, scale_modifier=scale_modifier, M=M)
    return res


def backward(res: OracleResult, dL_dcolor, dL_dsemantics, dL_ddepth, dL_dalpha, wide=False):
    """CudaRasterizer::Rasterizer::backward restated on the CPU.  wide=True accumulates the
    per-Gaussian sums in double (order-independent reference value)."""
    L = lib()
    i = res._inputs
    P, S, M, W, H = res.P, res.S, i["M"], res.W, res.H
    z = lambda *s: np.zeros(s, np.float32)
    g = dict(dL_dmeans2D=z(P, 3), dL_dconic=z(P, 4), dL_dopacity=z(P, 1), dL_dcolors=z(P, 3),
             dL_dsemantics=z(P, S), dL_ddepths=z(P, 1), dL_dmeans3D=z(P, 3), dL_dcov3D=z(P, 6),
             dL_dsh=z(P, M, 3), dL_dscales=z(P, 3), dL_drotations=z(P, 4))
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    gc, pc = _f(dL_dcolor if dL_dcolor is not None else z(3, H, W))
    gs, ps = _f(dL_dsemantics if dL_dsemantics is not None else z(S, H, W))
    gd, pd = _f(dL_ddepth if dL_ddepth is not None else z(1, H, W))
    ga, pa = _f(dL_dalpha if dL_dalpha is not None else z(1, H, W))
    alphas, palpha = _f(res.alpha)
    L.oracle_backward(C.c_void_p(res.handle), C.c_int(i["sh_degree"]), C.c_int(M), vp(i["bg"]), vp(i["means3D"]),
                      vp(i["shs"]), vp(i["colors_precomp"]), vp(i["semantics"]), palpha, vp(i["scales"]),
                      C.c_float(i["scale_modifier"]), vp(i["rotations"]), vp(i["cov3D_precomp"]),
                      vp(i["viewmatrix"]), vp(i["projmatrix"]), vp(i["campos"]), C.c_float(i["tanfovx"]),
                      C.c_float(i["tanfovy"]), pc, ps, pd, pa,
                      vp(g["dL_dmeans2D"]), vp(g["dL_dconic"]), vp(g["dL_dopacity"]), vp(g["dL_dcolors"]),
                      vp(g["dL_dsemantics"]), vp(g["dL_ddepths"]), vp(g["dL_dmeans3D"]), vp(g["dL_dcov3D"]),
                      vp(g["dL_dsh"]), vp(g["dL_dscales"]) if i["scales"] is not None else None,
                      vp(g["dL_drotations"]) if i["scales"] is not None else None, C.c_int(int(wide)))
    return g


def trace(*, means3D, opacities, W, H, viewmatrix, projmatrix, campos, tanfovx, tanfovy, bg, img_sem,
          shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None, sh_degree=3,
          scale_modifier=1.0, count_per_channel=True):
    L = lib()
    means3D, p_means = _f(means3D)
    P = means3D.shape[0]
    shs, p_shs = _f(shs)
    M = shs.shape[1] if shs is not None else 0
    colors_precomp, p_col = _f(colors_precomp)
    img_sem, p_img = _f(img_sem)
    S = img_sem.shape[0]
    opacities, p_opa = _f(np.asarray(opacities).reshape(-1))
    scales, p_sc = _f(scales)
    rotations, p_rot = _f(rotations)
    cov3D_precomp, p_cov = _f(cov3D_precomp)
    viewmatrix, p_view = _f(np.asarray(viewmatrix).reshape(-1))
    projmatrix, p_proj = _f(np.asarray(projmatrix).reshape(-1))
    campos, p_cam = _f(np.asarray(campos).reshape(-1))
    bg, p_bg = _f(np.asarray(bg).reshape(-1))
    out_color = np.zeros((3, H, W), np.float32)
    gau_sem = np.zeros((P, S), np.float32)
    num_gsem = np.zeros((P,), np.int32)
    radii = np.zeros((P,), np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    h = L.oracle_trace(C.c_int(P), C.c_int(sh_degree), C.c_int(M), C.c_int(S), p_bg, C.c_int(W), C.c_int(H),
                       p_means, p_shs, p_col, p_img, p_opa, p_sc, C.c_float(scale_modifier), p_rot, p_cov,
                       p_view, p_proj, p_cam, C.c_float(tanfovx), C.c_float(tanfovy),
                       vp(out_color), vp(gau_sem), vp(num_gsem), vp(radii), C.c_int(int(count_per_channel)))
    L.oracle_free(C.c_void_p(h))
    return dict(color=out_color, gau_sem=gau_sem, num_gsem=num_gsem, radii=radii)


def mark_visible(means3D, viewmatrix, projmatrix):
    L = lib()
    means3D, pm = _f(means3D)
    v, pv = _f(np.asarray(viewmatrix).reshape(-1))
    p, pp = _f(np.asarray(projmatrix).reshape(-1))
    out = np.zeros((means3D.shape[0],), np.uint8)
    L.oracle_mark_visible(C.c_int(means3D.shape[0]), pm, pv, pp, out.ctypes.data_as(C.c_void_p))
    return out.astype(bool)


def mask(x, mlp_weight, mlp_bias, lut, w, *, mode=0, hyperplane_b=0.0, log_scale=0.0, thresh=0.86):
    """x: [N,S] numpy.  Returns dict(sim, bg_mask, idx, top2_gap, sim_table)."""
    L = lib()
    x, px = _f(x)
    N, S = x.shape
    mlp_weight, pw = _f(mlp_weight)
    mlp_bias, pb = _f(mlp_bias)
    lut, pl = _f(lut)
    K, D = lut.shape
    w, pwv = _f(np.asarray(w).reshape(-1))
    table = np.zeros((K,), np.float32)
    sim = np.zeros((N,), np.float32)
    bg = np.zeros((N,), np.uint8)
    idx = np.zeros((N,), np.int32)
    gap = np.zeros((N,), np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    L.oracle_mask(C.c_int64(N), C.c_int(S), C.c_int(K), C.c_int(D), C.c_int(mode), C.c_int64(S), C.c_int64(1),
                  px, pw, pb, pl, pwv, C.c_float(hyperplane_b), C.c_float(log_scale), C.c_float(thresh),
                  vp(table), vp(sim), vp(bg), vp(idx), vp(gap))
    return dict(sim=sim, bg_mask=bg.astype(bool), idx=idx, top2_gap=gap, sim_table=table)
