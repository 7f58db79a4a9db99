"""Seeded synthetic splat scenes for the BASELINE.json configs and the
62-float `RichPoint` .ply format GSRast loads.

The reference ships no data (it expects a user-supplied data.ply, GSRastWindow.cpp:24), so
the generator is ours.  Attributes are produced *pre-activation* in the PLY convention and
activated exactly as SplatData::loadFromPly does (apps/gsrast/SplatData.cpp:48-58):
scale = exp(log_scale), rotation = normalize(q), opacity = sigmoid(logit).

Layouts returned by `SplatScene` follow the CudaRasterizer contract
(float3 means, float3 scales, float4 rot (r,x,y,z), float opacity, float[16][3] SH);
`SplatScene.gsrast_layout()` gives the vec4-padded / raw-PLY-order buffers the in-tree
viewer uploads instead (GSGaussians.cpp:121-134, SplatData.hpp:59-69).
"""
from __future__ import annotations

import dataclasses

import numpy as np

# name -> (P, W, H, seed, log-scale mean, opacity-logit (mean, std), colours precomputed)
CONFIGS = {
    "C1": dict(P=100_000, W=1280, H=720, seed=1, mu=-4.3, op=(1.0, 2.0), precomp=False),
    "C2": dict(P=3_300_000, W=1920, H=1080, seed=2, mu=-4.9, op=(1.0, 2.0), precomp=False),
    "C3": dict(P=6_000_000, W=3840, H=2160, seed=3, mu=-5.5, op=(1.0, 2.0), precomp=False),
    "C4": dict(P=3_300_000, W=1920, H=1080, seed=2, mu=-4.9, op=(1.0, 2.0), precomp=False),
    "C5": dict(P=2_000_000, W=1920, H=1080, seed=5, mu=-4.2, op=(-2.5, 0.5), precomp=True),
}


@dataclasses.dataclass
class SplatScene:
    means3D: np.ndarray        # float32 [P,3]
    scales: np.ndarray         # float32 [P,3]  (activated)
    rotations: np.ndarray      # float32 [P,4]  (r,x,y,z), normalised
    opacities: np.ndarray      # float32 [P]    (activated)
    shs: np.ndarray | None     # float32 [P,16,3] or None when colours are precomputed
    colors_precomp: np.ndarray | None  # float32 [P,3] or None
    sh_degree: int = 3
    max_coeffs: int = 16

    @property
    def P(self) -> int:
        return int(self.means3D.shape[0])

    def span(self) -> float:
        ext = self.means3D.max(axis=0) - self.means3D.min(axis=0)
        return float(ext.max())

    def gsrast_layout(self):
        """Buffers as the in-tree viewer uploads them: vec4 means (w=1), vec4 scales (w=e^1),
        vec4 rotations, opacities, SH in raw PLY order (f_dc[3], f_rest[45])."""
        P = self.P
        means4 = np.ones((P, 4), dtype=np.float32)
        means4[:, :3] = self.means3D
        scales4 = np.full((P, 4), np.float32(np.exp(np.float32(1.0))), dtype=np.float32)  # SplatData.cpp:51
        scales4[:, :3] = self.scales
        if self.shs is not None:
            shs_raw = contract_sh_to_ply_order(self.shs)
        else:
            shs_raw = np.zeros((P, 48), dtype=np.float32)
        return means4, scales4, self.rotations.copy(), self.opacities.copy(), shs_raw


def contract_sh_to_ply_order(shs: np.ndarray) -> np.ndarray:
    """[P,16,3] coefficient-major RGB-interleaved -> PLY order: f_dc_0..2, then f_rest with
    channel-major blocks (f_rest[c*15 + (k-1)] = sh[k][c])."""
    P = shs.shape[0]
    out = np.empty((P, 48), dtype=np.float32)
    out[:, 0:3] = shs[:, 0, :]
    out[:, 3:48] = np.transpose(shs[:, 1:, :], (0, 2, 1)).reshape(P, 45)
    return out


def ply_order_to_contract_sh(raw: np.ndarray) -> np.ndarray:
    """Inverse of contract_sh_to_ply_order."""
    P = raw.shape[0]
    shs = np.empty((P, 16, 3), dtype=np.float32)
    shs[:, 0, :] = raw[:, 0:3]
    shs[:, 1:, :] = np.transpose(raw[:, 3:48].reshape(P, 3, 15), (0, 2, 1))
    return shs


def _sigmoid(x):
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(-x.astype(np.float32)))).astype(np.float32)


def make_scene(P: int, seed: int, mu: float = -4.9, op=(1.0, 2.0), precomp: bool = False,
               scale_sigma: float = 0.6) -> SplatScene:
    rng = np.random.default_rng(seed)
    n_shell = int(0.7 * P)
    n_ball = P - n_shell
    # 70 % on noisy concentric shells r in {1,2,3} * U(0.95,1.05)
    d = rng.normal(size=(n_shell, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = rng.integers(1, 4, size=n_shell) * rng.uniform(0.95, 1.05, size=n_shell)
    shell = d * r[:, None]
    # 30 % uniform in the ball r < 3.5
    d2 = rng.normal(size=(n_ball, 3))
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    r2 = 3.5 * rng.uniform(0.0, 1.0, size=n_ball) ** (1.0 / 3.0)
    ball = d2 * r2[:, None]
    means = np.concatenate([shell, ball], axis=0)
    means = means[rng.permutation(P)].astype(np.float32)

    log_scales = rng.normal(mu, scale_sigma, size=(P, 3)).astype(np.float32)
    scales = np.exp(log_scales).astype(np.float32)
    q = rng.normal(size=(P, 4)).astype(np.float32)
    q /= np.sqrt((q * q).sum(axis=1, keepdims=True)).astype(np.float32)
    opac = _sigmoid(rng.normal(op[0], op[1], size=P).astype(np.float32))
    if precomp:
        colors = rng.uniform(0.0, 1.0, size=(P, 3)).astype(np.float32)
        return SplatScene(means, scales, q.astype(np.float32), opac, None, colors, sh_degree=0, max_coeffs=0)
    shs = np.empty((P, 16, 3), dtype=np.float32)
    shs[:, 0, :] = rng.normal(0.0, 1.0, size=(P, 3))
    for k in range(1, 16):
        ell = 1 if k < 4 else (2 if k < 9 else 3)
        shs[:, k, :] = rng.normal(0.0, 0.3 / ell, size=(P, 3))
    return SplatScene(means, scales, q.astype(np.float32), opac, shs, None, sh_degree=3, max_coeffs=16)


def make_config_scene(name: str, P: int | None = None) -> tuple[SplatScene, dict]:
    """Scene for a BASELINE.json config; P may be overridden for scaled-down parity tests
    (the scale law compensates so the per-tile load stays comparable)."""
    cfg = dict(CONFIGS[name])
    mu = cfg["mu"]
    if P is not None and P != cfg["P"]:
        # keep expected screen coverage constant: area ~ P * s^2  =>  s ~ P^-1/2
        mu = mu + 0.5 * float(np.log(cfg["P"] / P))
        cfg["P"] = P
    scene = make_scene(cfg["P"], cfg["seed"], mu=mu, op=cfg["op"], precomp=cfg["precomp"])
    return scene, cfg


# ----------------------------------------------------------------------------------------
# 62-float RichPoint .ply (apps/gsrast/SplatData.hpp:17-25, SplatData.cpp:114-156)
# ----------------------------------------------------------------------------------------
PLY_FLOATS = 62


def write_ply(path: str, scene: SplatScene) -> None:
    """Write the scene *pre-activation* in the RichPoint record layout:
    position[3] normal[3] shs[48] (PLY order) opacity_logit scale_log[3] rotation[4]."""
    P = scene.P
    rec = np.zeros((P, PLY_FLOATS), dtype=np.float32)
    rec[:, 0:3] = scene.means3D
    if scene.shs is not None:
        rec[:, 6:54] = contract_sh_to_ply_order(scene.shs)
    elif scene.colors_precomp is not None:
        rec[:, 6:9] = (scene.colors_precomp - 0.5) / 0.28209479177387814
    o = np.clip(scene.opacities.astype(np.float64), 1e-7, 1 - 1e-7)
    rec[:, 54] = np.log(o / (1 - o)).astype(np.float32)
    rec[:, 55:58] = np.log(scene.scales)
    rec[:, 58:62] = scene.rotations
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"]
    names += [f"f_rest_{i}" for i in range(45)]
    names += ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    # SplatData.cpp:129-136 parses the element count from the THIRD header line.
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % P
    header += "".join("property float %s\n" % n for n in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rec.tobytes())


def load_ply_native(path: str) -> SplatScene:
    """The same staging through the library's host loader (gsr_ply_count / gsr_ply_load, csrc/ply.cu) — what a C++
    host calls instead of SplatData::loadFromSplatsPly (SplatData.cpp:114-156).  Raises on any GSR_ERR_PLY_* code."""
    import ctypes as C

    from . import _lib

    L = _lib.lib()
    n = C.c_int(0)
    rc = L.gsr_ply_count(path.encode(), C.byref(n))
    if rc < 0:
        raise ValueError("%s (%s)" % (L.gsr_error_string(rc).decode(), path))
    P = n.value
    means = np.empty((P, 3), np.float32); scales = np.empty((P, 3), np.float32); rot = np.empty((P, 4), np.float32)
    opac = np.empty(P, np.float32); shs = np.empty((P, 16, 3), np.float32)
    bbox = np.zeros(6, np.float32); center = np.zeros(3, np.float32)
    fp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = L.gsr_ply_load(path.encode(), P, fp(means), fp(scales), fp(rot), fp(opac), fp(shs), fp(bbox), fp(center))
    if rc < 0:
        raise ValueError("%s (%s)" % (L.gsr_error_string(rc).decode(), path))
    sc = SplatScene(means, scales, rot, opac, shs, None, 3, 16)
    sc.bbox = bbox.reshape(2, 3)
    sc.center = center
    return sc


def read_ply(path: str) -> SplatScene:
    """SplatData::loadFromSplatsPly + loadFromPly activation, returning contract layouts."""
    with open(path, "rb") as f:
        lines = [f.readline() for _ in range(3)]
        P = int(lines[2].split()[2])
        while True:
            ln = f.readline()
            if not ln:
                raise ValueError("no end_header in %s" % path)
            if ln.strip() == b"end_header":
                break
        raw = f.read(P * PLY_FLOATS * 4)
    if len(raw) < P * PLY_FLOATS * 4:
        raise ValueError("Reader is EOF? (%s)" % path)  # SplatData.cpp:147-151
    rec = np.frombuffer(raw, dtype=np.float32).reshape(P, PLY_FLOATS)
    means = rec[:, 0:3].copy()
    shs = ply_order_to_contract_sh(rec[:, 6:54])
    opac = _sigmoid(rec[:, 54])
    scales = np.exp(rec[:, 55:58]).astype(np.float32)
    q = rec[:, 58:62].astype(np.float32)
    q = q / np.sqrt((q * q).sum(axis=1, keepdims=True)).astype(np.float32)
    return SplatScene(means, scales, q.astype(np.float32), opac, shs, None, 3, 16)
