"""View-batch rendering on a resident scene and its sharding across GPUs.

`ViewRenderer` wraps the native gsr_renderer (gsrast_b200/csrc/views.cu), the C form of the
reference's GSGaussians object (apps/gsrast/GSGaussians.{hpp,cpp}: configureFromSplatData +
draw()).  Multi-GPU: a batch of camera views is split across ranks, every rank
holds the full Gaussian set, no collective sits on the data path; an optional gather brings
the finished frames to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests).  A single
frame is never split.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _tensor_from_ptr(addr: int, nbytes: int, device) -> torch.Tensor:
    """uint8 tensor over `nbytes` of device memory owned by the native library (no copy; valid while the owner lives)."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(addr), False), "version": 2}
    with torch.cuda.device(device):
        return torch.as_tensor(h, device=device)


class PinnedFrames:
    """Page-locked host landing buffer [n,3,H,W] from gsr_pinned_alloc (optionally write-combined); `.tensor` is a
    CPU tensor over it.  Keep the object alive while the tensor is in use."""

    def __init__(self, shape, dtype=torch.float32, write_combined=False):
        self.nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        self.ptr = _lib.lib().gsr_pinned_alloc(self.nbytes, (1 if write_combined else 0) | 2)
        if not self.ptr:
            raise MemoryError("gsr_pinned_alloc(%d) failed" % self.nbytes)
        raw = (C.c_ubyte * self.nbytes).from_address(self.ptr)
        self.tensor = torch.frombuffer(raw, dtype=torch.uint8).view(dtype).view(*shape)

    def close(self):
        if getattr(self, "ptr", None):
            self.tensor = None
            _lib.lib().gsr_pinned_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_cameras(cameras) -> np.ndarray:
    """list[Camera] -> float32 [n,36] (view[16] proj[16] cam_pos[3] pad): the per-frame upload of
    GSGaussians::draw (GSGaussians.cpp:171-173)."""
    if isinstance(cameras, np.ndarray):
        return np.ascontiguousarray(cameras, dtype=np.float32).reshape(-1, 36)
    return np.ascontiguousarray(np.stack([c.packed() for c in cameras]).astype(np.float32))


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process (one per GPU) to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host
    buffer is allocated: first-touch then places the frame staging memory next to the GPU's PCIe root, which is what
    the device->host frame copies of `render_host` stream into.  Best effort: returns {"node": n, "cpus": k} or
    {"skipped": why} (no sysfs, single node, cgroup without those CPUs) and never raises."""
    import os

    try:
        import torch

        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return {"skipped": "numa_node = -1"}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return {"skipped": "node %d has no CPU in this process's affinity mask" % node}
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed)}
    except Exception as ex:  # noqa: BLE001 - placement is an optimisation, never an error
        return {"skipped": "%s: %s" % (type(ex).__name__, ex)}


def shard_views(n_views: int, rank: int, world_size: int) -> list[int]:
    """Round-robin assignment view v -> rank v % world_size (balanced to within one view)."""
    return list(range(rank, n_views, world_size))


def views_per_rank(n_views: int, world_size: int) -> int:
    return (n_views + world_size - 1) // world_size


class ViewRenderer:
    def __init__(self, *, P, D, M, means3D, shs, colors_precomp, opacities, scales, rotations, background, width,
                 height, scale_modifier=1.0, compat=False, flags=0, stream=None, keep_state=False):
        """keep_state: GSR_FLAG_KEEP_STATE — every geometry-state field stays materialised so that
        `map_geometry_state` serves the Inspector's panel (Inspector.cpp:174-188); default is the lean state."""
        self._keep = (means3D, shs, colors_precomp, opacities, scales, rotations, background)
        self.width, self.height = int(width), int(height)
        self.device = means3D.device
        self.P = int(P)
        p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        self._stream = stream
        sp = (stream.cuda_stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream)
        fl = int(flags) | (_lib.FLAG_GSRAST_COMPAT if compat else 0) | (_lib.FLAG_KEEP_STATE if keep_state else 0)
        # the native renderer binds to the CURRENT device: make that the tensors' device, not whatever torch's default is
        with torch.cuda.device(self.device):
            self._h = _lib.lib().gsr_renderer_create(int(P), int(D), int(M), p(means3D), p(shs), p(colors_precomp),
                                                     p(opacities), p(scales), p(rotations), p(background),
                                                     float(scale_modifier), self.width, self.height, sp, fl)
        if not self._h:
            raise RuntimeError("gsr_renderer_create failed")

    def map_geometry_state(self, lane: int = 0) -> dict:
        """mapGeometryState (GSGaussians.cpp:214-219) over the renderer's private scratch: zero-copy views of the nine
        reference fields of `lane` (view v of a call ran on lane v % num_lanes; with timings, lane 0)."""
        st = _lib.GeometryState()
        _lib.check(_lib.lib().gsr_renderer_map_geometry_state(self._h, int(lane), C.byref(st)))
        P = self.P

        def view(addr, nbytes, dtype, shape):
            return _tensor_from_ptr(addr, nbytes, self.device).view(dtype).view(*shape)

        f32, i32, u8 = torch.float32, torch.int32, torch.uint8
        return dict(depths=view(st.depths, 4 * P, f32, (P,)), clamped=view(st.clamped, 3 * P, u8, (P, 3)),
                    internal_radii=view(st.internal_radii, 4 * P, i32, (P,)), means2D=view(st.means2D, 8 * P, f32, (P, 2)),
                    cov3D=view(st.cov3D, 24 * P, f32, (P, 6)), conic_opacity=view(st.conic_opacity, 16 * P, f32, (P, 4)),
                    rgb=view(st.rgb, 12 * P, f32, (P, 3)), tiles_touched=view(st.tiles_touched, 4 * P, i32, (P,)),
                    point_offsets=view(st.point_offsets, 4 * P, i32, (P,)))

    @staticmethod
    def num_lanes() -> int:
        return int(_lib.lib().gsr_renderer_num_lanes())

    @classmethod
    def from_scene(cls, scene, width, height, device="cuda", background=(0.0, 0.0, 0.0), **kw):
        dev = torch.device(device)
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        return cls(P=scene.P, D=scene.sh_degree, M=max(scene.max_coeffs, 1), means3D=t(scene.means3D), shs=t(scene.shs),
                   colors_precomp=t(scene.colors_precomp), opacities=t(scene.opacities), scales=t(scene.scales),
                   rotations=t(scene.rotations), background=torch.tensor(background, dtype=torch.float32, device=dev),
                   width=width, height=height, **kw)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().gsr_renderer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, cameras, tan_fovx, tan_fovy, out=None, timings=False):
        """Render n views into a device tensor [n,3,H,W].  Returns (out, num_rendered[n][, times])."""
        cams = pack_cameras(cameras)
        n = cams.shape[0]
        if out is None:
            out = torch.empty((n, 3, self.height, self.width), dtype=torch.float32, device=self.device)
        nr = (C.c_int * max(n, 1))()
        times = _lib.StageTimes() if timings else None
        rc = _lib.lib().gsr_renderer_render(self._h, cams.ctypes.data, n, float(tan_fovx), float(tan_fovy),
                                            out.data_ptr(), C.cast(nr, C.c_void_p),
                                            C.cast(C.pointer(times), C.c_void_p) if timings else None)
        _lib.check(rc)
        res = (out, [int(nr[i]) for i in range(n)])
        return res + (times.as_dict(),) if timings else res

    def render_host(self, cameras, tan_fovx, tan_fovy, out_host=None):
        """Render n views and deliver them to (pinned) host memory [n,3,H,W]; the copy of view k
        overlaps the render of view k+1.  Returns (out_host, num_rendered[n])."""
        cams = pack_cameras(cameras)
        n = cams.shape[0]
        if out_host is None:
            out_host = torch.empty((n, 3, self.height, self.width), dtype=torch.float32).pin_memory()
        nr = (C.c_int * max(n, 1))()
        rc = _lib.lib().gsr_renderer_render_host(self._h, cams.ctypes.data, n, float(tan_fovx), float(tan_fovy),
                                                 out_host.data_ptr(), C.cast(nr, C.c_void_p))
        _lib.check(rc)
        return out_host, [int(nr[i]) for i in range(n)]

    def render_host_u8(self, cameras, tan_fovx, tan_fovy, out_host=None):
        """Same with 8-bit frames (quantised on the device): uint8 [n,3,H,W] in (pinned) host memory."""
        cams = pack_cameras(cameras)
        n = cams.shape[0]
        if out_host is None:
            out_host = torch.empty((n, 3, self.height, self.width), dtype=torch.uint8).pin_memory()
        nr = (C.c_int * max(n, 1))()
        rc = _lib.lib().gsr_renderer_render_host_u8(self._h, cams.ctypes.data, n, float(tan_fovx), float(tan_fovy),
                                                    out_host.data_ptr(), C.cast(nr, C.c_void_p))
        _lib.check(rc)
        return out_host, [int(nr[i]) for i in range(n)]


def frames_to_u8(frames: torch.Tensor) -> torch.Tensor:
    """Device float frames -> uint8, round(clamp(x, 0, 1) * 255), same shape (gsr_frames_to_u8)."""
    frames = frames.contiguous()
    out = torch.empty(frames.shape, dtype=torch.uint8, device=frames.device)
    rc = _lib.lib().gsr_frames_to_u8(frames.data_ptr(), out.data_ptr(), frames.numel(),
                                     torch.cuda.current_stream(frames.device).cuda_stream)
    _lib.check(rc)
    return out


def gather_frames(local_frames: torch.Tensor, n_views: int, rank: int, world_size: int, dst: int = 0, group=None):
    """Bring the frames of a round-robin-sharded batch to `dst` in view order.

    local_frames: [len(shard_views(n_views, rank, world_size)), 3, H, W] on this rank.
    Returns [n_views,3,H,W] on dst, None elsewhere.  One torch.distributed gather (NCCL on GPU
    ranks, gloo on CPU); shards are padded to views_per_rank so every rank sends the same size."""
    import torch.distributed as dist

    per = views_per_rank(n_views, world_size)
    shape = (per,) + tuple(local_frames.shape[1:])
    send = local_frames
    if local_frames.shape[0] != per:
        send = torch.zeros(shape, dtype=local_frames.dtype, device=local_frames.device)
        send[: local_frames.shape[0]] = local_frames
    send = send.contiguous()
    if world_size == 1:
        return send[:n_views]
    # one [world, per, ...] landing buffer; view v lives at [v % world, v // world] (round-robin shards), so the
    # view-ordered batch is its transpose — one strided device copy instead of a per-rank index scatter
    buf = torch.empty((world_size,) + shape, dtype=send.dtype, device=send.device) if rank == dst else None
    dist.gather(send, list(buf.unbind(0)) if rank == dst else None, dst=dst, group=group)
    if rank != dst:
        return None
    return buf.transpose(0, 1).reshape((per * world_size,) + shape[1:])[:n_views]
