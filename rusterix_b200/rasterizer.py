"""Rasterizer: the drop-in for the reference's `Rasterizer::setup(..).rasterize(..)`
(src/rasterizer.rs:35-193).  Same constructor, builder methods, public fields and call signature;
the work happens on the GPU through the C ABI (include/rxcuda.h).  No CPU path exists here."""
import ctypes as C

import numpy as np

from . import _abi, _lib, marshal, vekmath
from .types import Assets, MapMini, MatVecMode, RenderMode, SampleMode, Scene


class DeviceContext:
    """One rxc_ctx (one GPU, one stream) plus the host-side cache keys of what is resident."""

    _by_device = {}

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        st = self.lib.rxc_create(int(device), C.byref(h))
        if st != 0:
            raise _lib.RxcError(st, "rxc_create failed (no sm_100 GPU visible?)")
        self.handle = h
        self.device = device
        self._assets_key = None
        self._scene_key = None
        self._lights_key = None
        self._mapmini_key = ((), ())
        self._geometry_keys = None      # per 3D batch in submission order: identity of its arrays at the last upload
        self.last_upload_kept = 0       # leading 3D batches the last scene upload reused (rxc_update_scene)

    @classmethod
    def get(cls, device=0) -> "DeviceContext":
        if device not in cls._by_device:
            cls._by_device[device] = DeviceContext(device)
        return cls._by_device[device]

    def check(self, st):
        if st != 0:
            raise _lib.RxcError(st, self.lib.rxc_last_error(self.handle).decode())

    def check_static(self, st):
        if st != 0:
            raise _lib.RxcError(st, "call without a context failed")

    def close(self):
        if self.handle:
            self.lib.rxc_destroy(self.handle)
            self.handle = None
        self._by_device.pop(self.device, None)

    def set_stream(self, cuda_stream):
        self.check(self.lib.rxc_set_stream(self.handle, C.c_void_p(cuda_stream or 0)))

    def upload(self, scene: Scene, assets: Assets, index_bytes=4):
        akey = (assets._uid, assets._generation, len(assets.tile_list))
        if akey != self._assets_key:
            m = marshal.marshal_tiles(assets.tile_list)
            self.check(self.lib.rxc_set_assets(self.handle, m.struct, len(assets.tile_list)))
            self._assets_key = akey
            self._scene_key = None
        skey = (scene._uid, scene._generation, index_bytes, scene.structure_key())
        lights = scene.all_lights()
        lkey = tuple(
            (int(l.light_type), tuple(l.position), tuple(l.color), l.intensity, l.emitting, l.start_distance,
             l.end_distance, l.flicker, tuple(l.direction), l.cone_angle, tuple(l.normal), l.width, l.height,
             l.from_linedef) for l in lights)
        if skey != self._scene_key:
            m = marshal.marshal_scene(scene, index_bytes, assets)
            # an engine's frame loop replaces the dynamic batches and keeps the world: the leading 3D batches whose arrays are the
            # ones already resident are not uploaded again (rxc_update_scene)
            gkeys = marshal.geometry_keys(scene)
            keep = 0
            prev = self._scene_key
            if prev is not None and prev[:3] == skey[:3] and self._geometry_keys:
                while keep < min(len(gkeys), len(self._geometry_keys)) and gkeys[keep] == self._geometry_keys[keep]:
                    keep += 1
            self._geometry_keys = None
            st = self.lib.rxc_update_scene(self.handle, C.byref(m.struct), keep) if keep else _abi.RXC_ERR_INVALID
            if st == _abi.RXC_ERR_INVALID:   # nothing to keep, or the library no longer holds that prefix (an earlier upload failed): everything
                keep = 0
                st = self.lib.rxc_set_scene(self.handle, C.byref(m.struct))
            self.check(st)
            self._geometry_keys = gkeys
            self.last_upload_kept = keep
            self._scene_key = skey
            self._lights_key = lkey
        elif lkey != self._lights_key:
            m = marshal.marshal_lights(lights)
            self.check(self.lib.rxc_set_lights(self.handle, m.struct, len(lights)))
            self._lights_key = lkey

    def set_mapmini(self, mapmini):
        key = mapmini.key() if mapmini is not None else ((), ())
        if key != self._mapmini_key:
            if mapmini is None:
                self.check(self.lib.rxc_set_mapmini(self.handle, None))
            else:
                m = marshal.marshal_mapmini(mapmini)
                self.check(self.lib.rxc_set_mapmini(self.handle, C.byref(m.struct)))
            self._mapmini_key = key

    def selftest_div(self, n_pairs=1 << 30, seed=0x52555354) -> int:
        """Mismatches of the raster kernel's residual-corrected division against div.rn (must be 0)."""
        bad = C.c_uint64(0)
        pair = (C.c_uint32 * 2)()
        self.check(self.lib.rxc_selftest_div(self.handle, C.c_uint64(seed), C.c_uint64(n_pairs), C.byref(bad), pair))
        self.last_bad_pair = (int(pair[0]), int(pair[1]))
        return int(bad.value)

    def vm_execute(self, program: int, records: np.ndarray):
        """Diagnostics (rxc_vm_execute): run shader `program` of the resident scene on records[n, 18]
        (uv, color, normal, hitpoint, time, opacity); returns (out[n, 24], faults)."""
        rec = np.ascontiguousarray(records, dtype=np.float32).reshape(-1, 18)
        out = np.zeros((len(rec), 24), dtype=np.float32)
        faults = C.c_uint32(0)
        self.check(self.lib.rxc_vm_execute(self.handle, int(program), len(rec), rec.ctypes.data, out.ctypes.data, C.byref(faults)))
        return out, int(faults.value)

    def set_vm_jit(self, mode: int):
        """rxc_set_vm_jit: 0 = interpreter only, 1 = batch-shader kernels compiled in the background, 2 = compiled before
        the first frame that needs them.  Applies to scenes uploaded afterwards (the resident one is uploaded again)."""
        self.check(self.lib.rxc_set_vm_jit(self.handle, int(mode)))
        self._scene_key = None

    def vm_jit_info(self) -> dict:
        nt, nk, pend, used = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint64(0)
        log = C.create_string_buffer(1 << 14)
        self.check(self.lib.rxc_vm_jit_info(self.handle, C.byref(nt), C.byref(nk), C.byref(pend), C.byref(used), log, len(log)))
        return {"translated": int(nt.value), "kernels": int(nk.value), "pending": bool(pend.value), "launches": int(used.value),
                "log": log.value.decode(errors="replace")}

    def set_vm_state_mode(self, mode: int):
        """rxc_set_vm_state_mode: 0 = a fresh Execution per fragment (fast, default), 1 = the reference's order and per-tile
        Execution for every scene with programs, 2 = for the scenes whose state report flags a program."""
        self.check(self.lib.rxc_set_vm_state_mode(self.handle, int(mode)))

    def ordered_frames(self) -> int:
        mode, n = C.c_int32(0), C.c_uint64(0)
        self.check(self.lib.rxc_get_vm_state_mode(self.handle, C.byref(mode), C.byref(n)))
        return int(n.value)

    def vm_state_report(self):
        """rxc_vm_scene_state_report: per program of the resident scene 0 = the device and the reference (one Execution per
        tile, never reset) compute the same thing, 1 = the program can observe the difference, 2 = not analysable."""
        n = C.c_uint32(0)
        self.check(self.lib.rxc_vm_scene_state_report(self.handle, None, 0, C.byref(n)))
        out = (C.c_uint32 * max(1, n.value))()
        self.check(self.lib.rxc_vm_scene_state_report(self.handle, out, n.value, C.byref(n)))
        return list(out)[:n.value]

    def pin_host(self, buf):
        """rxc_pin_host on a writable buffer (numpy array, bytearray ...): frames written into it then drain by DMA."""
        p, keep = _buffer_pointer(buf, 1)
        nbytes = keep.nbytes if hasattr(keep, "nbytes") else keep.numel() * keep.element_size()
        self.check(self.lib.rxc_pin_host(self.handle, C.c_void_p(p), nbytes))
        return buf

    def unpin_host(self, buf):
        p, _keep = _buffer_pointer(buf, 1)
        self.check(self.lib.rxc_unpin_host(self.handle, C.c_void_p(p)))

    def stats(self) -> _abi.rxc_stats:
        s = _abi.rxc_stats()
        self.check(self.lib.rxc_get_stats(self.handle, C.byref(s)))
        return s

    def reset_stats(self):
        self.check(self.lib.rxc_reset_stats(self.handle))

    def set_profiling(self, on: bool):
        self.check(self.lib.rxc_set_profiling(self.handle, 1 if on else 0))

    def synchronize(self):
        self.check(self.lib.rxc_synchronize(self.handle))

    def kernel_names(self):
        return [self.lib.rxc_kernel_name(i).decode() for i in range(_abi.RXC_N_KERNELS)]


def _buffer_pointer(buf, nbytes):
    """numpy array / torch tensor (cpu or cuda) / bytearray -> raw address, with a length check
    (the reference panics on a short slice, src/rasterizer.rs:572)."""
    if buf is None:
        return None, None
    if hasattr(buf, "data_ptr"):  # torch tensor
        if buf.numel() * buf.element_size() < nbytes:
            raise ValueError("output buffer too small")
        if not buf.is_contiguous():
            raise ValueError("output tensor must be contiguous")
        return buf.data_ptr(), buf
    if isinstance(buf, np.ndarray):
        if buf.nbytes < nbytes:
            raise ValueError("output buffer too small")
        if not buf.flags["C_CONTIGUOUS"]:
            raise ValueError("output array must be contiguous")
        return buf.ctypes.data, buf
    mv = memoryview(buf)
    if mv.nbytes < nbytes:
        raise ValueError("output buffer too small")
    arr = np.frombuffer(mv, dtype=np.uint8)
    return arr.ctypes.data, arr


def _check_band(band, width, height):
    """`band` = (y0, y1) or (y0, y1, x0, x1).  In the C ABI a band of 0/0 means "the whole frame", so an EMPTY band
    must never reach it (it would render the full frame into a band-sized buffer): callers skip the render for a
    rank whose band is empty, and anything else that is not a proper sub-rectangle is an error here."""
    if band is None:
        return
    if len(band) not in (2, 4):
        raise ValueError("band is (y0, y1) or (y0, y1, x0, x1)")
    y0, y1 = int(band[0]), int(band[1])
    if not (0 <= y0 < y1 <= height):
        raise ValueError(f"empty or out-of-range row band {band} for a frame of {height} rows (skip the render for an empty band)")
    if len(band) == 4:
        x0, x1 = int(band[2]), int(band[3])
        if not (0 <= x0 < x1 <= width):
            raise ValueError(f"empty or out-of-range column band {band} for a frame of {width} columns")


class Rasterizer:
    def __init__(self, projection_matrix_2d, view_matrix, projection_matrix):
        self.render_mode_ = RenderMode.render_all()
        self.projection_matrix_2d = None if projection_matrix_2d is None else np.asarray(projection_matrix_2d, np.float32).reshape(3, 3)
        self.view_matrix = np.asarray(view_matrix, dtype=np.float32).reshape(4, 4)
        self.projection_matrix = np.asarray(projection_matrix, dtype=np.float32).reshape(4, 4)
        self.inverse_view_matrix = vekmath.inverted(self.view_matrix)              # :97
        self.inverse_projection_matrix = vekmath.inverted(self.projection_matrix)  # :116
        self.camera_pos = self.inverse_view_matrix[:3, 3].copy()                   # :98-102
        self.width = 0.0
        self.height = 0.0
        self.sample_mode_ = SampleMode.Nearest
        self.hash_anim = 0
        self.background_color = None
        self.ambient_color = None
        self.brush_preview = None         # src/rasterizer.rs:50 (BrushPreview)
        self.render_graph = None          # :56: a types.RenderGraph
        self.render_miss = []             # :78, rebuilt by every rasterize()
        self.sun_dir = None               # :86-87
        self.day_factor = 0.0
        self.mapmini = MapMini.default()  # src/rasterizer.rs:71
        self.preserve_transparency = False
        self.hour = 12.0
        self.time_ = 0.0
        self.matvec_mode = MatVecMode.FmaColumns
        self.device = 0
        self.index_bytes = 4

    @staticmethod
    def setup(projection_matrix_2d, view_matrix, projection_matrix) -> "Rasterizer":
        return Rasterizer(projection_matrix_2d, view_matrix, projection_matrix)

    # builder methods, src/rasterizer.rs:154-182
    def render_mode(self, m: RenderMode):
        self.render_mode_ = m
        return self

    def sample_mode(self, m):
        self.sample_mode_ = SampleMode(m)
        return self

    def background(self, pixel):
        self.background_color = tuple(int(c) for c in pixel)
        return self

    def ambient(self, v4):
        self.ambient_color = tuple(float(c) for c in v4)
        return self

    def time(self, t):
        self.time_ = float(t)
        return self

    def on_device(self, device: int):
        self.device = int(device)
        return self

    def _check_supported(self):
        """What the device path does not do is reported by the library (RXC_ERR_UNSUPPORTED from rxc_set_scene /
        rxc_rasterize: the Sky node's cloud layer, texture-baking VM ops, device VM limits); the only host-side
        check is the one the library cannot make, a render graph holding something other than the mirrored nodes."""
        g = self.render_graph
        if g is not None and not hasattr(g, "miss_nodes"):
            raise _lib.RxcError(_abi.RXC_ERR_UNSUPPORTED, "render_graph must be a rusterix_b200.types.RenderGraph")

    def prepare_render_graph(self):
        """src/rasterizer.rs:227-253: collect the miss nodes, render_setup them (the last Sky node's sun wins) and
        let a Sky node replace the ambient colour.  Host-side, like in the reference."""
        self.render_miss = list(self.render_graph.miss_nodes) if self.render_graph is not None else []
        for node in self.render_miss:
            r = node.render_setup(self.hour)
            if r is not None:
                self.sun_dir, self.day_factor = r
        for node in self.render_miss:
            a = node.render_ambient_color()
            if a is not None:
                self.ambient_color = a

    def rasterize(self, scene: Scene, pixels, width: int, height: int, tile_size: int, assets: Assets,
                  owner=None, depth=None, band=None, sync=True):
        """Same positional arguments as the reference (src/rasterizer.rs:185-193).  `pixels` is a
        writable uint8 buffer of at least width*height*4 bytes: a numpy array, bytearray, or a torch
        tensor on the CPU or on the context's GPU.  `owner` (uint32) / `depth` (float32) are
        optional parity outputs.  `band=(y0,y1)` renders only those rows into a band-sized buffer,
        `band=(y0,y1,x0,x1)` only that rectangle (x0 a multiple of 32)."""
        self._check_supported()
        self.prepare_render_graph()
        self.width, self.height = float(width), float(height)
        # "We append the in-scope chunk lights to the dynamic lights" -- on every call, never cleared
        # (src/rasterizer.rs:219-223)
        for chunk in scene.chunks.values():
            scene.dynamic_lights.extend(chunk.lights)
        ctx = DeviceContext.get(self.device)
        ctx.upload(scene, assets, self.index_bytes)
        ctx.set_mapmini(self.mapmini)
        _check_band(band, width, height)
        frame = marshal.make_frame(self, scene, width, height, tile_size, band)
        rows = height if band is None else band[1] - band[0]
        width = width if band is None or len(band) < 4 else band[3] - band[2]   # the buffer holds the rendered rectangle
        p, _k1 = _buffer_pointer(pixels, width * rows * 4)
        if p is None:
            raise ValueError("pixels is required")
        o, _k2 = _buffer_pointer(owner, width * rows * 4)
        d, _k3 = _buffer_pointer(depth, width * rows * 4)
        fn = ctx.lib.rxc_rasterize if sync else ctx.lib.rxc_rasterize_async
        ctx.check(fn(ctx.handle, C.byref(frame), C.c_void_p(p), C.c_void_p(o or 0), C.c_void_p(d or 0)))
        return pixels

    def rasterize_projected(self, scene: Scene, projected, pixels, width: int, height: int, tile_size: int, assets: Assets,
                            owner=None, depth=None):
        """The pre-projected entry (rxc_rasterize_projected): like rasterize(), but the 3D batches arrive the way the
        host's own `Scene::project` left them (src/scene.rs:154-200) -- `projected` holds, per 3D batch in submission
        order, projected_vertices / clipped_uvs / clipped_normals / clipped_indices / edges / visible / bounding_box
        (marshal.marshal_projected) -- and the device uses those bits verbatim instead of projecting itself."""
        self._check_supported()
        self.prepare_render_graph()
        self.width, self.height = float(width), float(height)
        for chunk in scene.chunks.values():
            scene.dynamic_lights.extend(chunk.lights)
        ctx = DeviceContext.get(self.device)
        ctx.upload(scene, assets, self.index_bytes)
        ctx.set_mapmini(self.mapmini)
        frame = marshal.make_frame(self, scene, width, height, tile_size, None)
        m = marshal.marshal_projected(projected)
        p, _k1 = _buffer_pointer(pixels, width * height * 4)
        o, _k2 = _buffer_pointer(owner, width * height * 4)
        d, _k3 = _buffer_pointer(depth, width * height * 4)
        ctx.check(ctx.lib.rxc_rasterize_projected(ctx.handle, C.byref(frame), m.struct, len(projected), C.c_void_p(p), C.c_void_p(o or 0), C.c_void_p(d or 0)))
        return pixels

    @staticmethod
    def prepare_batch(rasterizers, scene: Scene, width, height, tile_size, assets: Assets, band=None, device=0):
        """Marshal a camera sweep once: uploads the scene if needed and returns a FrameBatch that
        `rasterize_batch` replays with a single C call per step."""
        ctx = DeviceContext.get(device)
        for chunk in scene.chunks.values():  # once per prepared sweep (the reference appends per rasterize call)
            scene.dynamic_lights.extend(chunk.lights)
        ctx.upload(scene, assets, 4)
        if rasterizers:
            ctx.set_mapmini(rasterizers[0].mapmini)
        _check_band(band, width, height)
        n = len(rasterizers)
        frames = (_abi.rxc_frame * n)()
        for i, r in enumerate(rasterizers):
            r._check_supported()
            r.prepare_render_graph()
            frames[i] = marshal.make_frame(r, scene, width, height, tile_size, band)
        rows = height if band is None else band[1] - band[0]
        cols = width if band is None or len(band) < 4 else band[3] - band[2]
        return FrameBatch(ctx, frames, n, cols * rows * 4)

    @staticmethod
    def rasterize_batch(rasterizers, scene: Scene, pixels, width, height, tile_size, assets: Assets, band=None,
                        sync=True, device=0):
        """Camera sweep: one Rasterizer (camera) per frame, one scene, one launch sequence.
        `rasterizers` may be a list of Rasterizer or a FrameBatch from `prepare_batch`."""
        fb = rasterizers if isinstance(rasterizers, FrameBatch) else Rasterizer.prepare_batch(
            rasterizers, scene, width, height, tile_size, assets, band, device)
        return fb.run(pixels, sync)


class FrameBatch:
    """Pre-marshalled rxc_frame array of a camera sweep (host-side convenience, no device state)."""

    def __init__(self, ctx, frames, n, stride):
        self.ctx, self.frames, self.n, self.stride = ctx, frames, n, stride

    def run(self, pixels, sync=True):
        p, _k = _buffer_pointer(pixels, self.stride * self.n)
        fn = self.ctx.lib.rxc_rasterize_batch if sync else self.ctx.lib.rxc_rasterize_batch_async
        self.ctx.check(fn(self.ctx.handle, self.frames, self.n, C.c_void_p(p), self.stride))
        return pixels
