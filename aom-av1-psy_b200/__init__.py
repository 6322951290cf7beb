"""aom-av1-psy_b200: host-side mirror of the reference's temporal-filter interface
on top of the C ABI of libtf_gpu.so (include/tf_gpu.h).

The reference is C; the product's host side is the C++ inside libtf_gpu.so.
This module is only the thin ctypes binding used by tests, bench.py and
__graft_entry__: it names things the way av1/encoder/temporal_filter.{c,h} does
(YV12 buffers, TemporalFilterCtx fields, av1_temporal_filter,
av1_estimate_noise_from_single_plane, FRAME_DIFF) so parity tests read like the
reference's own.  There is no CPU fallback: importing works anywhere, but every
call raises TfGpuError unless libtf_gpu.so is built and a CUDA device exists.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtf_gpu.so")
# development only: point at another build of the same library (A/B runs of two kernels on one box)
LIB_PATH = os.environ.get("TF_GPU_LIB", LIB_PATH)

TF_GPU_MAX_FRAMES = 24
NOISE_ESTIMATION_EDGE_THRESHOLD = 50  # temporal_filter.h:79
TF_BLOCK = 32  # TF_BLOCK_SIZE = BLOCK_32X32, temporal_filter.h:31

ERRORS = {0: "OK", -1: "INVALID", -2: "MEM", -3: "CUDA", -4: "NO_DEVICE", -5: "UNSUPPORTED"}

EXPORTS = [
    "tf_gpu_abi_version", "tf_gpu_create", "tf_gpu_destroy", "tf_gpu_last_error", "tf_gpu_estimate_noise",
    "tf_gpu_filter", "tf_gpu_filter_dump", "tf_gpu_submit", "tf_gpu_wait", "tf_gpu_cache_frame",
    "tf_gpu_evict_frame", "tf_gpu_filter_resident", "tf_gpu_download_output", "tf_gpu_output_device_plane",
    "tf_gpu_host_register", "tf_gpu_host_unregister", "tf_gpu_last_stats", "tf_gpu_event_record",
    "tf_gpu_event_elapsed_ms", "tf_gpu_synchronize", "tf_gpu_microbench", "tf_gpu_last_kernel_times", "tf_gpu_filter_resident_async", "tf_gpu_filter_resident_result", "tf_gpu_collect_counters", "tf_gpu_read_counters",
    "tf_gpu_cache_frame_async", "tf_gpu_debug_read_plane", "tf_gpu_device_border", "tf_gpu_fullpel_search_batch",
    "tf_gpu_output_ipc_export", "tf_gpu_output_ipc_import",
]


class TfGpuError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"tf_gpu error {code} ({ERRORS.get(code, '?')}): {msg}")
        self.code = code


class DeviceCfg(C.Structure):
    _fields_ = [("device", C.c_int), ("max_cached_frames", C.c_int), ("reserved", C.c_int * 6)]


class Frame(C.Structure):
    """tf_gpu_frame: POD mirror of YV12_BUFFER_CONFIG (aom_scale/yv12config.h:43-123)."""
    _fields_ = [
        ("plane", C.c_void_p * 3), ("stride", C.c_int * 2), ("crop_w", C.c_int * 2), ("crop_h", C.c_int * 2),
        ("aligned_w", C.c_int * 2), ("aligned_h", C.c_int * 2), ("border", C.c_int), ("ss_x", C.c_int),
        ("ss_y", C.c_int), ("is_hbd", C.c_int), ("frame_id", C.c_uint64),
    ]


class Params(C.Structure):
    """tf_gpu_params: TemporalFilterCtx (temporal_filter.h:93-144) + the speed features the search reads."""
    _fields_ = [
        ("num_frames", C.c_int), ("filter_frame_idx", C.c_int), ("num_planes", C.c_int), ("bit_depth", C.c_int),
        ("noise_levels", C.c_double * 3), ("q_factor", C.c_int), ("filter_strength", C.c_int),
        ("mi_rows", C.c_int), ("mi_cols", C.c_int), ("border_in_pixels", C.c_int), ("force_integer_mv", C.c_int),
        ("allow_hp", C.c_int), ("subpel_method", C.c_int), ("subpel_iters_per_step", C.c_int),
        ("prune_mesh_level", C.c_int), ("mesh_patterns", (C.c_int * 2) * 4), ("use_downsampled_sad", C.c_int),
        ("compute_frame_diff", C.c_int), ("out_row_begin", C.c_int), ("out_row_end", C.c_int),
        ("extend_output_borders", C.c_int), ("cm_width", C.c_int), ("cm_height", C.c_int),
        ("reserved", C.c_int * 5),
    ]


class SearchItem(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("start_row", C.c_int16), ("start_col", C.c_int16)]


class SearchResult(C.Structure):
    _fields_ = [("row", C.c_int16), ("col", C.c_int16), ("var", C.c_int32)]


class Dump(C.Structure):
    _fields_ = [("subblock_mvs", C.c_void_p), ("subblock_mses", C.c_void_p), ("pred", C.c_void_p),
                ("accum", C.c_void_p), ("count", C.c_void_p)]


_lib = None


def load_library():
    """dlopen libtf_gpu.so; fails loudly when the CUDA extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TfGpuError(-4, f"{LIB_PATH} is not built (run __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, i, u64 = C.c_void_p, C.c_int, C.c_uint64
    lib.tf_gpu_abi_version.restype = i
    lib.tf_gpu_create.argtypes = [C.POINTER(vp), C.POINTER(DeviceCfg)]
    lib.tf_gpu_destroy.argtypes = [vp]
    lib.tf_gpu_destroy.restype = None
    lib.tf_gpu_last_error.argtypes = [vp]
    lib.tf_gpu_last_error.restype = C.c_char_p
    lib.tf_gpu_estimate_noise.argtypes = [vp, C.POINTER(Frame), i, i, i, C.POINTER(C.c_double)]
    lib.tf_gpu_filter.argtypes = [vp, C.POINTER(Params), C.POINTER(Frame), C.POINTER(Frame), C.POINTER(C.c_int64)]
    lib.tf_gpu_filter_dump.argtypes = lib.tf_gpu_filter.argtypes + [C.POINTER(Dump)]
    lib.tf_gpu_submit.argtypes = lib.tf_gpu_filter.argtypes + [C.POINTER(u64)]
    lib.tf_gpu_wait.argtypes = [vp, u64]
    lib.tf_gpu_cache_frame.argtypes = [vp, C.POINTER(Frame)]
    lib.tf_gpu_evict_frame.argtypes = [vp, u64]
    lib.tf_gpu_cache_frame_async.argtypes = [vp, C.POINTER(Frame)]
    lib.tf_gpu_debug_read_plane.argtypes = [vp, u64, i, vp, i, i, i, i, i]
    lib.tf_gpu_device_border.restype = i
    lib.tf_gpu_output_ipc_export.argtypes = [vp, i, C.c_char_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.tf_gpu_output_ipc_import.argtypes = [vp, i, C.c_char_p, C.c_size_t, C.c_size_t]
    lib.tf_gpu_fullpel_search_batch.argtypes = [vp, C.POINTER(Params), C.POINTER(Frame), C.POINTER(Frame), i,
                                                C.POINTER(SearchItem), i, C.POINTER(SearchResult)]
    lib.tf_gpu_filter_resident.argtypes = [vp, C.POINTER(Params), C.POINTER(u64), C.POINTER(C.c_int64),
                                           C.POINTER(C.c_float)]
    lib.tf_gpu_download_output.argtypes = [vp, C.POINTER(Frame), i, i]
    lib.tf_gpu_filter_resident_async.argtypes = [vp, C.POINTER(Params), C.POINTER(u64)]
    lib.tf_gpu_filter_resident_result.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_float)]
    lib.tf_gpu_output_device_plane.argtypes = [vp, i, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(i),
                                               C.POINTER(i)]
    lib.tf_gpu_host_register.argtypes = [vp, vp, C.c_size_t]
    lib.tf_gpu_host_unregister.argtypes = [vp, vp]
    lib.tf_gpu_last_stats.argtypes = [vp, C.POINTER(i), C.POINTER(C.c_float)]
    lib.tf_gpu_event_record.argtypes = [vp, i]
    lib.tf_gpu_event_elapsed_ms.argtypes = [vp, i, i, C.POINTER(C.c_float)]
    lib.tf_gpu_synchronize.argtypes = [vp]
    lib.tf_gpu_microbench.argtypes = [vp, i, C.POINTER(C.c_double)]
    lib.tf_gpu_last_kernel_times.argtypes = [vp, C.POINTER(C.c_float)]
    lib.tf_gpu_collect_counters.argtypes = [vp, i]
    lib.tf_gpu_read_counters.argtypes = [vp, C.POINTER(u64)]
    _lib = lib
    return lib


def _align(v, n):
    return (v + n - 1) // n * n


class Yv12Buffer:
    """Host frame with the reference's YV12_BUFFER_CONFIG layout
    (aom_realloc_frame_buffer, aom_scale/generic/yv12config.c:223-258): 8-aligned
    sizes, stride = align32(aligned_w + 2*border), planes at border*stride+border."""

    def __init__(self, width, height, ss_x=1, ss_y=1, use_hbd=False, border=160, monochrome=False, frame_id=0):
        self.width, self.height, self.ss_x, self.ss_y = width, height, ss_x, ss_y
        self.use_hbd, self.border, self.monochrome, self.frame_id = bool(use_hbd), border, bool(monochrome), frame_id
        self.dtype = np.uint16 if use_hbd else np.uint8
        aw, ah = _align(width, 8), _align(height, 8)
        self.aligned = [(aw, ah), (aw >> ss_x, ah >> ss_y)]
        self.crop = [(width, height), ((width + ss_x) >> ss_x, (height + ss_y) >> ss_y)]
        self.stride = [_align(aw + 2 * border, 32), _align(aw + 2 * border, 32) >> ss_x]
        self.borders = [(border, border), (border >> ss_x, border >> ss_y)]
        self.num_planes = 1 if monochrome else 3
        self.alloc = []
        for p in range(self.num_planes):
            k = 1 if p else 0
            rows = self.aligned[k][1] + 2 * self.borders[k][1]
            self.alloc.append(np.zeros((rows, self.stride[k]), self.dtype))

    def plane(self, p):
        """View whose [0,0] is pixel (0,0) (y_buffer / u_buffer / v_buffer)."""
        k = 1 if p else 0
        bx, by = self.borders[k]
        return self.alloc[p][by:, bx:]

    def set_planes(self, y, u=None, v=None, extend=True):
        """Copy crop-sized planes in; with extend, replicate edges with the extents of
        av1_copy_and_extend_frame (extend.c:113-131): top/left = border, right/bottom =
        max(aligned + border, align64(aligned)) - crop (chroma: luma extents >> ss)."""
        aw, ah = self.aligned[0]
        er_y = max(aw + self.border, _align(aw, 64)) - self.width
        eb_y = max(ah + self.border, _align(ah, 64)) - self.height
        for p, src in enumerate((y, u, v)[: self.num_planes]):
            k = 1 if p else 0
            cw, ch = self.crop[k]
            bx, by = self.borders[k]
            er = er_y >> self.ss_x if p else er_y
            eb = eb_y >> self.ss_y if p else eb_y
            a = self.alloc[p]
            a[by:by + ch, bx:bx + cw] = src
            if extend:
                x1 = min(bx + cw + er, a.shape[1])
                y1 = min(by + ch + eb, a.shape[0])
                a[by:by + ch, :bx] = a[by:by + ch, bx:bx + 1]
                a[by:by + ch, bx + cw:x1] = a[by:by + ch, bx + cw - 1:bx + cw]
                a[:by, :x1] = a[by:by + 1, :x1]
                a[by + ch:y1, :x1] = a[by + ch - 1:by + ch, :x1]
        return self

    def c_frame(self):
        f = Frame()
        for p in range(self.num_planes):
            k = 1 if p else 0
            bx, by = self.borders[k]
            f.plane[p] = self.alloc[p].ctypes.data + (by * self.stride[k] + bx) * self.alloc[p].itemsize
        for k in range(2):
            f.stride[k] = self.stride[k]
            f.crop_w[k], f.crop_h[k] = self.crop[k]
            f.aligned_w[k], f.aligned_h[k] = self.aligned[k]
        f.border, f.ss_x, f.ss_y = self.border, self.ss_x, self.ss_y
        f.is_hbd, f.frame_id = int(self.use_hbd), self.frame_id
        return f

    def full_blocks(self, p):
        """Plane region covered by whole 32x32 blocks (what temporal_filter.c:740-777 writes)."""
        k = 1 if p else 0
        w = _align(self.width, 32) >> (self.ss_x if p else 0)
        h = _align(self.height, 32) >> (self.ss_y if p else 0)
        return self.plane(p)[:h, :w]


def make_params(p):
    """dict (tests/_params.tf_params) -> tf_gpu_params."""
    c = Params()
    c.num_frames, c.filter_frame_idx = p["num_frames"], p["filter_frame_idx"]
    c.num_planes = 1 if p["monochrome"] else 3
    c.bit_depth = p["bit_depth"]
    for i in range(3):
        c.noise_levels[i] = float(p["noise_levels"][i])
    c.q_factor, c.filter_strength = p["q_factor"], p["filter_strength"]
    c.mi_rows, c.mi_cols = _align(p["height"], 8) // 4, _align(p["width"], 8) // 4
    c.border_in_pixels = p["border"]
    c.force_integer_mv, c.allow_hp = p["force_integer_mv"], p["allow_hp"]
    c.subpel_method, c.subpel_iters_per_step = p["subpel_method"], p["subpel_iters_per_step"]
    c.prune_mesh_level = p["prune_mesh_level"]
    for i in range(4):
        c.mesh_patterns[i][0], c.mesh_patterns[i][1] = p["mesh"][i]
    c.use_downsampled_sad, c.compute_frame_diff = p["use_downsampled_sad"], p["compute_frame_diff"]
    c.out_row_begin, c.out_row_end = p.get("out_row_begin", 0), p.get("out_row_end", 0)
    c.extend_output_borders = p.get("extend_output_borders", 0)
    c.cm_width, c.cm_height = p.get("cm_width", 0), p.get("cm_height", 0)
    return c


class TemporalFilterGpu:
    """One tf_gpu_ctx (CUDA device + stream + frame cache)."""

    def __init__(self, device=-1, max_cached_frames=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        cfg = DeviceCfg(device=device, max_cached_frames=max_cached_frames)
        rc = self.lib.tf_gpu_create(C.byref(self.h), C.byref(cfg))
        if rc:
            self.h = None
            raise TfGpuError(rc, "tf_gpu_create failed (no CUDA device? there is no CPU fallback)")

    def _check(self, rc):
        if rc:
            raise TfGpuError(rc, (self.lib.tf_gpu_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.tf_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # av1_estimate_noise_from_single_plane (temporal_filter.c:1150-1194)
    def estimate_noise_from_single_plane(self, frame, plane, bit_depth, edge_thresh=NOISE_ESTIMATION_EDGE_THRESHOLD):
        out = C.c_double()
        cf = frame.c_frame()
        self._check(self.lib.tf_gpu_estimate_noise(self.h, C.byref(cf), plane, bit_depth, edge_thresh, C.byref(out)))
        return out.value

    # av1_temporal_filter (temporal_filter.c:1276-1312)
    def temporal_filter(self, params, frames, out, dump=False):
        """frames: list of Yv12Buffer; out: Yv12Buffer.  Returns dict(diff=[sum, sse], mvs, mses, pred, accum, count)."""
        cp = make_params(params) if isinstance(params, dict) else params
        n = cp.num_frames
        arr = (Frame * max(len(frames), 1))(*[f.c_frame() for f in frames])
        co = out.c_frame()
        diff = (C.c_int64 * 2)()
        res = {}
        if dump:
            n = max(n, 1)
            mb_rows, mb_cols = (out.height + 31) // 32, (out.width + 31) // 32
            nb = mb_rows * mb_cols
            num_pels = 1024 + (0 if cp.num_planes == 1 else 2 * (1024 >> (out.ss_x + out.ss_y)))
            res["mvs"] = np.zeros((nb, n, 4, 2), np.int16)
            res["mses"] = np.zeros((nb, n, 4), np.int32)
            res["pred"] = np.zeros((nb, n, num_pels), np.uint16)
            res["accum"] = np.zeros((nb, num_pels), np.uint32)
            res["count"] = np.zeros((nb, num_pels), np.uint16)
            d = Dump(*[res[k].ctypes.data for k in ("mvs", "mses", "pred", "accum", "count")])
            self._check(self.lib.tf_gpu_filter_dump(self.h, C.byref(cp), arr, C.byref(co), diff, C.byref(d)))
        else:
            self._check(self.lib.tf_gpu_filter(self.h, C.byref(cp), arr, C.byref(co), diff))
        res["diff"] = np.array([diff[0], diff[1]], np.int64)
        return res

    def submit(self, params, frames, out):
        cp = make_params(params) if isinstance(params, dict) else params
        arr = (Frame * max(len(frames), 1))(*[f.c_frame() for f in frames])
        co = out.c_frame()
        diff = (C.c_int64 * 2)()
        t = C.c_uint64()
        self._check(self.lib.tf_gpu_submit(self.h, C.byref(cp), arr, C.byref(co), diff, C.byref(t)))
        return t.value, diff, (arr, co)

    def wait(self, ticket):
        self._check(self.lib.tf_gpu_wait(self.h, ticket))

    def cache_frame(self, frame):
        cf = frame.c_frame()
        self._check(self.lib.tf_gpu_cache_frame(self.h, C.byref(cf)))

    def cache_frame_async(self, frame):
        """Upload without waiting (at most one outstanding; see tf_gpu.h)."""
        cf = frame.c_frame()
        self._check(self.lib.tf_gpu_cache_frame_async(self.h, C.byref(cf)))

    def debug_read_plane(self, frame, plane, x0, y0, w, h):
        """Rectangle of a cached device plane, border included (test hook for the upload path)."""
        dst = np.zeros((h, w), frame.dtype)
        self._check(self.lib.tf_gpu_debug_read_plane(self.h, frame.frame_id, plane, dst.ctypes.data, w, x0, y0, w, h))
        return dst

    def fullpel_search_batch(self, params, src, ref, block_size, items):
        """items: sequence of (x, y, start_row, start_col); returns an int array [n, 3] of (row, col, var) --
        av1_full_pixel_search() as tf_motion_search() configures it, per block (see tf_gpu.h)."""
        cp = make_params(params) if isinstance(params, dict) else params
        n = len(items)
        arr = (SearchItem * max(n, 1))(*[SearchItem(*map(int, it)) for it in items])
        res = (SearchResult * max(n, 1))()
        cs, cr = src.c_frame(), ref.c_frame()
        self._check(self.lib.tf_gpu_fullpel_search_batch(self.h, C.byref(cp), C.byref(cs), C.byref(cr), block_size, arr, n, res))
        return np.array([(r.row, r.col, r.var) for r in res[:n]], np.int64).reshape(n, 3)

    def device_border(self):
        return self.lib.tf_gpu_device_border()

    def evict_frame(self, frame_id):
        self._check(self.lib.tf_gpu_evict_frame(self.h, frame_id))

    def filter_resident(self, params, frame_ids):
        cp = make_params(params) if isinstance(params, dict) else params
        ids = (C.c_uint64 * len(frame_ids))(*frame_ids)
        diff = (C.c_int64 * 2)()
        ms = C.c_float()
        self._check(self.lib.tf_gpu_filter_resident(self.h, C.byref(cp), ids, diff, C.byref(ms)))
        return ms.value, np.array([diff[0], diff[1]], np.int64)

    def filter_resident_async(self, params, frame_ids):
        cp = make_params(params) if isinstance(params, dict) else params
        ids = (C.c_uint64 * len(frame_ids))(*frame_ids)
        self._check(self.lib.tf_gpu_filter_resident_async(self.h, C.byref(cp), ids))

    def filter_resident_result(self):
        diff = (C.c_int64 * 2)()
        ms = C.c_float()
        self._check(self.lib.tf_gpu_filter_resident_result(self.h, diff, C.byref(ms)))
        return ms.value, np.array([diff[0], diff[1]], np.int64)

    def download_output(self, out, row_begin=0, row_end=0):
        co = out.c_frame()
        self._check(self.lib.tf_gpu_download_output(self.h, C.byref(co), row_begin, row_end))

    def output_device_plane(self, plane):
        p, pitch, rows, rb = C.c_void_p(), C.c_size_t(), C.c_int(), C.c_int()
        self._check(self.lib.tf_gpu_output_device_plane(self.h, plane, C.byref(p), C.byref(pitch), C.byref(rows),
                                                        C.byref(rb)))
        return p.value, pitch.value, rows.value, rb.value

    def output_ipc_export(self, plane):
        """(handle bytes, byte offset of pixel (0,0), pitch bytes) of an output plane, for another rank's import."""
        buf = C.create_string_buffer(64)
        off, pitch = C.c_size_t(), C.c_size_t()
        self._check(self.lib.tf_gpu_output_ipc_export(self.h, plane, buf, C.byref(off), C.byref(pitch)))
        return buf.raw, off.value, pitch.value

    def output_ipc_import(self, plane, handle, offset=0, pitch=0):
        """Store this context's output rows into another rank's plane (handle=None: back to its own planes)."""
        self._check(self.lib.tf_gpu_output_ipc_import(self.h, plane, handle, offset, pitch))

    def host_register(self, arr):
        self._check(self.lib.tf_gpu_host_register(self.h, arr.ctypes.data, arr.nbytes))

    def host_unregister(self, arr):
        self._check(self.lib.tf_gpu_host_unregister(self.h, arr.ctypes.data))

    def event_record(self, slot):
        self._check(self.lib.tf_gpu_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float()
        self._check(self.lib.tf_gpu_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def synchronize(self):
        self._check(self.lib.tf_gpu_synchronize(self.h))

    def microbench(self, kind):
        v = C.c_double()
        self._check(self.lib.tf_gpu_microbench(self.h, kind, C.byref(v)))
        return v.value

    def collect_counters(self, enable):
        self._check(self.lib.tf_gpu_collect_counters(self.h, int(enable)))

    def read_counters(self):
        c = (C.c_uint64 * 4)()
        self._check(self.lib.tf_gpu_read_counters(self.h, c))
        return [int(x) for x in c]

    def last_kernel_times(self):
        ms = (C.c_float * 3)()
        self._check(self.lib.tf_gpu_last_kernel_times(self.h, ms))
        return [ms[0], ms[1], ms[2]]

    def last_stats(self):
        n, ms = C.c_int(), C.c_float()
        self._check(self.lib.tf_gpu_last_stats(self.h, C.byref(n), C.byref(ms)))
        return n.value, ms.value
