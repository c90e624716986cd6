"""jpegb200 -- Python (ctypes) binding of libjpegb200.so, the B200-native baseline-JPEG
encode path with the per-stage function surface of rusty-electron/jpeg-encoder-opencl
(reference: src/utils.hpp:77-137, driver order src/OpenCLProject_JpegEncoder.cpp:59-225).

The directory name contains a hyphen, so import it through `load()` in
__graft_entry__.py / tests/conftest.py (module name `jpegb200`).  This module is a thin
binding: all work happens in the CUDA library, and there is no CPU fallback -- a missing
library or a missing GPU raises.
"""
import ctypes as C
import os
import subprocess
import weakref

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libjpegb200.so")

SUB_444, SUB_REPL420, SUB_420 = 0, 1, 2
FLAG_REF_INPLACE_DCT, FLAG_REF_TYPO_TABLES, FLAG_REF_ALWAYS_EOB = 1, 2, 4
FLAG_CLAMP_SOF, FLAG_NO_TIE_FIXUP, FLAG_TENSOR_DCT, FLAG_FMA_DCT = 8, 16, 32, 64
FLAG_OPTIMIZE_HUFFMAN = 128
FLAG_TMA = 256
FLAG_ENTROPY_LDG = 512
OK, E_INVALID, E_CUDA, E_NOSPACE, E_NOMEM, E_UNSUPPORTED, E_INTERNAL = range(7)

# every symbol include/jpegb200.h declares (tests check that the library exports them all)
SYMBOLS = [
    "jb_create", "jb_destroy", "jb_last_error", "jb_sync", "jb_stream", "jb_set_profiling", "jb_get_timings",
    "jb_reset_counters", "jb_version", "jb_host_alloc", "jb_host_free", "jb_device_alloc", "jb_device_free",
    "jb_memcpy_h2d", "jb_memcpy_d2h", "jb_csc_rgb8_aos", "jb_cds_aos", "jb_padded_size", "jb_pad_mirror_aos",
    "jb_u8_to_f64", "jb_levelshift_f64", "jb_dct_f64", "jb_quantize_f64", "jb_blockify", "jb_zigzag", "jb_rle",
    "jb_huffman", "jb_quality_tables", "jb_num_mcus", "jb_blocks_per_mcu", "jb_header_bytes", "jb_required_bytes",
    "jb_transform", "jb_entropy", "jb_encode_jfif", "jb_encode_batch", "jb_encode_batch_device", "jb_encode_strip",
    "jb_write_header", "jb_synth_rgb_device", "jb_planar_u32_from_aos", "jb_planar_u32_interleave",
    "jb_planar_u32_to_rgb8_device", "jb_encode_jfif_planar_u32", "jb_optimal_huffman_spec",
    "jb_encode_strip_begin", "jb_encode_strip_finish", "jb_copy_bytes_device", "jb_ipc_export", "jb_ipc_open", "jb_ipc_close",
    "jb_stitch_exchange", "jb_stitch_complete", "jb_encode_tiles", "jb_encode_nv12_device", "jb_encode_nv12_batch", "jb_rgb8_to_nv12_device",
    "jb_jfif_info_host", "jb_jfif_info_device", "jb_decode_jfif_device", "jb_decode_jfif", "jb_psnr_device",
    "jb_pad_mirror_planar_u32", "jb_blockify_planar_i32", "jb_f64_to_u8", "jb_remove_red_aos", "jb_value_categories",
]


class JbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"jpegb200 error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("subsampling", C.c_int32), ("restart_interval", C.c_int32), ("flags", C.c_uint32),
                ("qlum", C.c_uint32 * 64), ("qchrom", C.c_uint32 * 64)]


class JfifInfo(C.Structure):
    _fields_ = [("W", C.c_uint32), ("H", C.c_uint32), ("subsampling", C.c_int32), ("restart_interval", C.c_uint32),
                ("scan_offset", C.c_uint64)]


class Timings(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("CSCTime", "CDSTime", "levelShiftTime", "DCTTime", "QuantTime",
                                          "TotalCopyTime", "zigZagTime", "RLETime", "HuffmanTime", "transform_us",
                                          "fixup_us", "entropy_us", "h2d_us", "d2h_us", "edge_us")] + \
               [(n, C.c_uint64) for n in ("transform_launches", "total_launches", "tie_fixups")] + \
               [(n, C.c_double) for n in ("staged_dct_us", "staged_copy_us", "staged_huffman_us")]


def build(verbose=False):
    """Compile libjpegb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", HERE, "-j8"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libjpegb200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))


_lib = None


def lib():
    """The loaded C library (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, sz, u64 = C.c_void_p, C.c_size_t, C.c_uint64
    PP = C.POINTER(Params)
    L.jb_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.jb_destroy.argtypes = [vp]
    L.jb_destroy.restype = None
    L.jb_last_error.argtypes = [vp]
    L.jb_last_error.restype = C.c_char_p
    L.jb_sync.argtypes = [vp]
    L.jb_stream.argtypes = [vp]
    L.jb_stream.restype = vp
    L.jb_set_profiling.argtypes = [vp, C.c_int]
    L.jb_get_timings.argtypes = [vp, C.POINTER(Timings)]
    L.jb_reset_counters.argtypes = [vp]
    L.jb_host_alloc.argtypes = [C.POINTER(vp), sz]
    L.jb_host_free.argtypes = [vp]
    L.jb_device_alloc.argtypes = [vp, C.POINTER(vp), sz]
    L.jb_device_free.argtypes = [vp, vp]
    L.jb_memcpy_h2d.argtypes = [vp, vp, vp, sz]
    L.jb_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    L.jb_csc_rgb8_aos.argtypes = [vp, vp, sz, sz]
    L.jb_cds_aos.argtypes = [vp, vp, sz, sz]
    L.jb_padded_size.argtypes = [sz, sz, sz, C.POINTER(sz), C.POINTER(sz)]
    L.jb_pad_mirror_aos.argtypes = [vp, vp, sz, sz, vp, sz, sz]
    L.jb_u8_to_f64.argtypes = [vp, vp, vp, sz]
    L.jb_levelshift_f64.argtypes = [vp, vp, sz, C.c_double]
    L.jb_dct_f64.argtypes = [vp, vp, sz, sz, C.c_uint32]
    L.jb_quantize_f64.argtypes = [vp, vp, sz, sz, vp, vp]
    L.jb_blockify.argtypes = [vp, vp, sz, sz, vp]
    L.jb_zigzag.argtypes = [vp, vp, vp, sz]
    L.jb_rle.argtypes = [vp, vp, sz, C.c_uint32, vp, vp]
    L.jb_planar_u32_from_aos.argtypes = [vp, vp, sz, sz, vp]
    L.jb_planar_u32_interleave.argtypes = [vp, vp, sz, sz, vp]
    L.jb_planar_u32_to_rgb8_device.argtypes = [vp, vp, sz, sz, vp, sz]
    L.jb_encode_jfif_planar_u32.argtypes = [vp, vp, sz, sz, PP, vp, sz, C.POINTER(sz)]
    L.jb_pad_mirror_planar_u32.argtypes = [vp, vp, sz, sz, vp, sz, sz]
    L.jb_blockify_planar_i32.argtypes = [vp, vp, sz, sz, vp]
    L.jb_f64_to_u8.argtypes = [vp, vp, vp, sz]
    L.jb_remove_red_aos.argtypes = [vp, vp, sz, sz]
    L.jb_value_categories.argtypes = [vp, vp, sz, vp, vp]
    L.jb_huffman.argtypes = [vp, vp, sz, C.c_uint32, vp, sz, C.POINTER(u64)]
    L.jb_quality_tables.argtypes = [C.c_int, vp, vp]
    L.jb_optimal_huffman_spec.argtypes = [vp, vp, vp, C.POINTER(C.c_int)]
    L.jb_num_mcus.argtypes = [sz, sz, C.c_int]
    L.jb_num_mcus.restype = sz
    L.jb_blocks_per_mcu.argtypes = [C.c_int]
    L.jb_header_bytes.argtypes = [PP]
    L.jb_header_bytes.restype = sz
    L.jb_required_bytes.argtypes = [vp]
    L.jb_required_bytes.restype = sz
    L.jb_transform.argtypes = [vp, vp, sz, sz, sz, PP, vp]
    L.jb_entropy.argtypes = [vp, vp, sz, PP, vp, sz, C.POINTER(sz)]
    L.jb_encode_jfif.argtypes = [vp, vp, sz, sz, sz, PP, vp, sz, C.POINTER(sz)]
    L.jb_encode_batch.argtypes = [vp, vp, sz, sz, sz, sz, sz, PP, vp, sz, vp, vp]
    L.jb_encode_nv12_device.argtypes = [vp, vp, sz, sz, vp, sz, sz, sz, sz, sz, PP, vp, sz, vp, vp, vp]
    L.jb_encode_nv12_batch.argtypes = [vp, vp, sz, sz, vp, sz, sz, sz, sz, sz, PP, vp, sz, vp, vp]
    L.jb_rgb8_to_nv12_device.argtypes = [vp, vp, sz, sz, sz, vp, sz, vp, sz]
    L.jb_jfif_info_host.argtypes = [vp, sz, C.POINTER(JfifInfo)]
    L.jb_jfif_info_device.argtypes = [vp, vp, sz, C.POINTER(JfifInfo)]
    L.jb_decode_jfif_device.argtypes = [vp, vp, sz, vp, sz, vp]
    L.jb_decode_jfif.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz), C.POINTER(sz)]
    L.jb_psnr_device.argtypes = [vp, vp, sz, vp, sz, sz, sz, C.POINTER(C.c_double), C.POINTER(u64)]
    L.jb_encode_tiles.argtypes = [vp, vp, sz, sz, sz, sz, sz, PP, vp, sz, vp, vp, C.POINTER(sz)]
    L.jb_encode_batch_device.argtypes = [vp, vp, sz, sz, sz, sz, sz, PP, vp, sz, vp, vp, vp]
    L.jb_encode_strip.argtypes = [vp, vp, sz, sz, sz, PP, u64, C.c_int, C.c_int, vp, sz, C.POINTER(sz)]
    L.jb_encode_strip_begin.argtypes = [vp, vp, sz, sz, sz, PP, u64, C.c_int, vp]
    L.jb_encode_strip_finish.argtypes = [vp, vp, sz, vp]
    L.jb_copy_bytes_device.argtypes = [vp, vp, sz, vp, vp, vp]
    L.jb_stitch_exchange.argtypes = [vp, vp, C.c_int, C.c_int, u64, u64, vp, vp]
    L.jb_stitch_complete.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, u64]
    L.jb_ipc_export.argtypes = [vp, vp, vp]
    L.jb_ipc_open.argtypes = [vp, vp, C.POINTER(vp)]
    L.jb_ipc_close.argtypes = [vp, vp]
    L.jb_write_header.argtypes = [PP, sz, sz, vp, sz, C.POINTER(sz)]
    L.jb_synth_rgb_device.argtypes = [vp, u64, sz, sz, sz, sz, vp]
    _lib = L
    return L


def quality_tables(quality):
    """IJG-scaled versions of the reference's q50 tables (utils.hpp:42-62)."""
    ql = np.zeros(64, np.uint32)
    qc = np.zeros(64, np.uint32)
    lib().jb_quality_tables(int(quality), ql.ctypes.data, qc.ctypes.data)
    return ql, qc


def make_params(subsampling=SUB_420, quality=None, qlum=None, qchrom=None, restart_interval=0, flags=0):
    p = Params()
    p.subsampling, p.restart_interval, p.flags = subsampling, restart_interval, flags
    if quality is not None:
        qlum, qchrom = quality_tables(quality)
    for i in range(64):
        p.qlum[i] = int(qlum[i])
        p.qchrom[i] = int(qchrom[i])
    return p


def optimal_huffman_spec(counts):
    """(bits[16], vals[n]) of the optimal Huffman table for 256 symbol counts (host only)."""
    c = np.ascontiguousarray(counts, np.uint64)
    bits, vals, n = np.zeros(16, np.uint8), np.zeros(256, np.uint8), C.c_int()
    rc = lib().jb_optimal_huffman_spec(c.ctypes.data, bits.ctypes.data, vals.ctypes.data, C.byref(n))
    if rc:
        raise JbError(rc, "jb_optimal_huffman_spec")
    return bits, vals[: n.value]


def jfif_info(data):
    """Host-only marker parser: (W, H, subsampling, restart_interval, scan_offset) of a baseline JFIF file."""
    buf = np.frombuffer(bytes(data), np.uint8)
    info = JfifInfo()
    rc = lib().jb_jfif_info_host(buf.ctypes.data, buf.size, C.byref(info))
    if rc:
        raise JbError(rc, "jb_jfif_info_host: not a baseline 3-component JFIF file the decoder handles")
    return info


def header_bytes(params):
    return lib().jb_header_bytes(C.byref(params))


def _ptr(a):
    return a.ctypes.data if isinstance(a, np.ndarray) else int(a)


class Encoder:
    """One jb_ctx on one GPU.  Methods mirror the reference's stage functions (utils.hpp:81-137)."""

    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.jb_create(device, C.byref(h))
        if rc != OK:
            raise JbError(rc, "jb_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.jb_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc != OK:
            raise JbError(rc, self.L.jb_last_error(self.h).decode())

    # ---- staged: the reference's per-stage functions ------------------------------------
    def performCSC(self, img):
        """In place on an (H, W, 3) uint8 array; utils.hpp:81."""
        assert img.dtype == np.uint8 and img.flags.c_contiguous
        self._ck(self.L.jb_csc_rgb8_aos(self.h, _ptr(img), img.shape[1], img.shape[0]))
        return img

    def performCDS(self, img):
        assert img.dtype == np.uint8 and img.flags.c_contiguous
        self._ck(self.L.jb_cds_aos(self.h, _ptr(img), img.shape[1], img.shape[0]))
        return img

    def padMirror(self, img, mult=8):
        H, W, _ = img.shape
        nW, nH = -(-W // mult) * mult, -(-H // mult) * mult
        out = np.empty((nH, nW, 3), np.uint8)
        self._ck(self.L.jb_pad_mirror_aos(self.h, _ptr(np.ascontiguousarray(img)), W, H, _ptr(out), nW, nH))
        return out

    def copyUIntToDoubleImage(self, img):
        out = np.empty(img.shape, np.float64)
        self._ck(self.L.jb_u8_to_f64(self.h, _ptr(np.ascontiguousarray(img)), _ptr(out), img.size))
        return out

    def substractfromAll(self, imgd, val=128.0):
        self._ck(self.L.jb_levelshift_f64(self.h, _ptr(imgd), imgd.size, float(val)))
        return imgd

    def performDCT(self, imgd, flags=0):
        self._ck(self.L.jb_dct_f64(self.h, _ptr(imgd), imgd.shape[1], imgd.shape[0], flags))
        return imgd

    def performQuantization(self, imgd, qlum, qchrom):
        ql = np.ascontiguousarray(qlum, np.uint32)
        qc = np.ascontiguousarray(qchrom, np.uint32)
        self._ck(self.L.jb_quantize_f64(self.h, _ptr(imgd), imgd.shape[1], imgd.shape[0], _ptr(ql), _ptr(qc)))
        return imgd

    def everyMCUisnow2DArray(self, imgd):
        H, W, _ = imgd.shape
        out = np.empty((3 * W * H // 64, 64), np.int32)
        self._ck(self.L.jb_blockify(self.h, _ptr(imgd), W, H, _ptr(out)))
        return out

    def performZigZag(self, linear):
        out = np.empty_like(linear)
        self._ck(self.L.jb_zigzag(self.h, _ptr(np.ascontiguousarray(linear)), _ptr(out), linear.shape[0]))
        return out

    def performRLE(self, zz, flags=0):
        """Returns a list of 1-D int32 arrays (run, value, ...) like vector<vector<int>>; utils.hpp:132."""
        rows = zz.shape[0]
        pairs = np.empty((rows, 128), np.int32)
        counts = np.empty(rows, np.uint32)
        self._ck(self.L.jb_rle(self.h, _ptr(np.ascontiguousarray(zz)), rows, flags, _ptr(pairs), _ptr(counts)))
        return [pairs[i, :counts[i]].copy() for i in range(rows)]

    def removeRedChannel(self, img):
        """In place; utils.hpp:79."""
        assert img.dtype == np.uint8 and img.flags.c_contiguous
        self._ck(self.L.jb_remove_red_aos(self.h, _ptr(img), img.shape[1], img.shape[0]))
        return img

    def copyDoubleToUIntImage(self, imgd):
        out = np.empty(imgd.shape, np.uint8)
        self._ck(self.L.jb_f64_to_u8(self.h, _ptr(np.ascontiguousarray(imgd, np.float64)), _ptr(out), imgd.size))
        return out

    def valueCategories(self, values):
        """(categories uint8, value bits uint16) of int16 values: getValueCategory / valueToBitString, utils.hpp:134-135."""
        v = np.ascontiguousarray(values, np.int16).ravel()
        cat, bits = np.empty(v.size, np.uint8), np.empty(v.size, np.uint16)
        self._ck(self.L.jb_value_categories(self.h, _ptr(v), v.size, _ptr(cat), _ptr(bits)))
        return cat, bits

    def copyOntoLargerVectorWithPadding(self, planar, W, H, nW, nH):
        out = np.empty(3 * nW * nH, np.uint32)
        self._ck(self.L.jb_pad_mirror_planar_u32(self.h, _ptr(np.ascontiguousarray(planar, np.uint32)), W, H, _ptr(out), nW, nH))
        return out

    def everyMCUisnow1DArray(self, planar_i32, W, H):
        out = np.empty((3 * W * H // 64, 64), np.int32)
        self._ck(self.L.jb_blockify_planar_i32(self.h, _ptr(np.ascontiguousarray(planar_i32, np.int32)), W, H, _ptr(out)))
        return out

    def HuffmanEncoder(self, zz, rows_per_channel, flags=0):
        """Returns (packed MSB-first bytes, nbits): the bit sequence of utils.hpp:137."""
        cap = zz.shape[0] * 64 * 4 + 64
        out = np.zeros(cap, np.uint8)
        nbits = C.c_uint64()
        self._ck(self.L.jb_huffman(self.h, _ptr(np.ascontiguousarray(zz, np.int32)), rows_per_channel, flags, _ptr(out),
                                   cap, C.byref(nbits)))
        return out[: (nbits.value + 7) // 8].copy(), nbits.value

    # ---- fused ---------------------------------------------------------------------------
    def transform(self, rgb, params):
        """(H, W, 3) uint8 -> int16 [n_mcu, blocks_per_mcu, 64] zigzag coefficients in scan order."""
        H, W, _ = rgb.shape
        rgb = np.ascontiguousarray(rgb)
        n = self.L.jb_num_mcus(W, H, params.subsampling)
        out = np.empty((n, self.L.jb_blocks_per_mcu(params.subsampling), 64), np.int16)
        self._ck(self.L.jb_transform(self.h, _ptr(rgb), W, H, W * 3, C.byref(params), _ptr(out)))
        return out

    def entropy(self, coef, params):
        coef = np.ascontiguousarray(coef, np.int16)
        cap = coef.size * 4 + 4096
        out = np.empty(cap, np.uint8)
        n = C.c_size_t()
        self._ck(self.L.jb_entropy(self.h, _ptr(coef), coef.shape[0], C.byref(params), _ptr(out), cap, C.byref(n)))
        return out[: n.value].copy()

    def encode_jfif(self, rgb, params, cap=None):
        H, W, _ = rgb.shape
        rgb = np.ascontiguousarray(rgb)
        cap = cap if cap is not None else W * H * 3 + 65536
        out = np.empty(cap, np.uint8)
        n = C.c_size_t()
        self._ck(self.L.jb_encode_jfif(self.h, _ptr(rgb), W, H, W * 3, C.byref(params), _ptr(out), cap, C.byref(n)))
        return out[: n.value].tobytes()

    # the planar uint32 image layout of the reference's OpenCL half (utils.hpp:116-119)
    def copyImageToVector(self, rgb):
        H, W, _ = rgb.shape
        out = np.empty(3 * W * H, np.uint32)
        self._ck(self.L.jb_planar_u32_from_aos(self.h, _ptr(np.ascontiguousarray(rgb)), W, H, _ptr(out)))
        return out

    def switchVectorChannelOrdering(self, planar, W, H):
        out = np.empty(3 * W * H, np.uint32)
        self._ck(self.L.jb_planar_u32_interleave(self.h, _ptr(np.ascontiguousarray(planar)), W, H, _ptr(out)))
        return out

    def planar_u32_to_rgb8_device(self, d_planar, W, H, d_rgb, pitch):
        self._ck(self.L.jb_planar_u32_to_rgb8_device(self.h, d_planar, W, H, d_rgb, pitch))

    def encode_jfif_planar_u32(self, planar, W, H, params, cap=None):
        cap = cap if cap is not None else W * H * 3 + 65536
        out = np.empty(cap, np.uint8)
        n = C.c_size_t()
        self._ck(self.L.jb_encode_jfif_planar_u32(self.h, _ptr(np.ascontiguousarray(planar, dtype=np.uint32)), W, H,
                                                  C.byref(params), _ptr(out), cap, C.byref(n)))
        return out[: n.value].tobytes()

    def encode_batch(self, frames, params, out=None):
        """frames: (N, H, W, 3) uint8 (numpy, ideally pinned).  Returns (out, offsets, sizes)."""
        N, H, W, _ = frames.shape
        assert frames.flags.c_contiguous
        if out is None:
            out = np.empty(N * (W * H * 3 + 65536) if N * W * H < (1 << 26) else N * (W * H + 4096), np.uint8)
        offs = np.zeros(N, np.uint64)
        sizes = np.zeros(N, np.uint64)
        self._ck(self.L.jb_encode_batch(self.h, _ptr(frames), N, W, H, W * 3, W * H * 3, C.byref(params), _ptr(out),
                                        out.size, _ptr(offs), _ptr(sizes)))
        return out, offs, sizes

    def encode_tiles(self, rgb, tile_w, tile_h, params, cap=None):
        """(H, W, 3) uint8 of any size -> list of rows of JFIF files (bytes), tiles of tile_w x tile_h pixels."""
        H, W, _ = rgb.shape
        rgb = np.ascontiguousarray(rgb)
        ntx, nty = -(-W // tile_w), -(-H // tile_h)
        cap = cap if cap is not None else W * H * 3 + ntx * nty * 65536
        out = np.empty(cap, np.uint8)
        offs, sizes, n = np.zeros(ntx * nty, np.uint64), np.zeros(ntx * nty, np.uint64), C.c_size_t()
        self._ck(self.L.jb_encode_tiles(self.h, _ptr(rgb), W, H, W * 3, tile_w, tile_h, C.byref(params), _ptr(out), cap, _ptr(offs),
                                        _ptr(sizes), C.byref(n)))
        assert n.value == ntx * nty
        return [[out[int(offs[ty * ntx + tx]): int(offs[ty * ntx + tx] + sizes[ty * ntx + tx])].tobytes() for tx in range(ntx)]
                for ty in range(nty)]

    def encode_batch_ptr(self, rgb_ptr, N, W, H, pitch, frame_stride, params, out_ptr, cap, offs, sizes):
        self._ck(self.L.jb_encode_batch(self.h, rgb_ptr, N, W, H, pitch, frame_stride, C.byref(params), out_ptr, cap,
                                        _ptr(offs), _ptr(sizes)))

    def encode_batch_device(self, d_rgb, N, W, H, pitch, frame_stride, params, d_out, cap, d_offs, d_sizes, d_total):
        """All pointers are device addresses (ints); asynchronous, finish with sync()."""
        self._ck(self.L.jb_encode_batch_device(self.h, d_rgb, N, W, H, pitch, frame_stride, C.byref(params), d_out, cap,
                                               d_offs, d_sizes, d_total))

    def encode_nv12_device(self, d_y, pitch_y, fs_y, d_uv, pitch_uv, fs_uv, N, W, H, params, d_out, cap, d_offs, d_sizes, d_total):
        """NV12-style frames in HBM (device addresses as ints) -> JFIF files; asynchronous, finish with sync()."""
        self._ck(self.L.jb_encode_nv12_device(self.h, d_y, pitch_y, fs_y, d_uv, pitch_uv, fs_uv, N, W, H, C.byref(params), d_out, cap,
                                              d_offs, d_sizes, d_total))

    def encode_nv12_batch(self, y, uv, W, params, out=None):
        """NV12-style HOST frames: y (N, H, pitch_y >= W) uint8, uv (N, ceil(H/2), pitch_uv >= 2*ceil(W/2)) uint8 (numpy, ideally
        pinned) -> (out, offsets, sizes) as encode_batch."""
        N, H, py = y.shape
        _, ch, puv = uv.shape
        assert y.flags.c_contiguous and uv.flags.c_contiguous and uv.shape[0] == N and ch == (H + 1) // 2
        if out is None:
            out = np.empty(N * (W * H * 3 + 65536) if N * W * H < (1 << 26) else N * (W * H + 4096), np.uint8)
        offs = np.zeros(N, np.uint64)
        sizes = np.zeros(N, np.uint64)
        self._ck(self.L.jb_encode_nv12_batch(self.h, _ptr(y), py, py * H, _ptr(uv), puv, puv * ch, N, W, H, C.byref(params), _ptr(out),
                                             out.size, _ptr(offs), _ptr(sizes)))
        return out, offs, sizes

    def encode_nv12_batch_ptr(self, y_ptr, pitch_y, fs_y, uv_ptr, pitch_uv, fs_uv, N, W, H, params, out_ptr, cap, offs, sizes):
        self._ck(self.L.jb_encode_nv12_batch(self.h, y_ptr, pitch_y, fs_y, uv_ptr, pitch_uv, fs_uv, N, W, H, C.byref(params), out_ptr, cap,
                                             _ptr(offs), _ptr(sizes)))

    def rgb8_to_nv12_device(self, d_rgb, W, H, pitch, d_y, pitch_y, d_uv, pitch_uv):
        self._ck(self.L.jb_rgb8_to_nv12_device(self.h, d_rgb, W, H, pitch, d_y, pitch_y, d_uv, pitch_uv))

    def encode_strip(self, rgb, params, first_interval, last_strip, W=None, rows=None, pitch=None, device_io=False,
                     out=None, cap=None):
        if not device_io:
            rows, W, _ = rgb.shape
            rgb = np.ascontiguousarray(rgb)
            pitch = W * 3
            cap = W * rows * 3 + 65536
            out = np.empty(cap, np.uint8)
        n = C.c_size_t()
        self._ck(self.L.jb_encode_strip(self.h, _ptr(rgb), W, rows, pitch, C.byref(params), first_interval,
                                        int(last_strip), int(device_io), _ptr(out), cap, C.byref(n)))
        return out[: n.value].copy() if not device_io else n.value

    def encode_strip_begin(self, d_rgb, params, first_interval, last_strip, W, rows, pitch, d_len):
        """Asynchronous first half of a strip (device pointers as ints): *d_len = its byte count."""
        self._ck(self.L.jb_encode_strip_begin(self.h, d_rgb, W, rows, pitch, C.byref(params), first_interval, int(last_strip), d_len))

    def encode_strip_finish(self, d_out, cap, d_off):
        """Asynchronous second half: the strip is written to d_out + *d_off (d_out may be peer memory)."""
        self._ck(self.L.jb_encode_strip_finish(self.h, d_out, cap, d_off))

    def copy_bytes_device(self, d_dst, cap, d_dst_off, d_src, d_len):
        self._ck(self.L.jb_copy_bytes_device(self.h, d_dst, cap, d_dst_off, d_src, d_len))

    def stitch_exchange(self, d_ctl, rank, world, epoch, base, d_len, d_off):
        self._ck(self.L.jb_stitch_exchange(self.h, d_ctl, rank, world, epoch, base, d_len, d_off))

    def stitch_complete(self, d_ctl, rank, world, dst, epoch):
        self._ck(self.L.jb_stitch_complete(self.h, d_ctl, rank, world, dst, epoch))

    def ipc_export(self, d_ptr):
        h = np.zeros(64, np.uint8)
        self._ck(self.L.jb_ipc_export(self.h, d_ptr, _ptr(h)))
        return h

    def ipc_open(self, handle):
        p = C.c_void_p()
        self._ck(self.L.jb_ipc_open(self.h, _ptr(np.ascontiguousarray(handle, np.uint8)), C.byref(p)))
        return p.value

    def ipc_close(self, d_ptr):
        self._ck(self.L.jb_ipc_close(self.h, d_ptr))

    # ---- decode path ---------------------------------------------------------------------
    def decode_jfif(self, data):
        """JFIF bytes -> (H, W, 3) uint8 RGB, decoded on the GPU (libjpeg-exact reconstruction)."""
        buf = np.frombuffer(bytes(data), np.uint8)
        W, H = C.c_size_t(), C.c_size_t()
        rc = self.L.jb_decode_jfif(self.h, _ptr(buf), buf.size, None, 0, C.byref(W), C.byref(H))
        if rc != E_NOSPACE:
            self._ck(rc if rc else E_INTERNAL)
        out = np.empty((H.value, W.value, 3), np.uint8)
        self._ck(self.L.jb_decode_jfif(self.h, _ptr(buf), buf.size, _ptr(out), out.size, C.byref(W), C.byref(H)))
        return out

    def jfif_info_device(self, d_jfif, length):
        info = JfifInfo()
        self._ck(self.L.jb_jfif_info_device(self.h, d_jfif, length, C.byref(info)))
        return info

    def decode_jfif_device(self, d_jfif, length, d_rgb=None, pitch=0, d_coef=None):
        self._ck(self.L.jb_decode_jfif_device(self.h, d_jfif, length, d_rgb, pitch, d_coef))

    def psnr_device(self, d_a, pitch_a, d_b, pitch_b, W, H):
        """(PSNR in dB, sum of squared differences) of two RGB8 images in HBM."""
        ps, se = C.c_double(), C.c_uint64()
        self._ck(self.L.jb_psnr_device(self.h, d_a, pitch_a, d_b, pitch_b, W, H, C.byref(ps), C.byref(se)))
        return ps.value, se.value

    def write_header(self, params, W, H):
        out = np.empty(1024, np.uint8)
        n = C.c_size_t()
        rc = self.L.jb_write_header(C.byref(params), W, H, _ptr(out), 1024, C.byref(n))
        assert rc == OK
        return out[: n.value].tobytes()

    def synth_device(self, seed, W, y0, rows, pitch, d_out):
        self._ck(self.L.jb_synth_rgb_device(self.h, seed, W, y0, rows, pitch, d_out))

    def synth(self, seed, W, H):
        """Synthetic image generated on the GPU, returned as numpy (tests)."""
        d = self.device_alloc(W * H * 3)
        try:
            self.synth_device(seed, W, 0, H, W * 3, d)
            out = np.empty((H, W, 3), np.uint8)
            self._ck(self.L.jb_memcpy_d2h(self.h, _ptr(out), d, out.size))
        finally:
            self.device_free(d)
        return out

    # ---- plumbing ------------------------------------------------------------------------
    def sync(self):
        self._ck(self.L.jb_sync(self.h))

    def stream(self):
        return self.L.jb_stream(self.h)

    def set_profiling(self, on):
        self._ck(self.L.jb_set_profiling(self.h, int(on)))

    def reset_counters(self):
        self._ck(self.L.jb_reset_counters(self.h))

    def timings(self):
        t = Timings()
        self._ck(self.L.jb_get_timings(self.h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in Timings._fields_}

    def device_alloc(self, nbytes):
        p = C.c_void_p()
        self._ck(self.L.jb_device_alloc(self.h, C.byref(p), nbytes))
        return p.value

    def device_free(self, p):
        self._ck(self.L.jb_device_free(self.h, p))

    def h2d(self, dptr, arr):
        self._ck(self.L.jb_memcpy_h2d(self.h, dptr, _ptr(arr), arr.nbytes))

    def d2h(self, arr, dptr, nbytes=None):
        self._ck(self.L.jb_memcpy_d2h(self.h, _ptr(arr), dptr, arr.nbytes if nbytes is None else nbytes))


def pinned_empty(shape, dtype=np.uint8):
    """numpy array over cudaMallocHost memory; the block is freed (jb_host_free) once the array and all its
    views have been collected."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    if lib().jb_host_alloc(C.byref(p), max(n, 1)) != OK:
        raise MemoryError("jb_host_alloc failed")
    buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
    # views keep `buf` alive through their .base chain; when the last one is collected the pinned block is released
    weakref.finalize(buf, lib().jb_host_free, p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr
