"""ctypes bindings for the TEST-ONLY oracle (oracle/_build/liboracle.so) and, when it
was built, the reference's own code (oracle/_ref/libjpegref.so).  Used by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libjpegref.so")
REF_O0_SO = os.path.join(ROOT, "oracle", "_ref", "libjpegref_O0.so")

Q1, Q2, Q3 = 1, 2, 4
AS_WRITTEN = Q1 | Q2 | Q3
SUB_444, SUB_REPL420, SUB_420 = 0, 1, 2

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
sz = C.c_size_t


def build_oracle():
    """Compile oracle/ (and oracle/_ref when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)


_oracle = None
_ref = {}


def oracle():
    global _oracle
    if _oracle is not None:
        return _oracle
    if not os.path.exists(ORACLE_SO):
        build_oracle()
    L = C.CDLL(ORACLE_SO)
    L.orc_init.restype = None
    L.orc_csc.argtypes = [u8p, sz]
    L.orc_cds.argtypes = [u8p, sz, sz]
    L.orc_padded_size.argtypes = [sz, sz, sz, C.POINTER(sz), C.POINTER(sz)]
    L.orc_pad_mirror.argtypes = [u8p, sz, sz, u8p, sz, sz]
    L.orc_u8_to_double.argtypes = [u8p, f64p, sz]
    L.orc_subtract.argtypes = [f64p, sz, C.c_double]
    L.orc_dct_image.argtypes = [f64p, sz, sz, C.c_int]
    L.orc_quantize_image.argtypes = [f64p, sz, sz, u32p, u32p]
    L.orc_blockify.argtypes = [f64p, sz, sz, i32p]
    L.orc_zigzag.argtypes = [i32p, i32p, sz]
    L.orc_rle_block.argtypes = [i32p, i32p, C.c_int]
    L.orc_rle_block.restype = sz
    L.orc_category.argtypes = [C.c_int]
    L.orc_value_bits.argtypes = [C.c_int, C.POINTER(C.c_uint32)]
    L.orc_huffman_ref.argtypes = [i32p, sz, C.c_int, u8p, sz]
    L.orc_huffman_ref.restype = C.c_uint64
    L.orc_table_code.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32)]
    L.orc_quality_tables.argtypes = [C.c_int, u32p, u32p]
    L.orc_ycc_padded.argtypes = [u8p, sz, sz, C.c_int, u8p, C.POINTER(sz), C.POINTER(sz)]
    L.orc_transform.argtypes = [u8p, sz, sz, C.c_int, u32p, u32p, C.c_int, i16p]
    L.orc_transform_ycc.argtypes = [u8p, sz, sz, C.c_int, u32p, u32p, C.c_int, i16p]
    L.orc_num_mcus.argtypes = [sz, sz, C.c_int]
    L.orc_num_mcus.restype = sz
    L.orc_blocks_per_mcu.argtypes = [C.c_int]
    L.orc_entropy.argtypes = [i16p, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, sz,
                              C.POINTER(C.c_uint64)]
    L.orc_entropy.restype = sz
    L.orc_jfif_header.argtypes = [sz, sz, C.c_int, u32p, u32p, C.c_int, u8p, sz]
    L.orc_jfif_header.restype = sz
    L.orc_encode_jfif.argtypes = [u8p, sz, sz, C.c_int, u32p, u32p, C.c_int, C.c_int, u8p, sz]
    L.orc_encode_jfif.restype = sz
    L.orc_synth_rgb.argtypes = [C.c_uint64, sz, sz, sz, u8p]
    L.orc_aos_to_planar_u32.argtypes = [u8p, sz, sz, u32p]
    L.orc_encode_jfif_optimized.argtypes = [u8p, sz, sz, C.c_int, u32p, u32p, C.c_int, C.c_int, u8p, sz]
    L.orc_encode_jfif_optimized.restype = sz
    L.orc_optimal_spec.argtypes = [np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS"), u8p, u8p]
    L.orc_symbol_histogram.argtypes = [i16p, sz, C.c_int, C.c_int, C.c_int, np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")]
    L.orc_planar_u32_interleave.argtypes = [u32p, sz, sz, u32p]
    L.orc_init()
    _oracle = L
    return L


def have_ref():
    return os.path.exists(REF_SO)


def ref(o0=False):
    """The reference's own utils.cpp behind oracle/ref_harness.cpp (None if not built)."""
    path = REF_O0_SO if o0 else REF_SO
    if path in _ref:
        return _ref[path]
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    vp = C.c_void_p
    L.ref_padded_size.argtypes = [sz, sz, C.POINTER(sz), C.POINTER(sz)]
    L.ref_performCSC.argtypes = [u8p, sz, sz]
    L.ref_performCDS.argtypes = [u8p, sz, sz]
    L.ref_pad.argtypes = [u8p, sz, sz, u8p, sz, sz]
    L.ref_u8_to_double.argtypes = [u8p, f64p, sz, sz]
    L.ref_substractfromAll.argtypes = [f64p, sz, sz, C.c_double]
    L.ref_performDCT.argtypes = [f64p, sz, sz]
    L.ref_dct_from_copy.argtypes = [f64p, sz, sz]
    L.ref_performQuantization.argtypes = [f64p, sz, sz, vp, vp]
    L.ref_everyMCUisnow2DArray.argtypes = [f64p, sz, sz, i32p]
    L.ref_performZigZag.argtypes = [i32p, i32p, C.c_int]
    L.ref_performRLE.argtypes = [i32p, C.c_int, i32p, C.c_uint64, u32p]
    L.ref_performRLE.restype = C.c_uint64
    L.ref_HuffmanEncoder.argtypes = [i32p, C.c_int, vp, C.c_uint64]
    L.ref_HuffmanEncoder.restype = C.c_uint64
    L.ref_getValueCategory.argtypes = [C.c_int]
    L.ref_valueToBitString.argtypes = [C.c_int, C.c_char_p]
    L.ref_table_code.argtypes = [C.c_int, C.c_int, C.c_int]
    L.ref_table_code.restype = C.c_char_p
    L.ref_quant_tables.argtypes = [u32p, u32p]
    L.ref_copyImageToVector.argtypes = [u8p, sz, sz, u32p]
    L.ref_switchVectorChannelOrdering.argtypes = [u32p, sz, sz, u32p]
    for name, at in (("ref_removeRedChannel", [u8p, sz, sz]), ("ref_copyDoubleToUIntImage", [f64p, sz, sz, u8p]),
                     ("ref_copyOntoLargerVectorWithPadding", [u32p, sz, sz, u32p, sz, sz]),
                     ("ref_everyMCUisnow1DArray", [i32p, sz, sz, i32p])):
        if hasattr(L, name):  # a prebuilt harness from an earlier round lacks them
            getattr(L, name).argtypes = at
    L.ref_run_pipeline.argtypes = [u8p, sz, sz, C.c_int, vp, vp, vp, vp, vp, C.c_uint64, C.POINTER(C.c_uint64), vp]
    _ref[path] = L
    return L


# ---------------------------------------------------------------- conveniences

def read_ppm(path):
    """Binary P6 reader with the reference's rules (utils.cpp:11-65)."""
    with open(path, "rb") as f:
        if f.readline() != b"P6\n":
            raise ValueError("Invalid file format")
        line = f.readline()
        while line.startswith(b"#"):
            line = f.readline()
        w, h = (int(t) for t in line.split()[:2])
        if int(f.readline()) != 255:
            raise ValueError("Invalid maximum value")
        data = np.frombuffer(f.read(w * h * 3), np.uint8).reshape(h, w, 3).copy()
    return data


def q50():
    L = oracle()
    ql = np.array((C.c_uint * 64).in_dll(L, "orc_q50_lum"), np.uint32)
    qc = np.array((C.c_uint * 64).in_dll(L, "orc_q50_chrom"), np.uint32)
    return ql, qc


def quality_tables(q):
    ql = np.zeros(64, np.uint32)
    qc = np.zeros(64, np.uint32)
    oracle().orc_quality_tables(q, ql, qc)
    return ql, qc


def zigzag_order():
    L = oracle()
    return np.array((C.c_uint8 * 64).in_dll(L, "orc_zigzag_order"), np.uint8)


def synth(seed, W, H, y0=0):
    out = np.zeros((H, W, 3), np.uint8)
    oracle().orc_synth_rgb(seed, W, y0, H, out)
    return out


def ycc_padded(rgb, sub):
    H, W, _ = rgb.shape
    m = 16 if sub == SUB_420 else 8
    nW, nH = -(-W // m) * m, -(-H // m) * m
    out = np.zeros((nH, nW, 3), np.uint8)
    a, b = sz(), sz()
    rc = oracle().orc_ycc_padded(np.ascontiguousarray(rgb), W, H, sub, out, C.byref(a), C.byref(b))
    assert rc == 0 and (a.value, b.value) == (nW, nH)
    return out


def transform(rgb, sub, ql, qc, quirks=0):
    H, W, _ = rgb.shape
    L = oracle()
    n = L.orc_num_mcus(W, H, sub)
    coef = np.zeros((n, L.orc_blocks_per_mcu(sub), 64), np.int16)
    rc = L.orc_transform(np.ascontiguousarray(rgb), W, H, sub, ql, qc, quirks, coef)
    assert rc == 0
    return coef


def transform_ycc(ycc, sub, ql, qc, quirks=0):
    """orc_transform for an (H, W, 3) image that is already Y,Cb,Cr (the stages after performCDS)."""
    H, W, _ = ycc.shape
    L = oracle()
    n = L.orc_num_mcus(W, H, sub)
    coef = np.zeros((n, L.orc_blocks_per_mcu(sub), 64), np.int16)
    rc = L.orc_transform_ycc(np.ascontiguousarray(ycc), W, H, sub, ql, qc, quirks, coef)
    assert rc == 0
    return coef


def nv12_from_rgb(rgb):
    """(Y plane, interleaved CbCr plane) of the reference's CSC + CDS (utils.cpp:92-141): a complete 2x2 cell carries its
    truncated mean, a cell cut by an odd edge the chroma of its top-left pixel."""
    H, W, _ = rgb.shape
    ycc = np.ascontiguousarray(rgb).copy()
    oracle().orc_csc(ycc.reshape(-1), W * H)
    if W >= 2 and H >= 2:
        oracle().orc_cds(ycc.reshape(-1), W, H)
    return np.ascontiguousarray(ycc[:, :, 0]), np.ascontiguousarray(ycc[0::2, 0::2, 1:3]).reshape((H + 1) // 2, -1)


def ycc_from_nv12(y, uv):
    """Full-resolution Y,Cb,Cr (H, W, 3) with the chroma pairs replicated over their cells."""
    H, W = y.shape
    c = uv.reshape(uv.shape[0], -1, 2)
    full = np.repeat(np.repeat(c, 2, axis=0), 2, axis=1)[:H, :W]
    return np.ascontiguousarray(np.dstack([y, full[:, :, 0], full[:, :, 1]]))


def jfif_from_coef(coef, W, H, sub, ql, qc, restart_interval=0):
    """header + entropy segment + EOI for coefficients in scan order."""
    seg, _ = entropy(coef, sub, restart_interval)
    return bytes(jfif_header(W, H, sub, ql, qc, restart_interval)) + bytes(seg) + b"\xff\xd9"


def entropy(coef, sub, restart_interval=0, quirks=0, raw_bits=False, rst_phase=0, final_rst=False):
    """Returns (bytes, nbits)."""
    L = oracle()
    n_mcu = coef.shape[0]
    cap = coef.size * 4 + 1024
    out = np.zeros(cap, np.uint8)
    nbits = C.c_uint64()
    n = L.orc_entropy(np.ascontiguousarray(coef), n_mcu, sub, restart_interval, quirks, int(raw_bits), rst_phase,
                      int(final_rst), out, cap, C.byref(nbits))
    assert n != C.c_size_t(-1).value
    return out[:n].copy(), nbits.value


def jfif_header(W, H, sub, ql, qc, restart_interval=0):
    out = np.zeros(1024, np.uint8)
    n = oracle().orc_jfif_header(W, H, sub, ql, qc, restart_interval, out, 1024)
    return out[:n].copy()


def encode_jfif(rgb, sub, ql, qc, restart_interval=0, quirks=0):
    H, W, _ = rgb.shape
    cap = W * H * 6 + 65536
    out = np.zeros(cap, np.uint8)
    n = oracle().orc_encode_jfif(np.ascontiguousarray(rgb), W, H, sub, ql, qc, restart_interval, quirks, out, cap)
    assert n != C.c_size_t(-1).value
    return out[:n].tobytes()


def encode_jfif_optimized(rgb, sub, ql, qc, restart_interval=0, quirks=0):
    """Two-pass encode with per-image optimal Huffman tables (T.81 K.2 / libjpeg's jpeg_gen_optimal_table)."""
    H, W, _ = rgb.shape
    cap = W * H * 6 + 65536
    out = np.zeros(cap, np.uint8)
    n = oracle().orc_encode_jfif_optimized(np.ascontiguousarray(rgb), W, H, sub, ql, qc, restart_interval, quirks, out, cap)
    assert n != C.c_size_t(-1).value
    return out[:n].tobytes()


def optimal_spec(freq):
    """(bits[16], vals[n]) of the optimal table for 256 symbol frequencies."""
    bits, vals = np.zeros(16, np.uint8), np.zeros(256, np.uint8)
    n = oracle().orc_optimal_spec(np.ascontiguousarray(freq, np.uint64), bits, vals)
    assert n >= 0
    return bits, vals[:n]


def bits_to_ascii(packed, nbits):
    """MSB-first packed bits -> the reference's '0'/'1' string."""
    return np.unpackbits(np.asarray(packed, np.uint8))[:nbits].astype(np.uint8).__add__(48).tobytes()


def ref_pipeline(rgb, dct_mode, ql=None, qc=None, o0=False):
    """Run the reference's CPU path. Returns dict(ycc, zigzag, bits(ascii), stage_us)."""
    R = ref(o0)
    H, W, _ = rgb.shape
    a, b = sz(), sz()
    R.ref_padded_size(W, H, C.byref(a), C.byref(b))
    nW, nH = a.value, b.value
    rpc = nW * nH // 64
    ycc = np.zeros((nH, nW, 3), np.uint8)
    zz = np.zeros((3 * rpc, 64), np.int32)
    cap = rpc * 3 * 64 * 28
    bits = np.zeros(cap, np.uint8)
    nbits = C.c_uint64()
    us = np.zeros(9, np.float64)
    qlp = ql.ctypes.data if ql is not None else None
    qcp = qc.ctypes.data if qc is not None else None
    rc = R.ref_run_pipeline(np.ascontiguousarray(rgb), W, H, dct_mode, qlp, qcp, ycc.ctypes.data, zz.ctypes.data,
                            bits.ctypes.data, cap, C.byref(nbits), us.ctypes.data)
    assert rc == 0
    names = ["CSC", "CDS", "levelShift", "DCT", "Quant", "TotalCopy", "zigZag", "RLE", "Huffman"]
    return dict(ycc=ycc, zigzag=zz, bits=bits[: nbits.value].tobytes(), nbits=nbits.value,
                stage_us=dict(zip(names, us.tolist())))
