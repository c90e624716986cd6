"""CPU tests of the drop-in boundary: libjpegb200.so loads, exports every symbol that
include/jpegb200.h declares, its host-only entry points agree with the oracle, and without a
GPU the product fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "jpegb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(jb):
    L = jb.lib()
    names = declared_symbols()
    assert len(names) >= 38
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/jpegb200.h but not exported"
    assert sorted(jb.SYMBOLS) == names  # the Python binding covers the whole header


def test_no_oracle_in_product(jb):
    """The shipped library must not link or reference the test-only oracle."""
    blob = open(jb.LIB_PATH, "rb").read()
    assert b"orc_" not in blob and b"liboracle" not in blob and b"libjpegref" not in blob
    for root, _, files in os.walk(os.path.join(ROOT, "jpeg-encoder-opencl_b200")):
        for f in files:
            if f.endswith((".cu", ".cpp", ".h", ".cuh", ".hpp", ".py")):
                text = open(os.path.join(root, f)).read()
                assert "oracle_lib" not in text and "jpeg_oracle" not in text, f


def test_host_only_entry_points_match_oracle(jb):
    L = jb.lib()
    for q in (1, 10, 50, 75, 90, 100):
        a, b = jb.quality_tables(q)
        c, d = ol.quality_tables(q)
        assert np.array_equal(a, c) and np.array_equal(b, d)
    nW, nH = C.c_size_t(), C.c_size_t()
    assert L.jb_padded_size(253, 254, 8, C.byref(nW), C.byref(nH)) == 0 and (nW.value, nH.value) == (256, 256)
    assert L.jb_padded_size(1920, 1080, 16, C.byref(nW), C.byref(nH)) == 0 and (nW.value, nH.value) == (1920, 1088)
    for sub in (0, 1, 2):
        assert L.jb_blocks_per_mcu(sub) == ol.oracle().orc_blocks_per_mcu(sub)
        assert L.jb_num_mcus(253, 254, sub) == ol.oracle().orc_num_mcus(253, 254, sub)
        for ri in (0, 7):
            ql, qc = ol.quality_tables(75)
            p = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=ri)
            out = np.zeros(1024, np.uint8)
            n = C.c_size_t()
            assert L.jb_write_header(C.byref(p), 253, 254, out.ctypes.data, 1024, C.byref(n)) == 0
            want = ol.jfif_header(253, 254, sub, ql, qc, ri)
            assert n.value == len(want) == jb.header_bytes(p) and np.array_equal(out[: n.value], want)


def test_optimal_huffman_spec_matches_oracle(jb):
    """The product's table generator (JB_FLAG_OPTIMIZE_HUFFMAN, host code) == the oracle's restatement of T.81 K.2 /
    jpeg_gen_optimal_table on random histograms: skewed, flat, single-symbol, huge counts, more than 2^16 spread
    (code lengths above 16 folded back)."""
    rng = np.random.default_rng(4)
    for case in range(60):
        n_sym = int(rng.choice([1, 2, 3, 12, 40, 162, 256]))
        freq = np.zeros(256, np.uint64)
        idx = rng.choice(256, n_sym, replace=False)
        kind = case % 4
        if kind == 0:
            freq[idx] = rng.integers(1, 1000, n_sym)
        elif kind == 1:
            freq[idx] = (2.0 ** rng.uniform(0, 40, n_sym)).astype(np.uint64) + 1  # forces lengths > 16 before folding
        elif kind == 2:
            freq[idx] = 7
        else:
            freq[idx] = rng.integers(1, 2 ** 33, n_sym, dtype=np.uint64)
        gb, gv = jb.optimal_huffman_spec(freq)
        ob, ov = ol.optimal_spec(freq)
        assert np.array_equal(gb, ob) and np.array_equal(gv, ov), f"case {case}"
        assert int(gb.sum()) == n_sym and sum(int(gb[l]) * 2.0 ** -(l + 1) for l in range(16)) < 1.0


def test_jfif_marker_parser_host_only(jb):
    """The decode path's marker parser (host code, no device): the oracle's files in every mode, a file libjpeg wrote (other
    segment order, optimised tables), and what it must reject (progressive, truncated headers, garbage)."""
    import io
    from PIL import Image
    ql, qc = ol.quality_tables(75)
    img = ol.synth(3, 200, 120)
    for sub, ri in ((ol.SUB_444, 0), (ol.SUB_REPL420, 5), (ol.SUB_420, 13)):
        jf = ol.encode_jfif(img, sub, ql, qc, ri)
        info = jb.jfif_info(jf)
        assert (info.W, info.H, info.restart_interval) == (200, 120, ri)
        assert info.subsampling == (jb.SUB_420 if sub == ol.SUB_420 else jb.SUB_444)
        assert info.scan_offset == len(ol.jfif_header(200, 120, sub, ql, qc, ri)) == jb.header_bytes(jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=ri))
        with pytest.raises(jb.JbError):
            jb.jfif_info(jf[: info.scan_offset - 20])  # the header is cut short
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, "JPEG", quality=80, subsampling=2, optimize=True)
    info = jb.jfif_info(buf.getvalue())
    assert (info.W, info.H, info.subsampling, info.restart_interval) == (200, 120, jb.SUB_420, 0)
    assert buf.getvalue()[info.scan_offset - 14: info.scan_offset - 12] == b"\xff\xda"  # SOS (12-byte segment) ends at scan_offset
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, "JPEG", progressive=True)
    with pytest.raises(jb.JbError) as e:
        jb.jfif_info(buf.getvalue())
    assert e.value.code == jb.E_UNSUPPORTED
    for bad in (b"\xff\xd8\xff\xd9", b"hello, world", b"\xff\xd8" + b"\x00" * 64):
        with pytest.raises(jb.JbError):
            jb.jfif_info(bad)


def test_fails_loudly_without_gpu(jb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(jb.JbError) as e:
        jb.Encoder(0)
    assert e.value.code == jb.E_CUDA


REF_DRIVER = "/root/reference/src/OpenCLProject_JpegEncoder.cpp"
REF_HEADER = "/root/reference/src/utils.hpp"


@pytest.mark.skipif(not os.path.exists(REF_DRIVER), reason="the reference sources are not on this machine")
def test_reference_driver_compiles_against_the_compat_header(tmp_path):
    """The reference's own JpegEncoderHost (cpp:28-250), unmodified, compiles against host/utils_compat.hpp +
    host/core_compat.hpp (oracle/Makefile also links it with libjpegb200.so into oracle/_ref/ref_host_b200, which the
    GPU tests run)."""
    import subprocess
    lines = open(REF_DRIVER).read().split("\n")[27:250]
    src = tmp_path / "ref_host.cpp"
    src.write_text("#include <cmath>\n#include <fstream>\n#include <sstream>\n#include \"utils_compat.hpp\"\n"
                   "#include \"core_compat.hpp\"\n" + "\n".join(lines) + "\n")
    r = subprocess.run(["g++", "-std=c++17", "-w", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"),
                        "-I" + os.path.join(ROOT, "jpeg-encoder-opencl_b200", "host"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.skipif(not os.path.exists(REF_HEADER), reason="the reference sources are not on this machine")
def test_compat_header_declares_every_name_of_utils_hpp():
    """Every function and type name the reference's utils.hpp declares (utils.hpp:7-137) exists in utils_compat.hpp."""
    ref = re.sub(r"//.*", "", open(REF_HEADER).read())
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", ref)) - {"FIX", "defined"}
    names |= set(re.findall(r"\b(?:struct|typedef struct \w+)\s+(\w+)", ref))
    text = open(os.path.join(ROOT, "jpeg-encoder-opencl_b200", "host", "utils_compat.hpp")).read()
    missing = sorted(n for n in names if not re.search(rf"\b{n}\b", text))
    assert not missing, missing


def test_compat_header_covers_the_reference_surface():
    """utils_compat.hpp re-declares the reference's stage functions (utils.hpp:81-137) by name."""
    text = open(os.path.join(ROOT, "jpeg-encoder-opencl_b200", "host", "utils_compat.hpp")).read()
    for name in ("performCSC", "performCDS", "getNearest8x8ImageSize", "copyToLargerImage", "addReversedPadding",
                 "copyUIntToDoubleImage", "substractfromAll", "performDCT", "performQuantization",
                 "everyMCUisnow2DArray", "performZigZag", "performRLE", "HuffmanEncoder", "readPPMImage",
                 "writePPMImage", "quant_mat_lum", "quant_mat_chrom", "copyImageToVector", "switchVectorChannelOrdering",
                 "CPUTelemetry", "removeRedChannel", "getValueCategory", "valueToBitString", "everyMCUisnow1DArray",
                 "copyOntoLargerVectorWithPadding", "writeVectorToFile", "copyDoubleToUIntImage", "previewImage"):
        assert re.search(rf"\b{name}\b", text), name
