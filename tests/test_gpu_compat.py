"""GPU tests of the C++ host side: the reference's driver sequence written against
host/utils_compat.hpp (compat_demo) and the command-line driver (jpegb200_cli)."""
import io
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "jpeg-encoder-opencl_b200", "host")


def fnv1a(b):
    h = 0xCBF29CE484222325
    for x in bytes(b):
        h = ((h ^ x) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def tools():
    if not (os.path.exists(os.path.join(HOST, "compat_demo")) and os.path.exists(os.path.join(HOST, "jpegb200_cli"))):
        subprocess.run(["make", "-C", os.path.dirname(HOST), "tools"], check=True, capture_output=True)


def staged_oracle(rgb, quirks):
    L = ol.oracle()
    ql, qc = ol.q50()
    ycc = ol.ycc_padded(rgb, ol.SUB_REPL420)
    nH, nW, _ = ycc.shape
    d = np.zeros(ycc.size, np.float64)
    L.orc_u8_to_double(ycc.reshape(-1), d, ycc.size)
    L.orc_subtract(d, d.size, 128.0)
    L.orc_dct_image(d, nW, nH, 1 if quirks & ol.Q1 else 0)
    L.orc_quantize_image(d, nW, nH, ql, qc)
    rpc = nW * nH // 64
    lin = np.zeros((3 * rpc, 64), np.int32)
    zz = np.zeros_like(lin)
    L.orc_blockify(d, nW, nH, lin)
    L.orc_zigzag(lin, zz, 3 * rpc)
    packed = np.zeros(zz.size * 4, np.uint8)
    nb = L.orc_huffman_ref(zz, rpc, quirks, packed, packed.size)
    return zz, ol.bits_to_ascii(packed, nb)


@pytest.mark.parametrize("exact", [False, True])
def test_reference_driver_sequence_through_compat_header(exact):
    """JpegEncoderHost's call sequence (cpp:59-225) on the reference's function names, every stage on
    the GPU: zigzag array and Huffman bit string equal the oracle's (and, with --ref-exact, the
    reference as written: 307 829 bits on fruit.ppm)."""
    tools()
    args = [os.path.join(HOST, "compat_demo"), os.path.join(GOLDEN, "fruit.ppm")] + (["--ref-exact"] if exact else [])
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    for label in ("CSC", "CDS", "Total Copy", "Level Shifting", "DCT", "Quantization", "ZigZag", "RLE", "Huffman", "Total"):
        assert f"{label} Time B200:" in r.stdout
    m = re.search(r"padded (\d+)x(\d+)\s+zigzag_fnv ([0-9a-f]+)\s+nbits (\d+)\s+bits_fnv ([0-9a-f]+)", r.stdout)
    assert m, r.stdout
    zz, bits = staged_oracle(ol.read_ppm(os.path.join(GOLDEN, "fruit.ppm")), ol.AS_WRITTEN if exact else 0)
    assert (int(m.group(1)), int(m.group(2))) == (256, 256)
    assert int(m.group(3), 16) == fnv1a(zz.tobytes())
    assert int(m.group(4)) == len(bits) == (307829 if exact else 129097)
    assert int(m.group(5), 16) == fnv1a(bits)


def test_cli_writes_the_same_jfif(tmp_path, enc, jb, fruit):
    from PIL import Image
    tools()
    out = tmp_path / "fruit.jpg"
    r = subprocess.run([os.path.join(HOST, "jpegb200_cli"), os.path.join(GOLDEN, "fruit.ppm"), str(out), "--quality", "75",
                        "--sub", "420", "--restart", "16"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    data = out.read_bytes()
    ql, qc = ol.quality_tables(75)
    assert data == ol.encode_jfif(fruit, ol.SUB_420, ql, qc, 16)
    assert np.array(Image.open(io.BytesIO(data))).shape == fruit.shape
    assert "MP/s" in r.stdout


# ---- the rest of utils.hpp's surface (round 2): helpers of the reference's driver and of its OpenCL half ----------
def test_value_categories_equal_the_reference_for_every_int16(enc):
    """getValueCategory / valueToBitString (utils.cpp:623-653) on the device function the entropy coder uses, for every
    int16 the reference's functions are defined on, against the oracle and -- when it is built -- the reference itself."""
    v = np.arange(-32767, 32768, dtype=np.int16)
    cat, bits = enc.valueCategories(v)
    L = ol.oracle()
    import ctypes as C
    R = ol.ref() if ol.have_ref() else None
    buf = C.create_string_buffer(32)
    for i in list(range(0, v.size, 97)) + list(range(32767 - 2100, 32767 + 2100)):
        x = int(v[i])
        w = C.c_uint32()
        c = L.orc_value_bits(x, C.byref(w))
        assert c == int(cat[i]) == L.orc_category(x) and int(bits[i]) == w.value, x
        if R is not None:
            assert R.ref_getValueCategory(x) == c
            n = R.ref_valueToBitString(x, buf)
            assert n == c and buf.value[:n].decode() == (format(int(bits[i]), "0%db" % c) if c else "")


def test_helpers_of_the_opencl_half(enc):
    """copyOntoLargerVectorWithPadding (utils.cpp:710-741), everyMCUisnow1DArray (:501-515), removeRedChannel (:84-89),
    copyDoubleToUIntImage (:249-259) against the reference's own functions."""
    rng = np.random.default_rng(11)
    W, H, nW, nH = 45, 37, 48, 40
    rgb = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    planar = enc.copyImageToVector(rgb)
    padded = enc.copyOntoLargerVectorWithPadding(planar, W, H, nW, nH).reshape(3, nH, nW)
    want = np.stack([np.pad(rgb[:, :, c].astype(np.uint32), ((0, nH - H), (0, nW - W)), mode="symmetric") for c in range(3)])
    assert np.array_equal(padded, want)  # mirror including the edge, corner mirrored both ways (utils.cpp:211-233)
    ints = rng.integers(-2000, 2000, 3 * nW * nH).astype(np.int32)
    lin = enc.everyMCUisnow1DArray(ints, nW, nH)
    img = ints.reshape(3, nH // 8, 8, nW // 8, 8).transpose(0, 1, 3, 2, 4).reshape(-1, 64)
    assert np.array_equal(lin, img)
    d = rng.uniform(0, 255.9, (H, W, 3))
    assert np.array_equal(enc.copyDoubleToUIntImage(d), d.astype(np.uint8))
    nored = enc.removeRedChannel(rgb.copy())
    assert not nored[:, :, 0].any() and np.array_equal(nored[:, :, 1:], rgb[:, :, 1:])
    if ol.have_ref() and hasattr(ol.ref(), "ref_everyMCUisnow1DArray"):
        R = ol.ref()
        out = np.zeros(3 * nW * nH, np.uint32)
        R.ref_copyOntoLargerVectorWithPadding(planar, W, H, out, nW, nH)
        out = out.reshape(3, nH, nW)
        assert np.array_equal(out[:, :H, :], padded[:, :H, :]) and np.array_equal(out[:, H:, :W], padded[:, H:, :W])
        lin2 = np.zeros_like(lin)
        R.ref_everyMCUisnow1DArray(ints, nW, nH, lin2)
        assert np.array_equal(lin, lin2)
        a = rgb.copy()
        R.ref_removeRedChannel(a, W, H)
        assert np.array_equal(a, nored)
        b = np.zeros((H, W, 3), np.uint8)
        R.ref_copyDoubleToUIntImage(np.ascontiguousarray(d), W, H, b)
        assert np.array_equal(b, enc.copyDoubleToUIntImage(d))


REF_HOST = os.path.join(ROOT, "oracle", "_ref", "ref_host_b200")


@pytest.mark.skipif(not os.path.exists(REF_HOST), reason="oracle/_ref/ref_host_b200 was not built (needs /root/reference at build time)")
def test_the_references_own_driver_function_runs_on_the_gpu(tmp_path, golden):
    """JpegEncoderHost (src/OpenCLProject_JpegEncoder.cpp:28-250) compiled UNMODIFIED against utils_compat.hpp and linked
    with libjpegb200.so: it runs, fills CPUTelemetry, and the PPM dumps it writes between the stages carry the
    reference's own bytes (padded YCbCr after CSC + CDS + mirror padding: SURVEY 8c digest a19bc1f1...)."""
    import hashlib
    (tmp_path / "data").mkdir()
    (tmp_path / "run").mkdir()
    r = subprocess.run([REF_HOST, os.path.join(GOLDEN, "fruit.ppm")], capture_output=True, text=True, timeout=300,
                       cwd=str(tmp_path / "run"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for label in ("CSC", "CDS", "Total Copy", "Level Shifting", "DCT", "Quantization", "ZigZag", "RLE", "Huffman", "Total"):
        assert f"{label} Time CPU:" in r.stdout, label  # the reference's own output lines (it calls every stage "CPU")
    m = re.search(r"telemetry_us CSC (\d+) CDS (\d+) levelShift (\d+) DCT (\d+) Quant (\d+) TotalCopy (\d+) zigZag (\d+) RLE (\d+) Huffman (\d+)", r.stdout)
    assert m and int(m.group(4)) > 0 and int(m.group(9)) > 0
    padded = ol.read_ppm(str(tmp_path / "data" / "fruitCPU_copy_larger_padded.ppm"))
    assert padded.shape == (256, 256, 3)
    assert hashlib.sha256(padded.tobytes()).hexdigest() == "a19bc1f13b7d3327ddd8ca495dd704331c4852ec228d23b36a78e69590d289b3"
    nored = ol.read_ppm(str(tmp_path / "data" / "fruitCPU_no_blue.ppm"))
    fruit = ol.read_ppm(os.path.join(GOLDEN, "fruit.ppm"))
    assert not nored[:, :, 0].any() and np.array_equal(nored[:, :, 1:], fruit[:, :, 1:])


def test_cli_staged_mode_and_speedup_table(tmp_path, fruit):
    """jpegb200_cli --staged 1 --cpu-telemetry FILE: the reference's stage sequence through the staged entry points with a
    kernel time per stage (the reference's "<stage> time (GPU)" lines, cpp:362 ff.), Huffman and transfers included, the
    "## Speedups: ##" table of cpp:621-629 against the reference's own CPU times (measured here with oracle/_ref when it
    is built), and the check that the staged zigzag array equals the fused path's coefficients."""
    tools()
    tel = tmp_path / "cpu.txt"
    if ol.have_ref():
        us = ol.ref_pipeline(fruit, 0)["stage_us"]
    else:  # no reference build on this machine: any positive numbers exercise the table
        us = dict(CSC=300.0, CDS=100.0, levelShift=160.0, DCT=68000.0, Quant=2200.0, TotalCopy=1700.0, zigZag=1200.0, RLE=1900.0, Huffman=5500.0)
    tel.write_text("".join(f"{k}Time {v}\n" for k, v in us.items()))
    out = tmp_path / "fruit.jpg"
    r = subprocess.run([os.path.join(HOST, "jpegb200_cli"), os.path.join(GOLDEN, "fruit.ppm"), str(out), "--quality", "50", "--sub", "repl420",
                        "--staged", "1", "--cpu-telemetry", str(tel)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    for label in ("Color conversion", "Chroma subsampling", "Level shifting", "DCT", "Quantization", "ZigZag", "RLE", "Huffman"):
        m = re.search(rf"^{label} time \(GPU\): ([0-9.]+) us", r.stdout, re.M)
        assert m and float(m.group(1)) > 0, label
        assert re.search(rf"^{label}: [0-9.]+$", r.stdout, re.M), "speed-up line of " + label
    assert "## Speedups: ##" in r.stdout and "## Speedups (fused path): ##" in r.stdout
    assert "staged zigzag array == fused coefficients: yes" in r.stdout
    assert re.search(r"Huffman time \(GPU\): [0-9.]+ us   \(129097 bits", r.stdout)  # conformant bit count of fruit.ppm
    ql, qc = ol.q50()
    assert out.read_bytes() == ol.encode_jfif(fruit, ol.SUB_REPL420, ql, qc, 0)
