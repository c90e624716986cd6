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
