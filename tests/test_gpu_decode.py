"""Decode path: jb_decode_jfif* against independent decoders and against the encoder.
  * decoded pixels == PIL's and OpenCV's (both libjpeg-turbo: integer islow IDCT, fancy h2v2 upsampling, fixed-point colour
    conversion), bit for bit, for this library's files in every mode and for files PIL itself wrote;
  * decoded coefficients == the coefficients the encoder coded (jb_transform) == the oracle's: a round trip of the byte stream
    (Huffman codes, stuffing, restart markers, optimised tables);
  * PSNR computed on the GPU == the PSNR computed from PIL's decode (north_star: within 0.01 dB; here: the same pixels)."""
import io

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
SUBS = [ol.SUB_444, ol.SUB_REPL420, ol.SUB_420]


def pil_decode(data):
    from PIL import Image
    return np.array(Image.open(io.BytesIO(data)).convert("RGB"))


def cv_decode(data):
    import cv2
    return cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)[:, :, ::-1]


def psnr(a, b):
    se = float(((a.astype(np.int64) - b.astype(np.int64)) ** 2).sum())
    return 10 * np.log10(255.0 ** 2 * a.size / se) if se else float("inf")


@pytest.mark.parametrize("sub", SUBS)
def test_decoded_pixels_equal_libjpeg(enc, jb, fruit, sub):
    for img, q, ri in ((fruit, 75, 0), (fruit, 50, 7), (ol.synth(3, 640, 360), 90, 40), (ol.synth(4, 333, 211), 75, 1),
                       (ol.synth(5, 16, 16), 75, 0), (ol.synth(6, 1920, 1080), 75, 120)):
        ql, qc = ol.quality_tables(q)
        p = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=ri)
        jf = enc.encode_jfif(img, p)
        got = enc.decode_jfif(jf)
        want = pil_decode(jf)
        assert got.shape == want.shape == img.shape
        assert np.array_equal(got, want), f"sub {sub} {img.shape} q{q} ri {ri}: {int((got != want).sum())} differing samples"
        assert np.array_equal(got, cv_decode(jf))


def test_decodes_files_written_by_libjpeg(enc):
    """Files PIL (libjpeg-turbo) wrote: its own Huffman tables (optimize=True), its quantisation tables, 4:4:4 and 4:2:0."""
    from PIL import Image
    for seed, (W, H), kw in ((1, (320, 240), dict(quality=85, subsampling=0)), (2, (321, 243), dict(quality=60, subsampling=2)),
                             (3, (640, 480), dict(quality=95, subsampling=2, optimize=True)), (4, (64, 48), dict(quality=30, subsampling=0, optimize=True))):
        img = ol.synth(seed, W, H)
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, "JPEG", **kw)
        data = buf.getvalue()
        assert np.array_equal(enc.decode_jfif(data), pil_decode(data)), (W, H, kw)


@pytest.mark.parametrize("sub", SUBS)
def test_coefficient_round_trip(enc, jb, sub):
    """entropy decode (one thread per restart interval) returns exactly what was coded: == jb_transform == the oracle."""
    import torch
    rng = np.random.default_rng(9)
    for img, q, ri, flags in ((ol.synth(7, 512, 208), 75, 0, 0), (ol.synth(8, 512, 208), 75, 3, 0), (ol.synth(9, 253, 131), 90, 16, 0),
                              (rng.integers(0, 256, (96, 160, 3), dtype=np.uint8), 100, 10, 0),  # long codes, ZRL, stuffed bytes
                              (ol.synth(10, 640, 368), 75, 40, jb.FLAG_OPTIMIZE_HUFFMAN)):
        ql, qc = ol.quality_tables(q)
        p = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=ri, flags=flags)
        want = enc.transform(img, jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=ri))
        assert np.array_equal(want, ol.transform(img, sub, ql, qc))
        jf = np.frombuffer(enc.encode_jfif(img, p), np.uint8)
        d_jf = torch.from_numpy(jf.copy()).cuda()
        d_coef = torch.full((want.size,), 77, dtype=torch.int16, device="cuda")
        torch.cuda.synchronize()
        info = enc.jfif_info_device(d_jf.data_ptr(), jf.size)
        assert (info.W, info.H, info.restart_interval) == (img.shape[1], img.shape[0], ri)
        assert info.subsampling == (ol.SUB_420 if sub == ol.SUB_420 else ol.SUB_444)
        enc.decode_jfif_device(d_jf.data_ptr(), jf.size, None, 0, d_coef.data_ptr())
        assert np.array_equal(d_coef.cpu().numpy().reshape(want.shape), want), f"sub {sub} {img.shape} q{q} ri {ri}"


def test_psnr_on_the_gpu_agrees_with_the_cpu_decoder(enc, jb, fruit):
    """The parity report's PSNR (18.862 dB for the reference-mode q50 fruit.ppm) without a CPU decoder."""
    import torch
    ql, qc = ol.q50()
    p = jb.make_params(ol.SUB_REPL420, qlum=ql, qchrom=qc)
    jf = enc.encode_jfif(fruit, p)
    H, W, _ = fruit.shape
    d_jf = torch.from_numpy(np.frombuffer(jf, np.uint8).copy()).cuda()
    d_src = torch.from_numpy(fruit.copy()).cuda()
    d_rgb = torch.zeros(H * W * 3, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    enc.decode_jfif_device(d_jf.data_ptr(), len(jf), d_rgb.data_ptr(), W * 3, None)
    got, se = enc.psnr_device(d_rgb.data_ptr(), W * 3, d_src.data_ptr(), W * 3, W, H)
    ref = pil_decode(jf)
    assert se == int(((ref.astype(np.int64) - fruit.astype(np.int64)) ** 2).sum())
    assert abs(got - psnr(ref, fruit)) < 1e-9 and abs(got - 18.862) < 0.01
    # the encoder's file and the oracle's file are the same bytes, hence the same PSNR (north_star: within 0.01 dB)
    assert jf == ol.encode_jfif(fruit, ol.SUB_REPL420, ql, qc, 0)


def test_decode_rejects_what_it_does_not_handle(enc, jb):
    with pytest.raises(jb.JbError):
        enc.decode_jfif(b"\xff\xd8\xff\xd9")
    with pytest.raises(jb.JbError):
        enc.decode_jfif(b"not a jpeg at all")
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(ol.synth(1, 64, 64)).save(buf, "JPEG", progressive=True)
    with pytest.raises(jb.JbError) as e:
        enc.decode_jfif(buf.getvalue())
    assert e.value.code == jb.E_UNSUPPORTED
