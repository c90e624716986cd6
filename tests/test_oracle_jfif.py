"""CPU tests of the oracle's surface the reference does not have (byte packing, 0xFF00
stuffing, RSTn, JFIF segments, quality scaling, true 4:2:0): pinned by T.81/JFIF
conformance, i.e. by two independent decoders (PIL/libjpeg and OpenCV), SURVEY.md 8c."""
import io

import numpy as np
import pytest
from PIL import Image

import oracle_lib as ol
from conftest import noise_image

cv2 = pytest.importorskip("cv2")


def psnr(a, b):
    return 10 * np.log10(255.0 ** 2 / np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))


def decode_both(jf):
    a = np.array(Image.open(io.BytesIO(jf)).convert("RGB"))
    b = cv2.imdecode(np.frombuffer(jf, np.uint8), cv2.IMREAD_COLOR)[:, :, ::-1]
    return a, b


@pytest.mark.parametrize("sub,q,want_len,want_psnr", [
    (ol.SUB_REPL420, 50, 17006, 18.862),
    (ol.SUB_420, 50, None, 18.734),
    (ol.SUB_420, 75, None, 19.500),
    (ol.SUB_REPL420, 90, None, 20.18),
])
def test_fruit_decodes_with_survey_psnr(fruit, sub, q, want_len, want_psnr):
    ql, qc = ol.quality_tables(q)
    jf = ol.encode_jfif(fruit, sub, ql, qc)
    if want_len:
        assert len(jf) == want_len
    a, b = decode_both(jf)
    assert a.shape == fruit.shape and b.shape == fruit.shape
    assert abs(psnr(a, fruit) - want_psnr) < 0.01 and abs(psnr(a, fruit) - psnr(b, fruit)) < 0.01


@pytest.mark.parametrize("sub", [ol.SUB_444, ol.SUB_REPL420, ol.SUB_420])
def test_restart_intervals_do_not_change_pixels(fruit, sub):
    ql, qc = ol.quality_tables(75)
    base = decode_both(ol.encode_jfif(fruit, sub, ql, qc, 0))[0]
    for ri in (1, 4, 16, 300):
        jf = ol.encode_jfif(fruit, sub, ql, qc, ri)
        a, b = decode_both(jf)
        assert np.array_equal(a, base)
        assert abs(psnr(b, fruit) - psnr(base, fruit)) < 0.01
        n_rst = sum(1 for i in range(len(jf) - 1) if jf[i] == 0xFF and 0xD0 <= jf[i + 1] <= 0xD7)
        assert n_rst == -(-ol.oracle().orc_num_mcus(fruit.shape[1], fruit.shape[0], sub) // ri) - 1


def test_quality_tables_match_libjpeg():
    """IJG scaling of utils.hpp:42-62 == the tables libjpeg (PIL) writes at the same quality."""
    img = Image.fromarray(noise_image(1, 16, 16))
    zz = ol.zigzag_order()
    for q in (10, 25, 50, 75, 90, 95, 100):
        buf = io.BytesIO()
        img.save(buf, "JPEG", quality=q)
        tabs = Image.open(io.BytesIO(buf.getvalue())).quantization
        ql, qc = ol.quality_tables(q)
        # PIL >= 9.x returns tables in natural order; older versions in zigzag order
        cand_l = (list(ql), list(ql[zz]))
        cand_c = (list(qc), list(qc[zz]))
        assert list(tabs[0]) in cand_l, q
        assert list(tabs[1]) in cand_c, q


def test_true_420_is_close_to_libjpeg(fruit):
    ql, qc = ol.quality_tables(75)
    ours = ol.encode_jfif(fruit, ol.SUB_420, ql, qc)
    buf = io.BytesIO()
    Image.fromarray(fruit).save(buf, "JPEG", quality=75, subsampling=2)
    theirs = buf.getvalue()
    assert abs(len(ours) - len(theirs)) / len(theirs) < 0.02
    assert abs(psnr(decode_both(ours)[0], fruit) - psnr(decode_both(theirs)[0], fruit)) < 0.05


@pytest.mark.parametrize("shape", [(16, 16), (17, 31), (100, 37), (64, 200)])
def test_odd_sizes_decode(shape):
    H, W = shape
    img = ol.synth(5, W, H)
    for sub in (ol.SUB_444, ol.SUB_REPL420, ol.SUB_420):
        m = 16 if sub == ol.SUB_420 else 8
        if (-W) % m > W or (-H) % m > H:
            continue
        ql, qc = ol.quality_tables(85)
        a, b = decode_both(ol.encode_jfif(img, sub, ql, qc, 3))
        assert a.shape == img.shape and psnr(a, img) > 28 and abs(psnr(a, img) - psnr(b, img)) < 0.05


def test_strip_identity():
    """bits of restart interval k == the coder run on that interval's MCUs alone (predictors start
    at 0, utils.cpp:665) -- the identity the full-size GPU tests and the multi-GPU stitch rely on."""
    img = ol.synth(9, 96, 80)
    ql, qc = ol.quality_tables(75)
    mcux = 96 // 16
    coef = ol.transform(img, ol.SUB_420, ql, qc)
    whole, _ = ol.entropy(coef, ol.SUB_420, mcux)
    parts = []
    for row in range(80 // 16):
        strip_coef = ol.transform(img[row * 16:(row + 1) * 16], ol.SUB_420, ql, qc)
        assert np.array_equal(strip_coef, coef[row * mcux:(row + 1) * mcux])
        seg, _ = ol.entropy(strip_coef, ol.SUB_420, mcux, rst_phase=row, final_rst=row != 4)
        parts.append(seg)
    assert np.array_equal(np.concatenate(parts), whole)


def test_synth_generator_properties():
    a = ol.synth(0x4B3840, 128, 64)
    assert a.min() >= 28 and a.max() <= 228
    assert np.array_equal(a[10:20], ol.synth(0x4B3840, 128, 10, 10))  # row-addressable
    assert not np.array_equal(a, ol.synth(0x4B3841, 128, 64))


def test_optimised_huffman_tables(fruit):
    """The oracle's two-pass encode (T.81 K.2 tables as libjpeg builds them): legal tables (no code longer than 16
    bits, the all-ones code unused), optimal cost, smaller files that decode to the same pixels."""
    import heapq
    import io
    from PIL import Image
    rng = np.random.default_rng(2)
    for n_sym in (1, 2, 5, 40, 162, 256):
        freq = np.zeros(256, np.uint64)
        freq[rng.choice(256, n_sym, replace=False)] = rng.integers(1, 10 ** int(rng.integers(1, 9)), n_sym)
        bits, vals = ol.optimal_spec(freq)
        assert len(vals) == n_sym and sorted(vals) == sorted(np.nonzero(freq)[0])
        kraft = sum(int(bits[l]) * 2.0 ** -(l + 1) for l in range(16))
        assert kraft < 1.0  # the reserved all-ones code stays free
        lens, k = {}, 0
        for l in range(16):
            for _ in range(int(bits[l])):
                lens[int(vals[k])] = l + 1
                k += 1
        cost = sum(int(freq[s]) * l for s, l in lens.items())
        heap = [int(f) for f in freq if f] + [1]  # with the reserved symbol, unlimited length: a lower bound
        heapq.heapify(heap)
        bound = 0
        while len(heap) > 1:
            a, b = heapq.heappop(heap), heapq.heappop(heap)
            bound += a + b
            heapq.heappush(heap, a + b)
        assert cost <= bound  # (equal up to the reserved symbol's own code when no length was folded back)
    for sub in (ol.SUB_420, ol.SUB_444, ol.SUB_REPL420):
        for img, ri in ((fruit, 0), (ol.synth(5, 320, 200), 20)):
            ql, qc = ol.quality_tables(75)
            std, opt = ol.encode_jfif(img, sub, ql, qc, ri), ol.encode_jfif_optimized(img, sub, ql, qc, ri)
            assert len(opt) < len(std)
            assert np.array_equal(np.array(Image.open(io.BytesIO(std))), np.array(Image.open(io.BytesIO(opt))))
    # the Annex-K tables are back in force afterwards
    ql, qc = ol.q50()
    assert ol.encode_jfif(fruit, ol.SUB_REPL420, ql, qc) == ol.encode_jfif(fruit, ol.SUB_REPL420, ql, qc)
