"""Parity of the CUDA path (through the C ABI) with the CPU oracle, and -- where
oracle/_ref was built -- with the reference's own code.  Everything here needs a GPU.

Bars (BASELINE.json north_star): quantised coefficients bit-exact (the binary64 tie
fix-up removes even the <=1-LSB tolerance the spec allows), entropy stream byte-exact,
decoded PSNR identical.
"""
import hashlib
import io

import numpy as np
import pytest

import oracle_lib as ol
from conftest import noise_image

pytestmark = pytest.mark.gpu

sha = lambda b: hashlib.sha256(bytes(b)).hexdigest()
SUBS = [ol.SUB_444, ol.SUB_REPL420, ol.SUB_420]
SUBNAME = {0: "444", 1: "repl420", 2: "420"}


def mismatch_report(got, want):
    d = np.flatnonzero(got.reshape(-1) != want.reshape(-1))
    if d.size == 0:
        return "equal"
    g, w = got.reshape(-1), want.reshape(-1)
    head = ", ".join(f"[{i}] got {g[i]} want {w[i]}" for i in d[:8])
    return f"{d.size} of {g.size} differ (max abs {np.abs(g[d].astype(np.int64) - w[d]).max()}): {head}"


# ------------------------------------------------------------------ staged ------------

def test_csc_full_colour_cube(enc, golden):
    """performCSC over all 2^24 colours == the reference (sha256 pinned in tests/golden)."""
    h = hashlib.sha256()
    L = ol.oracle()
    for r in range(0, 256, 16):
        cube = np.zeros((16, 256, 256, 3), np.uint8)
        cube[..., 0] = np.arange(r, r + 16)[:, None, None]
        cube[..., 1] = np.arange(256)[None, :, None]
        cube[..., 2] = np.arange(256)[None, None, :]
        want = cube.copy().reshape(-1)
        L.orc_csc(want, want.size // 3)
        got = enc.performCSC(cube.reshape(16 * 256, 256, 3))
        assert np.array_equal(got.reshape(-1), want), mismatch_report(got, want)
        h.update(got.tobytes())
    assert h.hexdigest() == golden["csc_cube_sha256"]


@pytest.mark.parametrize("shape", [(254, 253), (64, 64), (17, 33), (2, 2), (1, 9)])
def test_cds_pad(enc, shape):
    H, W = shape
    img = noise_image(1, W, H)
    L = ol.oracle()
    want = img.copy()
    L.orc_cds(want.reshape(-1), W, H)
    got = enc.performCDS(img.copy())
    assert np.array_equal(got, want), mismatch_report(got, want)
    if W >= 8 and H >= 8:
        nW, nH = -(-W // 8) * 8, -(-H // 8) * 8
        wantp = np.zeros((nH, nW, 3), np.uint8)
        assert L.orc_pad_mirror(np.ascontiguousarray(want), W, H, wantp, nW, nH) == 0
        gotp = enc.padMirror(got, 8)
        assert np.array_equal(gotp, wantp), mismatch_report(gotp, wantp)


@pytest.mark.parametrize("inplace", [0, 1])
def test_staged_f64_pipeline_bit_exact(enc, jb, fruit, golden, inplace):
    """copyUIntToDoubleImage .. HuffmanEncoder on fruit.ppm: every stage equals the oracle
    bit for bit (binary64 included), and the digests equal the reference's."""
    L = ol.oracle()
    ql, qc = ol.q50()
    ycc = ol.ycc_padded(fruit, ol.SUB_REPL420)
    # the GPU's own CSC/CDS/pad of the source must give the same padded image
    g = enc.padMirror(enc.performCDS(enc.performCSC(fruit.copy())), 8)
    assert np.array_equal(g, ycc)
    nH, nW, _ = ycc.shape
    want = np.zeros(ycc.size, np.float64)
    L.orc_u8_to_double(ycc.reshape(-1), want, ycc.size)
    got = enc.copyUIntToDoubleImage(ycc)
    assert np.array_equal(got.reshape(-1), want)
    L.orc_subtract(want, want.size, 128.0)
    enc.substractfromAll(got, 128.0)
    assert np.array_equal(got.reshape(-1), want)
    L.orc_dct_image(want, nW, nH, inplace)
    enc.performDCT(got, jb.FLAG_REF_INPLACE_DCT if inplace else 0)
    assert np.array_equal(got.reshape(-1).view(np.uint64), want.view(np.uint64)), \
        f"binary64 DCT differs: max abs {np.abs(got.reshape(-1) - want).max()}"
    L.orc_quantize_image(want, nW, nH, ql, qc)
    enc.performQuantization(got, ql, qc)
    assert np.array_equal(got.reshape(-1), want)
    rpc = nW * nH // 64
    lin_w = np.zeros((3 * rpc, 64), np.int32)
    zz_w = np.zeros_like(lin_w)
    L.orc_blockify(want, nW, nH, lin_w)
    L.orc_zigzag(lin_w, zz_w, 3 * rpc)
    lin = enc.everyMCUisnow2DArray(got)
    assert np.array_equal(lin, lin_w)
    zz = enc.performZigZag(lin)
    assert np.array_equal(zz, zz_w)
    key = "as_written" if inplace else "dct_from_copy"
    assert sha(zz.tobytes()) == golden["fruit"][key]["zigzag_sha256"]
    # RLE (always-EOB as in the reference) and Huffman with the reference's quirks
    flags = jb.FLAG_REF_ALWAYS_EOB | jb.FLAG_REF_TYPO_TABLES
    rle = enc.performRLE(zz, flags)
    pairs = np.zeros(130, np.int32)
    for i in range(0, 3 * rpc, 7):
        n = L.orc_rle_block(np.ascontiguousarray(zz_w[i]), pairs, 1)
        assert np.array_equal(rle[i], pairs[:n]), f"RLE row {i}"
    packed, nbits = enc.HuffmanEncoder(zz, rpc, flags)
    assert nbits == golden["fruit"][key]["nbits"]
    assert sha(ol.bits_to_ascii(packed, nbits)) == golden["fruit"][key]["bits_sha256"]
    # conformant tables / EOB rule vs the oracle
    packed2, nbits2 = enc.HuffmanEncoder(zz, rpc, 0)
    wantp = np.zeros(zz.size * 4, np.uint8)
    nb = L.orc_huffman_ref(zz_w, rpc, 0, wantp, wantp.size)
    assert nbits2 == nb and np.array_equal(packed2, wantp[: (nb + 7) // 8])


def test_huffman_typo_codes_exercised(enc, jb):
    """Blocks built to hit luma AC 3/4..3/A (Q2) and full blocks (Q3): reference-exact bits."""
    rpc = 40
    rng = np.random.default_rng(5)
    zz = np.zeros((3 * rpc, 64), np.int32)
    for b in range(3 * rpc):
        zz[b, 0] = rng.integers(-1000, 1000)
        pos = 1
        while pos < 64:
            pos += int(rng.choice([0, 1, 3, 3, 3, 17, 35]))
            if pos >= 64:
                break
            mag = int(rng.choice([1, 2, 9, 17, 40, 100, 300, 700, 1023]))
            zz[b, pos] = mag if rng.random() < 0.5 else -mag
            pos += 1
        if b % 5 == 0:
            zz[b, 63] = 3
    L = ol.oracle()
    for flags in (0, ol.Q2, ol.Q3, ol.Q2 | ol.Q3):
        want = np.zeros(zz.size * 4, np.uint8)
        nb = L.orc_huffman_ref(zz, rpc, flags, want, want.size)
        got, nbits = enc.HuffmanEncoder(zz, rpc, flags)
        assert nbits == nb, f"flags {flags}: {nbits} vs {nb} bits"
        assert np.array_equal(got, want[: (nb + 7) // 8]), f"flags {flags}"
    R = ol.ref()
    if R is not None:  # the reference's own HuffmanEncoder, as written
        bits = np.zeros(zz.size * 28, np.uint8)
        n = R.ref_HuffmanEncoder(zz.copy(), rpc, bits.ctypes.data, bits.size)
        got, nbits = enc.HuffmanEncoder(zz, rpc, ol.Q2 | ol.Q3)
        assert nbits == n and ol.bits_to_ascii(got, nbits) == bits[:n].tobytes()


# ------------------------------------------------------------------- fused ------------

def _params(jb, sub, q, ri=0, flags=0):
    ql, qc = ol.quality_tables(q)
    return jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=ri, flags=flags), ql, qc


@pytest.mark.parametrize("sub", SUBS)
@pytest.mark.parametrize("q", [50, 75, 90])
def test_transform_fruit_bit_exact(enc, jb, fruit, sub, q):
    """Config #1: fruit.ppm coefficients == the binary64 reference formula, every coefficient."""
    p, ql, qc = _params(jb, sub, q)
    got = enc.transform(fruit, p)
    want = ol.transform(fruit, sub, ql, qc)
    assert np.array_equal(got, want), mismatch_report(got, want)
    t = enc.timings()
    assert t["tie_fixups"] > 0  # the near-tie replay did run (DC ties occur in ~1/64 of the blocks)


def test_transform_fruit_matches_reference_digest(enc, jb, fruit, golden):
    """REPLICATED_420 at the reference's q50: zigzag array digest of SURVEY 8c (DCT from a copy)."""
    p, ql, qc = _params(jb, ol.SUB_REPL420, 50)
    coef = enc.transform(fruit, p)
    planar = np.concatenate([coef[:, c, :] for c in range(3)]).astype(np.int32)
    assert sha(planar.tobytes()) == golden["fruit"]["dct_from_copy"]["zigzag_sha256"]


@pytest.mark.parametrize("sub", SUBS)
@pytest.mark.parametrize("shape", [(16, 16), (8, 24), (33, 47), (100, 300), (131, 253), (64, 1040), (255, 511)])
def test_transform_shapes(enc, jb, sub, shape):
    """Odd sizes, partial units, mirror padding, the odd-edge chroma rule; noise and smooth content."""
    H, W = shape
    m = 16 if sub == ol.SUB_420 else 8
    if (-W) % m > W or (-H) % m > H:
        pytest.skip("padding larger than the image (the reference would read out of bounds)")
    for img in (noise_image(W * 1000 + H, W, H), ol.synth(7, W, H)):
        for q in (75, 100):
            p, ql, qc = _params(jb, sub, q)
            got = enc.transform(img, p)
            want = ol.transform(img, sub, ql, qc)
            assert np.array_equal(got, want), f"{SUBNAME[sub]} {W}x{H} q{q}: " + mismatch_report(got, want)


def test_transform_grey_and_flat(enc, jb):
    """Grey pixels make every Y a CSC tie (table look-up path) and flat blocks make exact DC ties."""
    H, W = 64, 96
    g = np.repeat(np.random.default_rng(3).integers(0, 256, (H, W, 1), dtype=np.uint8), 3, axis=2)
    flat = np.zeros((H, W, 3), np.uint8)
    for by in range(H // 8):
        for bx in range(W // 8):
            flat[by * 8:(by + 1) * 8, bx * 8:(bx + 1) * 8] = (by * 12 + bx * 7) % 256
    for img in (g, flat):
        for sub in SUBS:
            p, ql, qc = _params(jb, sub, 50)
            got = enc.transform(img, p)
            want = ol.transform(img, sub, ql, qc)
            assert np.array_equal(got, want), mismatch_report(got, want)


def test_no_fixup_stays_within_one_lsb_at_ties(enc, jb, fruit):
    """Without the replay the binary32 path may differ by one LSB, and only where the kernel flags a near-tie."""
    p, ql, qc = _params(jb, ol.SUB_420, 75, flags=jb.FLAG_NO_TIE_FIXUP)
    got = enc.transform(fruit, p).astype(np.int32)
    want = ol.transform(fruit, ol.SUB_420, ql, qc).astype(np.int32)
    assert np.abs(got - want).max() <= 1
    assert (got != want).mean() < 1e-3


@pytest.mark.parametrize("sub", SUBS)
@pytest.mark.parametrize("ri", [0, 1, 3, 16, 1000])
def test_entropy_bytes_given_identical_coefficients(enc, jb, sub, ri):
    img = noise_image(11, 200, 120) // 2 + ol.synth(3, 200, 120) // 2
    for q in (30, 75, 95):
        p, ql, qc = _params(jb, sub, q, ri)
        coef = ol.transform(img, sub, ql, qc)
        want, _ = ol.entropy(coef, sub, ri)
        got = enc.entropy(coef, p)
        assert len(got) == len(want) and np.array_equal(got, want), \
            f"{SUBNAME[sub]} ri={ri} q{q}: {len(got)} vs {len(want)} bytes; " + \
            (mismatch_report(got, want) if len(got) == len(want) else "")


def test_entropy_stuffing_heavy(enc, jb):
    """Coefficients chosen to produce long runs of 1 bits (many 0xFF bytes) and tiny intervals."""
    rng = np.random.default_rng(9)
    n_mcu = 300
    coef = np.zeros((n_mcu, 3, 64), np.int16)
    coef[:, :, 0] = rng.integers(-1020, 1020, (n_mcu, 3))  # DC differences stay within category 11
    coef[:, :, 1:] = np.where(rng.random((n_mcu, 3, 63)) < 0.3, -1 * rng.integers(1, 1024, (n_mcu, 3, 63)), 0)
    coef[::7, :, 63] = 1023
    for ri in (0, 1, 2, 50):
        p = jb.make_params(ol.SUB_444, quality=75, restart_interval=ri)
        want, _ = ol.entropy(coef, ol.SUB_444, ri)
        got = enc.entropy(coef, p)
        assert (want == 0xFF).sum() > 100
        assert len(got) == len(want) and np.array_equal(got, want), f"ri={ri}"


@pytest.mark.parametrize("ri", [0, 1, 100, 256, 700, 7000])
def test_entropy_tile_streams_and_their_neighbours(enc, jb, ri):
    """The placement of round 2: tiles of 256 blocks leave k_encode as code streams and are placed with plain
    stores, completed with their neighbours' bits; tiles that span restart intervals or hold more than 4 KB of codes keep
    one slot per block and are placed with atomics on ranges k_zero clears (sparsely when such tiles are few: ri = 0, 256,
    7000 here; wholesale otherwise).  192 tiles of 4:4:4 blocks with every kind next to every other: sparse tiles, a band of
    dense MCUs (oversize tiles), blocks longer than a slot inside sparse tiles, intervals that end exactly at tile boundaries
    (ri = 256 MCUs = 3 tiles), inside tiles (100, 700, 7000) and everywhere (1)."""
    rng = np.random.default_rng(2024 + ri)
    n_mcu = 16384
    coef = np.zeros((n_mcu, 3, 64), np.int16)
    coef[:, :, 0] = rng.integers(-200, 200, (n_mcu, 3))
    sparse = rng.random((n_mcu, 3, 63)) < 0.08
    coef[:, :, 1:] = np.where(sparse, rng.integers(-6, 7, (n_mcu, 3, 63)), 0)
    dense = slice(500, 900)  # ~1300 bits per block: 330 kbit per tile
    coef[dense, :, 1:] = rng.integers(-1023, 1024, (400, 3, 63))
    for m in range(20, n_mcu, 97):  # ~200-bit blocks inside sparse tiles (and a few inside the dense band)
        coef[m, m % 3, 1:13] = rng.integers(300, 1023, 12) * rng.choice([-1, 1], 12)
    p = jb.make_params(ol.SUB_444, quality=75, restart_interval=ri)
    want, _ = ol.entropy(coef, ol.SUB_444, ri)
    got = enc.entropy(coef, p)
    assert len(got) == len(want) and np.array_equal(got, want), \
        f"ri={ri}: {len(got)} vs {len(want)} bytes; " + (mismatch_report(got, want) if len(got) == len(want) else "")
    # and the same blocks as 4:2:0 MCUs (6 blocks each: other tile / interval phases)
    coef6 = coef[: n_mcu // 2 * 2].reshape(n_mcu // 2, 6, 64)
    p = jb.make_params(ol.SUB_420, quality=75, restart_interval=ri)
    want, _ = ol.entropy(coef6, ol.SUB_420, ri)
    got = enc.entropy(coef6, p)
    assert len(got) == len(want) and np.array_equal(got, want), f"4:2:0 ri={ri}: {len(got)} vs {len(want)} bytes"


def test_entropy_tile_staging_without_tma_is_identical(enc, jb, fruit):
    """JB_FLAG_ENTROPY_LDG: k_encode stages its coefficient tiles with per-thread loads (the path a driver without
    cuTensorMapEncodeTiled would take) -- same bytes as the TMA box, for files and for the staged entropy call."""
    rng = np.random.default_rng(5)
    coef = np.zeros((700, 3, 64), np.int16)
    coef[:, :, 0] = rng.integers(-300, 300, (700, 3))
    coef[:, :, 1:] = np.where(rng.random((700, 3, 63)) < 0.15, rng.integers(-40, 41, (700, 3, 63)), 0)
    for ri in (0, 5):
        a = enc.entropy(coef, jb.make_params(ol.SUB_444, quality=75, restart_interval=ri))
        b = enc.entropy(coef, jb.make_params(ol.SUB_444, quality=75, restart_interval=ri, flags=jb.FLAG_ENTROPY_LDG))
        want, _ = ol.entropy(coef, ol.SUB_444, ri)
        assert np.array_equal(a, want) and np.array_equal(b, want)
    for sub in SUBS:
        for img in (fruit, ol.synth(4, 640, 360)):
            p0 = jb.make_params(sub, quality=80, restart_interval=7)
            p1 = jb.make_params(sub, quality=80, restart_interval=7, flags=jb.FLAG_ENTROPY_LDG)
            assert enc.encode_jfif(img, p0) == enc.encode_jfif(img, p1)


@pytest.mark.parametrize("sub", SUBS)
def test_jfif_fruit_byte_exact_and_decodes(enc, jb, fruit, sub):
    """Config #1: whole file == the oracle's, decodes in PIL and OpenCV with identical PSNR."""
    from PIL import Image
    import cv2
    for q, ri in ((75, 0), (75, 16), (50, 0), (90, 4)):
        p, ql, qc = _params(jb, sub, q, ri)
        got = enc.encode_jfif(fruit, p)
        want = ol.encode_jfif(fruit, sub, ql, qc, ri)
        assert got == want, f"{SUBNAME[sub]} q{q} ri={ri}: {len(got)} vs {len(want)} bytes"
        a = np.array(Image.open(io.BytesIO(got)).convert("RGB"))
        b = cv2.imdecode(np.frombuffer(got, np.uint8), cv2.IMREAD_COLOR)[:, :, ::-1]
        assert a.shape == fruit.shape and b.shape == fruit.shape
        psnr = lambda x: 10 * np.log10(255.0 ** 2 / np.mean((x.astype(np.float64) - fruit) ** 2))
        assert abs(psnr(a) - psnr(b)) < 0.01
    # REPLICATED_420 q50 without restarts is 17 006 bytes and decodes to 18.86 dB
    if sub == ol.SUB_REPL420:
        p, ql, qc = _params(jb, sub, 50)
        jf = enc.encode_jfif(fruit, p)
        assert len(jf) == 17006
        a = np.array(Image.open(io.BytesIO(jf)).convert("RGB")).astype(np.float64)
        assert abs(10 * np.log10(255.0 ** 2 / np.mean((a - fruit) ** 2)) - 18.862) < 0.01


def test_batch_equals_single_frames(enc, jb):
    N, H, W = 7, 72, 104
    frames = np.stack([ol.synth(100 + i, W, H) for i in range(N)])
    frames[3] = noise_image(4, W, H)
    p, ql, qc = _params(jb, ol.SUB_420, 75, 7)
    out, offs, sizes = enc.encode_batch(frames, p)
    assert offs[0] == 0 and np.all(offs[1:] == np.cumsum(sizes)[:-1])
    for i in range(N):
        want = ol.encode_jfif(frames[i], ol.SUB_420, ql, qc, 7)
        got = out[int(offs[i]): int(offs[i] + sizes[i])].tobytes()
        assert got == want, f"frame {i}: {len(got)} vs {len(want)}"


def test_batch_many_groups_pipelined(enc, jb):
    """More groups than pipeline slots (frames of ~0.4 MB -> several 96 MB groups would need thousands of
    frames; shrink by using large frames)."""
    N, H, W = 10, 2160, 3840  # 24.9 MB per frame -> 3 frames per group -> 4 groups > 3 slots
    base = ol.synth(1, W, 256)
    frames = np.empty((N, H, W, 3), np.uint8)
    for i in range(N):
        frames[i] = np.roll(np.tile(base, (H // 256 + 1, 1, 1))[:H], i * 37, axis=0)
    p, ql, qc = _params(jb, ol.SUB_420, 75, 240)
    out, offs, sizes = enc.encode_batch(frames, p)
    single = [enc.encode_jfif(frames[i], p) for i in range(N)]
    for i in range(N):
        assert out[int(offs[i]): int(offs[i] + sizes[i])].tobytes() == single[i], f"frame {i}"
    # one MCU row (= one restart interval) of frame 2 against the oracle, via the strip identity
    strip = frames[2][16 * 10: 16 * 11]
    coef = ol.transform(strip, ol.SUB_420, ql, qc)
    want, _ = ol.entropy(coef, ol.SUB_420, 240)
    hdr = jb.header_bytes(p)
    body = np.frombuffer(single[2], np.uint8)[hdr:-2]
    marks = np.flatnonzero((body[:-1] == 0xFF) & (body[1:] >= 0xD0) & (body[1:] <= 0xD7))
    seg = body[marks[9] + 2: marks[10]]
    assert np.array_equal(seg, want), f"interval 10: {len(seg)} vs {len(want)}"


def test_device_resident_api(enc, jb):
    N, H, W = 3, 48, 80
    frames = np.stack([ol.synth(50 + i, W, H) for i in range(N)])
    p, ql, qc = _params(jb, ol.SUB_420, 75)
    d_rgb = enc.device_alloc(frames.nbytes)
    cap = 1 << 20
    d_out = enc.device_alloc(cap)
    d_tab = enc.device_alloc(8 * (2 * N + 1))
    try:
        enc.h2d(d_rgb, frames)
        enc.encode_batch_device(d_rgb, N, W, H, W * 3, W * H * 3, p, d_out, cap, d_tab, d_tab + 8 * N, d_tab + 16 * N)
        enc.sync()
        tab = np.zeros(2 * N + 1, np.uint64)
        enc.d2h(tab, d_tab)
        out = np.zeros(int(tab[2 * N]), np.uint8)
        enc.d2h(out, d_out)
        for i in range(N):
            want = ol.encode_jfif(frames[i], ol.SUB_420, ql, qc, 0)
            assert out[int(tab[i]): int(tab[i] + tab[N + i])].tobytes() == want
        # too small an output buffer is reported, not overrun
        enc.encode_batch_device(d_rgb, N, W, H, W * 3, W * H * 3, p, d_out, 100, d_tab, d_tab + 8 * N, d_tab + 16 * N)
        with pytest.raises(jb.JbError) as ei:
            enc.sync()
        assert ei.value.code == jb.E_NOSPACE
    finally:
        for d in (d_rgb, d_out, d_tab):
            enc.device_free(d)


def test_strips_concatenate_to_whole_image(enc, jb):
    """RST strips (multi-GPU split of one image): header + strips + EOI == whole-image encode with DRI."""
    H, W = 200, 176
    img = ol.synth(77, W, H)
    mcux = W // 16
    p, ql, qc = _params(jb, ol.SUB_420, 75, mcux)  # one restart interval per MCU row
    whole = ol.encode_jfif(img, ol.SUB_420, ql, qc, mcux)
    assert enc.encode_jfif(img, p) == whole
    rows = [0, 64, 112, 200]
    parts = [enc.write_header(p, W, H)]
    for s in range(3):
        strip = img[rows[s]: rows[s + 1]]
        parts.append(enc.encode_strip(strip, p, first_interval=rows[s] // 16, last_strip=(s == 2)).tobytes())
    parts.append(b"\xff\xd9")
    assert b"".join(parts) == whole


def test_synth_generator_matches_oracle(enc):
    for seed, W, H in ((0x4B3840, 200, 50), (0xF000, 97, 61), (1, 1920, 8)):
        assert np.array_equal(enc.synth(seed, W, H), ol.synth(seed, W, H))


def test_error_paths(enc, jb):
    img = noise_image(1, 32, 32)
    p = jb.make_params(ol.SUB_420, quality=75)
    with pytest.raises(jb.JbError) as e:
        enc.encode_jfif(img, p, cap=64)
    assert e.value.code == jb.E_NOSPACE
    bad = jb.make_params(ol.SUB_420, quality=75)
    bad.qlum[5] = 0
    with pytest.raises(jb.JbError) as e:
        enc.encode_jfif(img, bad)
    assert e.value.code == jb.E_INVALID
    with pytest.raises(jb.JbError) as e:  # 4 px wide cannot be mirror-padded to 16 (utils.cpp:211-233)
        enc.encode_jfif(noise_image(1, 4, 32), p)
    assert e.value.code == jb.E_UNSUPPORTED


# ------------------------------------------------------- BASELINE.json full sizes -----

def _decode_psnr(jf, rgb):
    import cv2
    dec = cv2.imdecode(np.frombuffer(jf, np.uint8), cv2.IMREAD_COLOR)[:, :, ::-1]
    assert dec.shape == rgb.shape
    return 10 * np.log10(255.0 ** 2 / np.mean((dec.astype(np.float32) - rgb) ** 2))


def _interval_segments(jf, hdr):
    body = np.frombuffer(jf, np.uint8)[hdr:-2]
    marks = np.flatnonzero((body[:-1] == 0xFF) & (body[1:] >= 0xD0) & (body[1:] <= 0xD7))
    assert np.all((body[marks + 1] - 0xD0) == np.arange(len(marks)) % 8)  # RSTn count modulo 8
    bounds = np.concatenate([[0], marks + 2, [len(body) + 2]])
    return body, marks, bounds


def test_config2_4k_444_q90(enc, jb):
    """Config #2: synthetic 3840x2160 4:4:4 q90.  Size-independent checks: decodes, sane PSNR, and
    sampled MCU rows equal the oracle through the restart-interval identity."""
    W, H = 3840, 2160
    img = enc.synth(0x4B3840, W, H)
    assert np.array_equal(img[1000:1008], ol.synth(0x4B3840, W, 8, 1000))
    mcux = W // 8
    p, ql, qc = _params(jb, ol.SUB_444, 90, mcux)
    jf = enc.encode_jfif(img, p)
    assert _decode_psnr(jf, img) > 35
    body, marks, bounds = _interval_segments(jf, jb.header_bytes(p))
    assert len(marks) == H // 8 - 1
    for row in (0, 133, 269):
        coef = ol.transform(img[row * 8: row * 8 + 8], ol.SUB_444, ql, qc)
        want, _ = ol.entropy(coef, ol.SUB_444, mcux)
        seg = body[bounds[row]: bounds[row + 1] - 2]
        assert np.array_equal(seg, want), f"MCU row {row}"
    # the same image without restart markers: same coefficients -> same decoded pixels
    p0, _, _ = _params(jb, ol.SUB_444, 90, 0)
    import cv2
    a = cv2.imdecode(np.frombuffer(jf, np.uint8), cv2.IMREAD_COLOR)
    b = cv2.imdecode(np.frombuffer(enc.encode_jfif(img, p0), np.uint8), cv2.IMREAD_COLOR)
    assert np.array_equal(a, b)


def test_config3_8k_420_q75_restart(enc, jb):
    """Config #3: synthetic 7680x4320 4:2:0 q75, DRI = one MCU row (480 MCUs)."""
    W, H = 7680, 4320
    img = enc.synth(0x4B7680, W, H)
    p, ql, qc = _params(jb, ol.SUB_420, 75, 480)
    jf = enc.encode_jfif(img, p, cap=W * H)
    assert _decode_psnr(jf, img) > 30
    body, marks, bounds = _interval_segments(jf, jb.header_bytes(p))
    assert len(marks) == H // 16 - 1
    for row in (0, 100, 269):
        coef = ol.transform(img[row * 16: row * 16 + 16], ol.SUB_420, ql, qc)
        want, _ = ol.entropy(coef, ol.SUB_420, 480)
        seg = body[bounds[row]: bounds[row + 1] - 2]
        assert np.array_equal(seg, want), f"MCU row {row}"


# ------------------------------------------------ tensor-core variant of the fused kernel ------------

@pytest.mark.parametrize("sub", SUBS)
@pytest.mark.parametrize("q", [50, 75, 90, 100])
def test_tensor_core_transform_bit_exact(enc, jb, fruit, sub, q):
    """The default transform of every mode: tcgen05 contraction (fp16 2-split) + binary64 replay == the oracle,
    bit for bit (k_transform_tc for 4:2:0, k_transform_tc3 for 4:4:4 and the reference's replicated 4:2:0)."""
    ql, qc = ol.quality_tables(q)
    for img in (fruit, ol.synth(21, 1920, 128), noise_image(3, 200, 120), ol.synth(9, 1928, 70)):
        p = jb.make_params(sub, qlum=ql, qchrom=qc, flags=jb.FLAG_TENSOR_DCT)
        got = enc.transform(img, p)
        want = ol.transform(img, sub, ql, qc)
        assert np.array_equal(got, want), f"{SUBNAME[sub]} q{q} {img.shape}: " + mismatch_report(got, want)


@pytest.mark.parametrize("sub", SUBS)
def test_replay_list_stays_a_sliver(enc, jb, sub):
    """The binary64 replay is for near ties only: well under 1 % of the coefficients in every mode, also with the
    reference's in-place transform.  (A band that collapses keeps the output right and makes the encoder 50 x slower:
    round 2 shipped that for the replicated 4:2:0 mode, whose cell sums cancel for u = 4 / v = 4.)"""
    ql, qc = ol.quality_tables(50)
    img = ol.synth(31, 1920, 256)
    n_coef = 3 * 1920 * 256 if sub != ol.SUB_420 else 1920 * 256 * 3 // 2
    for flags in (jb.FLAG_TENSOR_DCT, jb.FLAG_TENSOR_DCT | jb.FLAG_REF_INPLACE_DCT):
        p = jb.make_params(sub, qlum=ql, qchrom=qc, flags=flags)
        enc.transform(img, p)
        ties = int(enc.timings()["tie_fixups"])
        assert ties < 0.01 * n_coef, f"{SUBNAME[sub]} flags {flags:#x}: {ties} of {n_coef} coefficients replayed"


def test_tensor_core_raw_error_is_small(enc, jb, fruit):
    """Without the replay the tensor-core path may differ by one LSB at (near) ties only."""
    ql, qc = ol.quality_tables(75)
    p = jb.make_params(ol.SUB_420, qlum=ql, qchrom=qc, flags=jb.FLAG_TENSOR_DCT | jb.FLAG_NO_TIE_FIXUP)
    img = ol.synth(5, 1920, 256)
    got = enc.transform(img, p).astype(np.int32)
    want = ol.transform(img, ol.SUB_420, ql, qc).astype(np.int32)
    assert np.abs(got - want).max() <= 1
    assert (got != want).mean() < 2e-3


@pytest.mark.parametrize("sub", SUBS)
@pytest.mark.parametrize("q", [50, 75, 100])
def test_fma_transform_bit_exact(enc, jb, fruit, sub, q):
    """JB_FLAG_FMA_DCT: the CUDA-core kernels (register AAN FDCT, analytic near-tie band) == the oracle."""
    ql, qc = ol.quality_tables(q)
    for img in (fruit, ol.synth(21, 1920, 128), noise_image(3, 200, 120), ol.synth(8, 333, 77)):
        p = jb.make_params(sub, qlum=ql, qchrom=qc, flags=jb.FLAG_FMA_DCT)
        got = enc.transform(img, p)
        want = ol.transform(img, sub, ql, qc)
        assert np.array_equal(got, want), f"{SUBNAME[sub]} q{q} {img.shape}: " + mismatch_report(got, want)


@pytest.mark.parametrize("sub", SUBS)
def test_tensor_core_matches_fma_at_bench_size(enc, jb, sub):
    """Frames of 1080p (smooth synthetic + full-range noise), q75 and q100: both kernels, same coefficients
    (16 frames for 4:2:0, 6 for the 8x8-MCU modes)."""
    rng = np.random.default_rng(77)
    n_s, n_r = (12, 4) if sub == ol.SUB_420 else (4, 2)
    frames = np.stack([ol.synth(900 + i, 1920, 1080) for i in range(n_s)]
                      + [rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8) for _ in range(n_r)])
    for q in (75, 100):
        a = jb.make_params(sub, quality=q, flags=jb.FLAG_FMA_DCT)
        b = jb.make_params(sub, quality=q)
        ca = np.stack([enc.transform(f, a) for f in frames[:: 5 if q == 100 else 1]])
        cb = np.stack([enc.transform(f, b) for f in frames[:: 5 if q == 100 else 1]])
        assert np.array_equal(ca, cb), f"{SUBNAME[sub]} q{q}: " + mismatch_report(ca, cb)


def test_tensor_core_jfif_equals_fma_path(enc, jb):
    frames = np.stack([ol.synth(300 + i, 640, 360) for i in range(4)])
    a = jb.make_params(ol.SUB_420, quality=75, restart_interval=40, flags=jb.FLAG_FMA_DCT)
    b = jb.make_params(ol.SUB_420, quality=75, restart_interval=40)
    oa, offa, sza = enc.encode_batch(frames, a)
    ob, offb, szb = enc.encode_batch(frames, b)
    assert np.array_equal(sza, szb) and np.array_equal(oa[: int(sza.sum())], ob[: int(szb.sum())])


# ------------------------------------------------ planar uint32 input (the reference's OpenCL-half layout) ----

def test_planar_u32_layout_and_fused_input(enc, jb, fruit):
    """copyImageToVector / switchVectorChannelOrdering (utils.hpp:116,119) == the oracle, and the fused encode fed
    with the planar uint32 image == the fused encode of the same pixels as RGB8 (host call and device-side hop)."""
    L = ol.oracle()
    for img in (fruit, noise_image(12, 37, 21), ol.synth(3, 640, 360)):
        H, W, _ = img.shape
        want = np.zeros(3 * W * H, np.uint32)
        L.orc_aos_to_planar_u32(np.ascontiguousarray(img).reshape(-1), W, H, want)
        planar = enc.copyImageToVector(img)
        assert np.array_equal(planar, want)
        inter = np.zeros_like(want)
        L.orc_planar_u32_interleave(want, W, H, inter)
        assert np.array_equal(enc.switchVectorChannelOrdering(planar, W, H), inter)
        for sub, ri in ((ol.SUB_420, 0), (ol.SUB_REPL420, 7), (ol.SUB_444, 0)):
            p = jb.make_params(sub, quality=75, restart_interval=ri)
            assert enc.encode_jfif_planar_u32(planar, W, H, p) == enc.encode_jfif(img, p)
    # device-side hop: planar words in HBM -> pitched RGB8 in HBM
    H, W, _ = fruit.shape
    planar = enc.copyImageToVector(fruit)
    pitch = (W * 3 + 15) // 16 * 16
    d_pl, d_rgb = enc.device_alloc(planar.nbytes), enc.device_alloc(pitch * H)
    try:
        enc.h2d(d_pl, planar)
        enc.planar_u32_to_rgb8_device(d_pl, W, H, d_rgb, pitch)
        back = np.zeros((H, pitch), np.uint8)
        enc.sync()
        enc.d2h(back, d_rgb)
        assert np.array_equal(back[:, : W * 3].reshape(H, W, 3), fruit)
    finally:
        enc.device_free(d_pl)
        enc.device_free(d_rgb)


def test_high_entropy_content_grows_the_workspace(jb):
    """Full-range noise at q100 needs more than the first entropy-workspace budget per block (64 bytes): the
    synchronous entry points enlarge it and run again (found by tests/tools/fuzz_gpu.py); bytes == the oracle."""
    enc = jb.Encoder(0)  # a fresh context: the budget is sticky per context
    try:
        img = noise_image(99, 200, 120)
        ql, qc = ol.quality_tables(100)
        for sub in SUBS:
            p = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=3)
            assert enc.encode_jfif(img, p) == ol.encode_jfif(img, sub, ql, qc, 3), SUBNAME[sub]
        frames = np.stack([noise_image(5 + i, 96, 64) for i in range(3)])
        p = jb.make_params(ol.SUB_444, qlum=ql, qchrom=qc)
        out, offs, sizes = enc.encode_batch(frames, p)
        for f in range(3):
            assert bytes(out[int(offs[f]): int(offs[f] + sizes[f])]) == ol.encode_jfif(frames[f], ol.SUB_444, ql, qc)
    finally:
        enc.close()


# ------------------------------------------------ optimised Huffman tables ------------------

@pytest.mark.parametrize("sub", SUBS)
def test_optimised_huffman_tables_equal_the_oracle(enc, jb, fruit, sub):
    """JB_FLAG_OPTIMIZE_HUFFMAN: GPU symbol histogram + T.81 K.2 tables + custom DHT == the oracle's two-pass
    encode byte for byte; the file is smaller and decodes to the same pixels as the Annex-K one."""
    import cv2
    for img, q, ri in ((fruit, 75, 0), (ol.synth(5, 640, 360), 50, 40), (noise_image(4, 120, 72), 100, 0), (ol.synth(6, 333, 77), 90, 3)):
        ql, qc = ol.quality_tables(q)
        p_std = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=ri)
        p_opt = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=ri, flags=jb.FLAG_OPTIMIZE_HUFFMAN)
        cap = img.shape[0] * img.shape[1] * 12 + 65536
        got = enc.encode_jfif(img, p_opt, cap=cap)
        assert got == ol.encode_jfif_optimized(img, sub, ql, qc, ri), f"{SUBNAME[sub]} {img.shape} q{q} ri{ri}"
        std = enc.encode_jfif(img, p_std, cap=cap)
        assert len(got) < len(std)
        a = cv2.imdecode(np.frombuffer(got, np.uint8), cv2.IMREAD_COLOR)
        b = cv2.imdecode(np.frombuffer(std, np.uint8), cv2.IMREAD_COLOR)
        assert a is not None and np.array_equal(a, b)


def test_optimised_huffman_batch_and_unsupported_paths(enc, jb):
    """A batch shares one table set per group of frames: every frame decodes to the pixels of its Annex-K file;
    strips and header-less entry points refuse the flag."""
    import cv2
    frames = np.stack([ol.synth(40 + i, 320, 176) for i in range(5)])
    p_std = jb.make_params(ol.SUB_420, quality=75)
    p_opt = jb.make_params(ol.SUB_420, quality=75, flags=jb.FLAG_OPTIMIZE_HUFFMAN)
    o1, off1, sz1 = enc.encode_batch(frames, p_std)
    o2, off2, sz2 = enc.encode_batch(frames, p_opt)
    assert int(sz2.sum()) < int(sz1.sum())
    for f in range(5):
        a = cv2.imdecode(o1[int(off1[f]): int(off1[f] + sz1[f])], cv2.IMREAD_COLOR)
        b = cv2.imdecode(o2[int(off2[f]): int(off2[f] + sz2[f])], cv2.IMREAD_COLOR)
        assert b is not None and np.array_equal(a, b)
    ps = jb.make_params(ol.SUB_420, quality=75, restart_interval=20, flags=jb.FLAG_OPTIMIZE_HUFFMAN)
    with pytest.raises(jb.JbError) as e:
        enc.encode_strip(frames[0], ps, 0, True)
    assert e.value.code == jb.E_UNSUPPORTED


# ------------------------------------------------ Q1 in the fused path ----------------------

def test_fused_path_reproduces_the_reference_as_written(enc, jb, fruit, golden):
    """JB_FLAG_REF_INPLACE_DCT in the fused path: the tcgen05 contraction runs the reference's in-place transform
    (utils.cpp:342-345, a different 64x64 matrix) and the binary64 replay re-enacts the overwriting loop, so the
    fused kernel's coefficients are the unmodified reference's zigzag array (SURVEY 8c digest), and with the two
    entropy quirks the bit string is the reference's HuffmanEncoder output."""
    ql, qc = ol.q50()
    flags = jb.FLAG_REF_INPLACE_DCT
    p = jb.make_params(ol.SUB_REPL420, qlum=ql, qchrom=qc, flags=flags)
    coef = enc.transform(fruit, p)
    assert np.array_equal(coef, ol.transform(fruit, ol.SUB_REPL420, ql, qc, ol.Q1))
    planar = np.concatenate([coef[:, c, :] for c in range(3)]).astype(np.int32)
    assert sha(planar.tobytes()) == golden["fruit"]["as_written"]["zigzag_sha256"]
    # other modes, sizes (edge MCUs go to the replay whole), qualities
    for sub in SUBS:
        for img, q in ((ol.synth(4, 320, 200), 75), (noise_image(6, 131, 77), 90), (ol.synth(9, 1928, 70), 50)):
            tl, tc = ol.quality_tables(q)
            pp = jb.make_params(sub, qlum=tl, qchrom=tc, flags=flags)
            got, want = enc.transform(img, pp), ol.transform(img, sub, tl, tc, ol.Q1)
            assert np.array_equal(got, want), f"{SUBNAME[sub]} {img.shape} q{q}: " + mismatch_report(got, want)
    # the whole reference pipeline as written, fused: JFIF-less bit string == the reference's digest
    p3 = jb.make_params(ol.SUB_REPL420, qlum=ql, qchrom=qc, flags=flags | jb.FLAG_REF_TYPO_TABLES | jb.FLAG_REF_ALWAYS_EOB)
    coef3 = enc.transform(fruit, p3)
    zz = np.ascontiguousarray(np.concatenate([coef3[:, c, :] for c in range(3)]).astype(np.int32))
    bits, nbits = enc.HuffmanEncoder(zz, coef3.shape[0], ol.Q2 | ol.Q3)
    assert nbits == golden["fruit"]["as_written"]["nbits"]
    assert sha(ol.bits_to_ascii(bits, nbits)) == golden["fruit"]["as_written"]["bits_sha256"]
    with pytest.raises(jb.JbError) as e:
        enc.transform(fruit, jb.make_params(ol.SUB_REPL420, qlum=ql, qchrom=qc, flags=flags | jb.FLAG_FMA_DCT))
    assert e.value.code == jb.E_UNSUPPORTED


# ------------------------------------------------ whole frames at the benchmark's sizes against the oracle (round 2) ----

@pytest.mark.parametrize("sub", SUBS)
def test_whole_1080p_frame_equals_the_oracle(enc, jb, sub):
    """One full 1920x1080 frame of the batch configuration per subsampling mode: coefficients and the JFIF file equal the
    oracle's, byte for byte (1080 = 67.5 MCU rows of 16: the last MCU row is mirrored inside the hot kernel)."""
    img = enc.synth(0xF000 + 17, 1920, 1080)
    assert np.array_equal(img, ol.synth(0xF000 + 17, 1920, 1080))
    for q, ri in ((75, 0), (50, 120)):
        ql, qc = ol.quality_tables(q)
        p = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=ri)
        got, want = enc.transform(img, p), ol.transform(img, sub, ql, qc)
        assert np.array_equal(got, want), f"{SUBNAME[sub]} q{q}: " + mismatch_report(got, want)
        assert enc.encode_jfif(img, p) == ol.encode_jfif(img, sub, ql, qc, ri), f"{SUBNAME[sub]} q{q} ri {ri}"


def test_config2_whole_4k_frame_equals_the_oracle(enc, jb):
    """Config #2 in full: the 3840x2160 4:4:4 q90 file (no restart markers, as the benchmark codes it) == the oracle's."""
    img = enc.synth(0x4B3840, 3840, 2160)
    ql, qc = ol.quality_tables(90)
    p = jb.make_params(ol.SUB_444, qlum=ql, qchrom=qc)
    assert enc.encode_jfif(img, p) == ol.encode_jfif(img, ol.SUB_444, ql, qc, 0)


def test_config3_whole_8k_frame_equals_the_oracle(enc, jb):
    """Config #3 in full: the 7680x4320 4:2:0 q75 file with DRI = one MCU row == the oracle's."""
    img = enc.synth(0x4B7680, 7680, 4320)
    ql, qc = ol.quality_tables(75)
    p = jb.make_params(ol.SUB_420, qlum=ql, qchrom=qc, restart_interval=480)
    got, want = enc.transform(img, p), ol.transform(img, ol.SUB_420, ql, qc)
    assert np.array_equal(got, want), mismatch_report(got, want)
    assert enc.encode_jfif(img, p, cap=7680 * 4320) == ol.encode_jfif(img, ol.SUB_420, ql, qc, 480)


@pytest.mark.parametrize("sub", [ol.SUB_420, ol.SUB_444])
def test_wider_and_taller_than_65536_with_clamped_sof(enc, jb, sub):
    """Dimensions above 2^16 (JB_FLAG_CLAMP_SOF): every coordinate of the pipeline -- including the block origins that
    the binary64 near-tie replay unpacks -- is 32 bits wide.  Noise at q100 flags thousands of coefficients beyond
    x = 65536 (resp. y = 65536); coefficients == oracle."""
    rng = np.random.default_rng(65536)
    ql, qc = ol.quality_tables(100)
    p = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=0, flags=jb.FLAG_CLAMP_SOF)
    for W, H in ((65536 + 2048, 16), (64, 65536 + 1024)):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        got, want = enc.transform(img, p), ol.transform(img, sub, ql, qc)
        assert enc.timings()["tie_fixups"] > 100
        assert np.array_equal(got, want), f"{W}x{H}: " + mismatch_report(got, want)
    with pytest.raises(jb.JbError) as e:
        enc.transform(np.zeros((16, 65536 + 48, 3), np.uint8), jb.make_params(sub, qlum=ql, qchrom=qc))  # no JB_FLAG_CLAMP_SOF
    assert e.value.code == jb.E_UNSUPPORTED


def test_back_to_back_device_calls_keep_their_own_headers(enc, jb):
    """jb_encode_batch_device is asynchronous: several calls with DIFFERENT headers (size, restart interval) queued
    without jb_sync must each get their own header and status (pinned staging ring, jb_api.cu)."""
    import torch
    shapes = [(64, 48, 0), (80, 32, 2), (48, 64, 3), (96, 16, 0), (32, 32, 1), (128, 48, 4)]
    ql, qc = ol.quality_tables(75)
    imgs = [ol.synth(900 + i, w, h) for i, (w, h, _) in enumerate(shapes)]
    d_in = [torch.from_numpy(im.copy()).cuda() for im in imgs]
    d_out = [torch.zeros(w * h * 3 + 4096, dtype=torch.uint8, device="cuda") for w, h, _ in shapes]
    d_tab = [torch.zeros(3, dtype=torch.int64, device="cuda") for _ in shapes]
    torch.cuda.synchronize()
    for rep in range(2):
        for i, (w, h, ri) in enumerate(shapes):
            p = jb.make_params(ol.SUB_420, qlum=ql, qchrom=qc, restart_interval=ri)
            enc.encode_batch_device(d_in[i].data_ptr(), 1, w, h, w * 3, w * h * 3, p, d_out[i].data_ptr(), d_out[i].numel(),
                                    d_tab[i].data_ptr(), d_tab[i].data_ptr() + 8, d_tab[i].data_ptr() + 16)
        enc.sync()
        for i, (w, h, ri) in enumerate(shapes):
            n = int(d_tab[i][2].item())
            assert bytes(d_out[i][:n].cpu().numpy()) == ol.encode_jfif(imgs[i], ol.SUB_420, ql, qc, ri), f"call {i}"
    # a too-small caller buffer in the middle of the queue is reported by jb_sync, the other calls are intact
    p = jb.make_params(ol.SUB_420, qlum=ql, qchrom=qc)
    enc.encode_batch_device(d_in[0].data_ptr(), 1, 64, 48, 192, 64 * 48 * 3, p, d_out[0].data_ptr(), 700, d_tab[0].data_ptr(),
                            d_tab[0].data_ptr() + 8, d_tab[0].data_ptr() + 16)
    enc.encode_batch_device(d_in[1].data_ptr(), 1, 80, 32, 240, 80 * 32 * 3, p, d_out[1].data_ptr(), d_out[1].numel(), d_tab[1].data_ptr(),
                            d_tab[1].data_ptr() + 8, d_tab[1].data_ptr() + 16)
    with pytest.raises(jb.JbError) as e:
        enc.sync()
    assert e.value.code == jb.E_NOSPACE
    n = int(d_tab[1][2].item())
    assert bytes(d_out[1][:n].cpu().numpy()) == ol.encode_jfif(imgs[1], ol.SUB_420, ql, qc, 0)


def test_batch_reports_the_size_of_the_whole_batch_on_overflow(enc, jb):
    """JB_E_NOSPACE from jb_encode_batch: jb_required_bytes() is what the WHOLE batch needs (several in-flight groups),
    and a second call with that capacity succeeds."""
    frames = np.stack([ol.synth(70 + f, 2048, 1536) for f in range(24)])  # 24 x 9.4 MB: three groups of 96 MB
    p = jb.make_params(ol.SUB_420, quality=75)
    small = np.empty(1 << 20, np.uint8)
    with pytest.raises(jb.JbError) as e:
        enc.encode_batch(frames, p, out=small)
    assert e.value.code == jb.E_NOSPACE
    need = enc.L.jb_required_bytes(enc.h)
    out, offs, sizes = enc.encode_batch(frames, p, out=np.empty(need, np.uint8))
    assert int(sizes.sum()) == need and int(offs[-1] + sizes[-1]) == need
    ql, qc = ol.quality_tables(75)
    for f in (0, 23):
        assert bytes(out[int(offs[f]): int(offs[f] + sizes[f])]) == ol.encode_jfif(frames[f], ol.SUB_420, ql, qc, 0)


def test_strip_in_two_asynchronous_halves(enc, jb):
    """jb_encode_strip_begin / _finish on one GPU: the strip written at a device-side offset equals jb_encode_strip, for
    aligned and odd offsets, chained calls (running offsets kept on the device) and jb_copy_bytes_device."""
    import torch
    W, H = 1024, 208
    img = ol.synth(0x77, W, H)
    ql, qc = ol.quality_tables(75)
    p = jb.make_params(ol.SUB_420, qlum=ql, qchrom=qc, restart_interval=W // 16)
    want = enc.encode_strip(img, p, 0, True)
    d_img = torch.from_numpy(img.copy()).cuda()
    ext = torch.cuda.ExternalStream(enc.stream())
    out = torch.zeros(W * H, dtype=torch.uint8, device="cuda")
    for off0 in (0, 1, 13, 16, 4099):
        ln = torch.zeros(1, dtype=torch.int64, device="cuda")
        off = torch.tensor([off0], dtype=torch.int64, device="cuda")
        out.zero_()
        torch.cuda.synchronize()
        enc.encode_strip_begin(d_img.data_ptr(), p, 0, True, W, H, W * 3, ln.data_ptr())
        enc.encode_strip_finish(out.data_ptr(), out.numel(), off.data_ptr())
        enc.sync()
        n = int(ln.item())
        assert n == len(want) and np.array_equal(out[off0: off0 + n].cpu().numpy(), want)
        assert not out[:off0].any() and not out[off0 + n: off0 + n + 64].any()
    # two chained halves with running offsets on the device == the whole strip; then pushed elsewhere with copy_bytes
    run = torch.zeros(3, dtype=torch.int64, device="cuda")
    run[0] = 5
    out.zero_()
    torch.cuda.synchronize()
    with torch.cuda.stream(ext):
        enc.encode_strip_begin(d_img.data_ptr(), p, 0, False, W, 96, W * 3, run.data_ptr() + 8)
        enc.encode_strip_finish(out.data_ptr(), out.numel(), run.data_ptr())
        run[1] += run[0]
        enc.encode_strip_begin(d_img.data_ptr() + 96 * W * 3, p, 6, True, W, H - 96, W * 3, run.data_ptr() + 16)
        enc.encode_strip_finish(out.data_ptr(), out.numel(), run.data_ptr() + 8)
        run[2] += run[1]
    enc.sync()
    end = int(run[2].item())
    assert end - 5 == len(want) and np.array_equal(out[5:end].cpu().numpy(), want)
    for dst_off, src_off in ((0, 5), (3, 5), (16, 6), (7, 8)):
        dst = torch.zeros(W * H, dtype=torch.uint8, device="cuda")
        o = torch.tensor([dst_off], dtype=torch.int64, device="cuda")
        ln = torch.tensor([end - src_off], dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        enc.copy_bytes_device(dst.data_ptr(), dst.numel(), o.data_ptr(), out.data_ptr() + src_off, ln.data_ptr())
        enc.sync()
        assert torch.equal(dst[dst_off: dst_off + end - src_off], out[src_off:end]) and not dst[dst_off + end - src_off:][:64].any()
        assert not dst[:dst_off].any()
    # a destination that is too small is reported at jb_sync and nothing is written
    enc.encode_strip_begin(d_img.data_ptr(), p, 0, True, W, H, W * 3, run.data_ptr())
    enc.encode_strip_finish(out.data_ptr(), 100, run.data_ptr() + 8)
    with pytest.raises(jb.JbError) as e:
        enc.sync()
    assert e.value.code == jb.E_NOSPACE


def test_tma_staged_transform_is_bit_identical(enc, jb, fruit):
    """JB_FLAG_TMA: the 4:2:0 tcgen05 transform with TMA-staged image tiles (k_transform_tma: cp.async.bulk.tensor boxes of a
    3-D tensor map, units of 8 x 2 MCUs, rows below the image fetched from their mirror image) == the oracle, for widths
    that are / are not multiples of the 128-pixel unit, heights with a mirrored last MCU row, odd numbers of MCU rows,
    batches (frame coordinate of the tensor map) and whole 1080p frames."""
    ql, qc = ol.quality_tables(75)
    p = jb.make_params(ol.SUB_420, qlum=ql, qchrom=qc, flags=jb.FLAG_TMA)
    p0 = jb.make_params(ol.SUB_420, qlum=ql, qchrom=qc)
    for W, H in ((1920, 1080), (128, 32), (256, 40), (1024, 16), (2048, 1000), (144, 72), (16, 16), (4096, 56), (336, 264)):
        img = ol.synth(W * 7 + H, W, H)
        got, want = enc.transform(img, p), ol.transform(img, ol.SUB_420, ql, qc)
        assert np.array_equal(got, want), f"{W}x{H}: " + mismatch_report(got, want)
        assert enc.encode_jfif(img, p) == enc.encode_jfif(img, p0) == ol.encode_jfif(img, ol.SUB_420, ql, qc, 0)
    frames = np.stack([ol.synth(300 + f, 640, 360) for f in range(7)])
    out, offs, sizes = enc.encode_batch(frames, p)
    for f in range(7):
        assert bytes(out[int(offs[f]): int(offs[f] + sizes[f])]) == ol.encode_jfif(frames[f], ol.SUB_420, ql, qc, 0), f
    noise = noise_image(5, 1280, 720)
    q100 = ol.quality_tables(100)
    pn = jb.make_params(ol.SUB_420, qlum=q100[0], qchrom=q100[1], flags=jb.FLAG_TMA)
    assert np.array_equal(enc.transform(noise, pn), ol.transform(noise, ol.SUB_420, *q100))


@pytest.mark.parametrize("sub", SUBS)
def test_tiles_are_independent_jfif_files_equal_to_the_oracle(enc, jb, sub):
    """jb_encode_tiles (images beyond SOF0's 16-bit dimensions): every tile is a complete JFIF file equal to the
    oracle's encode of the cropped tile -- partial last column / row mirror-padded like any image -- and decodes on its own."""
    import io
    from PIL import Image
    ql, qc = ol.quality_tables(75)
    p = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=3)
    img = ol.synth(0x7111, 200, 104)
    for tw, th in ((64, 48), (208, 112), (96, 16)):
        tiles = enc.encode_tiles(img, tw, th, p)
        assert len(tiles) == -(-104 // th) and len(tiles[0]) == -(-200 // tw)
        for ty, row in enumerate(tiles):
            for tx, jf in enumerate(row):
                crop = np.ascontiguousarray(img[ty * th: (ty + 1) * th, tx * tw: (tx + 1) * tw])
                assert jf == ol.encode_jfif(crop, sub, ql, qc, 3), (tw, th, tx, ty)
                assert np.array(Image.open(io.BytesIO(jf))).shape == crop.shape
    m = 16 if sub == ol.SUB_420 else 8
    with pytest.raises(jb.JbError):
        enc.encode_tiles(img, 60, 48, p)  # not a multiple of the MCU
    # wider than 65535 without JB_FLAG_CLAMP_SOF: two tiles, both exact
    wide = ol.synth(0x7112, 65536 + 4 * m, m)
    tiles = enc.encode_tiles(wide, 65536 - m, m, p)
    assert [len(r) for r in tiles] == [2]
    assert tiles[0][0] == ol.encode_jfif(np.ascontiguousarray(wide[:, : 65536 - m]), sub, ql, qc, 3)
    assert tiles[0][1] == ol.encode_jfif(np.ascontiguousarray(wide[:, 65536 - m:]), sub, ql, qc, 3)
    # a buffer that is too small reports the size of the whole grid
    with pytest.raises(jb.JbError) as e:
        enc.encode_tiles(img, 64, 48, p, cap=1000)
    assert e.value.code == jb.E_NOSPACE
    need = enc.L.jb_required_bytes(enc.h)
    assert need == sum(len(t) for r in enc.encode_tiles(img, 64, 48, p) for t in r)


def test_nv12_device_input(enc, jb):
    """jb_encode_nv12_device: frames that are YCbCr 4:2:0 already.  (1) jb_rgb8_to_nv12_device == the
    reference's CSC + CDS; (2) for even sizes NV12 made from RGB encodes to the very file the RGB path produces (the
    identity that pins the path to the reference); (3) arbitrary NV12 content, odd sizes and pitches, batches ==
    the oracle's stages from the mirror padding on (orc_transform_ycc)."""
    import torch
    ql, qc = ol.quality_tables(75)
    p = jb.make_params(ol.SUB_420, qlum=ql, qchrom=qc, restart_interval=5)
    rng = np.random.default_rng(12)

    def encode(y, uv, W, H, N=1, pitch_y=None, pitch_uv=None):
        pitch_y, pitch_uv = pitch_y or y.shape[-1], pitch_uv or uv.shape[-1]
        dy, duv = torch.from_numpy(np.ascontiguousarray(y)).cuda(), torch.from_numpy(np.ascontiguousarray(uv)).cuda()
        cap = N * (W * H * 3 + 4096)
        out, tab = torch.zeros(cap, dtype=torch.uint8, device="cuda"), torch.zeros(2 * N + 1, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        enc.encode_nv12_device(dy.data_ptr(), pitch_y, pitch_y * H, duv.data_ptr(), pitch_uv, pitch_uv * ((H + 1) // 2), N, W, H, p,
                               out.data_ptr(), cap, tab.data_ptr(), tab.data_ptr() + 8 * N, tab.data_ptr() + 16 * N)
        enc.sync()
        t = tab.cpu().numpy()
        return [bytes(out[int(t[f]): int(t[f] + t[N + f])].cpu().numpy()) for f in range(N)]

    # (16-byte aligned planes take the tcgen05 kernel -- complete MCU columns, rows below the image mirrored in the kernel,
    # odd heights included; the others the CUDA-core kernel)
    for W, H in ((64, 48), (1920, 1080), (250, 130), (253, 131), (37, 21), (16, 16), (64, 37), (208, 75), (48, 1081 // 7)):
        rgb = ol.synth(W + H, W, H)
        y, uv = ol.nv12_from_rgb(rgb)
        d_rgb = torch.from_numpy(rgb.copy()).cuda()
        dy = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
        duv = torch.zeros(((H + 1) // 2, 2 * ((W + 1) // 2)), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        enc.rgb8_to_nv12_device(d_rgb.data_ptr(), W, H, W * 3, dy.data_ptr(), W, duv.data_ptr(), duv.shape[1])
        enc.sync()
        assert np.array_equal(dy.cpu().numpy(), y) and np.array_equal(duv.cpu().numpy(), uv), (W, H)
        got = encode(y, uv, W, H)[0]
        want = ol.jfif_from_coef(ol.transform_ycc(ol.ycc_from_nv12(y, uv), ol.SUB_420, ql, qc), W, H, ol.SUB_420, ql, qc, 5)
        assert got == want, (W, H)
        if W % 2 == 0 and H % 2 == 0:
            assert got == ol.encode_jfif(rgb, ol.SUB_420, ql, qc, 5) == enc.encode_jfif(rgb, p), (W, H)
    # arbitrary content (full-range noise: many near ties -> the binary64 replay reads the planes), pitches, a batch
    W, H, N = 200, 72, 3
    py, puv = 208, 224
    y = rng.integers(0, 256, (N, H, py), dtype=np.uint8)
    uv = rng.integers(0, 256, (N, H // 2, puv), dtype=np.uint8)
    files = encode(y, uv, W, H, N, py, puv)
    for f in range(N):
        ycc = ol.ycc_from_nv12(np.ascontiguousarray(y[f, :, :W]), np.ascontiguousarray(uv[f, :, :W]))
        assert files[f] == ol.jfif_from_coef(ol.transform_ycc(ycc, ol.SUB_420, ql, qc), W, H, ol.SUB_420, ql, qc, 5), f
    assert enc.timings()["tie_fixups"] > 0
    with pytest.raises(jb.JbError) as e:
        pp = jb.make_params(ol.SUB_444, qlum=ql, qchrom=qc)
        d = torch.zeros(64 * 64, dtype=torch.uint8, device="cuda")
        enc.encode_nv12_device(d.data_ptr(), 64, 64 * 64, d.data_ptr(), 64, 64 * 32, 1, 64, 64, pp, d.data_ptr(), 64 * 64, 0, 0, 0)
    assert e.value.code == jb.E_UNSUPPORTED


def test_nv12_host_batch(enc, jb):
    """jb_encode_nv12_batch: NV12-style frames in HOST memory -> host JFIF files (the NV12 counterpart of jb_encode_batch).
    Files == the oracle's stages from the mirror padding on, for the smallest sizes the mirror padding allows, odd sizes, padded host pitches, and across several groups of the host pipeline; for even
    sizes also == the RGB path's file (NV12 made with the reference's CSC + CDS); too small an output reports what it needs."""
    ql, qc = ol.quality_tables(75)
    p = jb.make_params(ol.SUB_420, qlum=ql, qchrom=qc, restart_interval=3)
    rng = np.random.default_rng(77)

    def want(y, uv, W, H):
        ycc = ol.ycc_from_nv12(np.ascontiguousarray(y[:, :W]), np.ascontiguousarray(uv[:, :2 * ((W + 1) // 2)]))
        return ol.jfif_from_coef(ol.transform_ycc(ycc, ol.SUB_420, ql, qc), W, H, ol.SUB_420, ql, qc, 3)

    for W, H, N, pad in ((64, 48, 2, 0), (250, 131, 3, 6), (8, 8, 2, 0), (9, 11, 2, 1), (16, 16, 1, 0), (33, 17, 4, 0), (1920, 1080, 2, 0)):
        cw, ch = 2 * ((W + 1) // 2), (H + 1) // 2
        y = rng.integers(0, 256, (N, H, W + pad), dtype=np.uint8)
        uv = rng.integers(0, 256, (N, ch, cw + pad), dtype=np.uint8)
        out, offs, sizes = enc.encode_nv12_batch(y, uv, W, p)
        for f in range(N):
            assert bytes(out[int(offs[f]): int(offs[f] + sizes[f])]) == want(y[f], uv[f], W, H), (W, H, f)
    # even sizes: the RGB path's own files
    W, H = 96, 64
    rgbs = [ol.synth(40 + f, W, H) for f in range(3)]
    planes = [ol.nv12_from_rgb(r) for r in rgbs]
    y, uv = np.stack([a for a, _ in planes]), np.stack([b for _, b in planes])
    out, offs, sizes = enc.encode_nv12_batch(y, uv, W, p)
    for f in range(3):
        assert bytes(out[int(offs[f]): int(offs[f] + sizes[f])]) == enc.encode_jfif(rgbs[f], p) == ol.encode_jfif(rgbs[f], ol.SUB_420, ql, qc, 3)
    # several groups of the host pipeline (a group is ~96 MB of RGB-equivalent frames): 40 x 1080p = 3 groups
    W, H, N = 1920, 1080, 40
    y1, uv1 = ol.nv12_from_rgb(ol.synth(5, W, H))
    y = np.ascontiguousarray(np.broadcast_to(y1, (N, H, W))).copy()
    uv = np.ascontiguousarray(np.broadcast_to(uv1, (N, H // 2, W))).copy()
    y[:, 0, 0] = np.arange(N)  # every frame differs
    out, offs, sizes = enc.encode_nv12_batch(y, uv, W, p)
    assert [int(o) for o in offs] == [int(x) for x in np.concatenate(([0], np.cumsum(sizes[:-1])))]
    for f in (0, 17, 39):
        assert bytes(out[int(offs[f]): int(offs[f] + sizes[f])]) == want(y[f], uv[f], W, H), f
    small = np.empty(int(sizes.sum()) - 1, np.uint8)
    with pytest.raises(jb.JbError) as e:
        enc.encode_nv12_batch(y, uv, W, p, out=small)
    assert e.value.code == jb.E_NOSPACE and enc.L.jb_required_bytes(enc.h) == int(sizes.sum())
    with pytest.raises(jb.JbError) as e:
        enc.encode_nv12_batch(y[:1], uv[:1], W, jb.make_params(ol.SUB_444, qlum=ql, qchrom=qc))
    assert e.value.code == jb.E_UNSUPPORTED


@pytest.mark.parametrize("sub,q,ri", [(ol.SUB_420, 75, 0), (ol.SUB_444, 90, 0), (ol.SUB_REPL420, 50, 0), (ol.SUB_420, 95, 11)])
def test_fruit_tiled_stress(enc, jb, fruit, sub, q, ri):
    """SURVEY 8d's second input distribution: fruit.ppm tiled (a dithered image: very high-frequency, long codes, ZRL, blocks
    longer than a 128-bit slot, many 0xFF bytes).  Whole files == the oracle, batch == single frames."""
    H, W = 3 * 254 + 17, 4 * 253 + 5  # odd sizes: mirror padding on both edges, CDS edge rule
    img = np.tile(fruit, (4, 5, 1))[:H, :W].copy()
    p, ql, qc = _params(jb, sub, q, ri)
    got = enc.encode_jfif(img, p)
    assert got == ol.encode_jfif(img, sub, ql, qc, ri), (len(got),)
    frames = np.stack([np.roll(img, 7 * i, axis=1) for i in range(3)])
    out, offs, sizes = enc.encode_batch(frames, p)
    for i in range(3):
        assert out[int(offs[i]): int(offs[i] + sizes[i])].tobytes() == ol.encode_jfif(frames[i], sub, ql, qc, ri), i
