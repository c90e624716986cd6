"""Regenerates tests/golden/* from the reference itself.  Run in the build container
(needs /root/reference and oracle/_ref/libjpegref.so = the reference's src/utils.cpp
compiled unmodified, see oracle/Makefile):

    python tests/golden/make_golden.py

Outputs
  fruit.ppm            the reference's only fixture (data/fruit.ppm, 253x254 P6), copied verbatim
  reference_golden.json
      digests of the reference's outputs on fruit.ppm (as written, and with the DCT block
      read from a copy), its Huffman code strings (huffman.hpp), known answers of
      getValueCategory / valueToBitString, the CSC of the full 2^24 colour cube (sha256),
      and per-stage outputs on a small seeded random image.
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

sha = lambda b: hashlib.sha256(bytes(b)).hexdigest()


def main():
    R = ol.ref()
    assert R is not None, "build oracle/_ref first (make -C oracle)"
    shutil.copyfile("/root/reference/data/fruit.ppm", os.path.join(HERE, "fruit.ppm"))
    rgb = ol.read_ppm(os.path.join(HERE, "fruit.ppm"))
    g = {"fruit": {}}
    for name, mode in (("as_written", 0), ("dct_from_copy", 1)):
        r = ol.ref_pipeline(rgb, mode)
        g["fruit"][name] = dict(ycc_sha256=sha(r["ycc"].tobytes()), zigzag_sha256=sha(r["zigzag"].tobytes()),
                                nbits=r["nbits"], bits_sha256=sha(r["bits"]))
    # Huffman code strings exactly as huffman.hpp holds them
    tabs = {}
    for t, nrun in ((0, 1), (1, 1), (2, 16), (3, 16)):
        ncat = 12 if t < 2 else 11
        tabs[str(t)] = [[R.ref_table_code(t, run, cat).decode() for cat in range(ncat)] for run in range(nrun)]
    g["huffman_tables"] = tabs
    import ctypes as C
    buf = C.create_string_buffer(64)
    g["value_kat"] = {}
    for v in list(range(-40, 41)) + [-2047, -1024, -1023, -512, -255, 255, 511, 1023, 1024, 2047]:
        n = R.ref_valueToBitString(v, buf)  # fill the buffer first, then read it
        g["value_kat"][str(v)] = [R.ref_getValueCategory(v), buf.raw[:n].decode()]
    # CSC over the whole colour cube, r major / b minor, through the reference's performCSC
    h = hashlib.sha256()
    ydown = 0
    for r_ in range(256):
        cube = np.zeros((256, 256, 3), np.uint8)
        cube[:, :, 0] = r_
        cube[:, :, 1] = np.arange(256)[:, None]
        cube[:, :, 2] = np.arange(256)[None, :]
        R.ref_performCSC(cube.reshape(-1), 256, 256)
        h.update(cube.tobytes())
    g["csc_cube_sha256"] = h.hexdigest()
    # quantisation tables
    ql = np.zeros(64, np.uint32)
    qc = np.zeros(64, np.uint32)
    R.ref_quant_tables(ql, qc)
    g["quant_lum"], g["quant_chrom"] = ql.tolist(), qc.tolist()
    # a small seeded random image (odd sizes -> padding + odd-edge chroma rule)
    rng = np.random.default_rng(20261018)
    small = rng.integers(0, 256, (21, 19, 3), dtype=np.uint8)
    np.save(os.path.join(HERE, "small_rgb.npy"), small)
    for name, mode in (("as_written", 0), ("dct_from_copy", 1)):
        r = ol.ref_pipeline(small, mode)
        np.savez_compressed(os.path.join(HERE, f"small_{name}.npz"), ycc=r["ycc"], zigzag=r["zigzag"],
                            bits=np.frombuffer(r["bits"], np.uint8))
    with open(os.path.join(HERE, "reference_golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote golden fixtures:", {k: (v if isinstance(v, str) else "...") for k, v in g.items()})


if __name__ == "__main__":
    main()
