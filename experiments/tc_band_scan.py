"""Scan the tensor-core near-tie band: for each JB_TC_ERR_SCALE count coefficients that differ from the
oracle after the binary64 replay (0 = band wide enough) and the number of replayed coefficients."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    os.environ["JB_TC_ERR_SCALE"] = sys.argv[1]
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import __graft_entry__ as e
    import oracle_lib as ol
    jb = e.load(); enc = jb.Encoder(0)
    bad = ties = total = 0
    rng = np.random.default_rng(1)
    imgs = [ol.synth(3, 1920, 272), rng.integers(0, 256, (272, 960, 3), dtype=np.uint8),
            (rng.integers(0, 2, (272, 960, 3), dtype=np.uint8) * 255), ol.read_ppm(os.path.join(ROOT, "tests/golden/fruit.ppm"))]
    for q in (50, 75, 95, 100):
        ql, qc = ol.quality_tables(q)
        for img in imgs:
            p = jb.make_params(jb.SUB_420, qlum=ql, qchrom=qc, flags=jb.FLAG_TENSOR_DCT)
            got = enc.transform(img, p); ties += enc.timings()["tie_fixups"]
            want = ol.transform(img, jb.SUB_420, ql, qc)
            bad += int((got != want).sum()); total += got.size
    print(json.dumps({"scale": float(sys.argv[1]), "mismatches": bad, "fixups": int(ties), "coefs": total}))
else:
    for s in ("2.3e-5", "1e-5", "4e-6", "2e-6", "1e-6", "5e-7", "2e-7", "1e-7", "0"):
        subprocess.run([sys.executable, __file__, s])
