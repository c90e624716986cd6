"""Scan the tensor-core near-tie band: for each JB_TC_ERR_SCALE count coefficients that differ from the
oracle after the binary64 replay (0 = band wide enough) and the number of replayed coefficients."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
    subs = [int(x) for x in os.environ.get("SUBS", "2").split(",")]  # 2 = 4:2:0, 0 = 4:4:4, 1 = replicated 4:2:0
    for sub in subs:
        for q in (50, 75, 95, 100):
            ql, qc = ol.quality_tables(q)
            for img in imgs:
                p = jb.make_params(sub, qlum=ql, qchrom=qc, flags=jb.FLAG_TENSOR_DCT)
                got = enc.transform(img, p); ties += enc.timings()["tie_fixups"]
                want = ol.transform(img, sub, ql, qc)
                bad += int((got != want).sum()); total += got.size
    print(json.dumps({"scale": float(sys.argv[1]), "mismatches": bad, "fixups": int(ties), "coefs": total}))
else:
    for s in ("2e-6", "1e-6", "5e-7", "2e-7", "1e-7", "5e-8", "2e-8", "0"):
        subprocess.run([sys.executable, __file__, s])
