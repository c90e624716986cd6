"""Extreme shapes against the oracle: maximal SOF width / height, single MCU rows / columns, 8 Mpx odd sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
import oracle_lib as ol
jb = g.load(); enc = jb.Encoder(0)
bad = 0
for (W, H) in [(65535, 16), (16, 65535), (65535, 9), (9, 4000), (8191, 1001), (4000, 2003), (65528, 24), (40, 40)]:
    for sub in (ol.SUB_420, ol.SUB_444, ol.SUB_REPL420):
        m = 16 if sub == ol.SUB_420 else 8
        if (-W) % m > W or (-H) % m > H: continue
        img = ol.synth(W * 7 + H, W, H)
        ql, qc = ol.quality_tables(75)
        ri = -(-W // m) if W * H > 100000 else 0
        p = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=min(ri, 65535))
        t0 = time.time()
        got = enc.encode_jfif(img, p, cap=W * H * 3 + (1 << 20))
        want = ol.encode_jfif(img, sub, ql, qc, min(ri, 65535))
        ok = got == want
        bad += not ok
        print((W, H), sub, "ok" if ok else "MISMATCH", len(got), len(want), round(time.time() - t0, 1), flush=True)
print("extremes done,", bad, "bad")
