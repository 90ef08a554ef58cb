"""Small decode of every kernel family for compute-sanitizer runs (developer helper, run under gpurun):
   compute-sanitizer --tool memcheck python scripts/sanitize_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nbldpc  # noqa: E402
from common import matrix_path, product_frames  # noqa: E402

for rel, n_m, kw in (("matrices/N96_K48_GF64", 20, {}), ("matrices/KN/N96_K48_GF256.txt", 20, {}),
                     ("matrices/N96_K48_GF64", 20, dict(ecn_kind=1)), ("matrices/Mat28_N72_M18", 12, {})):
    code = nbldpc.Code(matrix_path(rel))
    fr, sigma = product_frames(code, 24, 2.5)
    noisy = np.stack([f["noisy"] for f in fr])
    d = nbldpc.Decoder(code, n_m, 25, 6, 0.3, max_batch=24, **kw)
    dec, synd, it = d.decode_noisy(noisy, sigma)
    rng = np.random.default_rng(0)
    d.select_nm((rng.random((8, code.q)) * 30).astype(np.float32))
    d.channel(noisy[:2], sigma, want_sorted=True)
    print(rel, kw, "iters", it.tolist()[:8], "launches", d.launch_count())
    d.close()
print("sanitize case done")
