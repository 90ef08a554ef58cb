"""Where the end-to-end path spends its time beyond the kernel (run under gpurun): wall time of nbgpu_decode_noisy on pinned host
buffers, the device span between the first launch and the end of the last kernel, with and without chunking."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nbldpc
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "AD_64800_R12_GF256"
matrix, n_m, nb_oper, offset, ebn, B = bench.WORKLOADS[wl]
code = nbldpc.Code(bench.find_matrix(matrix))
noisy, bits, sigma = bench.synth_frames(code, B, ebn, 0)
nbldpc.pin(noisy)
out = (nbldpc.pin(np.zeros((B, code.N), np.int32)), nbldpc.pin(np.zeros(B, np.int32)), nbldpc.pin(np.zeros(B, np.int32)))
for mode in ("chunks", "single"):
    if mode == "single":
        os.environ["NBGPU_NO_CHUNKS"] = "1"
    d = nbldpc.Decoder(code, n_m, nb_oper, bench.NB_ITER_MAX, offset, early_stop=False, max_batch=B)
    for it in range(4):
        t0 = time.perf_counter()
        d.decode_noisy_into(noisy, sigma, out)
        wall = 1e3 * (time.perf_counter() - t0)
        print(mode, "wall %.2f ms  device span first launch -> last kernel end %.2f ms" % (wall, d.last_kernel_ms()))
    if mode == "single":
        t0 = time.perf_counter(); d.upload_noisy(noisy, sigma); t1 = time.perf_counter(); d.run(); d.sync(); t2 = time.perf_counter(); d.download(out); t3 = time.perf_counter()
        print("single: upload %.2f ms, run %.2f ms (kernel %.2f), download %.2f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), d.last_kernel_ms(), 1e3 * (t3 - t2)))
    d.close()
