"""Frame sharding over the GPUs of one node (one process per GPU, torch.distributed) -- host logic only.

Frames of a Monte-Carlo run are independent: frame f consumes the drand48 draws
[f*D, (f+1)*D) with D = (K + 2N) * log2(q) (K*log2(q) information bits, tools.c:124-136, then two draws per
transmitted bit, channel.c:56-57), so every rank can generate its own frames by LCG jump-ahead.  There is no
collective on the decode path.  Results meet in ONE collective:

  * `allreduce_counters` -- sum of the int64 counters (throughput runs: frames, non-zero syndromes, iterations, ...);
  * `monte_carlo` -- gathers (bit errors, syndrome, iterations) per frame on rank 0, which applies the reference's
    sequential rule "stop after the 40th erroneous frame" (NB_LDPC.c:506) in frame order, so the statistics are
    the single-process ones whatever the number of ranks.

torch is used for the process group only (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def draws_per_frame(code):
    """drand48 draws one frame of the reference's main loop consumes (tools.c:124-136 + channel.c:56-57)"""
    return (code.K + 2 * code.N) * code.logq


def frame_range(total, rank, world):
    """Contiguous block of frame indices [lo, hi) of `rank` ([g*F/G, (g+1)*F/G))"""
    return (rank * total) // world, ((rank + 1) * total) // world


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def allreduce_counters(counters, device=None):
    """Sum an int64 counter vector over all ranks (the only collective of a throughput run)."""
    c = np.ascontiguousarray(counters, np.int64)
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return c.copy()
    import torch
    t = torch.from_numpy(c.copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t)
    return t.cpu().numpy()


def gather_rows(rows):
    """Concatenate per-rank int64 [n_r, k] arrays on rank 0 in rank order (None elsewhere)."""
    rows = np.ascontiguousarray(rows, np.int64)
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return rows
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(rows, out, dst=0)
    return np.concatenate(out, axis=0) if out is not None else None


def monte_carlo(code, frames, ebn, decode=None, batch=256, rank=0, world=1, max_err_frames=40, decoder=None):
    """The reference's Monte-Carlo loop (NB_LDPC.c:250-511) sharded by frame block over `world` ranks.

    Either decode(noisy[B, N, logq], sigma) -> (decide[B, N], synd[B], iters[B]) with frames made by the host source
    (Decoder.decode_noisy), or decoder=<Decoder>: frames are generated, decoded and scored on the rank's GPU
    (nbgpu_source_frames / nbgpu_run / nbgpu_source_results) and only 12 bytes per frame reach the host.
    Returns on rank 0 the dict of the reference's statistics, None on the other ranks."""
    lo, hi = frame_range(frames, rank, world)
    D = draws_per_frame(code)
    code.prepare_encoder()
    code.rng_default()
    code.rng_skip(lo * D)
    sigma = code.sigma(ebn)
    bing = code.tables()[0]
    rows = np.zeros((hi - lo, 3), np.int64)
    for b0 in range(lo, hi, batch):
        b1 = min(hi, b0 + batch)
        if decoder is not None:
            decoder.source_frames(b0, b1 - b0, ebn)
            decoder.run()
            rows[b0 - lo:b1 - lo] = np.stack(decoder.source_results(), axis=1)
            continue
        bits, noisy = [], []
        for _ in range(b0, b1):
            _, nbin = code.random_codeword()
            bits.append(nbin)
            noisy.append(code.noise(nbin, ebn))
        dec, synd, it = decode(np.stack(noisy), sigma)
        for i in range(b1 - b0):
            err = int((bing[dec[i, :code.K]] != bits[i][:code.K]).sum())                    # NB_LDPC.c:479-485
            rows[b0 - lo + i] = (err, int(synd[i]), int(it[i]))
    allrows = gather_rows(rows)
    if allrows is None:
        return None
    stats = dict(frames=0, err_frames=0, undetected=0, bit_errors=0, sum_it=0, stopped=False)
    for err, synd, it in allrows:                                                          # NB_LDPC.c:474-507, frame order
        stats["frames"] += 1
        stats["sum_it"] += int(it)
        stats["bit_errors"] += int(err)
        if err:
            stats["err_frames"] += 1
            if synd == 0:
                stats["undetected"] += 1
        if stats["err_frames"] == max_err_frames:
            stats["stopped"] = True
            break
    stats["frames_in_results_file"] = stats["frames"] if stats["stopped"] else frames + 1    # 'nb' after the loop, :576
    return stats
