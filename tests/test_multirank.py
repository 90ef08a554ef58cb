"""The N > 1 path on the CPU: two `gloo` ranks shard a Monte-Carlo run by frame block (drand48 jump-ahead), decode their
frames and meet in one collective.  The host logic under test is the product's (multigpu.py, Code frame source,
statistics rule); the decoder plugged in is the oracle, because no GPU exists here -- the GPU tests check the CUDA decoder
against the same oracle."""
import multiprocessing as mp
import os
import socket

import numpy as np
import pytest

import nbldpc
import oracle_lib as ol
from common import matrix_path

multigpu = __import__("importlib").import_module("ems-decoder-of-nb-ldpc-codes_b200.multigpu")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, rel, frames, ebn, n_m, nb_oper, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        path = matrix_path(rel)
        code = nbldpc.Code(path)
        o = ol.Oracle(path, code.dialect)

        def decode(noisy, sigma):
            out = [o.decode_frame(o.channel_llr(n, sigma), n_m, nb_oper, 10, 0.3) for n in noisy]
            return (np.stack([r["decide"] for r in out]), np.array([r["synd"] for r in out]), np.array([r["iters"] for r in out]))

        st = multigpu.monte_carlo(code, frames, ebn, decode, batch=64, rank=rank, world=world)
        tot = multigpu.allreduce_counters(np.array([1, rank, frames], np.int64))
        q.put((rank, st, tot.tolist()))
    finally:
        dist.destroy_process_group()


def _run(world, *args):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port) + args + (q,)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    return {r: (st, tot) for r, st, tot in res}


def test_frame_ranges_partition_the_run():
    for total in (1, 7, 2000, 2001):
        for world in (1, 2, 3, 8):
            r = [multigpu.frame_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total and all(r[k][1] == r[k + 1][0] for k in range(world - 1))


@pytest.mark.parametrize("frames,ebn,expect", [(2000, 3.0, dict(err_frames=24, bit_errors=153, frames=2000, file=2001, avr="1.56")),
                                               (400, 1.0, None)])
def test_two_ranks_reproduce_the_single_process_statistics(frames, ebn, expect):
    """known answer of the stock binary (`2000 10 matrices/N96_K48_GF64 3.0 20 0.3 25`), and a run that hits the
    40-erroneous-frames stop rule in the middle of rank 0's block"""
    rel, n_m, nb_oper = "matrices/N96_K48_GF64", 20, 25
    res = _run(2, rel, frames, ebn, n_m, nb_oper)
    st, tot = res[0]
    assert res[1][0] is None and tot == [2, 1, 2 * frames] and res[1][1] == tot
    o = ol.Oracle(matrix_path(rel))
    s = o.monte_carlo(frames, ebn, n_m, nb_oper, 10, 0.3)       # [frames shown, err frames, undetected, bit errors, sum_it, nb]
    o.close()
    assert [st["frames"], st["err_frames"], st["undetected"], st["bit_errors"], st["sum_it"], st["frames_in_results_file"]] == s
    if expect:
        assert st["err_frames"] == expect["err_frames"] and st["bit_errors"] == expect["bit_errors"]
        assert st["frames_in_results_file"] == expect["file"] and "%.2f" % (st["sum_it"] / st["frames"]) == expect["avr"]
    else:
        assert st["stopped"] and st["err_frames"] == 40 and st["frames"] < frames
