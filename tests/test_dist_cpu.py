"""N>1 host logic on CPU: world_size-2 gloo — utterance sharding + the single waveform all_gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ssr_speech_b200.dist import gather_waveforms, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_items, rank, world)
    # utterance i "generates" a waveform of length 100 + 7*i filled with i
    local = [torch.full((1, 100 + 7 * i), float(i)) for i in range(lo, hi)]
    out = gather_waveforms(local, device=torch.device("cpu"))
    ok = len(out) == n_items and all(o.shape == (1, 100 + 7 * i) and bool((o == i).all()) for i, o in enumerate(out))
    out_h = gather_waveforms(local, to_host=True)                  # default device (gloo: the tensors' own), one host copy
    ok = ok and len(out_h) == n_items and all(torch.equal(a, b) for a, b in zip(out, out_h))
    q.put((rank, ok, lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_items = 5                     # ragged: rank 0 gets 3, rank 1 gets 2
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, 0, 3), (1, True, 3, 5)]


def test_shard_and_gather_world2_with_an_empty_shard():
    """1 utterance over 2 ranks: rank 1 holds nothing and still takes part in both collectives."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, 0, 1), (1, True, 1, 1)]


def test_gather_single_process_passthrough():
    local = [torch.ones(1, 10), torch.zeros(1, 3)]
    out = gather_waveforms(local)
    assert len(out) == 2 and out[0].shape == (1, 10) and out[1].shape == (1, 3)
