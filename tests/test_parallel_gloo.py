"""world_size-2 `gloo` test of the multi-GPU plumbing (weight broadcast, image
sharding, detection all-gather) -- runs on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from luminoth_b200 import parallel as P


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    specs = [('a/w', (3, 3, 4, 8)), ('a/b', (8,)), ('fc/w', (16, 5))]
    rng = np.random.default_rng(0)
    weights = {n: rng.standard_normal(s).astype(np.float32) for n, s in specs} if rank == 0 else None
    got = P.broadcast_weights(weights, specs, torch.device('cpu'))
    ref = {n: np.random.default_rng(0).standard_normal((1,)) for n in ()}  # noqa: F841
    rng = np.random.default_rng(0)
    want = {n: rng.standard_normal(s).astype(np.float32) for n, s in specs}
    ok = all(np.array_equal(got[n], want[n]) for n, _ in specs)
    # sharding: 11 images over 2 ranks -> [0,6) and [6,11)
    lo, hi = P.shard_range(11, rank, world)
    B, K = hi - lo, 7
    g = torch.Generator().manual_seed(100 + rank)
    boxes = torch.rand((B, K, 4), generator=g); scores = torch.rand((B, K), generator=g)
    labels = torch.randint(0, 80, (B, K), generator=g, dtype=torch.int32)
    counts = torch.randint(0, K + 1, (B,), generator=g, dtype=torch.int32)
    # equal-size records are required by all_gather: pad the short shard
    Bmax = 6
    rec = torch.zeros((Bmax, P.record_width(K)))
    rec[:B] = P.pack_detections(boxes, scores, labels, counts)
    allrec = P.all_gather_detections(rec)
    ok = ok and allrec.shape == (world * Bmax, P.record_width(K))
    # EVERY rank's slice must hold that rank's detections: regenerate what rank r packed (same seeded generator)
    # and compare slice r of the gathered tensor with it -- the own slice and, crucially, the peers' slices
    for r in range(world):
        lo_r, hi_r = P.shard_range(11, r, world)
        Br = hi_r - lo_r
        gr = torch.Generator().manual_seed(100 + r)
        boxes_r = torch.rand((Br, K, 4), generator=gr); scores_r = torch.rand((Br, K), generator=gr)
        labels_r = torch.randint(0, 80, (Br, K), generator=gr, dtype=torch.int32)
        counts_r = torch.randint(0, K + 1, (Br,), generator=gr, dtype=torch.int32)
        b2, s2, l2, c2 = P.unpack_detections(allrec[r * Bmax:r * Bmax + Br], K)
        ok = ok and torch.equal(b2, boxes_r) and torch.equal(s2, scores_r) and torch.equal(l2, labels_r)
        ok = ok and torch.equal(c2, counts_r)
        ok = ok and bool((allrec[r * Bmax + Br:(r + 1) * Bmax] == 0).all())       # the padding rows stay zero
    ok = ok and not torch.equal(allrec[:Bmax], allrec[Bmax:])                      # the two slices really differ
    q.put((rank, bool(ok), (lo, hi)))
    dist.destroy_process_group()


def test_broadcast_shard_allgather_world2():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [(0, 6), (6, 11)]


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            spans = [P.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_pack_unpack_weights_roundtrip():
    specs = [('x', (2, 3)), ('y', (4,))]
    w = {'x': np.arange(6, dtype=np.float32).reshape(2, 3), 'y': np.ones(4, np.float32)}
    flat = P.pack_weights(w, specs)
    back = P.unpack_weights(flat, specs)
    assert all(np.array_equal(back[k], w[k]) for k in w)
