"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous batch sharding and the single
all-gather of packed per-image records.  The data path has no other collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spec_b200.pipeline import shard_range, all_gather_records, unpack_record, RecordGatherer, RECORD_FLOATS


def test_shard_range_covers_batch():
    for total in (1, 7, 8, 256, 2048, 2049):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    # record for image i is filled with a function of the GLOBAL image index
    idx = torch.arange(lo, hi, dtype=torch.float32)
    rec = idx[:, None] * 1000.0 + torch.arange(RECORD_FLOATS, dtype=torch.float32)[None, :] % 997
    full = all_gather_records(rec)
    ok = full.shape == (total, RECORD_FLOATS)
    exp = torch.arange(total, dtype=torch.float32)[:, None] * 1000.0 + torch.arange(RECORD_FLOATS, dtype=torch.float32)[None, :] % 997
    ok = ok and torch.equal(full, exp)
    d = unpack_record(full)
    ok = ok and d['smpl_vertices'].shape == (total, 6890, 3) and d['cam_angles'].shape == (total, 3)
    ok = ok and float(d['pred_cam_t'][total - 1, 0]) == float(exp[total - 1, 20670 + 147 + 98])
    # overlapped form used by the steady-state loop: results come back one step late, in order
    g = RecordGatherer(hi - lo, torch.device('cpu'))
    outs = []
    for step in range(3):
        prev = g.submit(rec + step)
        if prev is not None:
            outs.append(prev.clone())
    outs.append(g.flush().clone())
    ok = ok and len(outs) == 3 and all(torch.equal(o, exp + i) for i, o in enumerate(outs))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


@pytest.mark.timeout(180)
def test_all_gather_records_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, True), (1, True)]
