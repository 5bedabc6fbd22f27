"""Multi-GPU check of the record gather (run under torchrun, one rank per GPU; tests/test_gpu_multi.py spawns it):
every rank submits STEPS records with a (rank, step)-dependent pattern through the chosen gatherer and verifies that each
gathered buffer holds the blocks of ALL ranks for exactly that step (a late / early / torn PUT shows up as a wrong step id).
Consumers read the returned buffer behind a few milliseconds of other queued work (still before their next submit call,
as the contract demands) while other ranks run ahead, to exercise the slot-reuse protocol.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gather_check.py copy
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else 'copy'
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    os.environ['SPECB200_GATHER'] = mode
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    import spec_b200 as sb
    from spec_b200.constants import RECORD_FLOATS
    g, desc = sb.make_gatherer(B, dev)
    if rank == 0:
        print('gatherer:', desc, flush=True)
    if mode != 'nccl' and not isinstance(g, sb.PeerGatherer):
        raise SystemExit('peer gather was requested but the NCCL fallback was taken')
    rec = torch.empty(B, RECORD_FLOATS, device=dev)
    col = torch.arange(RECORD_FLOATS, device=dev, dtype=torch.float32) * 1e-3
    spin = torch.empty(8 << 20, device=dev)
    bad = torch.zeros(1, device=dev)

    def expected(step):
        r = torch.arange(world, device=dev, dtype=torch.float32).repeat_interleave(B)
        return (r * 1000 + step)[:, None] + col[None, :]

    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(steps):
        if rank == (i % world):
            for _ in range(3):
                spin.mul_(1.0001)                                   # skew the ranks: someone is always late
        rec.copy_((torch.full((B, 1), rank * 1000.0 + i, device=dev) + col[None, :]))
        prev = g.submit(rec)
        if prev is not None:
            if rank == ((i + 1) % world):
                for _ in range(2):
                    spin.mul_(1.0001)                               # slow consumer: the read sits behind queued work
            bad += (prev != expected(i - 1)).any().float()
    last = g.flush()
    bad += (last != expected(steps - 1)).any().float()
    t1.record()
    torch.cuda.synchronize(dev)
    dist.all_reduce(bad)
    if rank == 0:
        print(f'mode={mode} world={world} steps={steps} B={B}: mismatching buffers = {int(bad.item())}, {t0.elapsed_time(t1) / steps:.3f} ms/step', flush=True)
    if hasattr(g, 'close'):
        g.close()
    dist.barrier()
    dist.destroy_process_group()
    if bad.item() != 0:
        raise SystemExit(1)


if __name__ == '__main__':
    main()
