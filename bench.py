#!/usr/bin/env python
"""bench.py -- images/sec of the SPEC full forward (CamCalib -> (R,K) -> HMR -> SMPL -> projection),
224x224, batch 256 per GPU (BASELINE.json configs[2]; weak scaling over GPUs), synthetic data, random weights.

  python bench.py --gpus 1 --steps 20 --warmup 5              # our arm (B200, libspecb200)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W               # N ranks, one per GPU, NCCL all-gather of records
  python bench.py --impl reference --steps 3 --warmup 3       # the reference path's CPU restatement (oracle)

Prints ONE JSON line (contract in the task statement): value = whole-job images/s with inputs resident in HBM,
e2e = same metric through the public API with pinned-host inputs (H2D and D2H inside the timed region),
roofline = conv FLOPs / event-timed conv-kernel time against the measured bf16 peak, cpu_baseline = the oracle
on the host cores for a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
warnings.simplefilter('ignore')

import torch  # noqa: E402

METRIC = 'images/sec SPEC fwd (224^2, b256)'
UNIT = 'images/s'
CONV_FLOPS_PER_IMAGE = {'resnet50': 2 * 4087136256}           # SURVEY.md B.1, per trunk


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tf_burst': d['bf16_tflops'], 'tf_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region through NVML, inline from the timing loop (``sample()`` is
    called once in the middle of the K timed steps, while the GPU is busy with the steps already enqueued).  Every NVML query
    stalls the GPU for a moment: polling the nvidia-smi binary at 10 Hz slowed the timed loop by ~20 %, and a 20 Hz in-process
    sampler thread showed up as ~3 ms outlier steps (``step_ms_spread``), so exactly one sample is taken per timed region
    (plus one every 64 steps for long runs).  Falls back to a slow nvidia-smi poll thread when pynvml is missing."""
    BITS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self.t = None
        self.mode = None
        self._nvml = None

    def sample(self):
        if self._nvml is None:
            return
        h, nv = self._nvml
        try:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons') \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for name, bit in self.BITS:
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _smi_loop(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        while not self._stop.is_set():
            try:
                o = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(', ')
                self.sm.append(float(o[0])); self.max_mhz = float(o[1])
                for name, v in zip([b[0] for b in self.BITS], o[2:6]):
                    if v.strip().lower().startswith('active'):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.5)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = int(vis.split(',')[self.index]) if vis and vis.split(',')[0].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.mode = 'nvml'
            self._nvml = (h, nv)
        except Exception:
            self.mode = 'nvidia-smi'
            self.t = threading.Thread(target=self._smi_loop, daemon=True)
            self.t.start()

    def stop(self):
        self._stop.set()
        if self.t is not None:
            self.t.join(3)
        sm = sorted(self.sm)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(sm), 'via': self.mode}


# =============================================================================================== reference arm / cpu baseline
def build_oracle(backbone):
    from oracle import models as om
    from spec_b200.synthetic import synthetic_smpl_data, synthetic_mean_params, randomize_module_
    torch.manual_seed(0)
    hmr = om.HMR(backbone, use_cam=True, use_cam_feats=True, smpl_data=synthetic_smpl_data(0),
                 mean_params=synthetic_mean_params(0)).eval()
    randomize_module_(hmr.backbone, 0)
    cc = om.CameraRegressorNetwork('resnet50').eval()
    randomize_module_(cc.backbone, 1)
    return cc, hmr


def pick_cpu_threads(cc, sample):
    """All the host threads the oracle can USE: eager PyTorch convs stop scaling (and collapse under
    oversubscription / cgroup quotas) well before 128 threads, so calibrate on one CamCalib forward and keep
    the fastest count among {8,16,32,64,all}."""
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, avail) if c <= avail}) or [avail]
    x = torch.randn(sample, 3, 224, 224)
    best, best_t = cands[-1], float('inf')
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            cc(x)
            t0 = time.perf_counter()
            cc(x)
            dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    return best


def time_oracle(backbone, sample, steps, warmup):
    """Oracle (CPU restatement of the reference path) on the host cores; returns (images/s, ms per sample step, threads)."""
    from oracle.models import spec_full_forward
    from spec_b200.synthetic import synthetic_batch
    cc, hmr = build_oracle(backbone)
    cores = pick_cpu_threads(cc, sample)
    torch.set_num_threads(cores)
    b = synthetic_batch(sample, 0)
    args = (b['images'], b['bbox_scale'], b['bbox_center'], b['img_w'], b['img_h'])
    for _ in range(warmup):
        spec_full_forward(cc, hmr, *args)
    t0 = time.perf_counter()
    for _ in range(steps):
        spec_full_forward(cc, hmr, *args)
    dt = time.perf_counter() - t0
    return sample * steps / dt, dt / steps * 1e3, cores


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return                                                  # other ranks exit 0 without work
    sample = a.cpu_sample
    ips, ms, cores = time_oracle(a.backbone, sample, a.steps, a.warmup)
    desc = f'{sample} of {a.batch} images per step (bounded CPU sample), fp32 PyTorch oracle, {cores} threads'
    _emit({
        'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'SPEC full forward (CamCalib->{a.backbone}->HMR head->SMPL->proj), 224x224, batch {a.batch}',
                   'backbone': a.backbone, 'note': 'reference arithmetic (pare/smplx) is not installable offline; this is its CPU restatement (oracle/)'},
        'cpu_baseline': {'value': ips, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc},
        'e2e': {'value': ips, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    })


# =============================================================================================== our arm
def run_ours(a):
    import torch.distributed as dist
    import spec_b200 as sb
    from spec_b200.synthetic import synthetic_batch, randomize_module_
    from spec_b200.constants import RECORD_FLOATS

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != a.gpus:
        raise SystemExit(f'--gpus {a.gpus} but WORLD_SIZE={world}: launch N>1 with torch.distributed.run')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = sb.bind_process_to_gpu_numa(local)                      # pinned staging buffers on the GPU's NUMA node
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    B = a.batch                                                   # per-GPU batch (weak scaling)
    cc = sb.CameraRegressorNetwork('resnet50')
    randomize_module_(cc.backbone, 1)
    hmr = sb.HMR(a.backbone, use_cam=True, use_cam_feats=True)
    randomize_module_(hmr.backbone, 0)
    for m in (cc, hmr):
        m.backbone.set_precision(a.precision)
        m.backbone.chunk = a.chunk
        m.to(dev)
    pipe = sb.SPECPipeline(cc, hmr, use_graph=not a.no_graph)
    b = synthetic_batch(B, seed=rank, device=dev)                 # each rank generates its own shard
    args = (b['images'], b['bbox_scale'], b['bbox_center'], b['img_w'], b['img_h'])

    gatherer = sb.RecordGatherer(B, dev) if world > 1 else None

    def step():
        rec = pipe.forward_packed(*args)
        if world > 1:
            return gatherer.submit(rec)                           # the ONE collective of the data path; overlaps the next step
        return rec

    def timed(fn, steps, warmup, mid=None):
        t_w = time.perf_counter()
        n_w = 0
        while n_w < warmup or time.perf_counter() - t_w < 0.4:     # >= W steps AND ~0.4 s so the clocks have ramped
            fn()
            n_w += 1
            if n_w % 4 == 0:
                torch.cuda.synchronize(dev)
        # events are created (and each recorded once, which is where CUDA really allocates them) BEFORE the bracket, so that
        # nothing but the record of e0 sits between the synchronize and the first timed step
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]     # per-step spread (diagnostic only)
        for ev in [e0, e1] + marks:
            ev.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0.record()
        for i in range(steps):
            fn()
            marks[i].record()
            if mid is not None and i % 64 == min(steps // 2, 32):
                mid()                                              # clocks / throttle reasons, inside the timed region
        if gatherer is not None:
            gatherer.flush()                                       # the last gather completes inside the timed region
        e1.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        raw = [a_.elapsed_time(b_) for a_, b_ in zip([e0] + marks[:-1], marks)]
        per = sorted(raw)
        timed.step_ms = {'min': per[0], 'median': per[len(per) // 2], 'max': per[-1], 'argmax': raw.index(per[-1])}
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)             # max over ranks
        return ms.item()

    # ---- value: inputs resident in HBM
    if a.profile_range:
        for _ in range(a.warmup):
            step()
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        return None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms = timed(step, a.steps, a.warmup, mid=sampler.sample if rank == 0 else None)
    step_spread = dict(timed.step_ms)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / a.steps
    value = world * B * a.steps / (total_ms * 1e-3)
    launches_per_step = pipe.launches_per_step()

    # ---- e2e: public API with HOST (pinned) inputs; every step copies its inputs H2D and its results D2H inside the
    # timed region.  The copies ride on side streams (double-buffered staging) so they overlap the previous /
    # next step's kernels -- what a serving loop built on the public API does.
    host = {k: v.cpu().pin_memory() for k, v in b.items()}
    keys = ('images', 'bbox_scale', 'bbox_center', 'img_w', 'img_h')
    staging = [{k: torch.empty_like(host[k], device=dev) for k in keys} for _ in range(2)]
    rec_dev = [torch.empty(B, RECORD_FLOATS, dtype=torch.float32, device=dev) for _ in range(2)]
    host_rec = [torch.empty(B, RECORD_FLOATS, dtype=torch.float32).pin_memory() for _ in range(2)]
    s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)

    def e2e_run(steps):
        ev_h2d = [torch.cuda.Event() for _ in range(2)]
        ev_used = [None, None]
        ev_d2h = [None, None]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def issue_h2d(i):
            sl = i % 2
            with torch.cuda.stream(s_h2d):
                if ev_used[sl] is not None:
                    s_h2d.wait_event(ev_used[sl])                 # staging slot consumed by step i-2
                for k in keys:
                    staging[sl][k].copy_(host[k], non_blocking=True)
                ev_h2d[sl].record(s_h2d)

        with torch.cuda.stream(s_h2d):
            e0.record(s_h2d)
        issue_h2d(0)
        for i in range(steps):
            sl = i % 2
            if i + 1 < steps:
                issue_h2d(i + 1)
            main.wait_event(ev_h2d[sl])
            if ev_d2h[sl] is not None:
                main.wait_event(ev_d2h[sl])                       # rec_dev[sl] read back by step i-2
            st = staging[sl]
            rec = pipe.forward_packed(st['images'], st['bbox_scale'], st['bbox_center'], st['img_w'], st['img_h'])
            if world > 1:
                rec = sb.all_gather_records(rec)[rank * B:(rank + 1) * B]
            rec_dev[sl].copy_(rec, non_blocking=True)
            ev_used[sl] = torch.cuda.Event()
            ev_used[sl].record(main)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(ev_used[sl])
                host_rec[sl].copy_(rec_dev[sl], non_blocking=True)
                ev_d2h[sl] = torch.cuda.Event()
                ev_d2h[sl].record(s_d2h)
        with torch.cuda.stream(s_d2h):
            e1.record(s_d2h)
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1)

    e2e_steps = max(4, min(a.steps, 20))
    e2e_run(8)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e2e_ms_t = torch.tensor([e2e_run(e2e_steps)], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(e2e_ms_t, op=dist.ReduceOp.MAX)
    e2e_ms = e2e_ms_t.item()
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = host_rec[0].numel() * 4

    out = None
    if rank == 0:
        # ---- roofline: per-op CUDA-event timing of both trunks (live, eager, on the launching stream)
        peaks = load_peaks()
        conv_ms, conv_flops, step_ops_ms, conv_bytes = 0.0, 0.0, 0.0, 0
        for m in (cc, hmr):
            m.backbone.profile_ops(b['images'])                   # warm
            rows = m.backbone.profile_ops(b['images'])
            conv_ms += sum(r['ms'] for r in rows if r['flops'] > 0)
            conv_flops += sum(r['flops'] for r in rows)
            conv_bytes += sum(r.get('bytes', 0) for r in rows if r['flops'] > 0)
            step_ops_ms += sum(r['ms'] for r in rows)
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12
        n_conv = sum(1 for m in (cc, hmr) for o in m.backbone._program.ops if o['type'] == 1)
        # ---- cpu baseline (bounded sample, rank 0, N=1 only)
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            if numa.get('bound'):                                  # the CPU baseline may use every host core again:
                for tid in os.listdir('/proc/self/task'):          # reset every thread (OpenMP workers inherited the mask)
                    try:
                        os.sched_setaffinity(int(tid), numa['previous'])
                    except OSError:
                        pass
            ips, _, cores = time_oracle(a.backbone, a.cpu_sample, 2, 1)
            cpu = {'value': ips, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                   'sample': f'{a.cpu_sample}-image batches x2 of the same workload, fp32 PyTorch oracle (oracle/), {cores} threads'}
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms_per_step, 'step_ms_spread': step_spread, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': a.precision, 'data': 'synthetic',
            'config': {'workload': f'SPEC full forward (CamCalib resnet50 -> (R,K) -> HMR {a.backbone} -> SMPL -> projection), 224x224, batch {B}/GPU, random weights',
                       'backbone': a.backbone, 'batch_per_gpu': B, 'global_batch': B * world, 'cuda_graph': not a.no_graph,
                       'numa': {k: v for k, v in numa.items() if k != 'previous'},
                       'l2': 'inputs 154 MB/step/GPU > 126 MB L2, no explicit flush (weights stay L2-resident as in steady-state serving)',
                       'parallelism': f'dp{world} (batch shard + one all-gather of 85,176 B/image records)' if world > 1 else 'single GPU'},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'ms_per_step': e2e_ms / e2e_steps,
                    'steps': e2e_steps, 'how': 'pinned host inputs -> H2D -> SPECPipeline.forward_packed -> D2H of the packed records; copies on side streams, double-buffered'},
            'gpu_launches': int(launches_per_step * a.steps),
            'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peaks['tf_sustained'], 'unit': 'TFLOP/s',
                         'frac': achieved / peaks['tf_sustained'],
                         # dram__bytes_read.sum + dram__bytes_write.sum over the conv launches of one step (2 trunks), from the
                         # committed ncu --set full capture profiles/ncu_r01d.md (14.75 GB per trunk; resnet50, B=256, bf16)
                         'traffic': 29.5e9 if (a.backbone == 'resnet50' and B == 256 and a.precision == 'bf16') else None,
                         'algorithmic_bytes': conv_bytes,
                         'kernel': 'tcgen05 implicit-GEMM conv family: conv_tcp/conv_tcp2 (cta_group::2)/conv_tc/conv3x3_halo/conv_stem7, %d launches/step' % n_conv,
                         'how': 'sum of conv FLOPs of both trunks / sum of per-launch CUDA-event times (specb200_trunk_profile, eager, same stream)',
                         'peak_source': peaks['source'] + ' bf16 sustained', 'conv_ms_per_step': conv_ms,
                         'conv_share_of_trunk_ops': conv_ms / step_ops_ms if step_ops_ms else None,
                         'step_frac_of_peak': (world * 0 + conv_flops) / (ms_per_step * 1e-3) / 1e12 / peaks['tf_sustained']},
            'cpu_baseline': cpu,
        }
        _emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


_JSON_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout when the
    first communicator is created), so everything that goes to fd 1 from here on is sent to stderr and the JSON line is
    written to the saved descriptor by ``_emit``."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(obj):
    line = (json.dumps(obj) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, line)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=256, help='images per GPU per step')
    ap.add_argument('--backbone', default='resnet50')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp16', 'fp32'])
    ap.add_argument('--chunk', type=int, default=int(os.environ.get('SPECB200_CHUNK', '0')))
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile-range', action='store_true', help='only run the steps inside cudaProfilerStart/Stop (for ncu --profile-from-start off)')
    ap.add_argument('--cpu-sample', type=int, default=16, help='images per CPU-oracle step (bounded sample)')
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)


if __name__ == '__main__':
    main()
