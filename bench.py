#!/usr/bin/env python
"""Benchmark of the SELD feature-extraction hot path on B200 (metric of BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode foa|mic] [--clips C]

One "step" = the reference's whole `__main__` (feature_extractor.py:294-307) for a dev-set-shaped shard that is
already resident in HBM: fused extract (a1-a6) -> per-bin statistics (a7) -> all-reduce (N > 1) -> top_db clamp +
normalise (a8).  N = 1 workload = BASELINE.json configs[1]: 600 synthetic 60 s 4-channel 24 kHz FOA clips
-> [600, 3000, 64, 7].  For N > 1 every rank owns its own 600-clip shard (weak scaling; clips are independent,
the only collective is the <= 10.2 KB statistics all-reduce).

Prints ONE JSON line (rank 0).  `value` = audio-hours/s with inputs resident in HBM; `e2e` = the same through
HostDatasetExtractor with pinned HOST buffers (H2D of every clip and D2H of every feature inside the timed region);
`roofline` = the extract kernel's algorithmic bytes / its CUDA-event time against the measured HBM copy peak;
`cpu_baseline` = the oracle port of the reference's CPU path timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

PROD = dict(win_length=960, hop_length=480, n_fft=1024)
SR, L, T_OUT, N_MELS = 24000, 1_440_000, 3000, 64
CLIP_HOURS = 60.0 / 3600.0
ALG_BYTES = {'foa': 4 * 4 * L + 4 * T_OUT * N_MELS * 7, 'mic': 4 * 4 * L + 4 * T_OUT * N_MELS * 10}
METRIC, UNIT = 'feature_audio_hours_per_sec', 'audio-hours/s'


def measured_peak_gbs():
    try:
        with open(os.path.join(REPO, 'MEASURED_PEAKS.json')) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """SM clock / throttle-reason samples during the timed region, read through NVML in a background thread
    (the fields of the B200_PROFILING.md recipe).  NVML is used instead of a looping `nvidia-smi` process because
    the latter's driver queries measurably stall kernel launches (sporadic 2-10x slow steps in per-launch timings)."""
    REASONS = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40}

    def __init__(self, index, period_s=0.05):
        self.index, self.period, self.samples, self.err = index, period_s, [], None
        self._stop = threading.Event()
        self.thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(visible.split(',')[self.index]) if visible and visible.replace(',', '').isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as exc:           # noqa: BLE001
            self.err = str(exc)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:          # noqa: BLE001  older binding name
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((mhz, reasons))
            except Exception as exc:       # noqa: BLE001
                self.err = str(exc)
                return
            self._stop.wait(self.period)

    def stop(self):
        if self.thread is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [f'nvml unavailable: {self.err}']}
        self._stop.set()
        self.thread.join(timeout=2)
        sm = sorted(m for m, _ in self.samples)
        mask = 0
        for _, r in self.samples:
            mask |= r
        reasons = sorted(k for k, bit in self.REASONS.items() if mask & bit)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': self.max_mhz, 'reasons': reasons,
                'samples': len(sm), 'how': 'NVML, 50 ms period, during the timed region'}


def cpu_port_sample(mode, n_clips, threads=None):
    """Oracle port of the reference CPU path on `n_clips` full-size clips: extract + pad/cut + mean/std + normalise.
    Returns (audio-hours/s, seconds, threads)."""
    import numpy as np
    import torch
    from oracle import extractor as O
    from seld_b200.synth import make_clip
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    distinct = [make_clip(1000 + i) for i in range(min(4, n_clips))]
    O.extract_features_port(distinct[0][:, :48000], SR, mode=mode, **PROD)          # warm the thread pool
    t0 = time.perf_counter()
    feats = []
    for i in range(n_clips):
        f = O.extract_features_port(distinct[i % len(distinct)], SR, mode=mode, **PROD)
        feats.append(O.preprocess_features_port(f))
    allf = np.concatenate(feats, 0)
    mean, std = O.statistics_port(allf)
    for f in feats:
        O.normalize_port(f, mean, std)
    dt = time.perf_counter() - t0
    return n_clips * CLIP_HOURS / dt, dt, threads


def run_reference(args, rank):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port: the same torch / numpy library
    calls the reference makes; /root/reference itself is Python and does not exist on the GPU box)."""
    if rank != 0:
        return
    sample = max(1, args.ref_clips)
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, threads = cpu_port_sample(args.mode, sample)
        if i >= args.warmup:
            vals.append((v, dt))
    value = sample * CLIP_HOURS * len(vals) / sum(dt for _, dt in vals)
    ms = 1000.0 * sum(dt for _, dt in vals) / len(vals)
    desc = f'{sample} full-size clips per step (60 s x 4 ch x 24 kHz each), extract + mean/std + normalise, {threads} torch threads'
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args, args.clips),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': desc},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def workload_config(args, clips):
    return {'workload': f'{clips} synthetic 60 s 4-ch 24 kHz {args.mode.upper()} clips per GPU -> '
                        f'[{clips},3000,64,{7 if args.mode == "foa" else 10}] float32: fused extract + per-bin mean/std '
                        f'(all-reduce when N>1) + top_db clamp + normalise (BASELINE.json configs[1]/[3] shape)',
            'clips_per_gpu': clips, 'mode': args.mode, 'layout': args.layout, 'n_fft': 1024, 'win_length': 960, 'hop_length': 480, 'n_mels': 64,
            'l2': 'inputs (23 MB/clip, 13.8 GB/shard) and outputs (3.2 GB) are far larger than the 126 MB L2; no flush needed',
            'parallelism': f'clip-sharded x{args.gpus}'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mode', default='foa', choices=['foa', 'mic'])
    ap.add_argument('--clips', type=int, default=600, help='clips per GPU (dev-set shape: 600)')
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--e2e-clips', type=int, default=None)
    ap.add_argument('--cpu-clips', type=int, default=256, help='bounded CPU-baseline sample (full-size clips)')
    ap.add_argument('--ref-clips', type=int, default=8, help='clips per step of --impl reference')
    ap.add_argument('--layout', default='planar', choices=['planar', 'interleaved'],
                    help="planar [n,4,L] is the reference's (torchaudio.load) layout")
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from seld_b200 import pipeline
    from seld_b200.synth import make_clip

    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    n, mode = args.clips, args.mode
    n_ch = 7 if mode == 'foa' else 10
    # synthetic shard, generated on the device: 16 distinct seeded clips tiled with per-clip gains (all values stay
    # in (-1, 1)); resident in HBM before any timing starts
    wav = torch.empty(n, 4, L, dtype=torch.float32, device=dev)
    base = [make_clip(1000 + 100 * rank + i, device=dev) for i in range(min(16, n))]
    for i in range(n):
        wav[i] = base[i % len(base)] * (0.5 + 0.5 * ((i * 37) % 101) / 101.0)
    del base
    feat = torch.empty(n, T_OUT, N_MELS, n_ch, dtype=torch.float32, device=dev)
    t_raw = 1 + L // PROD['hop_length']
    wav_k = wav
    if args.layout == 'interleaved':
        wav_k = wav.transpose(1, 2).contiguous()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]

    def step(i=None):
        e = ev[i] if i is not None else None
        if e: e[0].record()
        f, key = pipeline.extract_batch(wav_k, SR, mode=mode, n_mels=N_MELS, t_out=T_OUT, out=feat, layout=args.layout, **PROD)
        if e: e[1].record()
        acc = pipeline.partial_statistics(f, key, t_raw)
        pipeline.allreduce_statistics(acc)
        mean, std = pipeline.finish_statistics(acc, N_MELS, n_ch)
        if e: e[2].record()
        pipeline.finalize_(f, key, t_raw, mean, std)
        if e: e[3].record()
        return mean, std

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_beg.record()
    for i in range(args.steps):
        step(i)
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_beg.elapsed_time(t_end)
    stage_ms = [sum(e[j].elapsed_time(e[j + 1]) for e in ev) / args.steps for j in range(3)]
    times = torch.tensor([total_ms] + stage_ms, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, ext_ms, stats_ms, fin_ms = times.tolist()
    ms_per_step = total_ms / args.steps
    value = world * n * CLIP_HOURS / (ms_per_step / 1000.0)

    peak, peak_src = measured_peak_gbs()
    achieved = n * ALG_BYTES[mode] / (ext_ms / 1000.0) / 1e9
    roofline = {'bound': 'hbm', 'kernel': f'seld::extract_kernel<32,{0 if mode == "foa" else 1}>', 'achieved': achieved,
                'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
                'alg_bytes_per_launch': n * ALG_BYTES[mode], 'kernel_ms': ext_ms,
                'stage_ms': {'extract': ext_ms, 'stats+allreduce': stats_ms, 'clamp+normalise': fin_ms}}
    traffic_file = os.path.join(REPO, 'profiles', f'traffic_{mode}.json')
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as fh:
                roofline['traffic'] = json.load(fh).get('dram_bytes_per_launch')
        except Exception:
            pass

    # ---- end to end with host buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        ne = args.e2e_clips or (n if world == 1 else min(n, 150))
        try:
            host_out = torch.empty(ne, T_OUT, N_MELS, n_ch, dtype=torch.float32, pin_memory=True)

            def run_e2e(host_in, layout, dtype, what):
                ex = pipeline.HostDatasetExtractor(ne, L, SR, mode=mode, n_mels=N_MELS, t_out=T_OUT, chunk_clips=24,
                                                   layout=layout, dtype=dtype, **PROD)
                ex.run(host_in, host_out)                                 # warm-up
                barrier()
                t0 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    ex.run(host_in, host_out)
                barrier()
                dt = torch.tensor([(time.perf_counter() - t0) / args.e2e_steps], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                return {'value': world * ne * CLIP_HOURS / dt.item(), 'unit': UNIT, 'h2d_bytes_per_step': ex.h2d_bytes,
                        'd2h_bytes_per_step': ex.d2h_bytes, 'clips_per_gpu': ne, 'ms_per_step': 1000.0 * dt.item(),
                        'host_input': what, 'api': 'seld_b200.pipeline.HostDatasetExtractor.run(pinned wav, pinned out)',
                        'checksum': float(host_out[0, :8].double().sum())}

            # (1) the reference's call surface: decoded float32 [clip, 4, L] tensors (what torchaudio.load returns)
            host_f32 = torch.empty(ne, 4, L, dtype=torch.float32, pin_memory=True)
            host_f32.copy_(wav[:ne])
            # (2) the same audio as it sits in the WAV files: 16-bit PCM frames [clip, L, 4]; decoded on the GPU
            host_pcm = torch.empty(ne, L, 4, dtype=torch.int16, pin_memory=True)
            for c0 in range(0, ne, 50):
                c1 = min(ne, c0 + 50)
                host_pcm[c0:c1].copy_(torch.clamp(torch.round(wav[c0:c1].transpose(1, 2) * 32768.0), -32768, 32767).to(torch.int16))
            del wav, feat, wav_k
            torch.cuda.empty_cache()
            e2e = run_e2e(host_f32, 'planar', torch.float32, 'float32 [clip,4,L] (torchaudio.load layout)')
            del host_f32
            e2e['pcm16'] = run_e2e(host_pcm, 'interleaved', torch.int16, 'int16 PCM [clip,L,4] (WAV frame order), decoded on the GPU')
        except RuntimeError as exc:                                       # e.g. not enough pinnable host memory
            e2e = {'value': None, 'unit': UNIT, 'error': str(exc).splitlines()[0][:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, dt, threads = cpu_port_sample(mode, args.cpu_clips)
        cpu = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': f'{args.cpu_clips} full-size clips, extract + mean/std + normalise, {dt:.1f} s of CPU work '
                         f'({os.cpu_count()} host cores)'}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
                'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic', 'config': workload_config(args, n), 'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e,
                'gpu_launches': (7 if mode == 'mic' else 6) * args.steps, 'clocks': clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
