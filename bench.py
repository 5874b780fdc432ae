#!/usr/bin/env python
"""Benchmark of the SELD feature-extraction hot path on B200 (metric of BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--clips C]

One "step" = the reference's whole `__main__` (feature_extractor.py:294-307) for a dev-set-shaped set of clips that is
already resident in HBM: fused extract (a1-a6) -> per-bin statistics (a7) -> all-reduce (N > 1) -> top_db clamp +
normalise (a8).  The line carries BASELINE.json's configs:

  configs[1]  `value`, `roofline`, `e2e`: 600 synthetic 60 s 4-channel 24 kHz FOA clips -> [600, 3000, 64, 7]
  configs[2]  `mic`: the same for MIC (log-mel + 6-pair GCC-PHAT, [600, 3000, 64, 10]); `foa_plus_mic`: both, 35.5 GB
  configs[3]  N > 1 (torchrun): the SAME 600 clips sharded clip i -> rank i mod N (strong scaling; `scaling: "strong"`),
              per-bin {sum, sum of squares, count} all-reduced over NCCL; `stats_check` compares the all-reduced mean / std
              with single-GPU statistics of all 600 clips; `weak` repeats the measurement with 600 clips PER GPU
  configs[4]  `config5`: masking of a 256 x [300, 64, 7] batch (train.py and trainv2.py parameters), the fused
              augment + mask launch, and the on-the-fly batch (256 wav chunks -> extract -> normalise -> mask)

`value` = audio-hours/s of the whole job with inputs resident in HBM, the step replayed as ONE CUDA graph (`graph`);
`roofline` = the extract kernel's algorithmic bytes / its CUDA-event time against the measured HBM copy peak; `e2e` =
the same metric through pipeline.HostDatasetExtractor with pinned HOST buffers (H2D of every clip, D2H of every feature
inside the timed region, `--e2e-steps` datasets pipelined); `e2e.copy_ceiling` = the same upload + download bytes as plain
concurrent pinned copies on every rank at once, no kernels -- what the host side of the box allows, measured in the same run;
`cpu_baseline` = the oracle port of the reference's CPU path on this box's host cores.
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
T_RAW = 1 + L // PROD['hop_length']
CLIP_HOURS = 60.0 / 3600.0
ALG_BYTES = {'foa': 4 * 4 * L + 4 * T_OUT * N_MELS * 7, 'mic': 4 * 4 * L + 4 * T_OUT * N_MELS * 10}
N_CH = {'foa': 7, 'mic': 10}
SEED_BASE = {'foa': 1000, 'mic': 2000}
METRIC, UNIT = 'feature_audio_hours_per_sec', 'audio-hours/s'


def measured_peak_gbs():
    try:
        with open(os.path.join(REPO, 'MEASURED_PEAKS.json')) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def cpu_model():
    try:
        with open('/proc/cpuinfo') as fh:
            for line in fh:
                if line.startswith('model name'):
                    return line.split(':', 1)[1].strip()
    except Exception:
        pass
    return 'unknown'


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
                'samples': len(sm), 'how': 'NVML, 50 ms period, during the timed regions'}


# ----------------------------------------------------------------------------------------------------------------- CPU arm
def cpu_port_sample(mode, n_clips, threads=None):
    """Oracle port of the reference CPU path on `n_clips` full-size clips: extract + pad/cut + mean/std + normalise.
    Returns (audio-hours/s, seconds, threads)."""
    import numpy as np
    import torch
    from oracle import extractor as O
    from seld_b200.synth import make_clip
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    distinct = [make_clip(SEED_BASE[mode] + i) for i in range(min(4, n_clips))]
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
        v, dt, threads = cpu_port_sample('foa', sample)
        if i >= args.warmup:
            vals.append((v, dt))
    value = sample * CLIP_HOURS * len(vals) / sum(dt for _, dt in vals)
    ms = 1000.0 * sum(dt for _, dt in vals) / len(vals)
    desc = (f'{sample} full-size FOA clips per step (60 s x 4 ch x 24 kHz each), extract + mean/std + normalise, '
            f'{threads} torch threads on {cpu_model()} ({os.cpu_count()} logical cores); per-clip rate, the reference loop is serial per file')
    mic_v, mic_dt, _ = cpu_port_sample('mic', max(1, sample // 2))
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args, 'foa', args.clips, args.gpus),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': desc, 'cpu_model': cpu_model()},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'mic': {'value': mic_v, 'unit': UNIT, 'sample': f'{max(1, sample // 2)} full-size MIC clips, {mic_dt:.1f} s'},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def workload_config(args, mode, total_clips, world):
    return {'workload': f'{total_clips} synthetic 60 s 4-ch 24 kHz {mode.upper()} clips (DCASE2021 dev-set shape) -> '
                        f'[{total_clips},3000,64,{N_CH[mode]}] float32: fused extract + per-bin mean/std (all-reduce when N>1) + '
                        f'top_db clamp + normalise (BASELINE.json configs[1]; configs[3] sharding at N>1)',
            'total_clips': total_clips, 'clips_per_gpu': -(-total_clips // world), 'mode': mode, 'layout': args.layout,
            'n_fft': 1024, 'win_length': 960, 'hop_length': 480, 'n_mels': 64,
            'l2': 'inputs (23 MB/clip; 1.7 GB per GPU even at 75 clips) and outputs are far larger than the 126 MB L2; no flush needed',
            'parallelism': f'clip i -> rank i mod {world}'}


# ----------------------------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--clips', type=int, default=600, help='clips of the whole job (dev-set shape: 600)')
    ap.add_argument('--e2e-steps', type=int, default=8, help='datasets pipelined through the host extractor in the e2e timed region')
    ap.add_argument('--cpu-clips', type=int, default=192, help='bounded CPU-baseline sample (full-size clips)')
    ap.add_argument('--ref-clips', type=int, default=32, help='clips per step of --impl reference')
    ap.add_argument('--layout', default='planar', choices=['planar', 'interleaved'],
                    help="planar [n,4,L] is the reference's (torchaudio.load) layout")
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of one CUDA graph per step')
    ap.add_argument('--no-mic', action='store_true')
    ap.add_argument('--no-weak', action='store_true')
    ap.add_argument('--no-peer', action='store_true', help='skip the peer-memory all-reduce variant (N > 1)')
    ap.add_argument('--no-config5', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from seld_b200 import _lib, pipeline, sharding
    from seld_b200.synth import make_clip

    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def make_shard(mode, indices, planar=False):
        """Clip i of the job = one of 16 seeded base clips times a per-clip gain (values stay in (-1, 1)): every rank can
        build any clip, so the sharded run and the single-GPU check see the same 600 clips."""
        base = [make_clip(SEED_BASE[mode] + i, device=dev) for i in range(16)]
        wav = torch.empty(len(indices), 4, L, dtype=torch.float32, device=dev)
        for j, i in enumerate(indices):
            wav[j] = base[i % 16] * (0.5 + 0.5 * ((i * 37) % 101) / 101.0)
        del base
        return wav.transpose(1, 2).contiguous() if (args.layout == 'interleaved' and not planar) else wav

    peak, peak_src = measured_peak_gbs()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    def measure(mode, indices, steps, peer=False):
        """Device-resident job over this rank's clips -> dict (times are max over ranks).  peer: the statistics all-reduce as this
        package's one-launch exchange over NVLink peer memory instead of NCCL."""
        wav = make_shard(mode, indices)
        step = pipeline.DatasetStep(wav, SR, mode=mode, n_mels=N_MELS, t_out=T_OUT, layout=args.layout, peer_allreduce=peer, **PROD)
        for _ in range(args.warmup):
            step.run()
        barrier()
        # (1) eager launches with per-stage events: the roofline of the extract kernel
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
        barrier()
        l0 = lib.seld_launch_count()
        for i in range(steps):
            step.run(ev[i])
        barrier()
        launches_per_step = (lib.seld_launch_count() - l0) / steps
        stage_ms = [sum(e[j].elapsed_time(e[j + 1]) for e in ev) / steps for j in range(3)]
        eager_ms = sum(e[0].elapsed_time(e[3]) for e in ev) / steps
        # (2) the step as one CUDA graph (NCCL all-reduce included)
        graph_ms, graph_err = None, None
        if not args.no_graph:
            try:
                step.capture()
                for _ in range(args.warmup):
                    step.replay()
                barrier()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                for _ in range(steps):
                    step.replay()
                t1.record()
                barrier()
                graph_ms = t0.elapsed_time(t1) / steps
            except Exception as exc:           # noqa: BLE001
                graph_err = str(exc).splitlines()[0][:200]
        # (3) eager, whole loop under one event pair (what round 1 reported)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0.record()
        for _ in range(steps):
            step.run()
        t1.record()
        barrier()
        loop_ms = t0.elapsed_time(t1) / steps
        vals = max_over_ranks([loop_ms, eager_ms] + stage_ms + [graph_ms if graph_ms is not None else -1.0])
        loop_ms, eager_ms, ext_ms, stats_ms, fin_ms, g_ms = vals
        res = {'ms_eager': loop_ms, 'ms_graph': g_ms if g_ms > 0 else None, 'graph_error': graph_err,
               'stage_ms': {'extract': ext_ms, 'stats+allreduce': stats_ms, 'clamp+normalise': fin_ms},
               'launches_per_step': launches_per_step, 'n_local': len(indices)}
        mean, std = step.mean.clone(), step.std.clone()
        del step, wav
        torch.cuda.empty_cache()
        return res, mean, std

    def roofline_of(mode, n_local, ext_ms):
        achieved = n_local * ALG_BYTES[mode] / (ext_ms / 1000.0) / 1e9
        r = {'bound': 'hbm', 'kernel': f'seld::extract_kernel<32,{0 if mode == "foa" else 1},...> (interior + edge launch)',
             'achieved': achieved, 'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s', 'frac': achieved / peak,
             'traffic': None, 'traffic_source': None, 'alg_bytes_per_launch': n_local * ALG_BYTES[mode], 'kernel_ms': ext_ms}
        tf = os.path.join(REPO, 'profiles', f'traffic_{mode}.json')
        if os.path.exists(tf):
            try:
                with open(tf) as fh:
                    t = json.load(fh)
                # static: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on
                # 600 clips (profiles/), scaled to this launch's clip count -- not measured in this run
                r['traffic'] = t.get('dram_bytes_per_launch') * n_local / float(t.get('clips', 600))
                r['traffic_source'] = 'static: ' + t.get('source', f'profiles/traffic_{mode}.json')
            except Exception:
                pass
        return r

    total = args.clips
    out = {}
    modes = ['foa'] + ([] if args.no_mic else ['mic'])
    stats_check = {}
    for mode in modes:
        mine = sharding.shard_indices(total, rank, world)
        res, mean, std = measure(mode, list(mine), args.steps)
        res['allreduce'] = {'used': 'nccl', 'nccl_ms_per_step': res['ms_graph'] or res['ms_eager']}
        if world > 1 and not args.no_peer:
            # the same job with the statistics all-reduce as ONE kernel over NVLink peer memory; the faster one is the headline
            try:
                res_p, mean_p, std_p = measure(mode, list(mine), args.steps, peer=True)
                ms_p = res_p['ms_graph'] or res_p['ms_eager']
                same = bool(torch.equal(mean_p, mean) and torch.equal(std_p, std))
                info = {'nccl_ms_per_step': res['allreduce']['nccl_ms_per_step'], 'peer_ms_per_step': ms_p,
                        'nccl_stats_stage_ms': res['stage_ms']['stats+allreduce'], 'peer_stats_stage_ms': res_p['stage_ms']['stats+allreduce'],
                        'peer_equals_nccl_bitwise': same}
                if ms_p < info['nccl_ms_per_step']:
                    res, mean, std = res_p, mean_p, std_p
                    info['used'] = 'peer (seld_stats_peer_allreduce: one launch over NVLink peer memory)'
                else:
                    info['used'] = 'nccl'
                res['allreduce'] = info
            except Exception as exc:           # noqa: BLE001  (no symmetric memory on this box, ...)
                res['allreduce']['peer_error'] = str(exc).splitlines()[0][:200]
        ms = res['ms_graph'] or res['ms_eager']
        res['value'] = total * CLIP_HOURS / (ms / 1000.0)
        res['ms_per_step'] = ms
        res['roofline'] = roofline_of(mode, res['n_local'], res['stage_ms']['extract'])
        out[mode] = res
        if world > 1:
            # the all-reduced statistics against single-GPU statistics of the same 600 clips (rank 0 extracts them all)
            if rank == 0:
                wav_all = make_shard(mode, list(range(total)))
                feat, key = pipeline.extract_batch(wav_all, SR, mode=mode, n_mels=N_MELS, t_out=T_OUT, layout=args.layout, **PROD)
                acc = pipeline.partial_statistics(feat, key, T_RAW)
                m1, s1 = pipeline.finish_statistics(acc, N_MELS, N_CH[mode])
                rel_m = float(((mean - m1).abs() / (m1.abs() + 1e-3)).max())
                rel_s = float(((std - s1).abs() / s1.abs().clamp_min(1e-6)).max())
                stats_check[mode] = {'max_rel_err_mean': rel_m, 'max_rel_err_std': rel_s, 'tolerance': 1e-5,
                                     'ok': bool(rel_m <= 1e-5 and rel_s <= 1e-5),
                                     'what': f'all-reduced mean/std over {world} ranks vs one GPU extracting all {total} clips'}
                del wav_all, feat, key
                torch.cuda.empty_cache()
            barrier()

    # ---- weak scaling as a secondary number (N > 1): every rank owns `total` clips of its own
    weak = None
    if world > 1 and not args.no_weak:
        res, _, _ = measure('foa', [rank * total + i for i in range(total)], max(5, args.steps // 2))
        ms = res['ms_graph'] or res['ms_eager']
        weak = {'value': world * total * CLIP_HOURS / (ms / 1000.0), 'unit': UNIT, 'ms_per_step': ms, 'clips_per_gpu': total,
                'scaling': 'weak', 'stage_ms': res['stage_ms']}

    # ---- BASELINE.json configs[4]: training-time form (rank 0; replicas only, nothing to shard)
    config5 = None
    if rank == 0 and not args.no_config5:
        config5 = run_config5(torch, dev)
    barrier()

    # ---- end to end with host buffers (pinned), H2D + D2H inside the timed region
    e2e = {}
    if not args.no_e2e:
        mine = list(sharding.shard_indices(total, rank, world))
        ne = len(mine)
        for mode in modes:
            try:
                wav = make_shard(mode, mine, planar=True)
                n_ch = N_CH[mode]
                host_out = [torch.empty(ne, T_OUT, N_MELS, n_ch, dtype=torch.float32, pin_memory=True) for _ in range(2)]

                def run_e2e(host_in, layout, dtype, what):
                    ex = pipeline.HostDatasetExtractor(ne, L, SR, mode=mode, n_mels=N_MELS, t_out=T_OUT, chunk_clips=24,
                                                       layout=layout, dtype=dtype, **PROD)
                    ex.run(host_in, host_out[0])                               # warm-up
                    barrier()
                    t0 = time.perf_counter()
                    ex.run(host_in, host_out[0])
                    torch.cuda.synchronize()
                    single = time.perf_counter() - t0
                    barrier()
                    l0 = lib.seld_launch_count()
                    t0 = time.perf_counter()
                    pending = [ex.submit(host_in, host_out[i & 1]) for i in range(args.e2e_steps)]
                    for _, _, done in pending:
                        done.synchronize()
                    torch.cuda.synchronize()
                    dt, single_max = max_over_ranks([(time.perf_counter() - t0) / args.e2e_steps, single])
                    launches = (lib.seld_launch_count() - l0) / args.e2e_steps
                    return {'value': total * CLIP_HOURS / dt, 'unit': UNIT, 'h2d_bytes_per_step': ex.h2d_bytes * world,
                            'd2h_bytes_per_step': ex.d2h_bytes * world, 'clips_per_gpu': ne, 'total_clips': total,
                            'ms_per_step': 1000.0 * dt, 'ms_single_dataset': 1000.0 * single_max,
                            'value_single_dataset': total * CLIP_HOURS / single_max, 'launches_per_step': launches,
                            'host_input': what, 'steps': args.e2e_steps,
                            'api': 'seld_b200.pipeline.HostDatasetExtractor.submit(pinned wav, pinned out): consecutive datasets in '
                                   'flight, upload of the next one overlapping the download of the previous one (PCIe is full duplex)',
                            'checksum': float(host_out[0][0, :8].double().sum())}

                def copy_ceiling(host_in, host_o, e2e_ms):
                    # what the box's host side allows: the SAME bytes moved by plain pinned copies on two streams (upload and
                    # download concurrently, all ranks at once), no kernels -- the ceiling of any e2e number on this box
                    d_in = torch.empty(host_in.shape, dtype=host_in.dtype, device=dev)
                    d_out = torch.empty(host_o.shape, dtype=host_o.dtype, device=dev)
                    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
                    times = []
                    for rep in range(3):
                        barrier()
                        t0 = time.perf_counter()
                        with torch.cuda.stream(s_up):
                            d_in.copy_(host_in, non_blocking=True)
                        with torch.cuda.stream(s_dn):
                            host_o.copy_(d_out, non_blocking=True)
                        s_up.synchronize()
                        s_dn.synchronize()
                        times.append(time.perf_counter() - t0)
                    dt = max_over_ranks([min(times[1:])])[0]
                    up, dn = host_in.numel() * host_in.element_size() * world, host_o.numel() * host_o.element_size() * world
                    del d_in, d_out
                    return {'ms': 1000.0 * dt, 'h2d_GBps_all_ranks': up / dt / 1e9, 'd2h_GBps_all_ranks': dn / dt / 1e9,
                            'e2e_frac_of_ceiling': 1000.0 * dt / e2e_ms,
                            'what': 'the same upload + download bytes as plain concurrent pinned cudaMemcpyAsync on every rank at '
                                    'once, no kernels: the host-side ceiling of this box for the e2e step'}

                # (1) the audio as it sits in the WAV files the reference reads: 16-bit PCM frames [clip, L, 4], decoded on the
                #     GPU exactly like torchaudio.load (sample / 32768) -- the path's first-class host input
                host_pcm = torch.empty(ne, L, 4, dtype=torch.int16, pin_memory=True)
                for c0 in range(0, ne, 50):
                    c1 = min(ne, c0 + 50)
                    host_pcm[c0:c1].copy_(torch.clamp(torch.round(wav[c0:c1].transpose(1, 2) * 32768.0), -32768, 32767).to(torch.int16))
                host_f32 = None
                if mode == 'foa':
                    # (2) decoded float32 [clip, 4, L] tensors (what torchaudio.load returns): twice the upload bytes
                    host_f32 = torch.empty(ne, 4, L, dtype=torch.float32, pin_memory=True)
                    host_f32.copy_(wav)
                del wav
                torch.cuda.empty_cache()
                r = run_e2e(host_pcm, 'interleaved', torch.int16, 'int16 PCM [clip,L,4] (WAV frame order), decoded on the GPU')
                if mode == 'foa':
                    r['copy_ceiling'] = copy_ceiling(host_pcm, host_out[0], r['ms_per_step'])
                del host_pcm
                if host_f32 is not None:
                    r['float32_input'] = run_e2e(host_f32, 'planar', torch.float32, 'float32 [clip,4,L] (torchaudio.load layout)')
                    del host_f32
                e2e[mode] = r
                del host_out
            except RuntimeError as exc:                                       # e.g. not enough pinnable host memory
                e2e[mode] = {'value': None, 'unit': UNIT, 'error': str(exc).splitlines()[0][:200]}
            torch.cuda.empty_cache()

    clocks = sampler.stop() if rank == 0 else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, dt, threads = cpu_port_sample('foa', args.cpu_clips)
        cpu = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'cpu_model': cpu_model(), 'logical_cores': os.cpu_count(),
               'sample': f'{args.cpu_clips} full-size FOA clips, extract + mean/std + normalise, {dt:.1f} s of CPU work'}
        if not args.no_mic:
            vm, dtm, _ = cpu_port_sample('mic', max(8, args.cpu_clips // 4))
            cpu['mic'] = {'value': vm, 'sample': f'{max(8, args.cpu_clips // 4)} full-size MIC clips, {dtm:.1f} s'}

    if rank == 0:
        foa = out['foa']
        line = {'metric': METRIC, 'value': foa['value'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': foa['ms_per_step'], 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic', 'config': workload_config(args, 'foa', total, world),
                'graph': {'used': foa['ms_graph'] is not None, 'ms_per_step_graph': foa['ms_graph'], 'ms_per_step_eager': foa['ms_eager'],
                          'error': foa['graph_error']},
                'allreduce': foa.get('allreduce'),
                'roofline': dict(foa['roofline'], stage_ms=foa['stage_ms']), 'cpu_baseline': cpu,
                'e2e': e2e.get('foa'), 'gpu_launches': int(round(foa['launches_per_step'] * args.steps)),
                'gpu_launches_per_step': foa['launches_per_step'], 'clocks': clocks}
        if 'mic' in out:
            mic = out['mic']
            line['mic'] = {'value': mic['value'], 'unit': UNIT, 'ms_per_step': mic['ms_per_step'], 'ms_per_step_eager': mic['ms_eager'],
                           'roofline': dict(mic['roofline'], stage_ms=mic['stage_ms']), 'e2e': e2e.get('mic'),
                           'gpu_launches_per_step': mic['launches_per_step'], 'allreduce': mic.get('allreduce'),
                           'config': workload_config(args, 'mic', total, world)['workload']}
            both_ms = foa['ms_per_step'] + mic['ms_per_step']
            alg = total * (ALG_BYTES['foa'] + ALG_BYTES['mic'] + 2 * 4 * T_OUT * N_MELS * 17)
            line['foa_plus_mic'] = {'value': 2 * total * CLIP_HOURS / (both_ms / 1000.0), 'unit': UNIT, 'ms': both_ms,
                                    'clips': 2 * total, 'alg_bytes_with_normalisation': alg,
                                    'frac_of_hbm_roofline': alg / world / (both_ms / 1000.0) / 1e9 / peak}
        if world > 1:
            line['stats_check'] = stats_check
            line['weak'] = weak
            bad = [m for m, c in stats_check.items() if not c['ok']]
            if bad:
                line['error'] = f'all-reduced statistics differ from the single-GPU ones for {bad}'
        if config5 is not None:
            line['config5'] = config5
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_config5(torch, dev):
    """BASELINE.json configs[4] on one GPU: (i) masking only, both parameter sets; (ii) fused augmentation + masks;
    (iii) on-the-fly batches.  Distinct batches cycle through > 126 MB so nothing is served from L2."""
    from seld_b200 import pipeline, transforms as T
    B, FR, M, C = 256, 300, 64, 7
    res = {'batch': f'{B} x [{FR},{M},{C}] float32 (6 s chunks)', 'alg_bytes_full_read_write': 2 * 4 * B * FR * M * C}

    def timed(fn, n):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    nb = 16
    x = torch.randn(nb, B, FR, M, C, device=dev)
    for name, tm, fm in (('mask_train_py_24x1_16x1', (24, 1), (16, 1)), ('mask_trainv2_py_6x10_8x6', (6, 10), (8, 6))):
        ms = timed(lambda i: T.mask_batch_(x[i % nb], tm, fm, seed=2, sample_offset=i * B), nb)
        res[name] = {'ms_per_batch': ms, 'GBps_vs_full_read_write': res['alg_bytes_full_read_write'] / ms / 1e6}
        x.normal_()
    if hasattr(T, 'augment_batch'):
        y = torch.rand(B, 60, 56, device=dev)
        ms = timed(lambda i: T.augment_batch(x[i % nb], y, spatial='foa', level_jitter=0.2, time_mask=(24, 1), freq_mask=(16, 1),
                                             seed=3, sample_offset=i * B), nb)
        res['augment_foa_iv_plus_jitter_plus_masks'] = {'ms_per_batch': ms, 'GBps': res['alg_bytes_full_read_write'] / ms / 1e6,
                                                        'what': 'foa_intensity_vec_aug + random_ups_and_downs + time/freq masks, '
                                                                'on-device draws, one launch over x + one over the labels'}
    del x
    Lc = (FR - 1) * 480 + 1024
    nb = 8
    chunks = (torch.rand(nb, B, 4, Lc, device=dev) - 0.5) * 0.2
    cmax = torch.full((B,), 20.0, device=dev)
    mean, std = torch.zeros(1, M, C, device=dev), torch.ones(1, M, C, device=dev)
    ms = timed(lambda i: pipeline.training_batch(chunks[i % nb], SR, cmax, mean, std, seed=2, sample_offset=i * B, **PROD), nb)
    res['onthefly_extract_normalise_mask'] = {'ms_per_batch': ms, 'batches_per_s': 1000.0 / ms,
                                              'audio_hours_per_s': B * 6.0 / 3600.0 / (ms / 1000.0)}
    return res


if __name__ == '__main__':
    main()
