#!/usr/bin/env python
"""Headline benchmark: valid mel frames/s of batched ``ForwardTacotron.generate`` (BASELINE.json configs[1]:
batch 64 x 200 synthetic phonemes, alpha 1.0, default config.yaml model) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one generate() call over one batch (stage A predictors, callbacks, length plan + its D2H, stage B).
Utterances are independent, so N GPUs run N replicas on different batches with NO data-path collective
(weak scaling); the only collectives are the timing barrier / max / sum.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

B, T = 64, 200                       # BASELINE.json configs[1]
CPU_SAMPLE_B = 64                    # the CPU arm times the whole batch (a few seconds per call on 16 cores)
METRIC, UNIT = 'mel_frames_per_s', 'frames/s'
WORKLOAD = 'ForwardTacotron.generate batch 64 x 200 phonemes, alpha 1.0, config.yaml defaults, synthetic weights'


def peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        d = json.loads(p.read_text())
        return {'hbm': float(d['hbm_gbs']), 'tensor': float(d.get('bf16_tflops_sustained', d['bf16_tflops'])),
                'tensor_burst': float(d['bf16_tflops']), 'src': 'measured'}
    return {'hbm': 6650.0, 'tensor': 1400.0, 'tensor_burst': 1590.0, 'src': 'fallback'}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_generate_rate(steps: int, warmup: int):
    """The reference's algorithm (oracle port, torch fp32 on all host cores; proven equal to the reference's own
    generate() by oracle/make_golden.py) on the SAME batch as the GPU arm: all 64 utterances, a few timed calls."""
    import torch
    from forwardtacotron_b200.utils import synth
    from oracle import model_oracle as mo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, _ = synth.synthetic_model('forward_tacotron')
    sd = model.state_dict()
    x = synth.synthetic_tokens(B, T, seed=1)[:CPU_SAMPLE_B]
    frames, times = 0, []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = mo.ft_generate(sd, x)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            frames = int((out['dur'] + 0.5).long().sum())
    total = sum(times)
    return {'value': frames * len(times) / total, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'all {CPU_SAMPLE_B} of the {B} utterances (T={T}): the whole cfg2 batch, {len(times)} timed '
                      f'generate() calls of oracle/model_oracle.py (torch {torch.__version__} fp32, {cores} threads), '
                      f'{frames} valid frames per call'}, total / len(times) * 1e3


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    cb, ms = cpu_generate_rate(steps, warmup)
    line = {'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': steps, 'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'global_batch': B * args.gpus, 'phonemes': T,
                       'note': 'CPU arm: every step is one generate() over the whole 64 x 200 batch (cpu_baseline.sample); '
                               'steps / warmup are capped at 3 / 1 so the run ends within minutes'},
            'cpu_baseline': cb,
            'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  An NVML thread polls every few ms (the timed
    region of the default run is ~100 ms, far shorter than one `nvidia-smi -lms` period); if NVML is not importable
    the `nvidia-smi` loop is the fallback."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, uuid: str, period_s: float = 0.004):
        import threading
        self.samples, self.reasons, self.sm_max, self.power = [], set(), None, []
        self.proc = self.tmp = self.thread = None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            masks = (('hw_slowdown', pynvml.nvmlClocksEventReasonHwSlowdown),
                     ('hw_thermal_slowdown', pynvml.nvmlClocksEventReasonHwThermalSlowdown),
                     ('sw_thermal_slowdown', pynvml.nvmlClocksEventReasonSwThermalSlowdown),
                     ('sw_power_cap', pynvml.nvmlClocksEventReasonSwPowerCap))

            def poll():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for name, m in masks:
                            if r & m:
                                self.reasons.add(name)
                        self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
                    except Exception:
                        pass
                    self._stop.wait(period_s)

            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.source = 'nvml thread'
        except Exception:
            self.thread = None
            self.source = 'nvidia-smi -lms 20'
            self.tmp = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
            try:
                self.proc = subprocess.Popen(['nvidia-smi', '-i', uuid, f'--query-gpu={self.FIELDS}',
                                              '--format=csv,noheader,nounits', '-lms', '20'], stdout=self.tmp,
                                             stderr=subprocess.DEVNULL)
            except Exception:
                self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': self.sm_max, 'reasons': [], 'samples': 0, 'source': self.source}
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
        elif self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.tmp.flush()
            rows = [r.split(',') for r in Path(self.tmp.name).read_text().strip().splitlines() if r.count(',') >= 6]
            os.unlink(self.tmp.name)
            for r in rows:
                try:
                    self.samples.append(float(r[0]))
                    self.sm_max = float(r[1])
                except ValueError:
                    continue
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.strip().lower().startswith('active'):
                        self.reasons.add(name)
            out['sm_max_mhz'] = self.sm_max
        if self.samples:
            out['sm_mhz'] = statistics.median(self.samples)
            out['samples'] = len(self.samples)
        if self.power:
            out['power_w_max'] = max(self.power)
        out['reasons'] = sorted(self.reasons)
        return out


# ------------------------------------------------------------------------------------------ our arm
def collect_profile(lib, steps):
    import ctypes as C
    n = lib.ftb_profile_families()
    ms, fl, by = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
    ln = (C.c_longlong * n)()
    from forwardtacotron_b200 import _lib
    _lib.check(lib.ftb_profile_collect(ms, fl, by, ln))
    fams = []
    for i in range(n):
        if ln[i]:
            fams.append({'name': lib.ftb_profile_family_name(i).decode(), 'ms_per_step': ms[i] / steps,
                         'launches_per_step': ln[i] / steps, 'flops_per_step': fl[i] / steps,
                         'bytes_per_step': by[i] / steps})
    return sorted(fams, key=lambda f: -f['ms_per_step'])


def roofline_of(fam, pk):
    hbm_bound = fam['name'] in ('length_regulator', 'elementwise', 'stft_mel')
    sec = fam['ms_per_step'] / 1e3
    ncu = {}
    p = ROOT / 'profiles' / 'ncu_traffic.json'  # per-launch DRAM bytes from the committed ncu --set full capture
    if p.exists():
        ncu = json.loads(p.read_text())
    if hbm_bound:
        ach = fam['bytes_per_step'] / sec / 1e9
        return {'kernel': fam['name'], 'bound': 'hbm', 'achieved': ach, 'peak': pk['hbm'], 'unit': 'GB/s',
                'frac': ach / pk['hbm'], 'traffic': ncu.get(fam['name']), 'peak_source': pk['src']}
    ach = fam['flops_per_step'] / sec / 1e12
    out = {'kernel': fam['name'], 'bound': 'tensor', 'achieved': ach, 'peak': pk['tensor'], 'unit': 'TFLOP/s',
           'frac': ach / pk['tensor'], 'traffic': ncu.get(fam['name']), 'peak_source': pk['src'] + ' (sustained bf16)'}
    if fam['name'].startswith('rnn_'):
        out['note'] = ('recurrence: a chain of dependent time steps (T or L per launch), bound by the per-step latency '
                       '(DSMEM hand-off + dependent MMA chain + gate maths, DESIGN.md 4), not by the tensor pipe; '
                       + ('exact-fp32 SIMT kernel, listed against the tensor peak only for scale'
                          if fam['name'] == 'rnn_gru_small' else 'the fraction is reported for completeness'))
    return out


def _one_blas_thread():
    """One BLAS / OpenMP thread per process: the pool already uses every core (16 workers x 16 BLAS threads thrash)."""
    try:
        import threadpoolctl
        _one_blas_thread.keep = threadpoolctl.threadpool_limits(1)
    except Exception:
        os.environ['OMP_NUM_THREADS'] = '1'


def _dsp_cpu_clip(args):
    """Pool worker: the oracle's wav_to_mel over one clip (mirrors the reference's per-file pool, preprocess.py:129-139)."""
    from oracle import dsp_oracle
    return dsp_oracle.wav_to_mel(args).shape[1]


def stft_cpu_baseline(audio_host, offs, budget_s=12.0):
    """The reference's CPU path for wav_to_mel -- librosa's algorithm as restated in oracle/dsp_oracle.py (librosa 0.7.2
    itself is not installable offline) -- over a bounded sample of the same clips: one thread, then a
    multiprocessing.Pool(cpu_count()) like preprocess.py:129.  audio-seconds per second."""
    import multiprocessing as mp
    import numpy as np
    from oracle import dsp_oracle
    cores = os.cpu_count() or 1
    clips = [np.asarray(audio_host[int(offs[i]):int(offs[i + 1])]) for i in range(min(len(offs) - 1, 16 * cores))]
    try:
        import threadpoolctl
        limit = threadpoolctl.threadpool_limits(1)
    except Exception:
        limit = None
    dsp_oracle.wav_to_mel(clips[0])
    t0 = time.perf_counter()
    n1 = 0
    for c in clips:                                   # single thread: as many clips as fit in ~1/3 of the budget
        dsp_oracle.wav_to_mel(c)
        n1 += 1
        if time.perf_counter() - t0 > budget_s / 3:
            break
    t1 = time.perf_counter() - t0
    secs1 = sum(len(c) for c in clips[:n1]) / 22050.0
    if limit is not None:
        limit.restore_original_limits()
    with mp.get_context('fork').Pool(cores, initializer=_one_blas_thread) as pool:
        pool.map(_dsp_cpu_clip, clips[:cores])         # warm the workers
        t0 = time.perf_counter()
        pool.map(_dsp_cpu_clip, clips)
        tp = time.perf_counter() - t0
    secs = sum(len(c) for c in clips) / 22050.0
    return {'value': secs / tp, 'unit': 'audio-s/s', 'cores': cores, 'kind': 'port',
            'single_thread_value': secs1 / t1,
            'sample': f'{len(clips)} of the clips through oracle/dsp_oracle.py (numpy float64 rFFT, Slaney mel) on a '
                      f'multiprocessing.Pool({cores}); single-thread figure over the first {n1} clips',
            'parity': 'unpinned against librosa 0.7.2 (not installable offline); the restatement agrees with '
                      "torchaudio's independent Slaney mel to 5.5e-6 (tests/test_oracle_dsp.py)"}


def stft_extra(torch, dev, pk, n_clips=1250, seed=7, cpu=True):
    """Second headline of BASELINE.json: STFT->log-mel audio-seconds/s (cfg4: 10 000 clips of 2-10 s over 8 GPUs =
    1 250 per GPU; ~165 M samples = 660 MB of fp32 >> 126 MB L2).  Three figures: the kernel alone (offsets planned
    once, CUDA events), the packed call as a user makes it (plans the offsets every call), and end to end from pinned
    HOST audio with the mel read back to the host."""
    from forwardtacotron_b200.utils import synth
    from forwardtacotron_b200.utils.config import default_config
    from forwardtacotron_b200.utils.dsp import DSP
    dsp = DSP.from_config(default_config())
    audio, offs = synth.synthetic_audio(n_clips, seed=seed)
    a_host = audio.pin_memory()
    a = a_host.to(dev)
    plan = dsp.plan_clips(offs, dev)
    out = torch.empty(80 * plan.total_frames, dtype=torch.float32, device=dev)
    for _ in range(3):
        dsp.wav_to_mel_packed(a, plan, out=out)
        dsp.wav_to_mel_packed(a, offs)
    torch.cuda.synchronize(dev)
    reps = 10

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    ms_kernel = timed(lambda: dsp.wav_to_mel_packed(a, plan, out=out))
    ms_call = timed(lambda: dsp.wav_to_mel_packed(a, offs))
    out_host = torch.empty(out.shape, dtype=torch.float32).pin_memory()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(3):
        ad = a_host.to(dev, non_blocking=True)
        o, _ = dsp.wav_to_mel_packed(ad, offs)
        out_host.copy_(o, non_blocking=True)
    torch.cuda.synchronize(dev)
    ms_e2e = (time.perf_counter() - t0) * 1e3 / 3
    secs = a.numel() / 22050.0
    nbytes = a.numel() * 4 + out.numel() * 4
    # The kernel is bound by instruction issue, not by HBM: warp instructions per frame from the committed
    # ncu capture (smsp__inst_executed.sum / frames), against 4 issue slots per SM and clock.
    wipf = 1499.0
    traffic = None
    p = ROOT / 'profiles' / 'ncu_traffic.json'
    if p.exists():
        d = json.loads(p.read_text())
        wipf = float(d.get('stft_mel_warp_instr_per_frame', wipf))
        traffic = d.get('stft_mel')
    frames = out.numel() / 80
    props = torch.cuda.get_device_properties(dev)
    issue_peak = props.multi_processor_count * 4 * 1.965e9
    res = {'metric': 'stft_mel_audio_seconds_per_s', 'value': secs / (ms_kernel / 1e3), 'unit': 'audio-s/s',
           'ms_per_step': ms_kernel, 'clips': n_clips, 'audio_seconds': secs,
           'packed_call': {'value': secs / (ms_call / 1e3), 'ms_per_step': ms_call,
                           'note': 'DSP.wav_to_mel_packed(audio, clip_offsets) as a user calls it: plans the frame offsets '
                                   'on the host and uploads them (one pinned H2D) every call'},
           'e2e': {'value': secs / (ms_e2e / 1e3), 'unit': 'audio-s/s', 'ms_per_step': ms_e2e,
                   'h2d_bytes_per_step': a.numel() * 4 + (2 * n_clips + 2) * 8, 'd2h_bytes_per_step': out.numel() * 4,
                   'note': 'pinned host audio -> device, packed call, mel -> pinned host, wall clock'},
           'workload': 'DSP.wav_to_mel, 22.05 kHz clips of 2-10 s (noise, sines, silence), n_fft 1024 / hop 256 / 80 mels',
           'issue_roofline': {'bound': 'issue', 'achieved': frames * wipf / (ms_kernel / 1e3) / 1e9, 'peak': issue_peak / 1e9,
                              'unit': 'G warp-instr/s', 'frac': frames * wipf / (ms_kernel / 1e3) / issue_peak,
                              'note': 'warp instructions per frame from the committed ncu capture; peak = SMs x 4 schedulers x 1965 MHz'},
           'roofline': {'kernel': 'stft_mel', 'bound': 'hbm', 'achieved': nbytes / (ms_kernel / 1e3) / 1e9,
                        'peak': pk['hbm'], 'unit': 'GB/s', 'frac': nbytes / (ms_kernel / 1e3) / 1e9 / pk['hbm'],
                        'traffic': traffic, 'note': 'kernel alone (CUDA events around the launches, offsets planned once)'}}
    if cpu:
        try:
            res['cpu_baseline'] = stft_cpu_baseline(audio.numpy(), offs.tolist())
        except Exception as e:
            res['cpu_baseline'] = {'error': str(e)}
    return res


def fastpitch_extra(torch, dev, seed=5):
    """BASELINE.json configs[2]: FastPitch batch 128 x 300 phonemes with pitch + energy callbacks (per GPU)."""
    from forwardtacotron_b200 import _lib
    from forwardtacotron_b200.utils import synth
    lib = _lib.lib()
    model, _ = synth.synthetic_model('fast_pitch')
    model = model.to(dev)
    x = synth.synthetic_tokens(128, 300, seed=seed).to(dev)
    pf, ef = (lambda p: p * 1.2), (lambda e: e + 0.1)
    out = model.generate(x, pitch_function=pf, energy_function=ef)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        out = model.generate(x, pitch_function=pf, energy_function=ef)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    frames = int(out['mel_len'].sum().item())
    res = {'metric': 'mel_frames_per_s', 'value': frames / (ms / 1e3), 'unit': 'frames/s', 'ms_per_step': ms,
           'workload': 'FastPitch.generate batch 128 x 300 phonemes, pitch*1.2 / energy+0.1 callbacks',
           'mel_frames_padded_L': int(out['mel'].shape[-1]), 'valid_frames': frames,
           'numerics': 'IEEE-half tcgen05 GEMMs (LayerNorm fused into the out_proj / conv2 epilogues) + tcgen05 / TMEM '
                       'attention, fp32 accumulate / residual stream / LayerNorm; duration predictor fp32-grade (DESIGN.md 2)'}
    try:  # per-family kernel time of one more call (events around every launch)
        lib.ftb_profile_enable(1)
        model.generate(x, pitch_function=pf, energy_function=ef)
        torch.cuda.synchronize(dev)
        fams = collect_profile(lib, 1)
        lib.ftb_profile_enable(0)
        pk = peaks()
        res['kernels'] = fams
        res['rooflines'] = [roofline_of(f, pk) for f in fams if f['flops_per_step'] > 0]
    except Exception as e:
        res['kernels'] = {'error': str(e)}
    return res


def _agree(torch, dist, dev, ok: bool) -> bool:
    """All ranks learn whether every rank got through the previous (collective-free) phase."""
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t[0]))


def sharded_corpus_extra(torch, dist, model, dev, world):
    """N > 1: the path's one exchange step.  A corpus of 256 x world utterances is bucketed,
    sharded over the ranks, synthesised (ragged batches, 3 in flight) and collected on rank 0 three ways:
      none        -- no exchange: every rank keeps its results (what the exchange is measured against)
      nccl        -- exact-size grouped ncclSend / ncclRecv to rank 0 (utils/batching.gather_mels)
      peer_window -- the post_proj epilogue of every rank stores straight into rank 0's HBM over NVLink
    Wall clock of the whole run incl. host bucketing, max over ranks; second pass of each mode (first = warm-up)."""
    from forwardtacotron_b200.utils import batching
    from forwardtacotron_b200.utils.peer_window import PeerWindow
    g = torch.Generator().manual_seed(3)
    n = 256 * world
    utts = [torch.randint(1, 135, (int(k),), generator=g).tolist() for k in torch.randint(40, 200, (n,), generator=g)]
    window = PeerWindow(min(8 << 30, (1 << 30) * world))
    res, frames, err = {}, 0, None
    for rep in range(2):
        for mode in ('none', 'nccl', 'peer_window'):
            torch.cuda.synchronize(dev)
            dist.barrier()
            t0 = time.perf_counter()
            out = None
            try:
                out = batching.synthesize_corpus(model, utts, max_tokens=8192, gather=mode != 'none',
                                                 window=window if mode == 'peer_window' else None)
                torch.cuda.synchronize(dev)
            except Exception as e:  # keep the ranks in step: everybody reaches the barrier below
                err = f'{mode}: {e}'
            dist.barrier()
            res[mode] = time.perf_counter() - t0
            if mode == 'nccl' and out is not None and dist.get_rank() == 0:
                frames = sum(int(m.shape[1]) for m in out)
    t = torch.tensor([res['none'], res['nccl'], res['peer_window']], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    window.close()
    if dist.get_rank() != 0:
        return None
    r = {'utterances': n, 'frames': frames, 'no_gather_ms': float(t[0]) * 1e3, 'nccl_gather_ms': float(t[1]) * 1e3,
         'peer_window_ms': float(t[2]) * 1e3, 'frames_per_s_peer_window': frames / float(t[2]) if frames else None,
         'note': 'whole sharded corpus run (bucket, ragged generate with 3 batches in flight, collect on rank 0); '
                 'nccl = exact-size grouped send/recv; peer window = post_proj epilogue stores into rank 0 HBM over '
                 'NVLink, no gather pass, no collective inside generate'}
    if err:
        r['error'] = err
    return r


def multi_gpu_configs_extra(torch, dist, dev, world, rank, pk):
    """N > 1: BASELINE.json configs[2..4] sharded over the ranks.  Every rank runs its share with NO collective inside
    the timed work; the per-rank (units, seconds) pairs are reduced afterwards (sum of units / max of seconds)."""
    from forwardtacotron_b200.utils import batching, synth
    res, local = {}, {}
    # cfg3: FastPitch 128 x 300 per rank (utterance-sharded, weak scaling like the headline)
    try:
        fp = fastpitch_extra(torch, dev, seed=5 + rank)
        local['fast_pitch'] = (float(fp['valid_frames']), fp['ms_per_step'] / 1e3)
    except Exception as e:
        local['fast_pitch'] = (0.0, 0.0)
        res['fast_pitch_error'] = str(e)
    # cfg4: 10 000 clips of 2-10 s sharded over the ranks (strong scaling): 10 000 / N clips here
    try:
        st = stft_extra(torch, dev, pk, n_clips=10000 // world, seed=7 + rank, cpu=False)
        local['stft_mel'] = (float(st['audio_seconds']), st['ms_per_step'] / 1e3)
        local['stft_mel_e2e'] = (float(st['audio_seconds']), st['e2e']['ms_per_step'] / 1e3)
    except Exception as e:
        local['stft_mel'] = local['stft_mel_e2e'] = (0.0, 0.0)
        res['stft_mel_error'] = str(e)
    # cfg5: 256 x ~2000-phoneme utterances, length-bucketed into batches of 32, dealt over the ranks, alpha 0.8 / 1.0 / 1.2
    try:
        model, _ = synth.synthetic_model('forward_tacotron')
        model = model.to(dev)
        g = torch.Generator().manual_seed(11)
        utts = [torch.randint(1, 135, (int(n),), generator=g).tolist() for n in torch.randint(1900, 2001, (256,), generator=g)]
        batching.synthesize_corpus(model, utts, alpha=0.8, max_tokens=65536, in_flight=2, gather=False)
        torch.cuda.synchronize(dev)
        for alpha in (0.8, 1.0, 1.2):
            batching.synthesize_corpus(model, utts, alpha=alpha, max_tokens=65536, in_flight=2, gather=False)  # untimed (see long_article_extra)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            mels = batching.synthesize_corpus(model, utts, alpha=alpha, max_tokens=65536, in_flight=2, gather=False)
            torch.cuda.synchronize(dev)
            local[f'long_article_alpha_{alpha}'] = (float(sum(int(m.shape[1]) for m in mels if m is not None)),
                                                    time.perf_counter() - t0)
        del model
    except Exception as e:
        for alpha in (0.8, 1.0, 1.2):
            local.setdefault(f'long_article_alpha_{alpha}', (0.0, 0.0))
        res['long_article_error'] = str(e)
    keys = sorted(local)
    units = torch.tensor([local[k][0] for k in keys], dtype=torch.float64, device=dev)
    secs = torch.tensor([local[k][1] for k in keys], dtype=torch.float64, device=dev)
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    dist.all_reduce(secs, op=dist.ReduceOp.MAX)
    for k, u, t in zip(keys, units.tolist(), secs.tolist()):
        res[k] = {'units': u, 'max_seconds_over_ranks': t, 'value': (u / t) if t > 0 else None,
                  'unit': 'audio-s/s' if k.startswith('stft') else 'frames/s'}
    res['note'] = (f'configs[2..4] on {world} GPUs: fast_pitch = 128 x 300 per rank (weak); stft_mel = 10 000 clips / '
                   f'{world} ranks (strong), kernel-only and host-buffer e2e; long_article = 256 utterances sharded '
                   f'by length bucket (strong), wall clock incl. host bucketing; value = sum(units) / max(seconds)')
    return res


def long_article_extra(torch, dev):
    """BASELINE.json configs[4]: ForwardTacotron on 256 x 2000-phoneme utterances, length-bucketed into batches of 32
    (utils/batching.synthesize_corpus, batches in flight on separate streams), alpha sweep 0.8 / 1.0 / 1.2.
    Wall clock including the host-side bucketing and the per-row slicing; one untimed pass per alpha first."""
    from forwardtacotron_b200.utils import batching, synth
    model, _ = synth.synthetic_model('forward_tacotron')
    model = model.to(dev)
    g = torch.Generator().manual_seed(11)
    utts = [torch.randint(1, 135, (int(n),), generator=g).tolist() for n in torch.randint(1900, 2001, (256,), generator=g)]
    res = {}
    # warm-up at the LONGEST setting (alpha 0.8): lanes and packed weights are created and the workspaces sized once
    batching.synthesize_corpus(model, utts, alpha=0.8, max_tokens=65536, in_flight=2)
    torch.cuda.synchronize(dev)
    for alpha in (0.8, 1.0, 1.2):
        # one untimed pass per setting: every alpha has its own frame counts, i.e. its own first-touch allocations in
        # torch's caching allocator (a synchronising cudaMalloc inside the timed region otherwise)
        batching.synthesize_corpus(model, utts, alpha=alpha, max_tokens=65536, in_flight=2)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        mels = batching.synthesize_corpus(model, utts, alpha=alpha, max_tokens=65536, in_flight=2)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        frames = sum(int(m.shape[1]) for m in mels)
        res[f'alpha_{alpha}'] = {'frames': frames, 'ms': dt * 1e3, 'frames_per_s': frames / dt}
    return {'metric': 'mel_frames_per_s', 'unit': 'frames/s', 'utterances': 256, 'phonemes_per_utterance': '1900-2000',
            'batching': 'length-bucketed, 32 utterances per batch, 2 batches in flight', **res}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from forwardtacotron_b200 import _lib
    from forwardtacotron_b200.utils import synth

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the ftb200 kernels have no CPU fallback')
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.lib()
    pk = peaks()
    if args.stft_only:
        print(json.dumps(stft_extra(torch, dev, pk)), flush=True)
        return

    model, _ = synth.synthetic_model('forward_tacotron')
    model.gemm_mode = args.gemm_mode
    model = model.to(dev)
    x_host = synth.synthetic_tokens(B, T, seed=1 + rank).pin_memory()      # a different batch per rank
    x = x_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- S batches in flight: step k runs on CUDA stream k % S (the model keeps one native lane -- packed weights,
    # workspace, internal streams -- per stream).  The sequential recurrences of one batch leave most SMs idle; the
    # neighbour's GEMMs fill them.  S = 1 reproduces strictly back-to-back generate() calls.
    S = max(1, args.in_flight)
    main_stream = torch.cuda.current_stream(dev)
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    xs = [x] + [synth.synthetic_tokens(B, T, seed=101 * (i + 1) + rank).to(dev) for i in range(1, S)]
    xs_host = [x_host] + [t.cpu().pin_memory() for t in xs[1:]]

    def run_steps(n, from_host=False, sinks=None):
        for st in streams:
            st.wait_stream(main_stream)
        outs = [None] * S
        for k in range(n):
            i = k % S
            with torch.cuda.stream(streams[i]):
                xin = xs_host[i].to(dev, non_blocking=True) if from_host else xs[i]
                outs[i] = model.generate(xin)
                if sinks is not None:
                    sinks[i].copy_(outs[i]['mel_post'], non_blocking=True)
        for st in streams:
            main_stream.wait_stream(st)
        return outs

    outs = run_steps(max(args.warmup, 3) * S)
    torch.cuda.synchronize(dev)
    frames_lane = [int(o['mel_len'].sum().item()) for o in outs]
    frames_per_step = sum(frames_lane[k % S] for k in range(args.steps)) / args.steps   # valid frames of an average step
    out = outs[0]
    L = int(out['mel'].shape[-1])
    with torch.cuda.stream(streams[0]):
        ws_bytes = int(lib.ftb_ft_workspace_bytes(model._handle, B, T, L))

    # ---- timed region: K steps, inputs resident in HBM, CUDA events on the launching stream
    sampler = ClockSampler('GPU-' + str(torch.cuda.get_device_properties(dev).uuid)) if rank == 0 else None
    launches0 = lib.ftb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run_steps(args.steps)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = lib.ftb_launch_count() - launches0
    # the timed region above is the reported one; four more identical regions show the spread (it is ~0.1 s long)
    repeats = [ms_total / args.steps]
    for _ in range(4):
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        r0.record()
        run_steps(args.steps)
        r1.record()
        barrier()
        repeats.append(r0.elapsed_time(r1) / args.steps)
    clocks = sampler.stop() if sampler else None

    # ---- per-family kernel durations.  In the timed region kernels of several streams overlap (stage A forks onto
    # side streams, S batches are in flight), so an event pair around a launch also measures queueing.  Two extra
    # (untimed) steps on ONE stream with the handle's SERIALIZE option give every launch's own duration.
    _lib.check(lib.ftb_ft_set_option(model._get_handle(dev), _lib.FTB_OPT_SERIALIZE, 1))
    lib.ftb_profile_enable(1)
    PROF_STEPS = 2
    for _ in range(PROF_STEPS):
        out = model.generate(x)
    torch.cuda.synchronize(dev)
    fams = collect_profile(lib, PROF_STEPS)
    lib.ftb_profile_enable(0)
    _lib.check(lib.ftb_ft_set_option(model._get_handle(dev), _lib.FTB_OPT_SERIALIZE, 0))
    for f in fams:
        f['source'] = 'serialised profile pass (CUDA events around every launch, nothing overlapping)'
    fams.sort(key=lambda f: -f['ms_per_step'])

    # ---- end to end through the public API with HOST buffers: pinned H2D of the tokens + D2H of the result
    sinks = [torch.empty(o['mel_post'].shape, dtype=torch.float32).pin_memory() for o in outs]
    mel_host = sinks[0]
    if not args.no_extras:
        run_steps(2 * S, from_host=True, sinks=sinks)
    barrier()
    t0 = time.perf_counter()
    run_steps(1 if args.no_extras else args.steps, from_host=True, sinks=sinks)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3

    t = torch.tensor([ms_total, e2e_ms, float(frames_per_step)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, e2e_ms, frames_all = float(mx[0]), float(mx[1]), float(sm[2])
    else:
        frames_all = float(frames_per_step)

    # ---- latency of ONE generate() with nothing else in flight (the throughput figure keeps S batches in flight)
    lat_ms = None
    try:
        torch.cuda.synchronize(dev)
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        model.generate(x)
        l0.record()
        for _ in range(5):
            model.generate(x)
        l1.record()
        torch.cuda.synchronize(dev)
        lat_ms = l0.elapsed_time(l1) / 5
    except Exception:
        pass

    line = None
    if rank == 0:
        value = frames_all * args.steps / (ms_total / 1e3)
        e2e_steps = 1 if args.no_extras else args.steps  # --no-extras (profiling runs) times ONE e2e step only
        e2e_value = frames_all * e2e_steps / (e2e_ms / 1e3)
        top = fams[0]
        serial_ms = sum(f['ms_per_step'] for f in fams)
        kernels = [dict(f, share=f['ms_per_step'] / serial_ms) for f in fams]
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16' if args.gemm_mode == 0 else 'bf16', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'global_batch': B * world, 'phonemes': T, 'mel_frames_padded_L': L,
                       'valid_frames_per_gpu_step': frames_per_step, 'parallelism': f'utterance-sharded x{world}',
                       'batches_in_flight_per_gpu': S,
                       'numerics': ('IEEE-half' if args.gemm_mode == 0 else 'bf16') +
                                   ' operands on tcgen05, fp32 accumulate / state / epilogues; duration predictor '
                                   'fp32-grade (3 x bf16 split); output heads on two-part operands (hi + lo).  IEEE half '
                                   'instead of bf16 because only an 11-bit significand holds max-abs 1e-2 / mean-abs 1e-3 '
                                   'at trained-checkpoint mel magnitude (tests/test_gpu_forward_tacotron.py::'
                                   'test_trained_magnitude_stress); same tensor-core rate and bytes as bf16',
                       'l2': f'per-step working set {ws_bytes / 1e9:.2f} GB >> 126 MB L2, no explicit flush'},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': x_host.numel() * 8,
                    'd2h_bytes_per_step': mel_host.numel() * 4 + B * 4, 'ms_per_step': e2e_ms / e2e_steps},
            'ms_per_step_repeats': {'all': repeats, 'min': min(repeats), 'median': statistics.median(repeats),
                                    'note': 'this rank; [0] is the reported region, the others follow it back to back'},
            'latency_ms_in_flight_1': lat_ms,
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': roofline_of(top, pk),
            # the same figure for every kernel family that has algorithmic work attached (the dominant one above)
            'rooflines': [roofline_of(f, pk) for f in fams if f['flops_per_step'] > 0 or f['bytes_per_step'] > 0],
            'kernels': kernels,
            'kernels_note': f'share = fraction of the serialised step ({serial_ms:.3f} ms kernel time); the timed steps '
                            f'overlap stage A / prenet on side streams and keep {S} batches in flight',
        }

    # ---- extras.  The headline above is complete; a watchdog prints it as it stands if an extra stalls (a collective
    # that one rank never reaches must not cost the line), then the process exits.
    import threading
    printed = threading.Event()

    def emit():
        if rank == 0 and not printed.is_set():
            printed.set()
            print(json.dumps(line), flush=True)

    def watchdog():
        if rank == 0 and line is not None:
            line.setdefault('extra', {})['watchdog'] = f'extras exceeded {args.extras_timeout} s; line printed without the unfinished ones'
        emit()
        os._exit(0)

    timer = threading.Timer(args.extras_timeout, watchdog)
    timer.daemon = True
    if not args.no_extras:
        timer.start()
    extra = {}
    if not args.no_extras:
        if world > 1:
            for name, fn in (('final_gather', lambda: sharded_corpus_extra(torch, dist, model, dev, world)),
                             ('configs', lambda: multi_gpu_configs_extra(torch, dist, dev, world, rank, pk))):
                try:
                    r = fn()
                except Exception as e:
                    r = {'error': str(e)}
                if rank == 0 and r is not None:
                    extra[name] = r
                    line['extra'] = extra
        else:
            for name, fn in (('stft_mel', lambda: stft_extra(torch, dev, pk)),
                             ('fast_pitch', lambda: fastpitch_extra(torch, dev)),
                             ('long_article', lambda: long_article_extra(torch, dev))):
                try:
                    extra[name] = fn()
                except Exception as e:  # the headline line must still be printed
                    extra[name] = {'error': str(e)}
                line['extra'] = extra
            cb, _ = cpu_generate_rate(3, 1)
            line['cpu_baseline'] = cb
    timer.cancel()
    emit()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--in-flight', type=int, default=3, help='generate() calls in flight on as many CUDA streams')
    ap.add_argument('--gemm-mode', type=int, default=0, choices=[0, 2],
                    help='ForwardTacotron operand type: 0 IEEE half (default, holds the absolute tolerance at trained '
                         'magnitude), 2 bf16 (same kernels and rate)')
    ap.add_argument('--extras-timeout', type=float, default=420.0,
                    help='seconds the extras (other configs, CPU baselines, multi-GPU exchange) may take before the '
                         'headline line is printed without them')
    ap.add_argument('--no-extras', action='store_true', help='profiling runs: skip the e2e / STFT / CPU legs')
    ap.add_argument('--stft-only', action='store_true', help='profiling runs: only the STFT->mel leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
