"""Mel featurisation of a folder of wavs on the GPU: the step either side of ``DSP.wav_to_mel`` in the reference's
``preprocess.py`` (:41-76): load -> start / end silence trimming (:66-67, ``DSP.trim_silence``) -> peak scaling
(:70-73) -> ``wav_to_mel`` (:76) -> ``np.save(mel/{id}.npy)`` (:44, float32 ``(n_mels, frames)``).  The reference runs
one file per ``multiprocessing`` worker; here all clips of a chunk go through ONE trim launch and ONE mel launch.

Outside this module (CPU-only upstream, third-party): resampling on load (librosa / resampy), the webrtcvad
long-silence trimmer (:64-65; off in config.yaml), WORLD pitch (:79-80), quantised waveforms for WaveRNN (:83-90),
text cleaning.

    python -m forwardtacotron_b200.preprocess --path wavs/ --out data/ [--config config.yaml]
"""
from __future__ import annotations

import argparse
from pathlib import Path
from typing import Iterable, List, Tuple

import numpy as np

from .utils.config import default_config
from .utils.dsp import DSP


def load_wav(path: Path, sample_rate: int) -> np.ndarray:
    """float32 mono in [-1, 1]; the file must already be at ``sample_rate`` (no resampler in this package)."""
    from scipy.io import wavfile
    sr, y = wavfile.read(str(path))
    if sr != sample_rate:
        raise ValueError(f'{path}: sample rate {sr} != {sample_rate} (resample upstream)')
    if y.ndim > 1:
        y = y.mean(axis=1)
    if np.issubdtype(y.dtype, np.integer):
        y = y.astype(np.float32) / float(np.iinfo(y.dtype).max + 1)
    return np.ascontiguousarray(y, dtype=np.float32)


def peak_scale(y: np.ndarray, should_peak_norm: bool) -> np.ndarray:
    """preprocess.py:70-73: divide by the peak when peak_norm is set or the clip exceeds full scale."""
    peak = float(np.abs(y).max()) if y.size else 0.0
    if peak > 0 and (should_peak_norm or peak > 1.0):
        y = y / peak
    return y


def trim_clips(clips: List[np.ndarray], dsp: DSP) -> List[np.ndarray]:
    """``DSP.trim_silence`` (preprocess.py:66-67) over a list of clips in one launch."""
    import torch
    dev = dsp._default_device()
    offs = torch.zeros(len(clips) + 1, dtype=torch.int64)
    offs[1:] = torch.cumsum(torch.tensor([len(c) for c in clips]), 0)
    flat = torch.from_numpy(np.concatenate(clips).astype(np.float32, copy=False)).pin_memory().to(dev, non_blocking=True)
    b = dsp.trim_bounds(flat, offs).cpu().tolist()
    return [c[s:e] for c, (s, e) in zip(clips, b)]


def featurize(paths: Iterable[Path], dsp: DSP, out_dir: Path, chunk: int = 256) -> List[Tuple[str, int]]:
    """wav files -> ``out_dir/mel/{stem}.npy``; returns [(id, frames)] like the reference's dataset list
    (preprocess.py:49, 148)."""
    paths = list(paths)
    (out_dir / 'mel').mkdir(parents=True, exist_ok=True)
    done: List[Tuple[str, int]] = []
    for i in range(0, len(paths), chunk):
        part = paths[i:i + chunk]
        clips = [load_wav(p, dsp.sample_rate) for p in part]
        if dsp.should_trim_long_silences:
            raise NotImplementedError('trim_long_silences needs webrtcvad (CPU, third-party); set it to false')
        if dsp.should_trim_start_end_silence:
            clips = trim_clips(clips, dsp)
        clips = [peak_scale(y, dsp.should_peak_norm) for y in clips]
        mels = dsp.wav_to_mel_batch(clips)
        for p, m in zip(part, mels):
            m = np.asarray(m, dtype=np.float32)
            np.save(out_dir / 'mel' / f'{p.stem}.npy', m, allow_pickle=False)
            done.append((p.stem, int(m.shape[-1])))
    return done


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description='GPU mel featurisation (the wav_to_mel part of preprocess.py)')
    ap.add_argument('--path', '-p', required=True, help='folder with .wav files')
    ap.add_argument('--out', default='data')
    ap.add_argument('--config', default=None, help='config.yaml (default: built-in reference defaults)')
    args = ap.parse_args(argv)
    if args.config:
        import yaml
        config = yaml.safe_load(Path(args.config).read_text())
    else:
        config = default_config()
    dsp = DSP.from_config(config)
    files = sorted(Path(args.path).rglob('*.wav'))
    done = featurize(files, dsp, Path(args.out))
    print(f'{len(done)} wav files -> {args.out}/mel, {sum(n for _, n in done)} frames')


if __name__ == '__main__':
    main()
