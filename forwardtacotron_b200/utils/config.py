"""Hyper-parameters of the hot path as a plain dict (the reference reads config.yaml into one,
utils/files.py:22-25).  Values are the reference defaults: ``dsp`` config.yaml:9-34,
``forward_tacotron.model`` :78-106, ``fast_pitch.model`` :130-163.  A user's own YAML / the dict
embedded in a checkpoint works the same way."""
from __future__ import annotations

import copy
from typing import Any, Dict

_DEFAULT: Dict[str, Any] = {
    'tts_model': 'forward_tacotron',
    'dsp': dict(sample_rate=22050, n_fft=1024, num_mels=80, hop_length=256, win_length=1024, fmin=0, fmax=8000,
                peak_norm=False, trim_start_end_silence=True, trim_silence_top_db=60, pitch_max_freq=600,
                trim_long_silences=False, vad_window_length=30, vad_moving_average_width=8,
                vad_max_silence_length=12, vad_sample_rate=16000, voc_mode='RAW', bits=9, mu_law=True),
    'forward_tacotron': {'model': dict(
        embed_dims=256, series_embed_dims=64,
        durpred_conv_dims=256, durpred_rnn_dims=64, durpred_dropout=0.5,
        pitch_conv_dims=256, pitch_rnn_dims=128, pitch_dropout=0.5, pitch_strength=1.0,
        energy_conv_dims=256, energy_rnn_dims=64, energy_dropout=0.5, energy_strength=1.0,
        prenet_dims=256, prenet_k=16, prenet_dropout=0.5, prenet_num_highways=4,
        rnn_dims=512,
        postnet_dims=256, postnet_k=8, postnet_num_highways=4, postnet_dropout=0.0)},
    'fast_pitch': {'model': dict(
        durpred_d_model=128, durpred_n_heads=2, durpred_layers=4, durpred_d_fft=128, durpred_dropout=0.5,
        pitch_d_model=128, pitch_n_heads=2, pitch_layers=4, pitch_d_fft=128, pitch_dropout=0.5, pitch_strength=1.0,
        energy_d_model=128, energy_n_heads=2, energy_layers=4, energy_d_fft=128, energy_dropout=0.5,
        energy_strength=1.0,
        d_model=256, conv1_kernel=9, conv2_kernel=1,
        prenet_layers=4, prenet_heads=2, prenet_fft=1024, prenet_dropout=0.1,
        postnet_layers=4, postnet_heads=2, postnet_fft=1024, postnet_dropout=0.1)},
}


def default_config(tts_model: str = 'forward_tacotron') -> Dict[str, Any]:
    cfg = copy.deepcopy(_DEFAULT)
    cfg['tts_model'] = tts_model
    return cfg
