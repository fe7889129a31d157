"""Checkpoint format of the reference (utils/checkpoints.py:12-40): one ``torch.save`` dict
``{'model': state_dict, 'optim': state_dict, 'config': dict}``; model type chosen by
``config['tts_model']``."""
from __future__ import annotations

from pathlib import Path
from typing import Any, Dict, Union

import torch

from ..models.fast_pitch import FastPitch
from ..models.forward_tacotron import ForwardTacotron


def init_tts_model(config: Dict[str, Any]) -> Union[ForwardTacotron, FastPitch]:
    model_type = config.get('tts_model', 'forward_tacotron')
    if model_type == 'forward_tacotron':
        return ForwardTacotron.from_config(config)
    if model_type == 'fast_pitch':
        return FastPitch.from_config(config)
    raise ValueError(f'Model type not supported: {model_type}')


def save_checkpoint(model: torch.nn.Module, optim, config: Dict[str, Any], path: Union[Path, str]) -> None:
    torch.save({'model': model.state_dict(), 'optim': optim.state_dict() if optim is not None else {},
                'config': config}, str(path))


def load_tts_model(checkpoint_path: Union[Path, str]):
    """gen_forward.py:20-28: checkpoint -> (model on CPU, config)."""
    checkpoint = torch.load(checkpoint_path, map_location=torch.device('cpu'))
    config = checkpoint['config']
    model = init_tts_model(config)
    model.load_state_dict(checkpoint['model'])
    return model, config
