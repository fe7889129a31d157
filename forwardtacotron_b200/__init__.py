"""B200-native (sm_100a) implementation of ForwardTacotron / FastPitch batched inference and the
STFT->log-mel feature extraction, behind the reference's own Python API.

    from forwardtacotron_b200.models.forward_tacotron import ForwardTacotron
    from forwardtacotron_b200.models.fast_pitch import FastPitch
    from forwardtacotron_b200.utils.dsp import DSP
"""
__version__ = '0.1.0'
