from .forward_tacotron import ForwardTacotron  # noqa: F401
from .fast_pitch import FastPitch  # noqa: F401
