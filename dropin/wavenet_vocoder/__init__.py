"""`import wavenet_vocoder` shim: put <repo>/dropin ahead of the reference on sys.path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from wavenet_autoencoders_b200.wavenet_vocoder import WaveNet, receptive_field_size, __version__  # noqa: E402,F401
from wavenet_autoencoders_b200.wavenet_vocoder import conv, mixture, modules, upsample, util, wavenet  # noqa: E402,F401

for _name in ("conv", "mixture", "modules", "upsample", "util", "wavenet"):
    sys.modules[__name__ + "." + _name] = globals()[_name]
