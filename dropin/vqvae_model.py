"""`import vqvae_model` shim (optional): the reference's own vqvae_model.py also works on top of the other shims."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wavenet_autoencoders_b200.vqvae_model import VQVAE, ConvReLURes, Encoder  # noqa: E402,F401
