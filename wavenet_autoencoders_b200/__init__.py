"""B200-native hot path of MingjieChen/wavenet_autoencoders.

* ``wavenet_autoencoders_b200.wavenet_vocoder`` -- drop-in for the reference's ``wavenet_vocoder`` package
* ``wavenet_autoencoders_b200.vector_quantization`` -- drop-in for ``vector_quantization.py``
* ``wavenet_autoencoders_b200.vqvae_model`` -- the reference's encoder -> VQ -> WaveNet composition on top of them
* ``libwae_b200.so`` (include/wae_b200.h) -- the C-ABI CUDA library underneath; ``python -m wavenet_autoencoders_b200.build``

Put ``<repo>/dropin`` in front of the reference on ``sys.path`` to make ``import wavenet_vocoder`` /
``import vector_quantization`` resolve to this package.
"""
from ._lib import WaeError, launch_count  # noqa: F401

__version__ = "0.1.0"
