"""Drop-in for the reference's ``wavenet_vocoder`` package (wavenet_vocoder/__init__.py:6)."""
from .version import version as __version__
from .wavenet import WaveNet, receptive_field_size

__all__ = ["WaveNet", "receptive_field_size", "__version__"]
