"""`import vector_quantization` shim: put <repo>/dropin ahead of the reference on sys.path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wavenet_autoencoders_b200.vector_quantization import (  # noqa: E402,F401
    SlicedVectorQuantize, SlicedVectorQuantizeEMA, VectorQuantize, VectorQuantizeEMA)
