"""Input-type predicates (wavenet_vocoder/util.py of the reference)."""

_VALID = ("mulaw-quantize", "mulaw", "raw")


def _check(s):
    assert s in _VALID, f"input_type must be one of {_VALID}"


def is_mulaw_quantize(s):
    _check(s)
    return s == "mulaw-quantize"


def is_mulaw(s):
    _check(s)
    return s == "mulaw"


def is_raw(s):
    _check(s)
    return s == "raw"


def is_scalar_input(s):
    return is_raw(s) or is_mulaw(s)
