"""fundsp_b200 — B200-native block-processing engine for FunDSP's voice-bank hot path."""
from . import prelude  # noqa: F401
from .graph import An, ArityError  # noqa: F401
