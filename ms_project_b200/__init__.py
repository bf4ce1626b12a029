"""ms_project_b200 -- B200-native candidate-kernel scoring for autoks (lschlessinger1/MS-project).

Host side is Python and mirrors the reference's model-scoring API; every number is computed by
hand-written sm_100a CUDA in `libb200gp.so` through ctypes.  There is no CPU fallback."""
from . import _lib  # noqa: F401
from .program import ProgramBatch, compile_kernel  # noqa: F401

__all__ = ['ProgramBatch', 'compile_kernel']
