"""2-D discrete wavelet transform path (mirror of ``pytorch_wavelets.dwt``)."""
from . import lowlevel  # noqa: F401
from .transform2d import DWTForward, DWTInverse  # noqa: F401
