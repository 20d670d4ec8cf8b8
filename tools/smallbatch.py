"""Small-batch chains (few planes): chain kernels vs owner kernels (run with B200W_OWNER=0 / 2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from chainbench import case  # noqa: E402

for n in (4, 8, 16, 32):
    case(n, 304, 304, "db3", "symmetric", 3)
    case(n, 256, 256, "haar", "reflect", 3)
