"""One chainbench line: python tools/onecase.py n h w wave mode J   (env knobs are read by the library)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from chainbench import case  # noqa: E402

n, h, w = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
case(n, h, w, sys.argv[4], sys.argv[5], int(sys.argv[6]))
