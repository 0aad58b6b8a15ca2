import ctypes, sys, os, shutil
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from adell_mri_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), 'libadell_b200_prof.so')
sys.argv = ['bench.py'] + sys.argv[1:]
import bench
lib = _lib.load()
lib.adell_debug_prof.argtypes = [ctypes.c_void_p, ctypes.c_int]
import io, contextlib
bench.main()
out = (ctypes.c_ulonglong * 8)()
lib.adell_debug_prof(out, 1)
v = list(out)
print('cycles (sum over warps): prod wait-empty %d, issue %d, prepare %d | cons wait-full %d, compute %d' % tuple(v[:5]))
