"""Where does the CLI's start-up second go?  Times, in a fresh process: cuInit, the primary context (cudaFree(0)), pvp_ctx_create,
and the first launch of each kernel family (lazy module loading).  Prints one line per stage."""
import ctypes as C
import os
import sys
import time

t0 = time.perf_counter()
cuda = C.CDLL("libcuda.so.1")
cuda.cuInit(0)
t1 = time.perf_counter()
rt = C.CDLL("libcudart.so.12")
rt.cudaFree(None)
t2 = time.perf_counter()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pixeltovoxelprojector_b200", "libpvp_b200.so"))
t3 = time.perf_counter()
h = C.c_void_p()
lib.pvp_ctx_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
rc = lib.pvp_ctx_create(0, None, C.byref(h))
t4 = time.perf_counter()
print(f"cuInit {t1 - t0:.3f} s; primary context (cudaFree(0)) {t2 - t1:.3f} s; dlopen libpvp_b200.so {t3 - t2:.3f} s; pvp_ctx_create {t4 - t3:.3f} s (rc {rc}); "
      f"visible devices: {os.environ.get('CUDA_VISIBLE_DEVICES', 'all')}")
