import sys, ctypes
sys.path.insert(0, ".")
import torch
from twotowermlretrieval_b200 import _lib
torch.zeros(1, device="cuda")
lib = _lib.load()
out = ctypes.c_int(0)
rc = lib.ttr_debug_gru_tc_max_clusters(ctypes.byref(out))
print("rc", rc, "max active 8-CTA clusters of gru_fwd_tc_kernel:", out.value, lib.ttr_last_error())
