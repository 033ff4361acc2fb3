#!/bin/bash
timeout 60 python - <<'PY'
import ctypes
from modulations_b200 import _lib
lib=_lib.load(); e=ctypes.c_int(-1)
rc=lib.b200dvb_tmem_selftest(ctypes.byref(e))
print("tmem selftest rc", rc, "errors", e.value, lib.b200dvb_last_cuda_error())
PY
echo "exit $?"
