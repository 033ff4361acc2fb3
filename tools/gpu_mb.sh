#!/bin/bash
python - <<'PY'
import numpy as np
from modulations_b200 import _lib
lib=_lib.load(); r=np.zeros(8); lib.b200dvb_microbench(_lib.host_ptr(r))
print("microbench", " ".join(f"{n}={v:.1f}" for n,v in zip(["FADD","FMNMX","ACS","SHFL","DADD","F2F","FADD2","MHz"],r)))
PY
SKIP_TESTS=1 CFGS="3:448 3:320 3:192" bash tools/gpu_perf_only.sh
