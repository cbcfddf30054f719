#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_sampler.py -x -q -m gpu > gpurun_out/r02_c5_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_c5_pytest.log
for v in "BASE=1" "DBG=8" "NOSTATS=1" "DBG=1"; do echo "== $v"; env SHAPESET=n128 $v timeout 300 python scripts/ncu_conv.py 16 5; done
python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from oracle.weights import make_state_dict
from buddy_b200.engine import Engine
torch.cuda.synchronize()
sd = make_state_dict(0)
t0 = time.time(); e = Engine(sd, "cuda"); torch.cuda.synchronize(); print("engine build s:", time.time() - t0)
PY
