#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
$T 300 python -m pytest tests/test_node_parallel.py -m gpu -q --timeout=200 -k "reference_test" > gpurun_out/pytest_r2final5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2final5.log
grep -E "^E  |^FAILED|passed|failed|rc=" gpurun_out/pytest_r2final5.log | tail -6
$T 120 python - > gpurun_out/tiny_grid_r2final5.log 2>&1 <<'PY'
import numpy as np, sys
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
from pysdc_b200 import backend; backend.set_backend(backend.CudaBackend())
from pysdc_b200.problems import heatNd_unforced
import sdc_oracle
for n in (2, 4, 10000):
    kw = dict(nvars=n, nu=0.1, freq=2, bc="periodic", solver_type="direct")
    P, O = heatNd_unforced(**kw), sdc_oracle.HeatFD(forced=False, **kw)
    rng = np.random.default_rng(n); rhs, u0 = rng.standard_normal(n), rng.standard_normal(n)
    r = P.u_init; r[:] = rhs; x = P.u_init; x[:] = u0
    got = P.solve_system(r, 0.013, x, 0.0).get(); ref = O.solve_system(rhs, 0.013, u0, 0.0)
    f = P.eval_f(x, 0.0).get(); fr = O.eval_f(u0, 0.0)
    print("n", n, "solve rel err", float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))), "eval_f rel err", float(np.max(np.abs(f - fr)) / max(np.max(np.abs(fr)), 1e-300)))
PY
cat gpurun_out/tiny_grid_r2final5.log | tail -5
