import sys, torch, numpy as np
sys.path.insert(0, ".")
from tests.cases import CASES
from tests.test_gpu_kernel_variants import _run
for name, prec in (("het_cartpole_lean", "fp64"), ("c4_cartpole_rows", "fp64"), ("c4_frozenlake8_rows", "fp64")):
    case = CASES[name]
    spec = _run(case, prec, False, specialize=1, steps=12)
    lean = _run(case, prec, False, specialize=0, steps=12)
    for k, (x, y) in enumerate(zip(spec, lean)):
        for key in x:
            if not torch.equal(x[key], y[key]):
                a, b = x[key].double().flatten(), y[key].double().flatten()
                bad = (a != b).nonzero().flatten()
                print(name, "step", k, key, "n_bad", len(bad), "first", bad[:4].tolist(), a[bad[:3]].tolist(), b[bad[:3]].tolist(), "shape", tuple(x[key].shape))
                break
        else:
            continue
        break
    else:
        print(name, "identical")
