"""GPU: the host-buffer step (nsgym_step_host -- the end-to-end call bench.py times as `e2e`): actions
from pinned host memory, results copied back, chunked over several streams.  It must return exactly
what the device-resident step leaves in the buffers, for any chunking and any batch size."""
import pytest

from tests.cases import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,precision,n,chunks", [
    ("c1_cartpole_readme", "fp32", 5000, 8), ("c1_cartpole_readme", "fp32", 257, 3), ("c1_cartpole_readme", "fp64", 4097, 1),
    ("c3_pendulum", "fp32", 3000, 5), ("c3_acrobot", "fp64", 1025, 4), ("c2_frozenlake8_drift", "fp64", 6001, 7),
    ("c5_bridge_split", "fp64", 999, 16), ("c4_cartpole_rows", "fp32", 700, 3),
    # chunks of >= 2^21 envs: every chunk runs the tiled (TMA-prefetched) specialised gridworld kernel on its sub-range
    ("c5_bridge_uniform", "fp64", (1 << 22) + 512, 2), ("c1_cartpole_readme", "fp32", 1 << 17, 4)])
def test_host_step_equals_device_step(name, precision, n, chunks):
    import torch

    from tests import parity_util as pu

    case = CASES[name]
    dev_env, host_env = pu.gpu_env(case, n, precision=precision), pu.gpu_env(case, n, precision=precision)
    dev_env.reset(seed=8)
    host_env.reset(seed=8)
    h_act, h_out = host_env.make_host_io()
    for k in range(12 if n < (1 << 20) else 4):
        a = dev_env.action_space.sample()
        dev_env.step_raw(a)
        h_act.copy_(a.cpu())
        host_env.step_host(h_act, h_out, n_chunks=chunks)
        torch.cuda.synchronize()
        b = dev_env.buffers
        assert torch.equal(h_out["reward"], b["reward"].cpu()), f"reward, step {k}"
        assert torch.equal(h_out["flags"], b["flags"].cpu()), f"flags, step {k}"
        assert torch.equal(h_out["change"], b["change"].cpu()), f"change, step {k}"
        if "obs" in h_out:
            assert torch.equal(h_out["obs"], b["obs"].cpu()), f"obs, step {k}"
        if "state" in h_out:
            assert torch.equal(h_out["state"], b["state"].cpu()), f"state, step {k}"
        for key in ("state", "theta", "t"):
            assert torch.equal(host_env.buffers[key], b[key]), f"device {key}, step {k}"
    h2d, d2h = host_env.host_bytes_per_step(h_act, h_out)
    assert h2d == h_act.numel() * h_act.element_size() and d2h > 0
