"""Multi-GPU GPU tests (need >= 2 CUDA devices: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`; skipped on a
one-GPU box).  NCCL, one process per GPU, launched through torch.distributed.run exactly like bench.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

from tests.util import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2])
def test_stylize_step_gradients_world_n_equal_world_1(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = tmp_path / "dist.json"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tests", "dist_grad_check.py"), str(out)],
                       capture_output=True, text=True, timeout=300, env=dict(os.environ, PYTHONPATH=ROOT))
    assert r.returncode == 0, r.stderr[-3000:]
    rep = json.load(open(out))
    print(json.dumps(rep))
    for case in ("patches", "split_patch"):
        c = rep[case]
        assert c["identical_on_all_ranks"], case
        # same terms, summed in a different order (atomics; two partial sums per weight gradient): fp32 rounding only
        assert c["grad_rel_l2_max"] < 1e-4, (case, c["grad_rel_l2"])
        # parameters after the Adam step: its first step moves every entry by lr * g / (|g| + 1e-8), i.e. by +-lr whatever |g| is,
        # so table entries whose gradient is pure rounding noise (|g| ~ 1e-9, a different sum order when ONE patch is split) move
        # differently; with whole patches each entry's partial sums come from one rank and the update is reproduced
        if case == "patches":
            assert c["param_rel_l2"] < 1e-5, case
