"""GPU end-to-end parity: the NATIVE multi-GPU path against the single-GPU run — same seeded
init, same global batch, dropout 0, K optimizer steps of the bench model through ``GPT2Trainer.train_step`` — the
loss curve must agree within 2e-2.  Reference intent: ``tests/test_hybrid.py:10-19`` (placeholders there).

Each arm is a separate launch of ``bench.py --parity-steps`` (one process per GPU, NCCL + the in-kernel NVLink
collectives), compared by ``bench.py --check-parity``.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STEPS = 4


def _run(n, mesh, out, port, model):
    base = [sys.executable]
    if n > 1:
        base += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                 "--master-port", str(port)]
    cmd = base + [os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--mesh", mesh, "--model", model,
                  "--parity-steps", str(STEPS), "--parity-out", out]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return json.load(open(out))


@pytest.fixture(scope="module")
def single_gpu_losses(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("parity") / "n1.json")
    return _run(1, "1,1,1", out, 0, "gpt2-small"), out


@pytest.mark.gpu
@pytest.mark.parametrize("n,mesh,port", [
    pytest.param(2, "1,2,1", 29711, marks=pytest.mark.multigpu(2), id="tp2-fused-sp"),
    pytest.param(2, "2,1,1", 29712, marks=pytest.mark.multigpu(2), id="dp2-zero1"),
    pytest.param(2, "1,1,2", 29713, marks=pytest.mark.multigpu(2), id="pp2-1f1b"),
    pytest.param(8, "2,2,2", 29714, marks=pytest.mark.multigpu(8), id="3d-2x2x2"),
])
def test_native_multi_gpu_loss_matches_single_gpu(n, mesh, port, single_gpu_losses, tmp_path):
    ref, ref_path = single_gpu_losses
    out = str(tmp_path / f"n{n}.json")
    got = _run(n, mesh, out, port, "gpt2-small")
    assert "gemm" in got["native_ops"] and "gemm" in ref["native_ops"]      # both arms ran the native kernels
    if "1,2,1" in mesh or "2,2,2" in mesh:
        assert "tp_fused" in got["native_ops"]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--check-parity", ref_path, out],
                       cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert all(abs(a - b) < 2e-2 for a, b in zip(ref["losses"], got["losses"])), (ref["losses"], got["losses"])
    assert got["losses"][-1] < got["losses"][0]          # and it trains
