"""bench.py's driver contract, as far as it can be exercised without a GPU: the reference arm (`--impl reference`, the
CPU restatement of the ITensors.jl path timed on the host cores) prints ONE JSON line with the contract's keys, uses
every host core even when the launcher exports OMP_NUM_THREADS=1 (torchrun does), only rank 0 works under a multi-rank
launch, and the GPU arm refuses to run without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env_extra=None, timeout=300):
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        env.pop(k, None)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, env=env)


def test_reference_arm_line_has_the_contract_keys():
    out = _run(["--impl", "reference", "--chi", "192", "--steps", "2", "--warmup", "1", "--gpus", "1"])
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert j["impl"] == "reference" and j["metric"] == base["metric"]
    for k in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in j, k
    assert j["steps"] == 2 and j["warmup"] == 1 and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["dtype"] == "f64" and j["gpu_launches"] == 0
    assert j["config"]["chi"] == j["config"]["reference_sample_chi"] == 192          # same config as asked, never shrunk
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # value is flops / time: F = 2 d^2 w (2 chi^3) + 4 d^3 w^2 chi^2  
    F = 2.0 * 4 * 5 * 2 * 192 ** 3 + 4.0 * 8 * 25 * 192 ** 2
    assert j["config"]["flop_per_step"] == F
    assert abs(j["value"] - F / (j["ms_per_step"] * 1e-3) * 1e-12) < 1e-9 * max(1.0, j["value"])


def test_reference_arm_uses_all_cores_under_torchrun_and_only_rank0_works():
    ncores = len(os.sched_getaffinity(0))
    out = _run(["--impl", "reference", "--chi", "128", "--steps", "1", "--warmup", "0", "--gpus", "2"],
               {"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert out.returncode == 0, out.stderr
    j = json.loads(out.stdout.strip().splitlines()[-1])
    assert j["cpu_baseline"]["cores"] == ncores and j["config"]["host_cores"] == ncores and j["n_gpus"] == 2
    out = _run(["--impl", "reference", "--chi", "128", "--gpus", "2"], {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    out = _run(["--steps", "1", "--warmup", "1", "--no-sweep", "--no-tebd"])
    assert out.returncode != 0
    assert "no CPU path" in (out.stderr + out.stdout)
