"""CPU: the reference arm of bench.py (`--impl reference`: the oracle port of the training step on the
host cores) runs without a GPU and prints ONE JSON line with the keys the driver reads; the B200 arm
refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True,
                          text=True, timeout=timeout)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--workload", "kuka_b64", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "triplets/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["config"]["workload"] == "kuka_b64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert abs(d["e2e"]["value"] - d["value"]) < 1e-9


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("needs a machine without a GPU")
    r = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-reward", timeout=120)
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]


def test_bench_dataset_is_in_the_reference_formats(tmp_path):
    """bench.py writes its synthetic triplets the way the reference stores them (pickled records + FSC csv /
    GoogleCommand folders) and the audioLoader mirror reads them back through its reference-named loaders."""
    import pickle
    from importlib import import_module
    sys.path.insert(0, ROOT)
    import bench
    al = import_module("voicecontrolledrobot-var_b200.Envs.audioLoader")
    for name in ("ithor_b256", "kuka_b64"):
        wl = dict(bench.WORKLOADS[name], items=40, clips_per_list=3)
        root = str(tmp_path / name)
        bench.write_dataset(root, wl, name)
        cfg = bench.make_config(name, wl, root)
        a = al.audioLoader(cfg)
        a.loadData()
        assert a.fs == 16000
        if wl["net"] == "ithor":
            assert cfg.name == "AI2ThorConfig" and cfg.taskNum == 4
            lists = {(l, o, x): len(a.words[l][o][x]) for l in a.words for o in a.words[l] for x in a.words[l][o]}
            assert len(lists) == 6 and set(lists.values()) == {3}
        else:
            assert cfg.name == "ArmConfig"
            assert [len(a.words[i]["GoogleCommand"]) for i in range(4)] == [3, 3, 3, 3]
        recs = []
        for f in sorted(os.listdir(os.path.join(root, "data", "train"))):
            recs += pickle.load(open(os.path.join(root, "data", "train", f), "rb"))
        assert len(recs) == 40 and recs[0]["image"].shape == (3, 96, 96) and recs[0]["image"].dtype.name == "uint8"
        assert set(recs[0]) == {"image", "ground_truth"} and all(0 <= r["ground_truth"] <= 4 for r in recs)
