"""The reference's experiment drivers run UNCHANGED against this repository's `models` package (the
nn.Module / autograd.Function surface is the drop-in boundary).  CPU, 1-2 epochs, synthetic data; only where the
reference tree is mounted (the build container).  MonotonicMLP.py = config 1, ToyExperiments.py (train_toy :121-165) =
config 2, UCIExperiments.py (train_uci :54-192) = configs 3/4.  tests/driver_harness.py explains the stand-ins."""
import os
import re
import subprocess
import sys

import pytest

from conftest import REPO

REF = os.environ.get("UMNN_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "MonotonicMLP.py")), reason="reference tree not mounted")


def _run(driver, args, cwd, timer_calls=0, timeout=600):
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""
    env["OMP_NUM_THREADS"] = "4"
    env.pop("PYTHONPATH", None)
    if timer_calls:
        env["DRIVER_MAX_TIMER_CALLS"] = str(timer_calls)
    r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "driver_harness.py"), REF, driver] + args,
                       capture_output=True, text=True, env=env, cwd=str(cwd), timeout=timeout)
    assert r.returncode == 0 and "DRIVER_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout + r.stderr


def test_monotonic_mlp_driver_trains(tmp_path):
    out = _run("MonotonicMLP.py", ["-nb_train", "400", "-nb_test", "50", "-nb_epoch", "2"], tmp_path)
    losses = [float(v) for v in re.findall(r"Monotonic:\s+([-+0-9.eE]+)", out)]
    assert len(losses) == 2 and all(l == l and l < 1e6 for l in losses)
    assert losses[1] < losses[0]                 # the monotone network fits the data better after the second epoch


def test_toy_experiments_driver_trains_and_samples(tmp_path):
    # no epoch flag: stopped through the clock after the third epoch started (2 timer reads per epoch + 2 around invert)
    out = _run("ToyExperiments.py", ["-dataset", "moons", "-folder", str(tmp_path) + "/"], tmp_path, timer_calls=8)
    assert "DRIVER_STOPPED_BY_HARNESS" in out
    epochs = re.findall(r"epoch: (\d+) - Train loss: ([-+0-9.eE]+) - Test loss: ([-+0-9.eE]+)", out)
    assert len(epochs) >= 2 and all(float(tr) == float(tr) and float(te) == float(te) for _, tr, te in epochs)
    assert "Inversion time" in out                # summary_plots ran model.invert(z, 5, "ParallelSimpler") at epoch 0
    assert os.path.isfile(os.path.join(tmp_path, "moons", "model.pt"))


def test_uci_experiments_driver_trains(tmp_path):
    out = _run("UCIExperiments.py", ["--data", "power", "-nb_epoch", "2", "-b_size", "100", "-save", "run", "-steps", "10",
                                     "-solver", "CCParallel", "-hidden_embedding", "64", "64", "-hidden_derivative", "32", "32",
                                     "-nb_flow", "2", "-Lipshitz", "1.5"], tmp_path)
    epochs = re.findall(r"epoch: (\d+) - Train loss: ([-+0-9.eE]+) - Valid loss: ([-+0-9.eE]+)", out)
    assert [int(e) for e, _, _ in epochs] == [0, 1]
    assert all(float(tr) == float(tr) and float(va) == float(va) for _, tr, va in epochs)
    assert float(epochs[1][1]) < float(epochs[0][1])            # it trains
    assert os.path.isfile(os.path.join(tmp_path, "power", "run", "model_best_train.pt"))
