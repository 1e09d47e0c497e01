"""The C++ host layer (include/mmlb200.hpp: the reference's recommender classes over the C ABI, for hosts without a .NET
toolchain) driven by tests/cpp/host_test.cpp. CPU: it compiles against the header pair, System.Random / formatting /
ToString() / defaults / reader errors behave like the reference, and without a device the context throws. GPU: config 1
(example.train / example.test, k = 10, 30 epochs, seed 1) through the C++ classes against tests/golden/oracle_config1.json."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_config1.json")))


def unhex(h):
    return np.array([int(x, 16) for x in h], np.uint32).view(np.float32)


@pytest.fixture(scope="module")
def host_test():
    from mymedialite_b200 import build
    so = build.build()
    libdir = os.path.dirname(so)
    out_dir = os.path.join(ROOT, "tests", "cpp", "build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "host_test")
    src = os.path.join(ROOT, "tests", "cpp", "host_test.cpp")
    hdrs = [os.path.join(ROOT, "include", h) for h in ("mmlb200.h", "mmlb200.hpp")]
    if not os.path.exists(exe) or any(os.path.getmtime(f) > os.path.getmtime(exe) for f in [src, so] + hdrs):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src,
                               "-o", exe, "-L", libdir, "-lmmlb200", "-Wl,-rpath," + libdir])
    # libmmlb200.so needs libcudart / libnccl: same places the Python process finds them
    env = dict(os.environ)
    ldd = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    dirs = {os.path.dirname(l.split("=>")[1].split("(")[0].strip()) for l in ldd.splitlines()
            if "=>" in l and ("cudart" in l or "nccl" in l) and "not found" not in l}
    env["LD_LIBRARY_PATH"] = ":".join(sorted(dirs) + [env.get("LD_LIBRARY_PATH", "")])
    return exe, env


def test_cpp_host_layer_without_a_device(host_test):
    exe, env = host_test
    r = subprocess.run([exe, "cpu"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
    import torch
    if not torch.cuda.is_available():
        assert "Context() threw" in r.stdout        # no CPU fallback behind the C++ classes either


@pytest.mark.gpu
def test_cpp_host_one_process_two_gpus(host_test):
    """NumGpus = 2 from ONE process (mml_ctx_create(n_gpus = 2): ncclCommInitAll + one host thread per GPU inside the library):
    BiasedMatrixFactorization on a 10M-rating set of the config 2 shape (71.5k x 10.7k, k = 64) tracks the NumGpus = 1 run from
    the same initial model (per-epoch test RMSE within 1 %: two different parallel schedules, neither is the reference -- the
    gate against the reference's own run is tests/test_rmse_gate_gpu.py); WRMF lists are identical; 10^4 per-user Recommend()
    calls after the first are lookups."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    exe, env = host_test
    r = subprocess.run([exe, "multi", "2", "10000000"], capture_output=True, text=True, env=env, timeout=600)
    print(r.stdout)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
    rmse = {}
    for line in r.stdout.splitlines():
        p = line.split()
        if p[:2] == ["multi", "epoch"]:
            rmse[(int(p[2]), int(p[3]))] = float(p[4])
        if p[:2] == ["multi", "wrmf"]:
            assert float(p[6]) < 50.0, line                     # 10^4 Recommend() calls from the cache
        if p[:2] == ["multi", "wrmf_lists_identical"]:
            assert p[2] == p[4], line
    for e in range(6):
        assert abs(rmse[(2, e)] - rmse[(1, e)]) / rmse[(1, e)] < 0.01, (e, rmse)
    assert rmse[(1, 5)] < rmse[(1, 0)]


@pytest.mark.gpu
def test_cpp_host_layer_config1_against_the_fixture(host_test, tmp_path):
    exe, env = host_test
    g = os.path.join(ROOT, "tests", "golden")
    r = subprocess.run([exe, "gpu", os.path.join(g, "example.train"), os.path.join(g, "example.test"), str(tmp_path)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
    seen = set()
    for line in r.stdout.splitlines():
        p = line.split()
        if p[0] == "epoch":
            want = GOLD["config1"][p[1]]["rmse_train_test_per_epoch"][int(p[2])]
            np.testing.assert_allclose([float(p[3]), float(p[4])], want, rtol=2e-5, atol=2e-5, err_msg=line)
            seen.add((p[1], int(p[2])))
        elif p[0] == "predict":
            np.testing.assert_allclose(np.array(p[2:], float), unhex(GOLD["config1"][p[1]]["predict_test"]), rtol=1e-5, atol=1e-5)
        elif p[0] == "roundtrip":
            assert float(p[2]) < 1e-5, line         # the text format carries 7 significant digits
        elif p[0] == "scoreitems":
            assert int(p[2]) == 3 and all(np.isfinite(float(x)) for x in p[3:]), line
        elif p[0] == "retrain":
            assert np.isfinite(float(p[2])), line
        elif p[0] == "wrmf_top2":
            want = GOLD["wrmf_k4_3epochs"]["top2_ignoring_training_items"][p[1]]
            assert [int(x) for x in p[2::2]] == want["items"], line
            np.testing.assert_allclose(np.array(p[3::2], float), unhex(want["scores"]), rtol=1e-4, atol=1e-7)
        elif p[0] == "tostring":
            assert line == "tostring WRMF num_factors=4 regularization=0.015 alpha=1 num_iter=3"
    assert len(seen) == 60                         # 30 epochs of both rating predictors
    for name in ("BiasedMatrixFactorization", "MatrixFactorization"):
        first = open(os.path.join(str(tmp_path), name + ".model")).readline().strip()
        assert first == "MyMediaLite.RatingPrediction.Cuda" + name
