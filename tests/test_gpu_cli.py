"""The reference's CI integration test (examples/ffm/run_fw_with_prediction_tests.sh) against the `fwgpu`
command-line front end: same flags, same assertions -- inference-weights predictions equal full-weights predictions
line by line, predictions are not constant, balanced accuracy on the unseen-combination ("hard") set > 0.80 --
plus the cache path (-c) and --sequential."""
import os
import random
import subprocess

import numpy as np
import pytest

from fwumious_wabbit_b200 import build

pytestmark = pytest.mark.gpu
FW = os.path.join(os.path.dirname(build.OUT), "fwgpu")


def get_score(a, b):  # examples/ffm/generate.py:12-20
    return 1 if (a == "Herbivore" and b == "Plant") or (a == "Carnivore" and b == "Meat") else -1


def generate(d, n_train=30000, n_eval=3000, num_animals=300, num_foods=200, block_beyond=3, seed=1):
    """examples/ffm/generate.py:31-92 with the arguments of run_fw_with_prediction_tests.sh:48 (block_beyond keeps its default 3):
    train pairs always involve one of the few "bridge" ids; the hard set pairs ids never seen together."""
    rnd = random.Random(seed)
    open(os.path.join(d, "vw_namespace_map.csv"), "w").write("A,animal\nB,food\n")

    def ex(person, movie):
        a, b = rnd.choice(["Herbivore", "Carnivore"]), rnd.choice(["Plant", "Meat"])
        return f"{get_score(a, b)} |A {a}-{person} |B {b}-{movie}\n"

    with open(os.path.join(d, "train.vw"), "w") as f:
        for _ in range(n_train):
            if rnd.randint(0, 1):
                f.write(ex(rnd.randint(0, num_animals), rnd.randint(0, block_beyond)))
            else:
                f.write(ex(rnd.randint(0, block_beyond), rnd.randint(0, num_foods)))
    with open(os.path.join(d, "test-hard.vw"), "w") as f:
        for _ in range(n_eval):
            f.write(ex(rnd.randint(block_beyond + 1, num_animals), rnd.randint(block_beyond + 1, num_foods)))


def run(args):
    r = subprocess.run([FW] + args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return r


def labels_of(path):
    return np.array([1.0 if l.split()[0] == "1" else 0.0 for l in open(path)])


def balanced_accuracy(p, y, thr=0.5):
    pred = p > thr
    tpr = np.mean(pred[y == 1]) if np.any(y == 1) else 0.0
    tnr = np.mean(~pred[y == 0]) if np.any(y == 0) else 0.0
    return 0.5 * (tpr + tnr)


@pytest.mark.parametrize("mode", ["hogwild", "sequential"])
def test_reference_ffm_integration_script(tmp_path, mode):
    d = str(tmp_path)
    generate(d)
    ns = "--keep A --keep B --interactions AB --ffm_k 10 --ffm_field A --ffm_field B".split()
    rest = "-l 0.1 -b 25 -c --sgd --loss_function logistic --link logistic --power_t 0.0 --l2 0.0 --hash all --noconstant".split()
    extra = ["--sequential"] if mode == "sequential" else []
    tr, full, inf = f"{d}/train.vw", f"{d}/full.fw", f"{d}/inference.fw"
    run(ns + rest + extra + ["--data", tr, "-p", f"{d}/training.txt", "-f", full, "--save_resume"])
    assert os.path.exists(tr + ".fwcache")                                   # -c wrote the cache (cache.rs:69-70)
    run(ns + rest + ["-i", full, "--convert_inference_regressor", inf])
    assert os.path.getsize(inf) <= os.path.getsize(full)  # equal here: --sgd models carry no accumulators
    run(ns + rest + ["-i", full, "--data", tr, "-p", f"{d}/eval_full.txt", "-t"])      # reads the cache this time
    run(ns + rest + ["-i", inf, "-d", tr, "-t", "-p", f"{d}/eval_inf.txt"])
    run(ns + rest + ["-i", inf, "-d", f"{d}/test-hard.vw", "-t", "-p", f"{d}/hard.txt"])
    y = labels_of(tr)
    p_train = np.loadtxt(f"{d}/training.txt")
    p_full, p_inf = np.loadtxt(f"{d}/eval_full.txt"), np.loadtxt(f"{d}/eval_inf.txt")
    assert len(p_train) == len(p_full) == len(p_inf) == len(y) == 30000
    assert open(f"{d}/eval_full.txt").read() == open(f"{d}/eval_inf.txt").read()   # run_fw_with_prediction_tests.sh:130-137
    assert len(np.unique(p_inf)) > 100                                             # :143-160 predictions are not constant
    assert balanced_accuracy(p_full, y) > 0.95
    hard = np.loadtxt(f"{d}/hard.txt")
    assert balanced_accuracy(hard, labels_of(f"{d}/test-hard.vw")) > 0.80            # :56, :247-252


def test_cli_flag_errors(tmp_path):
    d = str(tmp_path)
    generate(d, n_train=10, n_eval=2)
    r = subprocess.run([FW, "--data", f"{d}/train.vw", "--keep", "A", "-f", f"{d}/m.fw"], capture_output=True, text=True)
    assert r.returncode != 0 and "You need to use --save_resume with --final_regressor" in r.stderr   # main.rs:112-115
    r = subprocess.run([FW, "--data", f"{d}/train.vw", "--keep", "Z"], capture_output=True, text=True)
    assert r.returncode != 0 and "Unknown namespace" in r.stderr
    r = subprocess.run([FW, "--data", f"{d}/nothere/train.vw", "--keep", "A"], capture_output=True, text=True)
    assert r.returncode != 0 and "Could not find vw_namespace_map.csv" in r.stderr


def test_cli_deep_head(tmp_path):
    """The same round trip with a dense head (--nn_layers / --nn, model_instance.rs:430-470; regressor.rs:191-320):
    train with AdaGrad, save, convert to inference weights, predict with both files -> identical lines; it learns."""
    d = str(tmp_path)
    generate(d, n_train=20000, n_eval=1000)
    ns = "--keep A --keep B --ffm_k 4 --ffm_field A --ffm_field B --nn_layers 2 --nn 0:width:16 --nn 0:activation:relu --nn 1:width:8 --nn 1:activation:relu".split()
    rest = "-l 0.1 --ffm_learning_rate 0.05 --nn_learning_rate 0.02 -b 18 --ffm_bit_precision 18 --adaptive --sgd --power_t 0.4 --nn_power_t 0.45 --loss_function logistic --link logistic".split()
    tr, full, inf = f"{d}/train.vw", f"{d}/full.fw", f"{d}/inference.fw"
    run(ns + rest + ["--data", tr, "-p", f"{d}/training.txt", "-f", full, "--save_resume"])
    run(ns + rest + ["-i", full, "--convert_inference_regressor", inf])
    assert os.path.getsize(inf) < 0.6 * os.path.getsize(full)            # accumulators dropped from every block
    run(ns + rest + ["-i", full, "--data", tr, "-p", f"{d}/eval_full.txt", "-t"])
    run(ns + rest + ["-i", inf, "-d", tr, "-t", "-p", f"{d}/eval_inf.txt"])
    assert open(f"{d}/eval_full.txt").read() == open(f"{d}/eval_inf.txt").read()
    y = labels_of(tr)
    p = np.loadtxt(f"{d}/eval_full.txt")
    assert len(np.unique(p)) > 100 and balanced_accuracy(p, y) > 0.9
    r = subprocess.run([FW] + ns + rest + ["--data", tr, "--nn", "0:dropout:0.5"], capture_output=True, text=True)
    assert r.returncode != 0 and "not implemented" in r.stderr


def test_cli_gz_input_and_lz4_cache(tmp_path):
    """`*.gz` text input (all gzip members, buffer_handler.rs:19-23) and its LZ4-framed cache (cache.rs:68-71, 89-125): the
    same model trained in sequential mode from train.vw and from train.vw.gz predicts identically, the second pass reads the
    compressed cache, and an unknown extension is refused like the reference does (buffer_handler.rs:33-35)."""
    import gzip
    import shutil

    d = str(tmp_path)
    generate(d, n_train=4000, n_eval=10)
    lines = open(f"{d}/train.vw", "rb").read().splitlines(keepends=True)
    with open(f"{d}/packed.vw.gz", "wb") as f:   # two gzip members
        f.write(gzip.compress(b"".join(lines[:1500])))
        f.write(gzip.compress(b"".join(lines[1500:])))
    ns = "--keep A --keep B --ffm_k 4 --ffm_field A --ffm_field B".split()
    rest = "-l 0.1 -b 18 --ffm_bit_precision 18 --adaptive --sgd --sequential -c".split()
    run(ns + rest + ["--data", f"{d}/train.vw", "-p", f"{d}/p_plain.txt"])
    run(ns + rest + ["--data", f"{d}/packed.vw.gz", "-p", f"{d}/p_gz.txt"])
    assert open(f"{d}/p_plain.txt").read() == open(f"{d}/p_gz.txt").read()
    z = open(f"{d}/packed.vw.gz.fwcache", "rb").read()
    assert z[:4] == bytes([0x04, 0x22, 0x4D, 0x18]) and len(z) < os.path.getsize(f"{d}/train.vw.fwcache")
    os.remove(f"{d}/packed.vw.gz")               # the cache alone is enough now (cache.rs:98: "ignoring text input")
    open(f"{d}/packed.vw.gz", "wb").close()
    run(ns + rest + ["--data", f"{d}/packed.vw.gz", "-p", f"{d}/p_cache.txt"])
    assert open(f"{d}/p_cache.txt").read() == open(f"{d}/p_plain.txt").read()
    shutil.copy(f"{d}/train.vw", f"{d}/train.txt")
    r = subprocess.run([FW] + ns + ["--data", f"{d}/train.txt"], capture_output=True, text=True)
    assert r.returncode != 0 and "Please specify a valid input format (.vw, .zst, .gz)" in r.stderr


def test_cli_multi_batch_and_holdout_equal_single_batch(tmp_path):
    """Inputs larger than --batch_size, and the split --holdout_after causes, feed every mini-batch the right records:
    in --sequential mode (bit-exact reference semantics) 7 batches of 1000 + a holdout split predict exactly what one batch
    does (the offsets handed to fwgpu_learn_records are absolute, so the record base must not be shifted)."""
    d = str(tmp_path)
    generate(d, n_train=6500, n_eval=10)
    ns = "--keep A --keep B --ffm_k 4 --ffm_field A --ffm_field B".split()
    rest = "-l 0.1 -b 18 --ffm_bit_precision 18 --adaptive --sgd --sequential --holdout_after 4321".split()
    run(ns + rest + ["--data", f"{d}/train.vw", "-p", f"{d}/p_one.txt", "--batch_size", "100000"])
    run(ns + rest + ["--data", f"{d}/train.vw", "-p", f"{d}/p_many.txt", "--batch_size", "1000"])
    assert open(f"{d}/p_one.txt").read() == open(f"{d}/p_many.txt").read()
    # examples from holdout_after on are scored, not learned: the model is frozen there, so re-scoring that tail with the
    # saved regressor (-t) gives the same lines
    run(ns + rest + ["--data", f"{d}/train.vw", "-f", f"{d}/m.fw", "--save_resume", "--batch_size", "1000"])
    run(ns + ["-i", f"{d}/m.fw", "-t", "--data", f"{d}/train.vw", "-p", f"{d}/p_t.txt"])
    # (training mode returns the training-order forward, -t the predict-order one: same value up to the last float bits)
    assert np.max(np.abs(np.loadtxt(f"{d}/p_many.txt")[4320:] - np.loadtxt(f"{d}/p_t.txt")[4320:])) <= 2e-6
    # Hogwild mode with small batches learns as well as the sequential run does on this short stream (this is the path that
    # read misaligned words before)
    run(ns + rest[:-3] + ["--data", f"{d}/train.vw", "-p", f"{d}/p_h.txt", "--batch_size", "512"])
    y = labels_of(f"{d}/train.vw")
    ll_seq = -np.mean(np.where(y[3000:4320] == 1, np.log(np.loadtxt(f"{d}/p_one.txt")[3000:4320]), np.log(1 - np.loadtxt(f"{d}/p_one.txt")[3000:4320])))
    ll_hog = -np.mean(np.where(y[3000:4320] == 1, np.log(np.loadtxt(f"{d}/p_h.txt")[3000:4320]), np.log(1 - np.loadtxt(f"{d}/p_h.txt")[3000:4320])))
    assert ll_hog < 0.693 and abs(ll_hog - ll_seq) / ll_seq < 0.03, (ll_hog, ll_seq)


def test_cli_testonly_save_writes_a_loadable_inference_file(tmp_path):
    """`-t -i model -f out --save_resume`: the immutable ctx exports weights only, so the file it writes must say SGD
    (persistence.rs:163-172) -- and must load again and predict identically."""
    d = str(tmp_path)
    generate(d, n_train=3000, n_eval=10)
    ns = "--keep A --keep B --ffm_k 4 --ffm_field A --ffm_field B".split()
    rest = "-l 0.1 -b 18 --ffm_bit_precision 18 --adaptive --sgd".split()
    run(ns + rest + ["--data", f"{d}/train.vw", "-f", f"{d}/full.fw", "--save_resume"])
    run(ns + ["-i", f"{d}/full.fw", "-t", "--data", f"{d}/train.vw", "-p", f"{d}/p1.txt", "-f", f"{d}/resaved.fw", "--save_resume"])
    assert os.path.getsize(f"{d}/resaved.fw") < 0.6 * os.path.getsize(f"{d}/full.fw")
    run(ns + ["-i", f"{d}/resaved.fw", "-t", "--data", f"{d}/train.vw", "-p", f"{d}/p2.txt"])
    assert open(f"{d}/p1.txt").read() == open(f"{d}/p2.txt").read()


def test_cli_prediction_model_delay(tmp_path):
    """--prediction_model_delay D (main.rs:200-258): example i is scored by a model that has not seen the last D examples.
    With D >= the file every prediction comes from the initial model; with a small D the model scores almost as well as
    the up-to-date one."""
    d = str(tmp_path)
    generate(d, n_train=20000, n_eval=10)
    ns = "--keep A --keep B --interactions AB --ffm_k 10 --ffm_field A --ffm_field B".split()
    rest = "-l 0.1 -b 25 --sgd --power_t 0.0 --noconstant".split()
    run(ns + rest + ["--data", f"{d}/train.vw", "-p", f"{d}/p_all.txt", "--prediction_model_delay", "50000"])
    run(ns + rest + ["--data", f"{d}/train.vw", "-p", f"{d}/p_t.txt", "-t"])
    assert open(f"{d}/p_all.txt").read() == open(f"{d}/p_t.txt").read()
    run(ns + rest + ["--data", f"{d}/train.vw", "-p", f"{d}/p_50.txt", "--prediction_model_delay", "50"])
    run(ns + rest + ["--data", f"{d}/train.vw", "-p", f"{d}/p_0.txt"])
    y = labels_of(f"{d}/train.vw")

    def ll(path):
        p = np.clip(np.loadtxt(path), 1e-6, 1 - 1e-6)
        assert len(p) == 20000
        return -np.mean(np.where(y[10000:] == 1, np.log(p[10000:]), np.log(1 - p[10000:])))

    assert ll(f"{d}/p_0.txt") < ll(f"{d}/p_t.txt") - 0.05                       # training helps on this stream ...
    assert abs(ll(f"{d}/p_50.txt") - ll(f"{d}/p_0.txt")) < 0.03, (ll(f"{d}/p_50.txt"), ll(f"{d}/p_0.txt"))  # ... and a 50-example lag costs little


def test_cli_weight_quantization(tmp_path):
    """--convert_inference_regressor with --weight_quantization (main.rs:109, 136-148; quantization.rs): the FFM weights are
    written as 16-bit buckets under a ModelInstance that says so, `-t -i` dequantizes them, and the predictions stay close
    to the unquantized inference regressor's."""
    import json

    d = str(tmp_path)
    generate(d, n_train=20000, n_eval=10)
    ns = "--keep A --keep B --ffm_k 4 --ffm_field A --ffm_field B".split()
    rest = "-l 0.1 --ffm_learning_rate 0.05 -b 18 --ffm_bit_precision 16 --adaptive --sgd --power_t 0.4".split()
    tr, full, inf, q = f"{d}/train.vw", f"{d}/full.fw", f"{d}/inf.fw", f"{d}/q.fw"
    run(ns + rest + ["--data", tr, "-f", full, "--save_resume"])
    run(ns + rest + ["-i", full, "--convert_inference_regressor", inf])
    run(ns + rest + ["-i", full, "--convert_inference_regressor", q, "--weight_quantization"])
    n_f = (1 << 16) + 2 * 4
    assert os.path.getsize(inf) - os.path.getsize(q) == 2 * n_f - 8 + 1   # half the FFM block, less the header, and "true" is a byte shorter than "false"
    raw = open(q, "rb").read()
    l1 = int.from_bytes(raw[8:16], "little")
    l2 = int.from_bytes(raw[16 + l1:24 + l1], "little")
    mi = json.loads(raw[24 + l1:24 + l1 + l2])
    assert mi["optimizer"] == "SGD" and mi["dequantize_weights"] is True
    run(ns + rest + ["-i", inf, "-d", tr, "-t", "-p", f"{d}/p_inf.txt"])
    run(ns + rest + ["-i", q, "-d", tr, "-t", "-p", f"{d}/p_q.txt"])
    p_inf, p_q = np.loadtxt(f"{d}/p_inf.txt"), np.loadtxt(f"{d}/p_q.txt")
    assert len(p_inf) == len(p_q) == 20000 and len(np.unique(p_inf)) > 100, (len(p_inf), len(p_q), len(np.unique(p_inf)))
    print("quantized vs plain inference regressor: max |dp| %.2e, mean |dp| %.2e, std p %.3f" % (np.max(np.abs(p_inf - p_q)), np.mean(np.abs(p_inf - p_q)), np.std(p_inf)))
    assert np.max(np.abs(p_inf - p_q)) < 0.02 and np.mean(np.abs(p_inf - p_q)) < 2e-3
