"""CLI (maxpro_b200/bin/maxpro): flag surface of src/main.rs:11-33 and Rust float formatting.
The pipeline itself needs a GPU; the argument handling and the formatter do not."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "maxpro_b200", "bin", "maxpro")


def run(*args):
    return subprocess.run([BIN, *args], capture_output=True, text=True)


def test_cli_exists_and_help_lists_reference_flags():
    assert os.path.exists(BIN), "build with __graft_entry__.build()"
    out = run("--help").stdout
    for flag in ("-s, --samples", "-i, --iterations", "-n, --ndims", "-p, --plot", "-o, --output-path", "--seed",
                 "-m, --metric", "--anneal-iterations", "--anneal-cooling", "--anneal-t", "max-pro, maxi-min"):
        assert flag in out
    assert "[default: 100000]" in out and "[default: 0.99]" in out


def test_cli_usage_errors_like_clap():
    r = run("-s", "5")
    assert r.returncode == 2 and "--iterations <ITERATIONS>" in r.stderr and "--ndims <NDIMS>" in r.stderr
    r = run("-s", "5", "-i", "3", "-n", "2", "--metric", "maxpro")
    assert r.returncode == 2 and "possible values: max-pro, maxi-min" in r.stderr
    r = run("-s", "x", "-i", "3", "-n", "2")
    assert r.returncode == 2 and "invalid value 'x'" in r.stderr
    r = run("-s", "0", "-i", "3", "-n", "2", "--seed", "1")   # main.rs:54: assert!(n_samples > 0)
    assert r.returncode == 101 and "n_samples must be positive and nonzero" in r.stderr
    assert r.stdout.startswith("Generating hypercube with seed 1")  # printed before the asserts (main.rs:51)


def test_rust_float_formatting():
    """println!("{:?}") / println!("{}") of f64 (main.rs:100-103): shortest round-trip digits."""
    cases = {"0.1": ("0.1", "0.1"), "1": ("1.0", "1"), "0.30000000000000004": ("0.30000000000000004",) * 2,
             "1e-7": ("1e-7", "0.0000001"), "1e-5": ("1e-5", "0.00001"), "0.00001234": ("1.234e-5", "0.00001234"),
             "0.0001": ("0.0001", "0.0001"),
             "1e16": ("1e16", "10000000000000000"), "72.39111811323956": ("72.39111811323956",) * 2,
             "-2.5": ("-2.5", "-2.5"), "inf": ("inf", "inf"), "5e-324": ("5e-324", "0." + "0" * 323 + "5"),
             "123456789012345680": ("1.2345678901234568e17", "123456789012345680")}
    for arg, (dbg, disp) in cases.items():
        assert run("--debug-fmt", arg).stdout.split() == [dbg, disp], arg
    rng = np.random.default_rng(0)
    for v in rng.random(200):  # in [0,1): repr() is Rust's Debug for this range above 1e-5
        if v >= 1e-4:
            assert run("--debug-fmt", repr(float(v))).stdout.split()[0] == repr(float(v))


@pytest.mark.gpu
def test_cli_pipeline_end_to_end():
    r = run("--iterations", "2000", "--samples", "20", "--ndims", "2", "--metric", "max-pro", "--seed", "42",
            "--anneal-iterations", "500")
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "Generating hypercube with seed 42"
    design = eval(lines[1])  # Rust Debug of Vec<Vec<f64>> is a valid Python literal
    assert len(design) == 20 and all(len(row) == 2 for row in design)
    vals = {ln.split(":")[0]: float(ln.split(":")[1]) for ln in lines[2:5]}
    assert list(vals) == ["Original metric", "Swapped metric", "Annealed metric"]
    assert vals["Annealed metric"] <= vals["Swapped metric"] <= vals["Original metric"]
    import oracle
    assert abs(oracle.maxpro_criterion(design) - vals["Annealed metric"]) <= 1e-9 * vals["Annealed metric"]
    r2 = run("-i", "1000", "-s", "12", "-n", "3", "-m", "maxi-min", "--seed", "7", "--anneal-iterations", "200")
    assert r2.returncode == 0 and "Annealed metric:" in r2.stdout


@pytest.mark.gpu
def test_cli_pipeline_on_a_design_larger_than_shared_memory():
    """--samples 5000 --ndims 8 is the design size of BASELINE config C5 (320 KB): the search takes the chunked
    path, both annealing stages keep the design in HBM.  The reference CLI runs at any size; so does this one."""
    r = run("--iterations", "8", "--samples", "5000", "--ndims", "8", "--metric", "maxi-min", "--seed", "3",
            "--anneal-iterations", "10")
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    design = np.array(eval(lines[1]))
    assert design.shape == (5000, 8)
    vals = {ln.split(":")[0]: float(ln.split(":")[1]) for ln in lines[2:5]}
    assert vals["Annealed metric"] >= vals["Swapped metric"] >= vals["Original metric"]
    import oracle
    assert oracle.maximin_criterion(design) == vals["Annealed metric"]


@pytest.mark.gpu
def test_cli_gpus_flag_gives_the_one_device_result():
    """--gpus N (engine-only flag; 0 = every device of the box): the search and the chains are spread over the devices
    inside the C ABI (maxpro_set_devices), index blocks + host fold, so the printed design and metrics are those of
    one device whatever N is.  Also with --anneal-chains, and through the two-stage search (100,000 candidates)."""
    base = ("--iterations", "100000", "--samples", "50", "--ndims", "3", "--metric", "max-pro", "--seed", "5",
            "--anneal-iterations", "300", "--anneal-chains", "64")
    one = run(*base, "--gpus", "1")
    every = run(*base, "--gpus", "0")
    assert one.returncode == 0 and every.returncode == 0, (one.stderr, every.stderr)
    assert one.stdout == every.stdout
    assert "--gpus" in run("--help").stdout
    r = run(*base, "--gpus", "999")
    assert r.returncode != 0 and "device" in r.stderr.lower()
