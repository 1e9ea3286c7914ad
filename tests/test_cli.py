"""The `calc` command line (tomahawk_b200/twkb_calc; reference lib/calc.h:56-240) and the C++
twk_ld mirror (include/twkb_ld.hpp). CPU tests cover option parsing and error behaviour (exit
code 1 + "[date][ERROR] message" on stderr, like the reference); the GPU tests run it end to end."""
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import twk_format as tf
from tests.helpers import assert_records_bitexact, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "tomahawk_b200", "twkb_calc")


def run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True, timeout=600)


def test_cli_is_built_and_prints_usage():
    assert os.access(CLI, os.X_OK), "build it: make -C tomahawk_b200/csrc"
    r = run()
    assert r.returncode == 1 and "Usage:" in r.stderr and "-I STRING" in r.stderr
    r = run("calc")
    assert r.returncode == 1 and "Usage:" in r.stderr


@pytest.mark.parametrize("args,msg", [
    (["-i", "a.twk", "-o", "b", "-r", "1.5"], "Cannot have minimum R-squared value > 1"),
    (["-i", "a.twk", "-o", "b", "-r", "-0.5"], "Cannot have a negative minimum R-squared value"),
    (["-i", "a.twk", "-o", "b", "-P", "2"], "Cannot have a cutoff P-value > 1"),
    (["-i", "a.twk", "-o", "b", "-t", "0"], "Cannot have a non-positive number of worker threads"),
    (["-i", "a.twk", "-o", "b", "-c", "0"], "Cannot have a negative or zero amount of partitions"),
    (["-i", "a.twk", "-o", "b", "-C", "0"], "Cannot have a non-positive start partition"),
    (["-i", "a.twk", "-o", "b", "-w", "12x"], "not an integer"),
    (["-i", "a.twk", "-o", "b", "-w", "0"], "Cannot have a non-positive window size"),
    (["-o", "b", "-p", "-u"], "No input value specified..."),
    (["-i", "does_not_exist.twk", "-p"], "Failed to open"),   # default -o is "-" (stdout) like the reference: only the input is wrong
    (["-i", "a.twk", "-o", "", "-p"], "No output value specified..."),
    (["-i", "a.twk", "-o", "b", "-g", "0,x"], "Illegal device list"),
    (["-i", "a.twk", "-o", "b", "-K", "avx"], "Unknown kernel"),
    (["-i", "a.twk", "-o", "b", "-w", "1000", "-c", "3", "-C", "1"], "Cannot use chunking in window mode!"),
    (["-i", "/nonexistent/a.twk", "-o", "b"], "Failed to open"),
])
def test_cli_rejects_like_the_reference(args, msg):
    r = run("calc", *args)
    assert r.returncode == 1
    assert msg in r.stderr
    if msg != "not an integer":
        assert re.search(r"\[\d{4}-\d\d-\d\d \d\d:\d\d:\d\d,\d{3}\]\[ERROR\] ", r.stderr)


def test_cli_reads_input_and_fails_loudly_without_a_device(tmpdir_repo):
    """No CPU fallback: on a box without a B200 the run must fail at context creation."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    s = tf.synth_genotypes(64, 700, seed=3)
    twk = os.path.join(tmpdir_repo, "cli_cpu.twk")
    tf.write_twk(twk, s)
    r = run("calc", "-i", twk, "-o", os.path.join(tmpdir_repo, "cli_cpu_out"), "-p", "-w", "1e4", "-I", "1:100-2000")
    assert r.returncode == 1
    assert "500 variants from 1 blocks" in r.stderr          # -I selected one block before the device was needed
    assert "window=TRUE" in r.stderr and "window_size=10000" in r.stderr
    assert "no CUDA device" in r.stderr and "no CPU path" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0", "0,0,0"])
def test_cli_end_to_end_matches_reference_golden(devices, tmpdir_repo):
    """File to file through the binary; with several contexts (here on one device) the tile grid
    is dealt between them and the union must still be the reference's record set."""
    s, ref, prm, pairs, cli = load_golden("phased_r01")
    twk = os.path.join(tmpdir_repo, "cli_gpu.twk")
    tf.write_twk(twk, s)
    out = os.path.join(tmpdir_repo, f"cli_gpu_{len(devices)}")
    r = run("calc", "-i", twk, "-o", out, "-g", devices, *cli.split())
    assert r.returncode == 0, r.stderr
    back = tf.read_two(out + ".two")
    assert len(back) == 2 * len(ref)
    assert_records_bitexact(tf.canonical(back, forward_only=True), ref, p_rtol=1e-9)
    m = re.search(r"Variants: ([\d,]+), genotypes: ([\d,]+), output: ([\d,]+)", r.stderr)
    assert int(m.group(1).replace(",", "")) == pairs
    assert int(m.group(3).replace(",", "")) == len(ref)


@pytest.mark.gpu
def test_cli_window_and_interval_modes(tmpdir_repo):
    s, ref, prm, pairs, cli = load_golden("window")
    twk = os.path.join(tmpdir_repo, "cli_w.twk")
    tf.write_twk(twk, s)
    out = os.path.join(tmpdir_repo, "cli_w_out")
    r = run("calc", "-i", twk, "-o", out, "-s", *cli.split())
    assert r.returncode == 0, r.stderr
    assert r.stderr.strip() == ""                            # -s: silent
    assert_records_bitexact(tf.canonical(tf.read_two(out + ".two"), forward_only=True), ref, p_rtol=1e-9)
    s, ref, prm, pairs, cli = load_golden("interval_two")
    tf.write_twk(twk, s)
    r = run("calc", "-i", twk, "-o", out, *cli.split())
    assert r.returncode == 0, r.stderr
    assert_records_bitexact(tf.canonical(tf.read_two(out + ".two"), forward_only=True), ref, p_rtol=1e-9)


@pytest.mark.gpu
def test_cli_streams_to_stdout_by_default(tmpdir_repo):
    """No -o (the reference's default out = "-", include/core.h:909-924): the complete .two goes to stdout, byte-compatible
    with the file the same run writes with -o (offsets in the index come from a byte counter, not ftell)."""
    s, ref, prm, pairs, cli = load_golden("phased_r01")
    twk = os.path.join(tmpdir_repo, "cli_so.twk")
    tf.write_twk(twk, s)
    r = subprocess.run([CLI, "calc", "-i", twk, "-s", *cli.split()], capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1000:]
    piped = os.path.join(tmpdir_repo, "cli_so_piped.two")
    open(piped, "wb").write(r.stdout)
    assert_records_bitexact(tf.canonical(tf.read_two(piped), forward_only=True), ref, p_rtol=1e-9)
    state, ents, meta = tf.read_two_index(piped)
    assert sum(e[1] for e in ents) == 2 * len(ref)          # the index covers every record of the stream


# ---- twkb_sort: the reference's `sort` command line (lib/sort.h) over twkb_two_sort ----
import tomahawk_b200 as tb  # noqa: E402

SORT_CLI = os.path.join(ROOT, "tomahawk_b200", "twkb_sort")


@pytest.mark.parametrize("args,msg", [
    (["-i", "", "-o", "x"], "No input value specified..."),
    (["-i", "in.two", "-o", "x", "-m", "0"], "Cannot set memory limit <= 0..."),
    (["-i", "in.two", "-o", "x", "-t", "0"], "Cannot set number of threads <= 0..."),
    (["-i", "in.two", "-o", "x", "-c", "0"], "Cannot set the compression level <= 0..."),
    (["-i", "does_not_exist.two", "-o", "x"], "Failed to open"),
])
def test_sort_cli_rejects_like_the_reference(args, msg):
    r = subprocess.run([SORT_CLI, "sort"] + args, capture_output=True, text=True)
    assert r.returncode == 1
    assert "[ERROR]" in r.stderr and msg in r.stderr


def test_sort_cli_sorts_a_file(tmpdir_repo):
    s, recs, prm, pairs, _ = load_golden("phased_r0")
    twk_path = os.path.join(tmpdir_repo, "sc.twk")
    tf.write_twk(twk_path, s)
    src = os.path.join(tmpdir_repo, "sc.two")
    w = tb.TwoWriter(src, tb.TwkFile(twk_path), "pytest", c_level=1, b_size=500)
    w.add(recs[np.random.default_rng(3).permutation(len(recs))])
    w.close()
    out = os.path.join(tmpdir_repo, "sc_sorted.two")
    r = subprocess.run([SORT_CLI, "sort", "-i", src, "-o", out, "-t", "2", "-c", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = tf.read_two(out)
    assert len(got) == 2 * len(recs)
    assert np.array_equal(np.lexsort((got["packB"], got["packA"], got["ridB"], got["ridA"])), np.arange(len(got)))
    assert tf.read_two_index(out)[0] == 2
