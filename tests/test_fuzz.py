"""Untrusted input at the C-ABI boundary (size fields of .twk / .two files are attacker-controlled and no C++
exception may cross `extern "C"`): seeded mutations of valid files -- random bytes, truncation, the footer / index region,
8-byte fields set to extreme values -- must each be either read or rejected with a TwkbError. The mutated files are handled in
a child process with a limited address space, so a crash, std::terminate or a 100 GB allocation fails the test instead of
taking pytest (or the machine) down. Found with the same harness: an index whose variant count was corrupted became a 137 GB
metadata allocation (now checked against the blocks' own sizes before anything is sized by it, hostio.cpp: read_twk)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("what,first,n", [("twk", 0, 250), ("twk", 1000, 150), ("two", 0, 120)])
def test_mutated_files_are_read_or_rejected(what, first, n, tmpdir_repo):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_worker.py"), what, str(first), str(n), tmpdir_repo],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-1500:])
    last = r.stdout.strip().splitlines()[-1]
    accepted, rejected = int(last.split()[1]), int(last.split()[3])
    assert accepted + rejected == n and rejected > n // 2, last
