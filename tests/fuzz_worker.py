"""Child process of tests/test_fuzz.py: opens / sorts mutated .twk and .two files through the C-ABI. Every case must come back
as a result or a TwkbError -- a crash, an uncaught C++ exception or an unbounded allocation (the address space is limited)
ends the process with a non-zero status, which the parent test reports.
   python tests/fuzz_worker.py twk|two <first_seed> <n_cases> <work_dir>"""
import os
import random
import resource
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import tomahawk_b200 as tb  # noqa: E402
from oracle import ldcore as lc  # noqa: E402
from oracle import twk_format as tf  # noqa: E402


def mutate(raw: bytes, rng: random.Random) -> bytes:
    b = bytearray(raw)
    kind = rng.randrange(4)
    if kind == 0:      # a few random bytes
        for _ in range(rng.randrange(1, 8)):
            b[rng.randrange(len(b))] = rng.randrange(256)
    elif kind == 1:    # truncation
        b = b[:rng.randrange(1, len(b))]
    elif kind == 2:    # the tail: footer, index offset, compressed index
        for _ in range(rng.randrange(1, 6)):
            b[len(b) - 1 - rng.randrange(min(200, len(b)))] = rng.randrange(256)
    else:              # an 8-byte field replaced by an extreme value
        o = rng.randrange(0, len(b) - 8)
        b[o:o + 8] = struct.pack("<Q", rng.choice([1 << 62, (1 << 64) - 1, 1 << 40, 0]))
    return bytes(b)


def main():
    what, seed0, n, work = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    src_twk, src_two = os.path.join(work, "fuzz_src.twk"), os.path.join(work, "fuzz_src.two")
    s = tf.synth_genotypes(40, 1200, seed=5, missing_rate=0.02 if what == "twk" else 0.0)
    tf.write_twk(src_twk, s)
    if what == "two":
        recs, _ = lc.calc(s, lc.default_params(force_phased=1, minR2=0.05, minP=1.0))
        twk = tb.TwkFile(src_twk)
        w = tb.TwoWriter(src_two, twk, "fuzz", c_level=1, b_size=500)
        w.add(recs)
        w.close()
        twk.close()
    raw = open(src_twk if what == "twk" else src_two, "rb").read()
    resource.setrlimit(resource.RLIMIT_AS, (6 << 30, 6 << 30))   # a size field must never become the allocation
    ok = errs = 0
    for k in range(seed0, seed0 + n):
        case = os.path.join(work, "fuzz_case." + what)
        with open(case, "wb") as f:
            f.write(mutate(raw, random.Random(k)))
        try:
            if what == "twk":
                h = tb.TwkFile(case, n_threads=2, runs=(k % 2 == 1), intervals=(["1:1000-50000"] if k % 5 == 0 else ()))
                h.close()
            else:
                tb.sort_two(case, os.path.join(work, "fuzz_sorted.two"), c_level=1, n_threads=2, memory_limit=(200000 if k % 2 else 0))
            ok += 1
        except tb.TwkbError as e:
            if e.code == -4:   # TWKB_ENOMEM: an allocation was sized by a corrupt field
                print(f"case {k}: {e}")
                sys.exit(3)
            errs += 1
    print(f"{what}: {ok} accepted, {errs} rejected")


if __name__ == "__main__":
    main()
