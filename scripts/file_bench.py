"""File-to-file timing of `calc` (twkb_calc_file): .twk on disk -> .two on disk, with the phases
that surround the kernels (SURVEY.md 8(f)1 device-side ingest, 8(f)2 writer throughput).

  python scripts/file_bench.py [--variants M] [--samples N] [--min-r2 R] [--threads T] [--reference M']

Prints one JSON line per arrangement: host unpack + 1 writer thread (the reference's arrangement:
twk_igt_vec::Build on the host), host unpack + T threads, device decode + T threads (default of
twkb_calc_file). With --reference M' the reference's own calc binary (oracle/_ref) is timed
file-to-file on the first M' variants. The three outputs are compared record for record."""
import argparse, json, os, sys, tempfile, time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tomahawk_b200 as tb
from oracle import twk_format as tf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", type=int, default=200_000)
    ap.add_argument("--samples", type=int, default=2504)
    ap.add_argument("--min-r2", type=float, default=0.1)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--reference", type=int, default=0)
    ap.add_argument("--seed", type=int, default=20)
    ap.add_argument("--rare-fraction", type=float, default=0.0)
    ap.add_argument("--sorted", action="store_true", help="also time the sorted-output arrangements (device sort vs calc + host sorter)")
    a = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="twkb_file_")
    t0 = time.perf_counter()
    s = tf.synth_genotypes(a.samples, a.variants, seed=a.seed, rare_fraction=a.rare_fraction)
    twk = os.path.join(tmp, "in.twk")
    tf.write_twk(twk, s)
    gen = time.perf_counter() - t0
    pairs = a.variants * (a.variants - 1) // 2
    outs = {}
    for name, host_unpack, threads in (("host_unpack_1thread", 1, 1), ("host_unpack", 1, a.threads), ("device_decode", 0, a.threads)):
        best = None
        for rep in range(3):
            ld = tb.twk_ld()
            st = tb.default_settings(force_phased=1, minR2=a.min_r2, host_unpack=host_unpack, n_threads=threads)
            out = os.path.join(tmp, name)
            t1 = time.perf_counter()
            assert ld.Compute(st, twk, out)
            wall = time.perf_counter() - t1
            x = ld.last_stats
            row = {"arrangement": name, "threads": threads, "wall_s": wall, "read_s": x.seconds_file_read, "load_s": x.seconds_file_load,
                   "compute_and_write_s": x.seconds_file_total - x.seconds_file_read - x.seconds_file_load,
                   "decode_kernel_ms": x.ms_decode_kernel, "count_kernel_ms": x.ms_count_kernel, "stats_kernel_ms": x.ms_stats_kernel,
                   "h2d_bytes": int(x.bytes_h2d), "records": int(x.records_out), "pairs_per_s_file_to_file": pairs / wall}
            if best is None or wall < best["wall_s"]:
                best = row
        outs[name] = tf.canonical(tf.read_two(os.path.join(tmp, name + ".two")), forward_only=False)
        best.update(workload=f"{a.samples} samples x {a.variants} SNVs, -p -r {a.min_r2}", twk_bytes=os.path.getsize(twk),
                    two_bytes=os.path.getsize(os.path.join(tmp, name + ".two")), gen_s=round(gen, 1))
        print(json.dumps(best), flush=True)
    names = list(outs)
    same = all(np.array_equal(outs[names[0]].view(np.uint8), outs[n].view(np.uint8)) for n in names[1:])
    print(json.dumps({"outputs_identical": bool(same), "records_fwd_plus_rev": int(len(outs[names[0]]))}), flush=True)
    if a.sorted:
        # queryable (sorted, indexed) output: records ordered on the device vs calc followed by the host sorter (and the reference's sort)
        import subprocess
        from oracle import ldcore as lc
        best = None
        for rep in range(3):
            ld = tb.twk_ld()
            st = tb.default_settings(force_phased=1, minR2=a.min_r2, n_threads=a.threads, sorted_output=1)
            t1 = time.perf_counter()
            assert ld.Compute(st, twk, os.path.join(tmp, "device_sorted"))
            wall = time.perf_counter() - t1
            best = wall if best is None else min(best, wall)
        plain = os.path.join(tmp, "device_decode.two")
        t1 = time.perf_counter()
        n = tb.sort_two(plain, os.path.join(tmp, "host_sorted.two"), c_level=1, n_threads=a.threads)
        t_host = time.perf_counter() - t1
        same = np.array_equal(tf.read_two(os.path.join(tmp, "device_sorted.two")).view(np.uint8), tf.read_two(os.path.join(tmp, "host_sorted.two")).view(np.uint8))
        row = {"arrangement": "sorted output", "records": int(n), "device_sort_file_to_file_s": best, "host_sorter_alone_s": t_host,
               "identical_to_calc_then_sort": bool(same), "threads": a.threads}
        if os.path.exists(lc.REF_SORT):
            t1 = time.perf_counter()
            r = subprocess.run([lc.REF_SORT, "sort", "-i", plain, "-o", os.path.join(tmp, "ref_sorted.two"), "-t", str(a.threads)], capture_output=True, text=True)
            row["reference_sort_alone_s"] = time.perf_counter() - t1 if r.returncode == 0 else None
        print(json.dumps(row), flush=True)
    if a.reference:
        from oracle import ldcore as lc
        sub = tf.synth_genotypes(a.samples, a.reference, seed=a.seed, rare_fraction=a.rare_fraction)
        rtwk = os.path.join(tmp, "ref.twk")
        tf.write_twk(rtwk, sub)
        t1 = time.perf_counter()
        info = lc.run_reference_calc(rtwk, os.path.join(tmp, "ref_out"), ["-p", "-r", str(a.min_r2)], threads=a.threads)
        wall = time.perf_counter() - t1
        rp = a.reference * (a.reference - 1) // 2
        print(json.dumps({"arrangement": "reference calc (CPU)", "threads": a.threads, "variants": a.reference, "wall_s": wall,
                          "pairs_per_s_file_to_file": rp / wall, "pairs_per_s_compute_phase": info.get("pairs_per_s")}), flush=True)
    for fn in os.listdir(tmp):
        os.unlink(os.path.join(tmp, fn))
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
