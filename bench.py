#!/usr/bin/env python
"""bench.py -- throughput of the B200 `tomahawk calc` path (variant-pairs/s).

A "step" is one complete pass of the LD hot path over the synthetic genotype matrix: count kernel (+ fused R2
screen / compaction) and statistics kernel for every pair of this rank's share of the tile grid.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--variants M] [--samples S]
                  [--min-r2 R] [--kernel auto|popc|umma|fp4] [--unphased] [--missing F] [--no-extra]

Primary workload. N = 1: BASELINE.json configs[1] -- phased all-pairs, 2,504 samples (5,008 haplotypes) x
200,000 SNVs, R2 >= 0.1. N > 1 (torchrun, one rank per GPU): weak scaling, M = 200,000 * sqrt(N) variants so
every GPU keeps the N = 1 pair count. The ranks form an NCCL communicator INSIDE libtwkb (twkb_comm_init);
every load sends 1/N of the rows over the rank's own PCIe link and completes the matrix with one ncclAllGather
over NVLink (twkb_load_matrix_sliced); tiles are then computed with no further collective.

Secondary workloads in `extra_configs` of the same JSON line (each with its own value / e2e / roofline):
  N = 1: configs[0] (R2 >= 0: every pair goes through Fisher) and configs[2] (unphased, 5 % missing);
  N > 1: configs[3] (1,000,000 haplotypes x 100,000 SNVs) strong-scaled over the N GPUs, generated on the device;
         a window-banded slice of configs[4] (1,000,000 haplotypes, -w 500kb, 80 % rare), position-sharded with halo.

One JSON line is printed by rank 0 (README / DESIGN.md section 7 describe every key).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "variant-pairs/s"
UNIT = "pairs/s"
BASE_SAMPLES = 2504
BASE_VARIANTS = 200_000
REF_SAMPLE_VARIANTS = 30_000  # bounded CPU sample of the same workload: its FIRST M' variants
NOMINAL_TFLOPS = {"fp4": 9000.0, "i8": 4500.0}  # dense tcgen05 rates, B200_PROFILING.md (bf16 2250 x 4 / x 2)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variants", type=int, default=BASE_VARIANTS)
    ap.add_argument("--samples", type=int, default=BASE_SAMPLES)
    ap.add_argument("--min-r2", type=float, default=0.1)
    ap.add_argument("--kernel", default="auto", choices=["auto", "popc", "umma", "i8", "fp4"])
    ap.add_argument("--seed", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-variants", type=int, default=REF_SAMPLE_VARIANTS)
    ap.add_argument("--unphased", action="store_true", help="-u: 3x3 genotype tables (BASELINE configs[2] with --missing 0.05)")
    ap.add_argument("--missing", type=float, default=0.0, help="per-genotype missing rate of the synthetic data")
    ap.add_argument("--no-mma-ceiling", action="store_true", help="skip the MMA-only ceiling pass of the roofline block")
    ap.add_argument("--no-extra", action="store_true", help="primary workload only (no extra_configs, no peak probes)")
    ap.add_argument("--biobank-variants", type=int, default=100_000, help="SNVs of the configs[3] run at N > 1")
    ap.add_argument("--c5-variants-per-gpu", type=int, default=50_000, help="own SNVs per GPU of the configs[4] window slice (N > 1)")
    ap.add_argument("--c5", action="store_true", help="N = 1: run the configs[4] window slice (one shard) instead of configs[0] / configs[2]")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------- synthetic workload (numpy, N = 1 and the reference arm)
def primary_synth(args, n_variants):
    """The N = 1 workload and, through its first rows, the reference arm's bounded sample: ONE draw."""
    from tomahawk_b200 import synth

    return synth.synth_genotypes(args.samples, n_variants, seed=args.seed, missing_rate=args.missing)


def head_of(s, n):
    from tomahawk_b200 import synth

    return synth.Synth(alleles=s.alleles[:n], pos=s.pos[:n], rid=s.rid[:n], n_samples=s.n_samples)


# --------------------------------------------------------------------- reference arm
def reference_run(sub, flags, steps, warmup, what, keep_two=False):
    """Times the reference's own `tomahawk calc` (oracle/_ref/tomahawk_calc, built from /root/reference by
    oracle/build_ref.sh) on this host's cores on the synthetic genotypes `sub`. One of the two places bench.py
    executes anything under oracle/ (the other is the parity comparison of the very records this run wrote).
    Returns (info dict, records or None)."""
    from oracle import ldcore as lc
    from oracle import twk_format as tf

    cores = os.cpu_count() or 1
    n_variants = sub.n_variants
    tmp = tempfile.mkdtemp(prefix="twkb_ref_")
    twk = os.path.join(tmp, "ref.twk")
    tf.write_twk(twk, sub)
    pairs = n_variants * (n_variants - 1) // 2
    sample = f"{what} ({pairs} pairs), same N, {' '.join(flags)} -t {cores}"
    recs = None
    if lc.have_reference():
        kind = "reference"
        rates, times = [], []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            info = lc.run_reference_calc(twk, os.path.join(tmp, "ref_out"), list(flags), threads=cores)
            dt = time.perf_counter() - t0
            if it >= warmup:
                rates.append(info.get("pairs_per_s", pairs / dt))
                times.append(dt)
        value = float(np.mean(rates))
        ms = float(np.mean(times)) * 1e3
        if keep_two:
            recs = tf.read_two(os.path.join(tmp, "ref_out.two"))
    else:
        kind = "port"
        cores = 1
        prm_kw = {"minR2": float(flags[flags.index("-r") + 1])} if "-r" in flags else {}
        prm_kw.update({"forced_unphased": 1} if "-u" in flags else {"force_phased": 1})
        t0 = time.perf_counter()
        got, _ = lc.calc(sub, lc.default_params(**prm_kw), cap=max(1 << 20, pairs // 4))
        dt = time.perf_counter() - t0
        value, ms = pairs / dt, dt * 1e3
        if keep_two:
            recs = got
    for fn in os.listdir(tmp):
        os.unlink(os.path.join(tmp, fn))
    os.rmdir(tmp)
    return {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "ms": ms, "pairs": pairs}, recs


def parity_block(ref_recs, gpu_recs, unphased):
    """The reference's records against the GPU's on the same variants: identical key sets; counts, flags, D, D', R, R2 and
    chi-squared bit-equal for phased math; P relative. (tests/ hold the full bars; this is the in-bench statement.)"""
    from oracle import twk_format as tf

    ref = tf.canonical(ref_recs, forward_only=True)
    got = tf.canonical(gpu_recs, forward_only=True)

    def keys(r):
        return set(zip(r["ridA"].tolist(), (r["packA"] >> 2).tolist(), r["ridB"].tolist(), (r["packB"] >> 2).tolist()))

    sa, sb = keys(ref), keys(got)
    out = {"checked": True, "n_ref": int(len(ref)), "n_gpu": int(len(got)), "only_ref": len(sa - sb), "only_gpu": len(sb - sa)}
    if out["only_ref"] == 0 and out["only_gpu"] == 0 and len(ref) == len(got):
        phased = (ref["controller"] & 1) == 1
        fields = ("controller", "cnt", "D", "Dprime", "R", "R2", "ChiSqFisher")
        out["bit_equal_fields"] = [f for f in fields if np.array_equal(got[f][phased], ref[f][phased])]
        out["fields_checked"] = list(fields)
        out["records_phased_math"] = int(phased.sum())
        big = ref["P"] > 1e-290
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = np.abs(got["P"][big] - ref["P"][big]) / ref["P"][big]
        out["max_rel_P"] = float(rel.max()) if rel.size else 0.0
        out["P_underflow_both_tiny"] = bool(np.all(got["P"][~big] <= 1e-289))
        if unphased and (~phased).any():
            u = ~phased
            with np.errstate(divide="ignore", invalid="ignore"):
                out["max_rel_R2_cubic"] = float(np.max(np.abs(got["R2"][u] - ref["R2"][u]) / np.maximum(ref["R2"][u], 1e-300)))
    return out


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line. Libraries that write to file descriptor 1 behind Python's back
    (NCCL prints its version banner there, whatever NCCL_DEBUG_FILE says) are sent to stderr: fd 1 is
    re-pointed at fd 2 for the whole run and the JSON line is written to a saved duplicate of the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def flop_per_pair(n_samples, unphased, missing):
    """GEMM view of SURVEY.md 8d: 2N bit-MACs per phased pair (x4 masked counts with missing data); 4 / 9 plane
    products over N samples for unphased tables."""
    if unphased:
        return 2.0 * (9 if missing else 4) * n_samples
    return 2.0 * 2 * n_samples * (4 if missing else 1)


class Bench:
    """One workload on one rank: resident steps, end-to-end steps."""

    def __init__(self, torch, tb, dist, rank, world, local_rank, flush):
        self.torch, self.tb, self.dist = torch, tb, dist
        self.rank, self.world, self.local_rank, self.flush = rank, world, local_rank, flush

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allreduce(self, vals, op):
        if self.dist is None:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return [float(x) for x in t]

    def resident(self, eng, steps, warmup):
        torch = self.torch

        def one():
            self.flush.zero_()  # L2 flush between iterations (256 MiB > 126 MB L2)
            torch.cuda.synchronize()
            eng.compute_resident()
            return eng.stats()

        for _ in range(warmup):
            one()
        self.barrier()
        t0 = time.perf_counter()
        acc = {"dev": [], "cnt": [], "sts": [], "sp": [], "launches": 0, "cnt_launches": 0}
        st = None
        for _ in range(steps):
            st = one()
            acc["dev"].append(st.ms_device_total); acc["cnt"].append(st.ms_count_kernel); acc["sts"].append(st.ms_stats_kernel)
            acc["sp"].append(st.ms_sparse_kernel)
            acc["launches"] += st.count_launches + st.stats_launches + st.other_launches + st.sparse_launches
            acc["cnt_launches"] += st.count_launches
        self.barrier()
        acc["wall"] = time.perf_counter() - t0
        step_ms = float(np.mean(acc["dev"]))
        step_ms_max, = self.allreduce([step_ms], "MAX")
        pairs_total, = self.allreduce([float(st.pairs_visited)], "SUM")
        acc.update(st=st, step_ms=step_ms_max, pairs_total=pairs_total, pairs_rank=float(st.pairs_visited),
                   value=pairs_total / (step_ms_max * 1e-3))
        return acc

    def e2e(self, eng, load_fn, reps):
        """host buffers -> C-ABI -> records back on the host; copies inside the timed region."""
        torch = self.torch
        ms, parts, s2 = [], None, None
        for it in range(1 + reps):
            self.flush.zero_()
            self.barrier()
            t1 = time.perf_counter()
            load_fn()                                # H2D from pinned memory (+ NCCL exchange at N > 1) + device layout
            t_load = time.perf_counter() - t1
            eng.compute_discard()                    # compute + D2H of every record into pinned staging
            torch.cuda.synchronize()
            dt = time.perf_counter() - t1
            s2 = eng.stats()
            if it > 0:
                ms.append(dt * 1e3)
                parts = {"load_wall_ms": t_load * 1e3, "load_device_ms": s2.ms_h2d, "compute_wall_ms": (dt - t_load) * 1e3,
                         "compute_device_ms": s2.ms_device_total}
        e2e_ms, = self.allreduce([float(np.mean(ms))], "MAX")
        h2d, d2h, recs = self.allreduce([float(s2.bytes_h2d), float(s2.bytes_d2h), float(s2.records_out)], "SUM")
        return {"ms": e2e_ms, "parts": parts, "h2d": int(h2d), "d2h": int(d2h), "records": int(recs),
                "launches": int(s2.count_launches + s2.stats_launches + s2.other_launches + s2.sparse_launches)}


def tensor_roofline(tb, acc, n_samples, unphased, missing, peaks, peak_src, fp4_measured, mma_only=None, traffic=None):
    st = acc["st"]
    fp4 = st.kernel_used == tb.KERNEL_UMMA_FP4
    fpp = flop_per_pair(n_samples, unphased, missing)
    n_launch = max(acc["cnt_launches"], 1)
    avg_launch_s = (sum(acc["cnt"]) / n_launch) * 1e-3
    pairs_per_launch = acc["pairs_rank"] * len(acc["dev"]) / n_launch
    if st.sparse_variants:
        # the list kernel served the pairs with a rare member: credit the tensor kernel with the MACs it issued for the
        # dense x dense triangle (tile padding included -- the only figure available without the dense pair count)
        achieved = 2.0 * (st.mma_macs / max(1, st.count_launches)) / avg_launch_s / 1e12
        note_pairs = "dense x dense tiles only (MACs issued, tile padding included); pairs with a rare member run on the list kernel"
    else:
        achieved = pairs_per_launch * fpp / avg_launch_s / 1e12
        note_pairs = f"algorithmic work = pairs x {fpp:g} flop (SURVEY 8d; K padding and tile edges are not counted)"
    peak = NOMINAL_TFLOPS["fp4" if fp4 else "i8"]
    ratio = 4.0 if fp4 else 2.0
    out = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
           "kernel": "count_umma3_kernel<%s,%s>" % ("e2m1" if fp4 else "int8", "planes" if (unphased or missing) else "1 plane"),
           "peak_source": f"nominal dense {'e2m1 (kind::mxf4)' if fp4 else 'int8 (kind::i8)'} tcgen05 rate; MEASURED_PEAKS.json ({peak_src}) has bf16 only",
           "frac_of_scaled_measured_bf16": achieved / (ratio * peaks["bf16_tflops_sustained"]),
           "scaled_measured_bf16_peak": ratio * peaks["bf16_tflops_sustained"],
           "ms_per_launch": avg_launch_s * 1e3, "note": note_pairs}
    if fp4 and fp4_measured:
        out["measured_fp4_gemm_tflops"] = fp4_measured
        out["frac_of_measured_fp4_gemm_sustained"] = achieved / fp4_measured["sustained"]
    if mma_only:
        out["mma_only_ceiling"] = mma_only
        out["frac_of_mma_only_ceiling"] = achieved / mma_only
    return out


def main():
    args = parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = max(args.gpus, 1)

    mode_flag = "-u" if args.unphased else "-p"
    mode_name = "-u (unphased 3x3)" if args.unphased else "-p"
    miss_name = f", {100 * args.missing:g}% missing genotypes" if args.missing > 0 else ""
    workload = (f"tomahawk calc {mode_name} all-pairs, synthetic {args.samples} samples ({2 * args.samples} haplotypes) x "
                f"{args.variants} SNVs{miss_name}, R2>={args.min_r2}")
    ref_flags = [mode_flag, "-r", str(args.min_r2)]

    if args.impl == "reference":
        if rank != 0:
            return 0
        nv = min(args.ref_variants, args.variants)
        s = primary_synth(args, args.variants)
        r, _ = reference_run(head_of(s, nv), ref_flags, args.steps, args.warmup, f"first {nv} of the workload's {args.variants} variants")
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64 popcount + f64", "data": "synthetic",
            "config": {"workload": workload, "reference_sample": r["sample"]},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        emit_json(line)
        return 0

    import torch
    import tomahawk_b200 as tb
    from tomahawk_b200 import synth, tools

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: tomahawk_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    uid = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        box = [tb.comm_unique_id() if rank == 0 else None]      # torch.distributed only carries the 128-byte id
        dist.broadcast_object_list(box, src=0)
        uid = box[0]

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    B = Bench(torch, tb, dist, rank, world, local_rank, flush)
    peaks, peak_src = measured_peaks()
    kernel = {"auto": tb.KERNEL_AUTO, "popc": tb.KERNEL_POPC, "umma": tb.KERNEL_UMMA, "i8": tb.KERNEL_UMMA,
              "fp4": tb.KERNEL_UMMA_FP4}[args.kernel]

    # ---- workload: weak scaling keeps pairs per GPU constant => M grows with sqrt(N)
    n_variants = int(round(args.variants * math.sqrt(world)))
    n_samples = args.samples
    stride = synth.words_per_variant(n_samples)
    t_gen0 = time.perf_counter()
    s_full = None
    if world == 1:
        # numpy generator: the very draw whose first rows the reference arm (and the parity block) use
        s_full = primary_synth(args, n_variants)
        data, mask = synth.pack_bits(s_full)
        meta = synth.variant_meta(s_full)
        host = torch.from_numpy(data.view(np.int64)).pin_memory()
        host_mask = torch.from_numpy(mask.view(np.int64)).pin_memory() if mask is not None else None
        del data, mask
        data_name = "synthetic (numpy generator, seed %d)" % args.seed
    else:
        # device generator: the stream is keyed on the global variant index, so every rank produces the same
        # matrix without a broadcast; a rank keeps ONLY its slice of the rows in (pinned) host memory
        d_full, m_full, meta = tools.synth_device(n_samples, n_variants, seed=args.seed, missing_rate=args.missing)
        b_row, e_row = tb.comm_slice(n_variants, rank, world)
        host = torch.empty((e_row - b_row, stride), dtype=torch.int64).pin_memory()
        host.copy_(d_full[b_row:e_row])
        host_mask = None
        if m_full is not None:
            host_mask = torch.empty((e_row - b_row, stride), dtype=torch.int64).pin_memory()
            host_mask.copy_(m_full[b_row:e_row])
        del d_full, m_full
        torch.cuda.empty_cache()
        data_name = "synthetic (device generator libtwkb_tools, seed %d)" % args.seed
    t_gen = time.perf_counter() - t_gen0
    host_np = host.numpy().view(np.uint64)
    host_mask_np = host_mask.numpy().view(np.uint64) if host_mask is not None else None

    eng = tb.Engine(force_phased=0 if args.unphased else 1, forced_unphased=1 if args.unphased else 0, minR2=args.min_r2,
                    kernel=kernel, device=local_rank, part_index=rank, part_count=world)
    if world > 1:
        eng.comm_init(uid, rank, world)

    def load_primary():
        if world > 1:
            eng.load_sliced(n_samples, n_variants, host_np, host_mask_np, meta)
        else:
            eng.load(n_samples, host_np, host_mask_np, meta)

    load_primary()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    acc = B.resident(eng, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    e2e = B.e2e(eng, load_primary, max(2, min(args.steps, 3)))
    st = acc["st"]
    value, step_ms_max, pairs_total = acc["value"], acc["step_ms"], acc["pairs_total"]
    e2e_value = pairs_total / (e2e["ms"] * 1e-3)
    tensor = st.kernel_used in (tb.KERNEL_UMMA, tb.KERNEL_UMMA_FP4)
    fp4 = st.kernel_used == tb.KERNEL_UMMA_FP4
    H = 2 * n_samples

    # ---- measured denominators (rank 0, N = 1): cuBLASLt block-scaled e2m1 GEMM and the POPC issue rate
    fp4_measured, popc_measured = None, None
    if rank == 0 and world == 1 and not args.no_extra:
        try:
            b, s_ = tools.fp4_gemm_tflops(8192, 2.0)
            fp4_measured = {"burst": b, "sustained": s_,
                            "how": "cuBLASLt NVFP4 (e2m1, 16-element UE4M3 block scales) GEMM 8192^3, best of 10 / 2 s back to back, same run"}
        except Exception as e:
            sys.stderr.write(f"[bench] fp4 gemm probe unavailable: {e}\n")
        try:
            r_, per, mhz = tools.popc_rate()
            popc_measured = {"popc_per_s": r_, "popc_per_clk_per_sm_at_nominal_clock": per, "nominal_sm_mhz": mhz}
        except Exception as e:
            sys.stderr.write(f"[bench] popc probe unavailable: {e}\n")

    # ---- roofline of the dominant (count) kernel
    roofline = None
    if rank == 0:
        if tensor:
            mma_only = None
            if (world == 1 and not args.no_mma_ceiling and not args.unphased and args.missing == 0 and os.path.exists(tb.PROF_LIB_PATH)):
                # the same launch with operand traffic and epilogue switched off, in the profiling build of the same sources
                # (libtwkb_prof.so; the product library has no such switch): what the tcgen05 pipe delivers for this
                # instruction stream on this board
                os.environ["TWKB_DEBUG_FLAGS"] = "3"
                try:
                    peng = tb.Engine(force_phased=1, minR2=args.min_r2, kernel=kernel, device=local_rank, profiling=True)
                    peng.load(n_samples, host_np, host_mask_np, meta)
                    ms = []
                    for _ in range(3):
                        flush.zero_(); torch.cuda.synchronize()
                        peng.compute_resident()
                        s3 = peng.stats()
                        ms.append(s3.ms_count_kernel / max(s3.count_launches, 1))
                    peng.close()
                    fpp = flop_per_pair(n_samples, args.unphased, args.missing > 0)
                    mma_only = (acc["pairs_rank"] * len(acc["dev"]) / max(acc["cnt_launches"], 1)) * fpp / (min(ms[1:]) * 1e-3) / 1e12
                except Exception as e:
                    sys.stderr.write(f"[bench] MMA-only ceiling skipped: {e}\n")
                finally:
                    del os.environ["TWKB_DEBUG_FLAGS"]
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "round2_traffic.json")
            if os.path.exists(tpath):  # dram bytes per launch from the committed ncu --set full capture of THIS configuration
                key = f"{n_samples}x{n_variants}:{'u' if args.unphased else 'p'}:{args.missing:g}:{args.min_r2:g}:n{world}"
                traffic = json.load(open(tpath)).get(key)
            roofline = tensor_roofline(tb, acc, n_samples, args.unphased, args.missing > 0, peaks, peak_src, fp4_measured, mma_only, traffic)
        else:
            w = math.ceil(H / 32)
            n_launch = max(acc["cnt_launches"], 1)
            avg_launch_s = (sum(acc["cnt"]) / n_launch) * 1e-3
            achieved_wops = (acc["pairs_rank"] * len(acc["dev"]) / n_launch) * w / avg_launch_s
            smax = (clocks or {}).get("sm_max_mhz") or 1965.0
            peak_wops = popc_measured["popc_per_s"] if popc_measured else 16 * 148 * smax * 1e6
            roofline = {"bound": "int_pipe_popc", "achieved": achieved_wops / 1e12, "peak": peak_wops / 1e12,
                        "unit": "Tword-op/s", "frac": achieved_wops / peak_wops, "traffic": None, "kernel": "count_popc_kernel<0>",
                        "peak_source": "measured POPC issue rate (libtwkb_tools popc_rate_kernel)" if popc_measured else
                                       "nominal 16 POPC/clk/SM x 148 SM x max SM clock",
                        "note": f"INT-pipe roofline; {w} AND+POPC word-ops per pair"}
        if popc_measured:
            roofline["measured_popc_rate"] = popc_measured

    # ---- parity: the reference's own records on the first M' variants against the GPU's on the same rows
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            nv = min(args.ref_variants, n_variants)
            r, ref_recs = reference_run(head_of(s_full, nv), ref_flags, 1, 0, f"first {nv} of the workload's {n_variants} variants", keep_two=True)
            cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
            peng = tb.Engine(force_phased=0 if args.unphased else 1, forced_unphased=1 if args.unphased else 0, minR2=args.min_r2,
                             kernel=kernel, device=local_rank)
            peng.load(n_samples, host_np[:nv], host_mask_np[:nv] if host_mask_np is not None else None, meta[:nv])
            got = peng.compute()
            peng.close()
            parity = parity_block(ref_recs, got, args.unphased)
            parity["sample"] = r["sample"]
            parity["against"] = "the reference binary's .two (oracle/_ref/tomahawk_calc)" if r["kind"] == "reference" else "the oracle port"
        except Exception as e:  # the baseline is a reported number, never a gate
            cpu_baseline = cpu_baseline or {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": str(e)[:200]}
            parity = parity or {"checked": False, "why": str(e)[:200]}
    del s_full

    # ---- secondary workloads
    extra = []
    if not args.no_extra and not args.unphased and args.missing == 0 and args.variants == BASE_VARIANTS and args.samples == BASE_SAMPLES:
        del host, host_np
        if world == 1:
            eng.close()
            torch.cuda.empty_cache()
            if not args.c5:
                extra.append(extra_config0(args, B, tb, synth, peaks, peak_src, fp4_measured))
                extra.append(extra_config2(args, B, tb, tools, peaks, peak_src, fp4_measured))
        else:
            # same context and communicator (an NCCL unique id serves ONE ncclCommInitRank): only the matrix changes
            extra.append(extra_config3(args, B, tb, tools, eng, peaks, peak_src))
        if world > 1 or args.c5:
            extra.append(extra_config4(args, B, tb, tools, peaks, peak_src))
        extra = [x for x in extra if x]

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": ("e2m1 x e2m1 -> f32 exact (tcgen05 kind::mxf4) + f64 statistics" if fp4 else
                  "int8 x int8 -> int32 (tcgen05) + f64 statistics" if tensor else "u32 popcount + f64 statistics"),
        "data": data_name,
        "config": {
            "workload": (f"tomahawk calc {mode_name} all-pairs, synthetic {n_samples} samples ({H} haplotypes) x {n_variants} SNVs"
                         f"{miss_name}, R2>={args.min_r2}" + (f" (weak scaling: {args.variants} x sqrt({world}) variants)" if world > 1 else "")),
            "baseline_config": baseline_config_name(args),
            "pairs_per_step": pairs_total, "haplotype_cmp_per_s": value * H, "records_per_step": e2e["records"],
            "kernel": "umma_fp4" if fp4 else "umma_i8" if tensor else "popc",
            "l2": "256 MiB device memset between steps (flush) and operands > L2",
            "seed": args.seed, "gen_seconds": round(t_gen, 2),
            "ms_count_kernel_per_step": float(np.mean(acc["cnt"])), "ms_stats_kernel_per_step": float(np.mean(acc["sts"])),
            "wall_seconds_timed_region": acc["wall"],
            "multi_gpu": (f"{world} ranks, NCCL communicator inside libtwkb (twkb_comm_init); each load: own row slice over PCIe + "
                          f"one in-place ncclAllGather; tiles dealt by part_index/part_count, no collective during compute") if world > 1 else None,
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                "ms_per_step": e2e["ms"], "gpu_launches_per_step": e2e["launches"], "parts_last_step": e2e["parts"],
                "path": ("every rank: twkb_load_matrix_sliced (its 1/N of the rows from pinned host memory, NCCL exchange) -> twkb_compute with a host sink"
                         if world > 1 else "twkb_load_matrix from pinned host memory -> twkb_compute with a host sink (record drain thread)")},
        "gpu_launches": int(acc["launches"]),
        "clocks": clocks,
        "roofline": roofline,
    }
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    if parity is not None:
        line["parity"] = parity
    if extra:
        line["extra_configs"] = extra
    emit_json(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def baseline_config_name(args):
    if args.unphased:
        return "BASELINE.json configs[2]" if (args.samples == 10000 and args.variants == 100000) else "configs[2] shape (unphased), other size"
    if args.samples == BASE_SAMPLES and args.variants == BASE_VARIANTS and args.min_r2 == 0.1:
        return "BASELINE.json configs[1]"
    if args.samples == BASE_SAMPLES and args.variants == 10000 and args.min_r2 == 0:
        return "BASELINE.json configs[0]"
    return "phased all-pairs, other size"


# ------------------------------------------------------------------------------- extra_configs
def _extra_entry(name, workload, acc, e2e, roofline, H, cpu=None, parity=None, **more):
    out = {"baseline_config": name, "workload": workload, "value": acc["value"], "unit": UNIT, "ms_per_step": acc["step_ms"],
           "steps": len(acc["dev"]), "pairs_per_step": acc["pairs_total"], "haplotype_cmp_per_s": acc["value"] * H,
           "ms_count_kernel_per_step": float(np.mean(acc["cnt"])), "ms_stats_kernel_per_step": float(np.mean(acc["sts"])),
           "ms_sparse_kernel_per_step": float(np.mean(acc["sp"])), "gpu_launches": int(acc["launches"]), "roofline": roofline}
    if e2e is not None:
        out["e2e"] = {"value": acc["pairs_total"] / (e2e["ms"] * 1e-3), "unit": UNIT, "ms_per_step": e2e["ms"], "h2d_bytes_per_step": e2e["h2d"],
                      "d2h_bytes_per_step": e2e["d2h"], "records_per_step": e2e["records"], "parts_last_step": e2e["parts"]}
    if cpu is not None:
        out["cpu_baseline"] = cpu
    if parity is not None:
        out["parity"] = parity
    out.update(more)
    return out


def extra_config0(args, B, tb, synth, peaks, peak_src, fp4_measured):
    """BASELINE configs[0]: 2,504 samples x 10,000 SNVs, -p, R2 >= 0 -- every pair is a record and goes through Fisher."""
    try:
        torch = B.torch
        n, m = BASE_SAMPLES, 10_000
        s = synth.synth_genotypes(n, m, seed=args.seed)
        data, _ = synth.pack_bits(s)
        meta = synth.variant_meta(s)
        host = torch.from_numpy(data.view(np.int64)).pin_memory()
        hn = host.numpy().view(np.uint64)
        eng = tb.Engine(force_phased=1, minR2=0.0, device=B.local_rank)
        eng.load(n, hn, None, meta)
        acc = B.resident(eng, 5, 3)
        e2e = B.e2e(eng, lambda: eng.load(n, hn, None, meta), 2)
        roof = tensor_roofline(tb, acc, n, False, False, peaks, peak_src, fp4_measured)
        # the step is bound by the statistics kernel (Fisher for every pair) and, end to end, by PCIe
        st_ms = float(np.mean(acc["sts"]))
        roof["dominant_kernel"] = {"kernel": "stats_kernel<phased>", "ms_per_step": st_ms, "candidates_per_step": int(acc["st"].pairs_screened),
                                   "ns_per_fisher_test": st_ms * 1e6 / max(1, int(acc["st"].pairs_screened)),
                                   "note": "fp64 pipe + instruction issue; profiles/round2_ncu_stats_c1_*.csv"}
        pcie = e2e["d2h"] / (e2e["ms"] * 1e-3) / 1e9
        cpu = parity = None
        if not args.no_cpu_baseline:
            nv = 2000
            r, ref_recs = reference_run(head_of(s, nv), ["-p", "-r", "0"], 1, 0, f"first {nv} of the workload's {m} variants", keep_two=True)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
            pe = tb.Engine(force_phased=1, minR2=0.0, device=B.local_rank)
            pe.load(n, hn[:nv], None, meta[:nv])
            parity = parity_block(ref_recs, pe.compute(), False)
            parity["sample"] = r["sample"]
            pe.close()
        eng.close()
        return _extra_entry("BASELINE.json configs[0]", f"tomahawk calc -p all-pairs, synthetic {n} samples x {m} SNVs, R2>=0 (every pair a record)",
                            acc, e2e, roof, 2 * n, cpu, parity, e2e_record_gbps=pcie)
    except Exception as e:
        return {"baseline_config": "BASELINE.json configs[0]", "error": str(e)[:300]}


def extra_config2(args, B, tb, tools, peaks, peak_src, fp4_measured):
    """BASELINE configs[2]: 10,000 samples x 100,000 SNVs, -u, 5 % missing genotypes (device generator: numpy needs minutes)."""
    try:
        torch = B.torch
        n, m, miss = 10_000, 100_000, 0.05
        d, mk, meta = tools.synth_device(n, m, seed=args.seed, missing_rate=miss)
        host = torch.empty(d.shape, dtype=torch.int64).pin_memory(); host.copy_(d)
        hostm = torch.empty(mk.shape, dtype=torch.int64).pin_memory(); hostm.copy_(mk)
        del d, mk
        torch.cuda.empty_cache()
        hn, hm = host.numpy().view(np.uint64), hostm.numpy().view(np.uint64)
        eng = tb.Engine(forced_unphased=1, minR2=0.1, device=B.local_rank)
        eng.load(n, hn, hm, meta)
        acc = B.resident(eng, 5, 3)
        e2e = B.e2e(eng, lambda: eng.load(n, hn, hm, meta), 2)
        roof = tensor_roofline(tb, acc, n, True, True, peaks, peak_src, fp4_measured)
        cpu = parity = None
        if not args.no_cpu_baseline:
            from oracle import twk_format as tf
            nv = 1000
            al = tools.rows_to_alleles(hn[:nv], hm[:nv], n)
            sub = tf.Synth(alleles=al, pos=meta["pos"][:nv].copy(), rid=meta["rid"][:nv].copy(), n_samples=n)
            r, ref_recs = reference_run(sub, ["-u", "-r", "0.1"], 1, 0, f"first {nv} of the workload's {m} variants", keep_two=True)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
            pe = tb.Engine(forced_unphased=1, minR2=0.1, device=B.local_rank)
            pe.load(n, hn[:nv], hm[:nv], meta[:nv])
            parity = parity_block(ref_recs, pe.compute(), True)
            parity["sample"] = r["sample"]
            parity["note"] = "unphased pass/fail flips are confined to decision boundaries (enumerated in tests/test_gpu_parity.py)"
            pe.close()
        eng.close()
        return _extra_entry("BASELINE.json configs[2]", f"tomahawk calc -u all-pairs, synthetic {n} samples x {m} SNVs, 5% missing genotypes, R2>=0.1, Fisher + chi2",
                            acc, e2e, roof, 2 * n, cpu, parity)
    except Exception as e:
        return {"baseline_config": "BASELINE.json configs[2]", "error": str(e)[:300]}


def extra_config3(args, B, tb, tools, eng, peaks, peak_src):
    """BASELINE configs[3]: 500,000 samples (1,000,000 haplotypes) x 100,000 SNVs, -p, R2 >= 0.1, STRONG-scaled over the N GPUs
    (every GPU holds the whole 12.5 GB matrix; the tile grid is dealt to the ranks). Generated on the device."""
    try:
        torch = B.torch
        n, m = 500_000, args.biobank_variants
        rank, world = B.rank, B.world
        stride = tools.words_per_variant(n)
        t0 = time.perf_counter()
        b, e = tb.comm_slice(m, rank, world)
        # this rank's slice only: generated on the device, kept in pinned host memory (the e2e "host buffer")
        d, _, meta_slice = tools.synth_device(n, m, seed=args.seed, first=b, n_rows=e - b)
        host = torch.empty(d.shape, dtype=torch.int64).pin_memory(); host.copy_(d)
        del d
        torch.cuda.empty_cache()
        # metadata of all variants: every rank computed its slice's allele counts; gathered through torch.distributed
        parts = [None] * world
        B.dist.all_gather_object(parts, meta_slice.tobytes())
        meta = np.concatenate([np.frombuffer(p, dtype=tb.VARIANT_DTYPE) for p in parts])
        t_gen = time.perf_counter() - t0
        hn = host.numpy().view(np.uint64)
        load = lambda: eng.load_sliced(n, m, hn, None, meta)
        load()
        acc = B.resident(eng, 2, 1)
        e2e = B.e2e(eng, load, 1)
        st = acc["st"]
        roof = tensor_roofline(tb, acc, n, False, False, peaks, peak_src, None)
        if st.sparse_variants:
            sp_ms = float(np.mean(acc["sp"]))
            roof["list_kernel"] = {"kernel": "count_sparse_kernel", "variants": int(st.sparse_variants), "ms_per_step": sp_ms,
                                   "word_ops_per_step": int(st.sparse_word_ops), "bound": "hbm",
                                   "achieved_gbs": st.sparse_word_ops * 4 / (sp_ms * 1e-3) / 1e9 if sp_ms > 0 else None,
                                   "peak_gbs": peaks["hbm_gbs"]}
        out = _extra_entry("BASELINE.json configs[3]",
                           f"tomahawk calc -p all-pairs, synthetic {n} samples ({2 * n} haplotypes) x {m} SNVs, R2>=0.1, strong scaling over {world} GPUs",
                           acc, e2e, roof, 2 * n, scaling="strong", n_gpus=world, gen_seconds=round(t_gen, 2),
                           matrix_bytes=int(m) * stride * 8)
        eng.close()
        return out if rank == 0 else None
    except Exception as e:
        return {"baseline_config": "BASELINE.json configs[3]", "error": str(e)[:300]}


def spot_check_counts(recs, rows, meta, n_samples, pos_step, first_variant, k=48):
    """Counts of k records against the bits of this rank's host rows: ALTALT = popcount(rowA & rowB), margins = the allele
    counts, cells sum to 2N. Records whose rows this rank does not hold on the host are skipped."""
    if len(recs) == 0:
        return {"records": 0, "checked": 0, "counts_equal_bits": None}
    ia = (recs["packA"] >> 2).astype(np.int64) // pos_step - first_variant
    ib = (recs["packB"] >> 2).astype(np.int64) // pos_step - first_variant
    ok = np.flatnonzero((ia >= 0) & (ia < len(rows)) & (ib >= 0) & (ib < len(rows)))
    pick = ok[np.linspace(0, len(ok) - 1, min(k, len(ok))).astype(np.int64)] if len(ok) else ok
    good = True
    for r in pick:
        a, b = int(ia[r]), int(ib[r])
        n11 = int(np.unpackbits((rows[a] & rows[b]).view(np.uint8)).sum())
        c = recs["cnt"][r]
        good = good and c[3] == n11 and c[1] + c[3] == meta["ac"][a] and c[2] + c[3] == meta["ac"][b] and c.sum() == 2 * n_samples
    return {"records": int(len(recs)), "checked": int(len(pick)), "counts_equal_bits": bool(good)}


def extra_config4(args, B, tb, tools, peaks, peak_src):
    """A window-banded slice of BASELINE configs[4] (1,000,000 haplotypes x 2,000,000 SNVs, -w 500kb, 80 % of the variants with
    MAF < 1 %): c5_variants_per_gpu x N SNVs at the configuration's density (500 kb = 10,000 variants), POSITION-SHARDED
    (SURVEY 8e): every rank generates, uploads and holds only its own .twk blocks + the halo the window reaches
    (twkb_plan_shards), nothing is exchanged. Weak scaling: the slice grows with N; the full configuration is 2,000,000 / N
    own variants per GPU of the same band."""
    name = "BASELINE.json configs[4] (window-banded slice)"
    try:
        torch = B.torch
        rank, world = B.rank, B.world
        n, per_gpu, pos_step, w, bs = 500_000, args.c5_variants_per_gpu, 50, 500_000, 500
        m = per_gpu * world
        stride = tools.words_per_variant(n)
        t0 = time.perf_counter()
        first = np.arange(0, m + bs, bs, dtype=np.uint32); first[-1] = m      # .twk blocks of 500 variants (lib/importer.h:36)
        pos_meta = np.zeros(m, dtype=tb.VARIANT_DTYPE); pos_meta["pos"] = np.arange(m, dtype=np.uint64) * pos_step
        own, halo = tb.plan_shards(first, pos_meta, w, world)
        v0, v_own, v1 = int(first[own[rank]]), int(first[own[rank + 1]]), int(first[halo[rank]])
        d, _, meta = tools.synth_device(n, m, seed=args.seed + 4, rare_fraction=0.8, pos_step=pos_step, first=v0, n_rows=v1 - v0)
        host = torch.empty(d.shape, dtype=torch.int64).pin_memory(); host.copy_(d)
        del d
        torch.cuda.empty_cache()
        t_gen = time.perf_counter() - t0
        hn = host.numpy().view(np.uint64)
        eng = tb.Engine(force_phased=1, minR2=0.1, window=1, l_window=w, shard_blocks=int(own[rank + 1] - own[rank]), device=B.local_rank)
        blocks = first[own[rank]:halo[rank]] - v0

        def load():
            eng.load(n, hn, None, meta)
            eng.set_blocks(blocks)
        load()
        acc = B.resident(eng, 2, 1)
        e2e = B.e2e(eng, load, 1)
        st = acc["st"]
        spot = spot_check_counts(eng.compute(), hn, meta, n, pos_step, v0)
        roof = tensor_roofline(tb, acc, n, False, False, peaks, peak_src, None)
        sp_ms = float(np.mean(acc["sp"]))
        roof["list_kernel"] = {"kernel": "count_sparse_kernel", "variants": int(st.sparse_variants), "ms_per_step": sp_ms,
                               "word_ops_per_step": int(st.sparse_word_ops), "bound": "hbm",
                               "achieved_gbs": st.sparse_word_ops * 4 / (sp_ms * 1e-3) / 1e9 if sp_ms > 0 else None,
                               "peak_gbs": peaks["hbm_gbs"]}
        loaded, = B.allreduce([float(v1 - v0)], "SUM")
        spot_ok, = B.allreduce([1.0 if spot["counts_equal_bits"] in (True, None) else 0.0], "MIN")
        spot["all_ranks_ok"] = bool(spot_ok)
        out = _extra_entry(name,
                           f"tomahawk calc -p -w {w}, synthetic {n} samples ({2 * n} haplotypes) x {m} SNVs ({per_gpu} per GPU, {pos_step} bp apart), "
                           f"80% of the variants MAF<1%, R2>=0.1, position-sharded over {world} GPU(s)",
                           acc, e2e, roof, 2 * n, scaling="weak", n_gpus=world, gen_seconds=round(t_gen, 2),
                           shard={"own_variants_rank0": v_own - v0, "halo_variants_rank0": v1 - v_own, "variants_loaded_all_ranks": int(loaded),
                                  "matrix_bytes_whole": int(m) * stride * 8, "matrix_bytes_rank0": int(v1 - v0) * stride * 8,
                                  "collectives": "none (own blocks + halo loaded by every rank)"},
                           parity_spot=spot)
        eng.close()
        return out if rank == 0 else None
    except Exception as e:
        return {"baseline_config": name, "error": str(e)[:300]}


if __name__ == "__main__":
    sys.exit(main())
