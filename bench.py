#!/usr/bin/env python
"""bench.py -- throughput of the B200 `tomahawk calc` path (variant-pairs/s).

A "step" is one complete pass of the LD hot path over the synthetic genotype
matrix: count kernel (+ fused R2 screen/compaction) and statistics kernel for
every pair of this rank's share of the tile grid.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--variants M] [--samples S] [--min-r2 R] [--kernel auto|popc|umma]

Workload at N=1: BASELINE.json configs[1] -- phased all-pairs, 2,504 samples
(5,008 haplotypes) x 200,000 SNVs, R2 >= 0.1. For N>1 (weak scaling, launched by
torchrun, one rank per GPU) the variant count grows as M*sqrt(N) so that every
GPU keeps the pair count of the N=1 run; rank 0 generates the matrix, it is
broadcast once over NCCL and each rank computes an interleaved share of the
tile grid with no further collectives.

One JSON line is printed by rank 0 (see README / DESIGN.md for every key).
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
REF_SAMPLE_VARIANTS = 30_000  # bounded CPU sample of the same workload (first M' variants)
# dram__bytes_read.sum + dram__bytes_write.sum of one count_umma3_kernel<e2m1> launch at the full C2 size
# (profiles/round1_ncu_c2_fp4_full_v2.csv: 20.256 GB + 0.049 GB; the operand is 0.51 GB, re-read from L2 misses)
TRAFFIC_C2_FP4 = 20.255835e9 + 48.632064e6


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
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
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
        time.sleep(0.25)
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


# --------------------------------------------------------------------- reference arm
def reference_run(args, n_variants, steps, warmup, as_baseline=False):
    """Times the reference's own `tomahawk calc` (oracle/_ref/tomahawk_calc, built
    from /root/reference by oracle/build_ref.sh) on this host's cores, on the first
    n_variants variants of the same synthetic workload. The only place bench.py
    executes anything under oracle/."""
    from oracle import ldcore as lc
    from oracle import twk_format as tf

    cores = os.cpu_count() or 1
    s = tf.synth_genotypes(args.samples, n_variants, seed=args.seed, missing_rate=args.missing)
    mode_flag = "-u" if args.unphased else "-p"
    tmp = tempfile.mkdtemp(prefix="twkb_ref_")
    twk = os.path.join(tmp, "ref.twk")
    tf.write_twk(twk, s)
    pairs = n_variants * (n_variants - 1) // 2
    sample = f"first {n_variants} of the workload's variants ({pairs} pairs), same N, {mode_flag} -r {args.min_r2} -t {cores}"
    if lc.have_reference():
        kind = "reference"
        rates, times = [], []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            info = lc.run_reference_calc(twk, os.path.join(tmp, "ref_out"), [mode_flag, "-r", str(args.min_r2)], threads=cores)
            dt = time.perf_counter() - t0
            if it >= warmup:
                rates.append(info.get("pairs_per_s", pairs / dt))
                times.append(dt)
            if as_baseline:
                break
        if as_baseline and not rates:
            rates.append(info.get("pairs_per_s", pairs / dt)); times.append(dt)
        value = float(np.mean(rates))
        ms = float(np.mean(times)) * 1e3
    else:
        kind = "port"
        cores = 1
        prm = lc.default_params(minR2=args.min_r2, **({"forced_unphased": 1} if args.unphased else {"force_phased": 1}))
        t0 = time.perf_counter()
        lc.calc(s, prm, cap=max(1 << 20, pairs // 4))
        dt = time.perf_counter() - t0
        value, ms = pairs / dt, dt * 1e3
    for fn in os.listdir(tmp):
        os.unlink(os.path.join(tmp, fn))
    os.rmdir(tmp)
    return {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "ms": ms, "pairs": pairs}


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


def main():
    args = parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = max(args.gpus, 1)

    mode_name = "-u (unphased 3x3)" if args.unphased else "-p"
    miss_name = f", {100 * args.missing:g}% missing genotypes" if args.missing > 0 else ""
    workload = (f"tomahawk calc {mode_name} all-pairs, synthetic {args.samples} samples ({2 * args.samples} haplotypes) x "
                f"{args.variants} SNVs{miss_name}, R2>={args.min_r2}")

    if args.impl == "reference":
        if rank != 0:
            return 0
        nv = min(args.ref_variants, args.variants)
        r = reference_run(args, nv, args.steps, args.warmup)
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
    from tomahawk_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: tomahawk_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    # ---- workload: weak scaling keeps pairs per GPU constant => M grows with sqrt(N)
    n_variants = int(round(args.variants * math.sqrt(world)))
    n_samples = args.samples
    stride = synth.words_per_variant(n_samples)
    t_gen0 = time.perf_counter()
    if rank == 0:
        s = synth.synth_genotypes(n_samples, n_variants, seed=args.seed, missing_rate=args.missing)
        data, mask = synth.pack_bits(s)
        meta = synth.variant_meta(s)
        del s
    else:
        data = np.zeros((n_variants, stride), dtype=np.uint64)
        mask = np.zeros((n_variants, stride), dtype=np.uint64) if args.missing > 0 else None
        meta = np.zeros(n_variants, dtype=synth.VARIANT_DTYPE)
    t_gen = time.perf_counter() - t_gen0

    kernel = {"auto": tb.KERNEL_AUTO, "popc": tb.KERNEL_POPC, "umma": tb.KERNEL_UMMA, "i8": tb.KERNEL_UMMA,
              "fp4": tb.KERNEL_UMMA_FP4}[args.kernel]
    eng = tb.Engine(force_phased=0 if args.unphased else 1, forced_unphased=1 if args.unphased else 0, minR2=args.min_r2,
                    kernel=kernel, device=local_rank, part_index=rank, part_count=world)
    # host copy in pinned memory (the e2e leg copies from here every step)
    host = torch.from_numpy(data.view(np.int64)).pin_memory()
    host_mask = torch.from_numpy(mask.view(np.int64)).pin_memory() if mask is not None else None
    if world > 1:
        # ONE broadcast of the packed matrix (+ metadata) over NCCL/NVLink, then no collectives
        dev = host.cuda(non_blocking=True) if rank == 0 else torch.empty_like(host, device="cuda")
        dist.broadcast(dev, src=0)
        meta_t = torch.from_numpy(meta.view(np.uint8).copy()).cuda()
        dist.broadcast(meta_t, src=0)
        meta = meta_t.cpu().numpy().view(synth.VARIANT_DTYPE)
        if rank != 0:
            host.copy_(dev.cpu())
        dev_mask = None
        if host_mask is not None:
            dev_mask = host_mask.cuda(non_blocking=True) if rank == 0 else torch.empty_like(host_mask, device="cuda")
            dist.broadcast(dev_mask, src=0)
            if rank != 0:
                host_mask.copy_(dev_mask.cpu())
        torch.cuda.synchronize()
        eng.load_device(n_samples, n_variants, dev.data_ptr(), dev_mask.data_ptr() if dev_mask is not None else None, stride, meta)
        del dev, dev_mask
    else:
        eng.load(n_samples, host.numpy().view(np.uint64), host_mask.numpy().view(np.uint64) if host_mask is not None else None, meta)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        flush.zero_()  # L2 flush between iterations (256 MiB > 126 MB L2)
        torch.cuda.synchronize()
        eng.compute_resident()
        return eng.stats()

    for _ in range(args.warmup):
        step_resident()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    dev_ms, cnt_ms, sts_ms, launches = [], [], [], 0
    cnt_launches = 0
    t0 = time.perf_counter()
    st = None
    for _ in range(args.steps):
        st = step_resident()
        dev_ms.append(st.ms_device_total)
        cnt_ms.append(st.ms_count_kernel)
        sts_ms.append(st.ms_stats_kernel)
        launches += st.count_launches + st.stats_launches + st.other_launches
        cnt_launches += st.count_launches
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    pairs_rank = st.pairs_visited
    step_ms = float(np.mean(dev_ms))
    t = torch.tensor([step_ms, float(pairs_rank)], dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        step_ms_max, pairs_total = float(tmax[0]), float(tsum[1])
    else:
        step_ms_max, pairs_total = step_ms, float(pairs_rank)
    value = pairs_total / (step_ms_max * 1e-3)

    # ---- e2e: host buffers -> C-ABI -> records back on the host, copies inside the timed region.
    # N > 1: every rank uploads the matrix over its own PCIe link. The alternative (rank 0 uploads, one NCCL broadcast
    # pipelined in 8 row slices, twkb_load_matrix_device) was measured and is slower on this box: e2e 41.0 vs 39.9 ms
    # at N = 2 and 46.8 vs 45.1 ms at N = 4 (the broadcast serialises behind rank 0's single PCIe link).
    e2e_ms = []
    h2d = d2h = 0
    host_np = host.numpy().view(np.uint64)
    host_mask_np = host_mask.numpy().view(np.uint64) if host_mask is not None else None
    for it in range(1 + max(2, min(args.steps, 3))):
        flush.zero_()
        barrier()
        t1 = time.perf_counter()
        eng.load(n_samples, host_np, host_mask_np, meta)   # H2D from pinned memory + device transpose
        t_load = time.perf_counter() - t1
        eng.compute_discard()                      # compute + D2H of every record into pinned staging
        torch.cuda.synchronize()
        dt = time.perf_counter() - t1
        s2 = eng.stats()
        if it > 0:
            e2e_ms.append(dt * 1e3)
            e2e_parts = {"load_wall_ms": t_load * 1e3, "load_device_ms": s2.ms_h2d, "compute_wall_ms": (dt - t_load) * 1e3,
                         "compute_device_ms": s2.ms_device_total}
            h2d, d2h = int(s2.bytes_h2d), int(s2.bytes_d2h)
            launches_e2e = s2.count_launches + s2.stats_launches + s2.other_launches
    e2e_t = torch.tensor([float(np.mean(e2e_ms))], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = pairs_total / (float(e2e_t[0]) * 1e-3)
    records = int(s2.records_out)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (count) kernel
    peaks, peak_src = measured_peaks()
    H = 2 * n_samples
    avg_launch_s = (sum(cnt_ms) / max(cnt_launches, 1)) * 1e-3
    pairs_per_launch = pairs_rank * args.steps / max(cnt_launches, 1)
    tensor = st.kernel_used in (tb.KERNEL_UMMA, tb.KERNEL_UMMA_FP4)
    fp4 = st.kernel_used == tb.KERNEL_UMMA_FP4
    if tensor:
        # GEMM view: 2N bit-MACs per visited pair = 2*2N flop. The tensor peak of the
        # operand type is the bf16 figure scaled by the nominal dense ratio (bf16 : int8/fp8 : fp4 =
        # 1 : 2 : 4; B200_PROFILING.md); the step is longer than a burst, so the sustained figure.
        if args.unphased:   # 9 (missing) / 4 plane products over N samples
            flop_per_pair = 2.0 * (9 if args.missing > 0 else 4) * n_samples
        else:               # 2N bit-MACs, x4 masked counts with missing data
            flop_per_pair = 2.0 * H * (4 if args.missing > 0 else 1)
        achieved = pairs_per_launch * flop_per_pair / avg_launch_s / 1e12
        # MEASURED_PEAKS.json holds bf16 only. The int8 / e2m1 tcgen05 kinds run at 2x / 4x the bf16 MAC
        # rate (nominal dense 2.25 : 4.5 : 9 PFLOP/s, B200_PROFILING.md). `peak` is that nominal figure of
        # the operand kind: ncu confirms it is the right denominator (profiles/round1_ncu_c2_fp4_full.csv:
        # tensor pipe 79.8 % active at 0.79 of nominal). 4x the measured *sustained bf16* number
        # underestimates the e2m1 pipe (frac would read 1.25: cuBLAS bf16 is power-capped near 1.3 GHz,
        # this kernel holds 1.84-1.97 GHz at ~760 W); it is reported beside it for reference.
        ratio = 4.0 if fp4 else 2.0
        peak = 9000.0 if fp4 else 4500.0
        c2 = (fp4 and world == 1 and n_variants == BASE_VARIANTS and n_samples == BASE_SAMPLES and not args.unphased
              and args.missing == 0)
        traffic = TRAFFIC_C2_FP4 if c2 else None
        # Measured ceiling of the tensor pipe for THIS kernel's instruction stream on THIS device: the same
        # launch with operand traffic and epilogue switched off (TWKB_DEBUG_FLAGS=3: only the first ring
        # fill is loaded, accumulators are not drained; results are discarded). What remains is the
        # tcgen05.mma issue rate under the board's power/clock behaviour.
        mma_only = None
        if not args.no_mma_ceiling and not args.unphased and args.missing == 0 and os.path.exists(tb.PROF_LIB_PATH):
            # profiling build of the same sources (libtwkb_prof.so); the product library has no such switch
            os.environ["TWKB_DEBUG_FLAGS"] = "3"
            try:
                peng = tb.Engine(force_phased=1, minR2=args.min_r2, kernel=kernel, device=local_rank, part_index=rank,
                                 part_count=world, profiling=True)
                peng.load(n_samples, host_np, host_mask_np, meta)
                ms = []
                for _ in range(3):
                    flush.zero_(); torch.cuda.synchronize()
                    peng.compute_resident()
                    s3 = peng.stats()
                    ms.append(s3.ms_count_kernel / max(s3.count_launches, 1))
                peng.close()
                mma_only = pairs_per_launch * flop_per_pair / (min(ms[1:]) * 1e-3) / 1e12
            finally:
                del os.environ["TWKB_DEBUG_FLAGS"]
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "count_umma3_kernel<%s>" % ("true" if fp4 else "false"),
                    "peak_source": f"nominal dense {'e2m1 (kind::mxf4)' if fp4 else 'int8 (kind::i8)'} tcgen05 rate; no entry "
                                   f"for this operand kind in MEASURED_PEAKS.json ({peak_src})",
                    "frac_of_scaled_measured_bf16": achieved / (ratio * peaks["bf16_tflops_sustained"]),
                    "scaled_measured_bf16_peak": ratio * peaks["bf16_tflops_sustained"],
                    "mma_only_ceiling": mma_only, "frac_of_mma_only_ceiling": (achieved / mma_only) if mma_only else None,
                    "note": f"algorithmic work = pairs x {flop_per_pair:g} flop (SURVEY 8d; K padding to 256 and the 256x240 tile edge "
                            f"are not counted); traffic = dram read+write bytes of one launch from the committed ncu --set full "
                            f"capture; mma_only_ceiling = same launch without operand loads and epilogue, measured in this run"}
    else:
        # LOP3+POPC kernel: INT-pipe bound. Algorithmic work = ceil(2N/32) AND+POPC word-ops per pair;
        # peak = 16 POPC lanes/clk/SM x 148 SMs x max SM clock (to be replaced by the measured issue rate).
        w = math.ceil(H / 32)
        achieved_wops = pairs_per_launch * w / avg_launch_s
        smax = (clocks or {}).get("sm_max_mhz") or 1965.0
        peak_wops = 16 * 148 * smax * 1e6
        # expressed in GB/s-equivalent of operand bits so the JSON keeps the contract's units
        roofline = {"bound": "int_pipe_popc", "achieved": achieved_wops / 1e12, "peak": peak_wops / 1e12,
                    "unit": "Tword-op/s", "frac": achieved_wops / peak_wops, "traffic": None,
                    "kernel": "count_popc_kernel<0>",
                    "note": "INT-pipe roofline: 16 POPC/clk/SM x 148 SM x max SM clock; 157 word-ops per pair"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": ("e2m1 x e2m1 -> f32 exact (tcgen05 kind::mxf4) + f64 statistics" if fp4 else
                  "int8 x int8 -> int32 (tcgen05) + f64 statistics" if tensor else "u32 popcount + f64 statistics"),
        "data": "synthetic",
        "config": {
            "workload": (f"tomahawk calc {mode_name} all-pairs, synthetic {n_samples} samples ({H} haplotypes) x {n_variants} SNVs"
                         f"{miss_name}, R2>={args.min_r2}" + (f" (weak scaling: {args.variants} x sqrt({world}) variants)" if world > 1 else "")),
            "baseline_config": "BASELINE.json configs[2]" if args.unphased else "BASELINE.json configs[1]",
            "pairs_per_step": pairs_total, "haplotype_cmp_per_s": value * H, "records_per_step": records,
            "kernel": "umma_fp4" if fp4 else "umma_i8" if tensor else "popc",
            "l2": "256 MiB device memset between steps (flush) and operands > L2",
            "seed": args.seed, "gen_seconds": round(t_gen, 2),
            "ms_count_kernel_per_step": float(np.mean(cnt_ms)), "ms_stats_kernel_per_step": float(np.mean(sts_ms)),
            "wall_seconds_timed_region": wall,
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": float(e2e_t[0]), "gpu_launches_per_step": int(launches_e2e), "parts_last_step": e2e_parts,
                "path": "every rank: twkb_load_matrix from its pinned host copy -> twkb_compute with a host sink"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = reference_run(args, min(args.ref_variants, n_variants), 1, 0, as_baseline=True)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        except Exception as e:  # the baseline is a reported number, never a gate
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": str(e)[:200]}
    emit_json(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
