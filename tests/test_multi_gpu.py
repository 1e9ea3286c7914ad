"""Multi-GPU data plane (NCCL inside libtwkb): every rank uploads its slice of the rows, the ranks exchange the
slices, each computes its share of the tiles. GPU tests need >= 2 devices (run with `gpurun --gpus 2`); the
CPU tests cover the slice arithmetic and a world_size-2 gloo rehearsal of the host-side protocol."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import tomahawk_b200 as tb
from oracle import ldcore as lc
from oracle import twk_format as tf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("m,n", [(10, 4), (200_000, 8), (7, 8), (1, 2), (1000, 1), (565_685, 8)])
def test_comm_slices_partition_the_rows(m, n):
    cuts = [tb.comm_slice(m, r, n) for r in range(n)]
    assert cuts[0][0] == 0 and cuts[-1][1] == m
    for (b0, e0), (b1, e1) in zip(cuts, cuts[1:]):
        assert e0 == b1 and b0 <= e0
    sizes = [e - b for b, e in cuts]
    assert max(sizes) == -(-m // n)                      # ceil(m / n) rows on the full slices


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, q):
    """Host-side protocol of a sliced load, rehearsed on CPU: rank 0 creates the communicator id and the
    metadata, both travel through torch.distributed; every rank derives its own slice."""
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    m = 1001
    b, e = tb.comm_slice(m, rank, world)
    rows = torch.arange(b, e, dtype=torch.int64)
    gathered = [None] * world
    dist.all_gather_object(gathered, rows.tolist())
    if rank == 0:
        q.put((uid[0], gathered))
    dist.destroy_process_group()


def test_two_rank_gloo_slice_protocol():
    import torch.multiprocessing as mp

    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    uid, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert uid == bytes(range(128))
    assert sum(gathered, []) == list(range(1001))


# ------------------------------------------------------------------ GPU (>= 2 devices)
def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


WORKER = r'''
import os, sys, pickle
sys.path.insert(0, {root!r})
import numpy as np
import torch, torch.distributed as dist
import tomahawk_b200 as tb
from oracle import twk_format as tf, ldcore as lc
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo")
kw = pickle.loads(bytes.fromhex(sys.argv[1]))
s = tf.synth_genotypes(**kw["synth"])
data, mask = tf.pack_bits(s)
meta = lc.variant_meta(s)
uid = [tb.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
eng = tb.Engine(device=rank, part_index=rank, part_count=world, **kw["prm"])
eng.comm_init(uid[0], rank, world)
b, e = tb.comm_slice(s.n_variants, rank, world)
if kw["runs"]:
    raw, desc = tf.encode_runs(s, seed=3)
    eng.load_runs_sliced(s.n_samples, raw, desc, meta)
else:
    eng.load_sliced(s.n_samples, s.n_variants, data[b:e], mask[b:e] if mask is not None else None, meta)
rows, mrows = eng.rows(s.n_variants, data.shape[1], with_mask=mask is not None)
assert np.array_equal(rows, data), "gathered rows differ from the full matrix"
if mask is not None:
    assert np.array_equal(mrows, mask)
recs = eng.compute()
st = eng.stats()
out = [None] * world
dist.all_gather_object(out, (recs.tobytes(), int(st.pairs_visited), int(st.bytes_h2d)))
if rank == 0:
    with open(sys.argv[2], "wb") as f:
        pickle.dump(out, f)
eng.close()
dist.destroy_process_group()
'''


@pytest.mark.gpu
@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("case", ["matrix", "matrix_missing_unphased", "runs"])
def test_sliced_load_and_exchange_equals_single_gpu(case, tmpdir_repo):
    import pickle

    world = min(_n_gpus(), 4)
    kw = {"matrix": dict(synth=dict(n_samples=2504, n_variants=5003, seed=41), prm=dict(force_phased=1, minR2=0.1), runs=False),
          "matrix_missing_unphased": dict(synth=dict(n_samples=500, n_variants=3001, seed=42, missing_rate=0.05),
                                          prm=dict(forced_unphased=1, minR2=0.1), runs=False),
          "runs": dict(synth=dict(n_samples=700, n_variants=2600, seed=43), prm=dict(force_phased=1, minR2=0.05), runs=True)}[case]
    script = os.path.join(tmpdir_repo, "mg_worker.py")
    open(script, "w").write(WORKER.format(root=ROOT))
    out = os.path.join(tmpdir_repo, f"mg_{case}.pkl")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), script, pickle.dumps(kw).hex(), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    parts = pickle.load(open(out, "rb"))
    s = tf.synth_genotypes(**kw["synth"])
    data, mask = tf.pack_bits(s)
    eng = tb.Engine(**kw["prm"])
    eng.load(s.n_samples, data, mask, lc.variant_meta(s))
    whole = tf.canonical(eng.compute(), forward_only=False)
    st = eng.stats()
    eng.close()
    got = np.concatenate([np.frombuffer(p[0], dtype=tb.TWO_DTYPE) for p in parts])
    got = tf.canonical(got, forward_only=False)
    assert np.array_equal(got.view(np.uint8), whole.view(np.uint8))          # disjoint parts, same records
    assert sum(p[1] for p in parts) == st.pairs_visited
    if not kw["runs"]:
        full = data.nbytes * (2 if mask is not None else 1)
        assert max(p[2] for p in parts) < 0.6 * full                         # every rank uploaded only its slice


@pytest.mark.gpu
@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
def test_cli_multi_device_uses_the_communicator(tmpdir_repo):
    """twkb_calc -g 0,1: the C++ mirror (include/twkb_ld.hpp) forms the communicator from threads of one process."""
    s = tf.synth_genotypes(600, 2300, seed=44)
    twk = os.path.join(tmpdir_repo, "mg_cli.twk")
    tf.write_twk(twk, s)
    devs = ",".join(str(d) for d in range(min(_n_gpus(), 4)))
    exe = os.path.join(ROOT, "tomahawk_b200", "twkb_calc")
    outs = {}
    for name, g in (("multi", devs), ("single", "0")):
        out = os.path.join(tmpdir_repo, f"mg_cli_{name}.two")
        r = subprocess.run([exe, "calc", "-p", "-r", "0.1", "-i", twk, "-o", out, "-g", g, "-t", "4"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        if name == "multi":
            assert "NCCL communicator" in r.stderr
        outs[name] = tf.canonical(tf.read_two(out), forward_only=False)
    assert len(outs["multi"]) > 100
    assert np.array_equal(outs["multi"].view(np.uint8), outs["single"].view(np.uint8))
