"""Synthetic genotype matrices in the layout ``twkb_load_matrix`` takes.

Bench / test data tooling only (numpy): a seeded generator with a skewed
allele-frequency spectrum and LD blocks, and the packer that
turns allele codes into the reference's ``twk_igt_vec`` rows
(reference lib/core.cpp:349-383). Nothing here computes LD.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

VARIANT_DTYPE = np.dtype(
    [("rid", "<u4"), ("pos", "<u4"), ("ac", "<u4"), ("an", "<u4"), ("hwe", "<f8"),
     ("gt_missing", "u1"), ("gt_phase", "u1"), ("pad", "u1", (6,))]
)

@dataclass
class Synth:
    """Genotypes as allele codes (0 ref, 1 alt, 2 missing; genotype_encoder.h:11-17)."""

    alleles: np.ndarray  # uint8 [n_variants, 2*n_samples]
    pos: np.ndarray  # uint32 [n_variants], 0-based, strictly increasing
    rid: np.ndarray  # uint32 [n_variants]
    n_samples: int
    phased: bool = True

    @property
    def n_variants(self) -> int:
        return int(self.alleles.shape[0])

    @property
    def ac(self) -> np.ndarray:
        return (self.alleles == 1).sum(axis=1).astype(np.uint32)

    @property
    def an(self) -> np.ndarray:
        """Number of missing alleles (what the reference stores in twk1_t::an)."""
        return (self.alleles == 2).sum(axis=1).astype(np.uint32)


def synth_genotypes(
    n_samples: int,
    n_variants: int,
    seed: int = 1,
    missing_rate: float = 0.0,
    rare_fraction: float = 0.0,
    pos_step: int = 100,
    p_copy: float = 0.7,
    redraw: float = 0.05,
    chunk: int = 2048,
) -> Synth:
    """LD-block generator.

    Variants come in blocks: a founder is drawn fresh with alt-allele frequency
    0.5*U^3 (floored at 1/2N); each following variant continues the block with
    probability ``p_copy`` and copies the founder's haplotypes with an
    independent per-haplotype re-draw probability 1-(1-redraw)^depth.
    ``rare_fraction`` of the founders are forced to MAF < 1 % (config 5).
    Every variant is resampled until 1 <= ac <= 2N-1 (the reference asserts on
    monomorphic sites, include/core.h:536-548). Missing genotypes knock out both
    alleles of a sample i.i.d. with ``missing_rate``.
    """
    rng = np.random.default_rng(seed)
    H = 2 * n_samples
    out = np.zeros((n_variants, H), dtype=np.uint8)
    # block structure
    new_block = rng.random(n_variants) >= p_copy
    new_block[0] = True
    founder = np.maximum.accumulate(np.where(new_block, np.arange(n_variants), 0))
    depth = np.arange(n_variants) - founder
    af = 0.5 * rng.random(n_variants) ** 3
    if rare_fraction > 0:
        rare = rng.random(n_variants) < rare_fraction
        af = np.where(rare, 0.01 * rng.random(n_variants), af)
    af = np.maximum(af, 1.0 / H)
    af = af[founder]
    founders_idx = np.flatnonzero(new_block)
    # founders first (fresh draws), in chunks
    for s in range(0, len(founders_idx), chunk):
        idx = founders_idx[s : s + chunk]
        u = rng.random((len(idx), H), dtype=np.float32)
        out[idx] = u < af[idx, None].astype(np.float32)
    # members: copy founder with per-haplotype redraw
    members_idx = np.flatnonzero(~new_block)
    for s in range(0, len(members_idx), chunk):
        idx = members_idx[s : s + chunk]
        pr = 1.0 - (1.0 - redraw) ** depth[idx]
        u = rng.random((len(idx), H), dtype=np.float32)
        fresh = rng.random((len(idx), H), dtype=np.float32) < af[idx, None].astype(np.float32)
        out[idx] = np.where(u < pr[:, None].astype(np.float32), fresh, out[founder[idx]])
    # guarantee 1 <= ac <= H-1
    ac = out.sum(axis=1)
    for v in np.flatnonzero(ac == 0):
        out[v, rng.integers(0, H)] = 1
    for v in np.flatnonzero(ac == H):
        out[v, rng.integers(0, H)] = 0
    if missing_rate > 0:
        for s in range(0, n_variants, chunk):
            m = rng.random((min(chunk, n_variants - s), n_samples), dtype=np.float32) < missing_rate
            m2 = np.repeat(m, 2, axis=1)
            blk = out[s : s + chunk]
            blk[m2] = 2
        # keep at least one alt allele (ac >= 1) after knocking genotypes out
        ac = (out == 1).sum(axis=1)
        for v in np.flatnonzero(ac == 0):
            out[v, 0] = 1
            out[v, 1] = 0 if out[v, 1] == 2 else out[v, 1]
    pos = (np.arange(n_variants, dtype=np.uint64) * pos_step).astype(np.uint32)
    rid = np.zeros(n_variants, dtype=np.uint32)
    return Synth(alleles=out, pos=pos, rid=rid, n_samples=n_samples)


def words_per_variant(n_samples: int) -> int:
    """Row stride (u64 words) of the packed matrix: ceil(2N/64) rounded up to a
    multiple of 2 so every row is 128-bit aligned (the reference aligns rows to
    SIMD_ALIGNMENT, include/core.h:52-60,126-136)."""
    w = (2 * n_samples + 63) // 64
    return (w + 1) // 2 * 2


def pack_bits(s: Synth):
    """-> (data[u64 M x W], mask[u64 M x W] or None). lib/core.cpp:365-383."""
    M, H = s.alleles.shape
    W = words_per_variant(s.n_samples)
    pad = W * 64 - H
    a = s.alleles
    data_bits = (a == 1).astype(np.uint8)
    miss = a == 2
    has_missing = bool(miss.any())

    def _pack(bits):
        bits = np.pad(bits, ((0, 0), (0, pad)))
        by = np.packbits(bits, axis=1, bitorder="little")
        return np.ascontiguousarray(by).view("<u8").reshape(M, W)

    data = _pack(data_bits)
    mask = None
    if has_missing:
        ms = miss.reshape(M, s.n_samples, 2).any(axis=2)
        mask = _pack(np.repeat(ms, 2, axis=1).astype(np.uint8))
    return data, mask




def variant_meta(s: Synth) -> np.ndarray:
    """twkb_variant array (subset of twk1_t, reference include/core.h:291-295)."""
    m = np.zeros(s.n_variants, VARIANT_DTYPE)
    m["rid"] = s.rid
    m["pos"] = s.pos
    m["ac"] = s.ac
    m["an"] = s.an
    m["hwe"] = 1.0
    m["gt_missing"] = s.an != 0
    m["gt_phase"] = 1 if s.phased else 0
    return m


def biobank_matrix(n_samples, n_variants, seed):
    """BASELINE configs[4] in miniature: 1M-haplotype rows, 80 % of the variants rare (MAF < 1 %:
    60 % with <= 400 carriers -- the list class under the automatic threshold -- and 20 % with up to
    10,000), 20 % common; neighbours share carriers (LD) so that records survive an R2 cut."""
    rng = np.random.default_rng(seed)
    nb = 2 * n_samples
    words = (nb + 127) // 128 * 2
    data = np.zeros((n_variants, words), np.uint64)
    ac = np.zeros(n_variants, np.uint32)
    prev_idx = None
    for v in range(n_variants):
        u = rng.random()
        if u < 0.8:
            k = int(rng.integers(2, 400)) if u < 0.6 else int(rng.integers(400, 10000))
            if prev_idx is not None and rng.random() < 0.6:     # copy most carriers of the previous rare variant
                keep = prev_idx[rng.random(len(prev_idx)) < 0.9]
                idx = np.unique(np.concatenate([keep, rng.integers(1, nb, max(1, k // 10))]))
            else:
                idx = np.unique(rng.integers(1, nb, k))
            prev_idx = idx
            row = np.zeros(words * 8, np.uint8)
            np.bitwise_or.at(row, idx >> 3, (1 << (idx & 7)).astype(np.uint8))
            data[v] = row.view(np.uint64)
            ac[v] = len(idx)
        else:
            bits = np.zeros(words * 64, np.uint8)
            bits[1:nb] = rng.random(nb - 1) < rng.uniform(0.05, 0.5)
            ac[v] = bits.sum()
            data[v] = np.packbits(bits, bitorder="little").view(np.uint64)
    meta = np.zeros(n_variants, VARIANT_DTYPE)
    meta["pos"] = 100 * (1 + np.arange(n_variants)); meta["ac"] = ac; meta["hwe"] = 1.0; meta["gt_phase"] = 1
    return data, meta
