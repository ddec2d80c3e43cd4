#!/usr/bin/env python
"""Synthetic FASTQ of a BASELINE config shape: uniform-random genome, reads from uniform positions,
50 % reverse-complemented, i.i.d. substitutions; quality 'I', '#' at substituted bases; optional N injection.
    python tools/make_fastq.py out.fastq --genome 200000 --reads 40000 --len 100 --err 0.005 [--seed 1] [--n-frac 0.01]
    [--pair out2.fastq]   (mate 2 from the opposite strand at a ~300 bp insert)"""
import argparse
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("out"); ap.add_argument("--pair")
ap.add_argument("--genome", type=int, default=200_000); ap.add_argument("--reads", type=int, default=40_000)
ap.add_argument("--len", type=int, default=100); ap.add_argument("--err", type=float, default=0.005)
ap.add_argument("--seed", type=int, default=1); ap.add_argument("--n-frac", type=float, default=0.0)
a = ap.parse_args()
rng = np.random.default_rng(a.seed)
ACGT = np.frombuffer(b"ACGT", np.uint8)
COMP = np.zeros(256, np.uint8); COMP[list(b"ACGTN")] = list(b"TGCAN")
g = ACGT[rng.integers(0, 4, a.genome)]
L, n = a.len, a.reads
ins = 300 if a.pair else 0
start = rng.integers(0, a.genome - L - ins + 1, n)
idx = start[:, None] + np.arange(L)[None, :]


def mutate(r):
    mask = rng.random(r.shape) < a.err
    code = np.searchsorted(ACGT, r)
    r = np.where(mask, ACGT[(code + rng.integers(1, 4, r.shape)) & 3], r)
    q = np.where(mask, ord("#"), ord("I")).astype(np.uint8)
    if a.n_frac > 0:
        sel = np.nonzero(rng.random(r.shape[0]) < a.n_frac)[0]
        pos = rng.integers(0, L, sel.size)
        r[sel, pos] = ord("N"); q[sel, pos] = ord("#")
    return r.astype(np.uint8), q


def write(path, reads, quals, tag):
    with open(path, "wb") as f:
        for i in range(reads.shape[0]):
            f.write(b"@r%d/%d\n" % (i, tag)); f.write(reads[i].tobytes()); f.write(b"\n+\n"); f.write(quals[i].tobytes()); f.write(b"\n")


r1 = g[idx]
flip = rng.random(n) < 0.5
if a.pair:
    r2 = COMP[g[idx + ins][:, ::-1]]
    a1 = np.where(flip[:, None], r2, r1); a2 = np.where(flip[:, None], r1, r2)
    m1, q1 = mutate(a1); m2, q2 = mutate(a2)
    write(a.out, m1, q1, 1); write(a.pair, m2, q2, 2)
else:
    r1 = np.where(flip[:, None], COMP[r1[:, ::-1]], r1)
    m1, q1 = mutate(r1)
    write(a.out, m1, q1, 1)
