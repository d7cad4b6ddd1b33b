"""Deterministic synthetic sequence databases for parity tests and bench.py.

Protein lengths follow a Swiss-Prot-like log-normal, residues are i.i.d. with Robinson-Robinson
background frequencies; a fraction of the records are mutated copies of the query (planted
homologs, so that scores cross the 8-bit tier limit) and a few exact duplicates are added so that
ties straddle the top-k cut-off.
"""
import os

import numpy as np

AA = "ARNDCQEGHILKMFPSTWYV"
# Robinson & Robinson (1991) amino-acid background frequencies, order of AA above
AA_FREQ = np.array([0.07805, 0.05129, 0.04487, 0.05364, 0.01925, 0.04264, 0.06295, 0.07377, 0.02199, 0.05142,
                    0.09019, 0.05744, 0.02243, 0.03856, 0.05203, 0.07120, 0.05841, 0.01330, 0.03216, 0.06441])
AA_FREQ = AA_FREQ / AA_FREQ.sum()
NT = "ACGT"


def _lengths(rng, n, mu, sigma, lo, hi):
    ln = np.rint(rng.lognormal(mu, sigma, n)).astype(np.int64)
    return np.clip(ln, lo, hi)


def random_protein(rng, length):
    idx = rng.choice(len(AA), size=int(length), p=AA_FREQ)
    return "".join(AA[i] for i in idx)


def mutate(rng, seq, rate, alphabet, freq=None):
    s = np.array(list(seq))
    mask = rng.random(len(s)) < rate
    repl = rng.choice(len(alphabet), size=int(mask.sum()), p=freq)
    s[mask] = [alphabet[i] for i in repl]
    return "".join(s)


def protein_db(n, seed, query=None, homolog_frac=0.01, n_dup=50, mu=5.6, sigma=0.62, min_len=30, max_len=35000):
    """List of n protein sequences (strings)."""
    rng = np.random.default_rng(seed)
    lens = _lengths(rng, n, mu, sigma, min_len, max_len)
    total = int(lens.sum())
    flat = rng.choice(len(AA), size=total, p=AA_FREQ).astype(np.uint8)
    table = np.frombuffer(AA.encode(), dtype=np.uint8)
    chars = table[flat].tobytes().decode()
    seqs, pos = [], 0
    for ln in lens:
        seqs.append(chars[pos:pos + int(ln)])
        pos += int(ln)
    if query:
        n_h = int(n * homolog_frac)
        where = rng.choice(n, size=n_h, replace=False)
        for w in where:
            rate = rng.uniform(0.10, 0.70)
            left = random_protein(rng, rng.integers(0, 60))
            right = random_protein(rng, rng.integers(0, 60))
            seqs[w] = left + mutate(rng, query, rate, AA, AA_FREQ) + right
    if n_dup and n > 2 * n_dup:
        src = rng.choice(n, size=n_dup, replace=False)
        dst = rng.choice(n, size=n_dup, replace=False)
        for s, d in zip(src, dst):
            seqs[d] = seqs[s]
    return seqs


def nucleotide_db(n, seed, query=None, n_homologs=0, mu=7.2, sigma=0.45, min_len=50, max_len=20000, iupac_rate=0.0):
    rng = np.random.default_rng(seed)
    lens = _lengths(rng, n, mu, sigma, min_len, max_len)
    total = int(lens.sum())
    flat = rng.integers(0, 4, size=total).astype(np.uint8)
    table = np.frombuffer(NT.encode(), dtype=np.uint8)
    chars = table[flat].tobytes().decode()
    seqs, pos = [], 0
    for ln in lens:
        seqs.append(chars[pos:pos + int(ln)])
        pos += int(ln)
    if query and n_homologs:
        where = rng.choice(n, size=n_homologs, replace=False)
        for w in where:
            rate = rng.uniform(0.05, 0.40)
            a = rng.integers(0, max(1, len(query) // 5))
            b = len(query) - rng.integers(0, max(1, len(query) // 5))
            seqs[w] = mutate(rng, query[a:b], rate, NT)
    return seqs


def random_nucleotide(seed, length, iupac_rate=0.02):
    rng = np.random.default_rng(seed)
    s = np.array(list("ACGT"))[rng.integers(0, 4, size=length)]
    if iupac_rate > 0:
        mask = rng.random(length) < iupac_rate
        s[mask] = np.array(list("MRSVWYHKDBN"))[rng.integers(0, 11, size=int(mask.sum()))]
    return "".join(s)


def write_fasta(path, seqs, prefix="seq"):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">%s%d\n" % (prefix, i))
            for k in range(0, len(s), 80):
                f.write(s[k:k + 80])
                f.write("\n")
            if len(s) == 0:
                f.write("\n")


# ---- large databases: chunked, multi-process, deterministic per (seed, chunk) ----

_CHUNK = 50000


def _protein_chunk(args):
    n, seed, chunk_idx, query, homolog_frac, n_dup = args
    seqs = protein_db(n, [seed, chunk_idx], query=query, homolog_frac=homolog_frac, n_dup=n_dup)
    base = chunk_idx * _CHUNK
    return "".join(">s%d\n%s\n" % (base + i, s) for i, s in enumerate(seqs)).encode(), sum(len(s) for s in seqs)


def write_protein_db(path, n, seed, query=None, homolog_frac=0.01, n_dup=50, workers=None):
    """Writes an n-record synthetic protein FASTA file (one line per sequence); returns the residue count.
    Record i lives in chunk i // 50000, generated from the RNG stream (seed, chunk)."""
    import multiprocessing as mp
    jobs = []
    for c in range((n + _CHUNK - 1) // _CHUNK):
        m = min(_CHUNK, n - c * _CHUNK)
        jobs.append((m, seed, c, query, homolog_frac, n_dup if m > 2 * n_dup else 0))
    workers = workers or min(len(jobs), max(1, (os.cpu_count() or 2) - 1))
    total = 0
    with open(path, "wb") as f:
        if workers > 1:
            with mp.get_context("fork").Pool(workers) as pool:
                for blob, res in pool.imap(_protein_chunk, jobs):
                    f.write(blob)
                    total += res
        else:
            for j in jobs:
                blob, res = _protein_chunk(j)
                f.write(blob)
                total += res
    return total
