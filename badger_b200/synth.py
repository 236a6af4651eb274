"""Synthetic reads / whitelists of the shapes BASELINE.json names.

Pure numpy, seeded with ``numpy.random.Generator(PCG64(seed))``; nothing here is
on the product path.  Used by ``bench.py``, the tests and ``oracle/make_golden.py``.

Model (SURVEY.md §8d "Synthetic inputs"): a whitelist of W distinct uniform-random
16-mers; the cells are the first ``n_cells`` entries of a seeded permutation of the
whitelist; every read picks a cell uniformly, appends a random 12-mer UMI, suffers
i.i.d. per-base errors with total rate ``perr`` split sub:ins:del = 1:1:1, and the
FIRST 16 bases are kept (this emulates slicing 16 bp after the R1 adapter,
reference ``barcode_extraction/barcode_callers.py:220-222``, and is why the
truncated distances of ``barcode_graph.py:243`` matter).  3 % of the reads get ``*``.

Barcodes are handled as uint32 in the reference's own 2-bit little-endian packing
(``common.py:21-25``): base i occupies bits 2i..2i+1, A0 C1 G2 T3.
"""
from __future__ import annotations

import numpy as np

from .common import sorted_unique

BC_LEN = 16
_ALPHABET = np.frombuffer(b"ACGT", dtype=np.uint8)

# name -> (reads R, n_cells, whitelist W, perr, threshold)  -- BASELINE.json configs 1..5
CONFIGS = {
    "C1": dict(reads=10_000, n_cells=5_000, whitelist=100_000, perr=0.05, threshold=1, seed=1000),
    "C2": dict(reads=1_000_000, n_cells=10_000, whitelist=3_000_000, perr=0.05, threshold=1, seed=1001),
    "C3": dict(reads=5_000_000, n_cells=5_000, whitelist=5_000, perr=0.01, threshold=1, seed=1002),
    "C4": dict(reads=20_000_000, n_cells=10_000, whitelist=3_000_000, perr=0.05, threshold=2, seed=1003),
    # C5 names ~5e7 distinct barcodes: at 1e4 reads per cell that takes a per-base error rate of 0.128 (bisected on the
    # scale-free form of the config, 1e7 reads over 1e3 cells: 4 014 / 4 552 / 4 816 / 4 946 / 5 011 distinct per cell at
    # perr 0.105 / 0.1175 / 0.1237 / 0.1269 / 0.1284).  C5r1 is the same config at ONT's 5 % (1.59e7 distinct),
    # the shape round 1 ran.
    "C5": dict(reads=100_000_000, n_cells=10_000, whitelist=3_000_000, perr=0.128, threshold=2, seed=1004),
    "C5r1": dict(reads=100_000_000, n_cells=10_000, whitelist=3_000_000, perr=0.05, threshold=2, seed=1004),
}


def rng_for(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def make_whitelist(W: int, rng: np.random.Generator) -> np.ndarray:
    """W distinct uniform-random packed 16-mers, in random (file) order."""
    got = np.empty(0, dtype=np.uint32)
    while got.size < W:
        draw = rng.integers(0, 1 << 32, size=int((W - got.size) * 1.05) + 16, dtype=np.uint64).astype(np.uint32)
        got = sorted_unique(np.concatenate([got, draw]))
    got = got[rng.permutation(got.size)[:W]]
    return np.ascontiguousarray(got)


def pick_cells(whitelist: np.ndarray, n_cells: int, rng: np.random.Generator) -> np.ndarray:
    return np.ascontiguousarray(whitelist[rng.permutation(whitelist.size)[:n_cells]])


def _simulate_chunk(cells, n, perr, rng, star_frac, umi_len):
    L = BC_LEN + umi_len
    bc = cells[rng.integers(0, cells.size, size=n)]
    src = np.empty((n, L), dtype=np.uint8)
    for i in range(BC_LEN):
        src[:, i] = (bc >> np.uint32(2 * i)) & np.uint32(3)
    src[:, BC_LEN:] = rng.integers(0, 4, size=(n, umi_len), dtype=np.uint8)
    u = rng.random((n, L), dtype=np.float32)
    ev = np.zeros((n, L), dtype=np.uint8)           # 0 keep, 1 sub, 2 del, 3 ins(before base)
    ev[u < perr] = 3
    ev[u < perr * (2.0 / 3.0)] = 2
    ev[u < perr * (1.0 / 3.0)] = 1
    subst = (src + rng.integers(1, 4, size=(n, L), dtype=np.uint8)) & 3
    insb = rng.integers(0, 4, size=(n, L), dtype=np.uint8)
    cur = np.zeros(n, dtype=np.int64)
    acc = np.zeros(n, dtype=np.uint64)
    for j in range(L):
        e = ev[:, j]
        m = (e == 3) & (cur < BC_LEN)
        sh = np.where(m, 2 * cur, 0).astype(np.uint64)
        acc |= np.where(m, insb[:, j].astype(np.uint64) << sh, np.uint64(0))
        cur += (e == 3)
        m = (e != 2) & (cur < BC_LEN)
        base = np.where(e == 1, subst[:, j], src[:, j]).astype(np.uint64)
        sh = np.where(m, 2 * cur, 0).astype(np.uint64)
        acc |= np.where(m, base << sh, np.uint64(0))
        cur += (e != 2)
    short = cur < BC_LEN                              # ≥13 deletions: pad with random bases
    if short.any():
        for idx in np.nonzero(short)[0]:
            c = int(cur[idx])
            while c < BC_LEN:
                acc[idx] |= np.uint64(int(rng.integers(0, 4)) << (2 * c))
                c += 1
    return acc.astype(np.uint32), rng.random(n) >= star_frac


def simulate_reads(cells: np.ndarray, R: int, perr: float, rng: np.random.Generator,
                   star_frac: float = 0.03, umi_len: int = 12, chunk: int = 1 << 20, workers: int = 1):
    """Return (observed uint32[R], valid bool[R]); ``valid[i]==False`` means the read got ``*``.

    workers > 1: the chunks are drawn from child generators of ``rng`` (``rng.spawn``) on a thread pool - deterministic
    for a given seed, but a DIFFERENT stream than the serial form (used for the 10^8-read config, where the serial
    synthesis takes two minutes)."""
    out = np.empty(R, dtype=np.uint32)
    valid = np.empty(R, dtype=bool)
    starts = list(range(0, R, chunk))
    if workers <= 1:
        for s in starts:
            n = min(chunk, R - s)
            out[s:s + n], valid[s:s + n] = _simulate_chunk(cells, n, perr, rng, star_frac, umi_len)
        return out, valid
    from concurrent.futures import ThreadPoolExecutor
    kids = rng.spawn(len(starts))

    def work(k):
        s = starts[k]
        n = min(chunk, R - s)
        out[s:s + n], valid[s:s + n] = _simulate_chunk(cells, n, perr, kids[k], star_frac, umi_len)

    with ThreadPoolExecutor(max_workers=workers) as pool:
        list(pool.map(work, range(len(starts))))
    return out, valid


def unrank_many(ranks: np.ndarray) -> np.ndarray:
    """uint32[n] -> numpy 'S16' array of ACGT strings (``common.py:27-38`` vectorised)."""
    ranks = np.asarray(ranks, dtype=np.uint32)
    codes = np.empty((ranks.size, BC_LEN), dtype=np.uint8)
    for i in range(BC_LEN):
        codes[:, i] = (ranks >> np.uint32(2 * i)) & np.uint32(3)
    return _ALPHABET[codes].view("S%d" % BC_LEN).reshape(-1)


def rank_many(strs) -> np.ndarray:
    """Iterable of 16-char ACGT str/bytes -> uint32[n] (``common.py:21-25`` vectorised; test helper)."""
    arr = np.asarray([s.encode() if isinstance(s, str) else s for s in strs], dtype="S%d" % BC_LEN)
    chars = arr.view(np.uint8).reshape(-1, BC_LEN)
    lut = np.full(256, 255, dtype=np.uint8)
    lut[_ALPHABET] = np.arange(4, dtype=np.uint8)
    codes = lut[chars].astype(np.uint32)
    if (codes > 3).any():
        raise KeyError("non-ACGT base")
    out = np.zeros(arr.size, dtype=np.uint32)
    for i in range(BC_LEN):
        out |= codes[:, i] << np.uint32(2 * i)
    return out


def make_dataset(name_or_cfg, reads: int | None = None, seed: int | None = None, workers: int = 1):
    """Generate (whitelist, cells, observed, valid) for a named config (optionally resized; workers: see simulate_reads)."""
    cfg = dict(CONFIGS[name_or_cfg]) if isinstance(name_or_cfg, str) else dict(name_or_cfg)
    if reads is not None:
        cfg["reads"] = reads
    if seed is not None:
        cfg["seed"] = seed
    rng = rng_for(cfg["seed"])
    wl = make_whitelist(cfg["whitelist"], rng)
    cells = pick_cells(wl, min(cfg["n_cells"], wl.size), rng)
    obs, valid = simulate_reads(cells, cfg["reads"], cfg["perr"], rng, workers=workers)
    return wl, cells, obs, valid, cfg


def write_whitelist(path: str, whitelist: np.ndarray) -> None:
    """One barcode per line with a trailing newline (the reference then holds '' in its set, badger.py:85)."""
    with open(path, "wb") as fh:
        fh.write(b"\n".join(unrank_many(whitelist).tolist()) + b"\n")


def write_extraction_tsv(path: str, observed: np.ndarray, valid: np.ndarray, rng: np.random.Generator,
                         extra17_frac: float = 0.0) -> None:
    """8-column extraction TSV (header per ``barcode_callers.py:62,119``).

    ``extra17_frac`` of the valid reads carry a 17th base, which the reference strips
    (``barcode_graph.py:195-196``, ``badger.py:108-109``).
    """
    strs = unrank_many(observed).tolist()
    n = observed.size
    add17 = rng.random(n) < extra17_frac if extra17_frac > 0 else np.zeros(n, dtype=bool)
    tail = _ALPHABET[rng.integers(0, 4, size=n)]
    with open(path, "w") as fh:
        fh.write("#read_id\tbarcode\tUMI\tBC_score\tvalid_UMI\tstrand\tpolyT_start\tR1_end\n")
        for i in range(n):
            if valid[i]:
                bc = strs[i].decode()
                if add17[i]:
                    bc += chr(tail[i])
                fh.write("read_%d\t%s\tACGTACGTACGT\t16\tTrue\t+\t40\t21\n" % (i, bc))
            else:
                fh.write("read_%d\t*\t*\t-1\tFalse\t.\t-1\t-1\n" % i)
