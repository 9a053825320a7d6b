"""Synthetic KMC databases for tests and benchmarks (support code, not the hot path).

The reference's counting stage is the external `kmc` binary (main.cpp:136-140), which is a
missing blob, so databases are produced here: a seeded random genome, reads sampled at a
given coverage with substitution errors, canonical k-mer counting (sort + run lengths),
and a writer for the `.kmc_pre/.kmc_suf` layout the reference reader accepts
(kmc_file.cpp:177-235, 428-515).

All randomness is a counter-based splitmix64 evaluated with wrapping int64 tensor
arithmetic, so the same seed gives the same database on CPU and on CUDA. torch is used as
an array library only (sort / unique on whichever device is available).
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np
import torch

_M64 = (1 << 64) - 1


def _i64(x: int) -> int:
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    """logical shift right on int64 tensors"""
    return (x >> s) & _i64((1 << (64 - s)) - 1)


def splitmix64(idx: torch.Tensor, seed: int, stream: int) -> torch.Tensor:
    """counter-based RNG: 64 random bits per index (int64 bit patterns)"""
    z = idx + _i64(seed * 0x9E3779B97F4A7C15 + stream * 0xD1B54A32D192ED03 + 0x632BE59BD9B4E019)
    z = z * _i64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def _uniform_int(idx: torch.Tensor, seed: int, stream: int, n: int) -> torch.Tensor:
    """integers in [0, n) (n < 2**31 keeps the tiny modulo bias irrelevant)"""
    return _lsr(splitmix64(idx, seed, stream), 1) % n


def _geometric_tail(x: torch.Tensor, dtype) -> torch.Tensor:
    """floor(log(x) / log(0.45)) + 1 clamped to 1..8 for x in (0, 1], i.e. the j with 0.45^j < x <= 0.45^(j-1),
    evaluated by comparing against the eight thresholds (exact on every device, unlike log())"""
    c = torch.ones(x.shape, dtype=torch.int64, device=x.device)
    for j in range(1, 8):
        c += (x <= torch.tensor(0.45 ** j, dtype=dtype, device=x.device)).to(torch.int64)
    return c


def revcomp_packed(v: torch.Tensor, k: int) -> torch.Tensor:
    """reverse complement of 2-bit packed k-mers held in int64 (k <= 31)"""
    x = ~v
    x = (_lsr(x, 2) & 0x3333333333333333) | ((x & 0x3333333333333333) << 2)
    x = (_lsr(x, 4) & 0x0F0F0F0F0F0F0F0F) | ((x & 0x0F0F0F0F0F0F0F0F) << 4)
    x = (_lsr(x, 8) & 0x00FF00FF00FF00FF) | ((x & 0x00FF00FF00FF00FF) << 8)
    x = (_lsr(x, 16) & 0x0000FFFF0000FFFF) | ((x & 0x0000FFFF0000FFFF) << 16)
    x = _lsr(x, 32) | (x << 32)
    return _lsr(x, 64 - 2 * k)


def canonical_packed(v: torch.Tensor, k: int) -> torch.Tensor:
    assert k <= 31
    return torch.minimum(v, revcomp_packed(v, k))


@dataclass
class Spectrum:
    k: int
    kmers: np.ndarray   # uint64, sorted ascending, canonical, unique
    counts: np.ndarray  # uint32
    genome: np.ndarray  # uint8 codes (for neighbour-rich query sets)


def genome_kmers(genome: torch.Tensor, k: int) -> torch.Tensor:
    """packed forward k-mer starting at every genome position (int64, len G-k+1)"""
    g = genome.to(torch.int64)
    n = g.numel() - k + 1
    out = torch.zeros(n, dtype=torch.int64, device=g.device)
    for j in range(k):
        out |= g[j:j + n] << (2 * (k - 1 - j))
    return out


def synth_reads_spectrum(genome_bp: int, coverage: float, read_len: int, k: int = 31, err_rate: float = 0.01,
                         seed: int = 1, ci: int = 1, cs: int = 1023, repeat_frac: float = 0.01,
                         device: str | None = None, chunk_reads: int = 1 << 20) -> Spectrum:
    """Simulate reads from a random genome and count canonical k-mers (counts saturate at cs,
    k-mers below ci dropped) -- what `kmc -k -ci -cs` would emit for such a FASTQ."""
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    G, L = int(genome_bp), int(read_len)
    pos = torch.arange(G, dtype=torch.int64, device=dev)
    genome = (splitmix64(pos, seed, 1) & 3)
    # a few repeated segments so that some k-mers have genomic multiplicity > 1
    n_rep = int(G * repeat_frac / 500)
    if n_rep > 0 and G > 4000:
        ridx = torch.arange(n_rep, dtype=torch.int64, device=dev)
        src = _uniform_int(ridx, seed, 2, G - 1000)
        dst = _uniform_int(ridx, seed, 3, G - 1000)
        off = torch.arange(500, dtype=torch.int64, device=dev)
        # every segment copies from the ORIGINAL genome; where destination segments overlap the segment with the
        # highest index wins (what a sequential scatter does) -- decided with an integer max-reduction, so the
        # result does not depend on the device or on the order a parallel scatter happens to take
        winner = torch.full((G,), -1, dtype=torch.int64, device=dev)
        winner.scatter_reduce_(0, (dst[:, None] + off[None, :]).reshape(-1), ridx[:, None].expand(-1, 500).reshape(-1), "amax")
        at = torch.nonzero(winner >= 0).squeeze(1)
        w = winner[at]
        genome[at] = genome.clone()[src[w] + (at - dst[w])]
    gk = genome_kmers(genome, k)
    per_read = L - k + 1
    n_reads = int(G * coverage / L)
    uniq_parts, cnt_parts = [], []
    joff = torch.arange(per_read, dtype=torch.int64, device=dev)
    for r0 in range(0, n_reads, chunk_reads):
        r1 = min(n_reads, r0 + chunk_reads)
        ridx = torch.arange(r0, r1, dtype=torch.int64, device=dev)
        start = _uniform_int(ridx, seed, 4, G - L + 1)
        inst = gk[(start[:, None] + joff[None, :]).reshape(-1)]
        # substitution errors: events (read, position, delta) at rate err_rate per base
        n_ev = int(round((r1 - r0) * L * err_rate))
        if n_ev > 0:
            eidx = torch.arange(n_ev, dtype=torch.int64, device=dev) + r0 * 1000003
            er = _uniform_int(eidx, seed, 5, r1 - r0)
            eq = _uniform_int(eidx, seed, 6, L)
            ed = _uniform_int(eidx, seed, 7, 3) + 1
            key = torch.unique(er * L + eq, sorted=True, return_inverse=False)   # one event per (read, pos)
            er, eq = key // L, key % L
            ed = _uniform_int(key + r0 * L, seed, 7, 3) + 1
            j = eq[:, None] - torch.arange(k, dtype=torch.int64, device=dev)[None, :]      # k-mer offsets covering pos
            ok = (j >= 0) & (j < per_read)
            sh = 2 * (k - 1 - (eq[:, None] - j))
            mask = (ed[:, None] << sh)[ok]
            tgt = (er[:, None] * per_read + j)[ok]
            acc = torch.zeros_like(inst)
            acc.index_add_(0, tgt, mask)        # disjoint bit fields: add == or
            inst = inst ^ acc
        inst = canonical_packed(inst, k)
        u, c = torch.unique(inst, sorted=True, return_counts=True)
        uniq_parts.append(u)
        cnt_parts.append(c)
        del inst
    u = torch.cat(uniq_parts)
    c = torch.cat(cnt_parts)
    if len(uniq_parts) > 1:
        order = torch.argsort(u, stable=True)
        u, c = u[order], c[order]
        uu, inv = torch.unique_consecutive(u, return_inverse=True)
        cc = torch.zeros(uu.numel(), dtype=torch.int64, device=dev)
        cc.index_add_(0, inv, c)
        u, c = uu, cc
    c = torch.clamp(c, max=cs)
    keep = c >= ci
    u, c = u[keep], c[keep]
    return Spectrum(k=k, kmers=u.cpu().numpy().astype(np.uint64), counts=c.cpu().numpy().astype(np.uint32),
                    genome=genome.to(torch.uint8).cpu().numpy())


def synth_direct_spectrum(genome_bp: int, coverage: float, read_len: int, k: int = 31, err_mult: float = 1.5,
                          seed: int = 1, ci: int = 1, cs: int = 1023, device: str | None = None) -> Spectrum:
    """Cheaper genome-derived spectrum for large shapes: every genomic k-mer gets a count around
    coverage*(L-k+1)/L, plus err_mult single-substitution neighbours per genomic k-mer with a
    short low-count tail (what sequencing errors leave behind)."""
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    G = int(genome_bp)
    pos = torch.arange(G, dtype=torch.int64, device=dev)
    genome = (splitmix64(pos, seed, 1) & 3)
    gk = genome_kmers(genome, k)
    n = gk.numel()
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    lam = coverage * (read_len - k + 1) / read_len
    # sum of 8 uniforms ~ normal: mean lam, sd sqrt(lam)
    s = torch.zeros(n, dtype=torch.float64, device=dev)
    for t in range(8):
        s += (_lsr(splitmix64(idx, seed, 10 + t), 11).to(torch.float64) / float(1 << 53))
    solid_c = torch.clamp((lam + (s - 4.0) * (lam ** 0.5) * 1.2247).round().to(torch.int64), min=1)
    n_err = int(n * err_mult)
    eidx = torch.arange(n_err, dtype=torch.int64, device=dev)
    src = gk[_uniform_int(eidx, seed, 20, n)]
    epos = _uniform_int(eidx, seed, 21, k)
    ed = _uniform_int(eidx, seed, 22, 3) + 1
    err = src ^ (ed << (2 * epos))
    r = _lsr(splitmix64(eidx, seed, 23), 11).to(torch.float64) / float(1 << 53)
    err_c = _geometric_tail(1 - r, torch.float64)                                                # geometric tail 1,2,3..
    allk = canonical_packed(torch.cat([gk, err]), k)
    allc = torch.cat([solid_c, err_c])
    order = torch.argsort(allk, stable=True)
    allk, allc = allk[order], allc[order]
    u, inv = torch.unique_consecutive(allk, return_inverse=True)
    c = torch.zeros(u.numel(), dtype=torch.int64, device=dev)
    c.index_add_(0, inv, allc)
    c = torch.clamp(c, max=cs)
    keep = c >= ci
    u, c = u[keep], c[keep]
    return Spectrum(k=k, kmers=u.cpu().numpy().astype(np.uint64), counts=c.cpu().numpy().astype(np.uint32),
                    genome=genome.to(torch.uint8).cpu().numpy())


def kmc_signatures(kmers: np.ndarray, k: int, signature_len: int) -> np.ndarray:
    """KMC's signature of every packed k-mer: the smallest normalised m-mer (m = signature_len); an m-mer normalises to the
    smaller of itself and its reverse complement among those that are allowed, to 4^m when neither is.  Not allowed: a TTT,
    TGT or TT? ending, an ACA beginning, AA anywhere but in front.  (What a real KMC run bins its records by; used to write
    test databases that CKMCFile::CheckKmer can search, kmc_file.cpp:339-340.)"""
    m = signature_len
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
    none = np.uint32(1 << (2 * m))

    def allowed(x: np.ndarray) -> np.ndarray:
        ok = ((x & 0x3F) != 0x3F) & ((x & 0x3F) != 0x3B) & ((x & 0x3C) != 0x3C)
        for j in range(m - 2):                                  # bases j and j + 1 (from the end) both A
            ok &= ((x >> (2 * j)) & 0xF) != 0
        ok &= (x >> (2 * (m - 3))) != 4
        return ok

    best = np.full(kmers.size, 0xFFFFFFFF, dtype=np.uint32)
    for i in range(k - m + 1):
        x = ((kmers >> np.uint64(2 * (k - m - i))) & np.uint64((1 << (2 * m)) - 1)).astype(np.uint32)
        rc = np.zeros_like(x)
        t = x.copy()
        for _ in range(m):
            rc = (rc << 2) | (3 - (t & 3))
            t >>= 2
        cand = np.minimum(np.where(allowed(x), x, none), np.where(allowed(rc), rc, none))
        best = np.minimum(best, cand)
    return best


def write_kmc_db(base: str, kmers: np.ndarray, counts: np.ndarray, k: int = 31, lut_prefix_length: int = 3,
                 n_bins: int = 1, counter_size: int = 2, min_count: int = 1, max_count: int = 1023,
                 signature_len: int = 7, signature_bins: bool = False, one_strand: bool = False) -> int:
    """Write <base>.kmc_pre / <base>.kmc_suf (KMC2/3 layout, version word 0x200).

    kmers must be unique packed values; they are split over n_bins by a hash of the value (KMC
    bins by minimiser signature; the listing reader only needs sorted records per bin and a
    per-bin LUT, kmc_file.cpp:439-449) and sorted within each bin. Returns the record count.
    signature_bins: bin by KMC signature and write the signature map, as a real KMC run does, so that the
    random-access API (CheckKmer) finds the records; one_strand sets the header's "forward strand only" byte."""
    assert (k - lut_prefix_length) % 4 == 0 and 1 <= k <= 32
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    n = kmers.size
    suf_bytes = (k - lut_prefix_length) // 4
    sig_map = np.zeros(4 ** signature_len + 1, dtype=np.uint32)
    if n >= (1 << 23) and torch.cuda.is_available() and not signature_bins and not one_strand:
        return _write_kmc_db_cuda(base, kmers, counts, k, lut_prefix_length, n_bins, counter_size, min_count, max_count, signature_len)
    if signature_bins:
        all_sigs = np.arange(4 ** signature_len + 1, dtype=np.uint64)
        sig_map = (((all_sigs * np.uint64(2654435761)) >> np.uint64(9)) % np.uint64(n_bins)).astype(np.uint32)
        bins = sig_map[kmc_signatures(kmers, k, signature_len)].astype(np.int64)
    elif n_bins > 1:
        h = (kmers * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(40)
        bins = (h % np.uint64(n_bins)).astype(np.int64)
    else:
        bins = np.zeros(n, dtype=np.int64)
    order = np.lexsort((kmers, bins))
    kmers, counts, bins = kmers[order], counts[order], bins[order]
    slots = 4 ** lut_prefix_length
    prefix = (kmers >> np.uint64(8 * suf_bytes)).astype(np.int64)
    slot_id = bins * slots + prefix
    per_slot = np.bincount(slot_id, minlength=n_bins * slots).astype(np.uint64)
    starts = np.zeros(n_bins * slots + 1, dtype=np.uint64)
    np.cumsum(per_slot, out=starts[1:])
    lut = starts.copy()          # n_bins*slots entries + the guard word
    rec = np.zeros((n, suf_bytes + counter_size), dtype=np.uint8)
    for b in range(suf_bytes):
        rec[:, b] = ((kmers >> np.uint64(8 * (suf_bytes - 1 - b))) & np.uint64(0xFF)).astype(np.uint8)
    for b in range(counter_size):
        rec[:, suf_bytes + b] = ((counts >> np.uint32(8 * b)) & np.uint32(0xFF)).astype(np.uint8)
    with open(base + ".kmc_suf", "wb") as f:
        f.write(b"KMCS")
        f.write(rec.tobytes())
        f.write(b"KMCS")
    header = struct.pack("<7IQB7x5I I", k, 0, counter_size, lut_prefix_length, signature_len, min_count, max_count,
                         n, 1 if one_strand else 0, 0, 0, 0, 0, 0, 0x200)
    with open(base + ".kmc_pre", "wb") as f:
        f.write(b"KMCP")
        f.write(lut.tobytes())
        f.write(sig_map.tobytes())
        f.write(header)
        f.write(struct.pack("<I", len(header)))
        f.write(b"KMCP")
    return n


def _write_kmc_db_cuda(base, kmers, counts, k, lut_prefix_length, n_bins, counter_size, min_count, max_count, signature_len) -> int:
    """same files as write_kmc_db, with the sort and the record packing done by torch on the GPU (large shapes)"""
    dev = torch.device("cuda")
    n = kmers.size
    suf_bytes = (k - lut_prefix_length) // 4
    km = torch.from_numpy(kmers.view(np.int64)).to(dev)
    ct = torch.from_numpy(counts.astype(np.int64)).to(dev)
    if n_bins > 1:
        h = _lsr(km * _i64(0x9E3779B97F4A7C15), 40)
        bins = h % n_bins
    else:
        bins = torch.zeros(n, dtype=torch.int64, device=dev)
    order = torch.argsort(km, stable=True)
    km, ct, bins = km[order], ct[order], bins[order]
    order = torch.argsort(bins, stable=True)
    km, ct, bins = km[order], ct[order], bins[order]
    del order
    slots = 4 ** lut_prefix_length
    prefix = _lsr(km, 8 * suf_bytes) if suf_bytes else km
    per_slot = torch.bincount(bins * slots + prefix, minlength=n_bins * slots)
    starts = torch.zeros(n_bins * slots + 1, dtype=torch.int64, device=dev)
    starts[1:] = torch.cumsum(per_slot, 0)
    lut = starts.cpu().numpy().astype(np.uint64)
    with open(base + ".kmc_suf", "wb") as f:
        f.write(b"KMCS")
        step = 1 << 25
        for a in range(0, n, step):
            b = min(n, a + step)
            rec = torch.empty((b - a, suf_bytes + counter_size), dtype=torch.uint8, device=dev)
            for j in range(suf_bytes):
                rec[:, j] = (_lsr(km[a:b], 8 * (suf_bytes - 1 - j)) & 0xFF).to(torch.uint8)
            for j in range(counter_size):
                rec[:, suf_bytes + j] = ((ct[a:b] >> (8 * j)) & 0xFF).to(torch.uint8)
            f.write(rec.cpu().numpy().tobytes())
        f.write(b"KMCS")
    header = struct.pack("<7IQB7x5I I", k, 0, counter_size, lut_prefix_length, signature_len, min_count, max_count,
                         n, 0, 0, 0, 0, 0, 0, 0x200)
    with open(base + ".kmc_pre", "wb") as f:
        f.write(b"KMCP")
        f.write(lut.tobytes())
        f.write(np.zeros(4 ** signature_len + 1, dtype=np.uint32).tobytes())
        f.write(header)
        f.write(struct.pack("<I", len(header)))
        f.write(b"KMCP")
    return n


def _kmers_at(genome: torch.Tensor, a: int, b: int, k: int) -> torch.Tensor:
    """packed forward k-mers starting at genome positions [a, b) (genome = uint8 codes)"""
    g = genome[a:b + k - 1].to(torch.int64)
    n = b - a
    out = torch.zeros(n, dtype=torch.int64, device=g.device)
    for j in range(k):
        out |= g[j:j + n] << (2 * (k - 1 - j))
    return out


def make_db_streamed(base: str, genome_bp: int, coverage: float, read_len: int, k: int = 31, err_frac: float = 0.3, seed: int = 1,
                     ci: int = 2, cs: int = 1023, lut_prefix_length: int = 7, n_bins: int = 512, n_groups: int = 16,
                     chunk: int = 1 << 26, n_present: int = 0, device: str | None = None) -> dict:
    """Genome-derived spectrum for shapes too large for the in-memory generators (NA12878-shaped: 3.1 Gbp).

    Every genomic k-mer gets a count around coverage*(L-k+1)/L; a fraction err_frac of the positions
    also contributes a single-substitution neighbour with a short low-count tail.  The (k-mer, count)
    pairs stay on the device (10 bytes each); the database is written bin group by bin group (sort,
    merge duplicates, -ci filter, pack records, append to .kmc_suf), so neither the host nor the
    device ever holds more than one group besides the pair buffer.  Returns a dict with the record
    count and, if n_present > 0, `present`: that many k-mers sampled uniformly from the records."""
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    G = int(genome_bp)
    assert (k - lut_prefix_length) % 4 == 0 and k <= 31 and n_bins % n_groups == 0
    genome = torch.empty(G, dtype=torch.uint8, device=dev)
    for a in range(0, G, chunk):
        b = min(G, a + chunk)
        genome[a:b] = (splitmix64(torch.arange(a, b, dtype=torch.int64, device=dev), seed, 1) & 3).to(torch.uint8)
    n_pos = G - k + 1
    lam = coverage * (read_len - k + 1) / read_len
    thr = int(err_frac * 65536)
    cap = n_pos + int(n_pos * err_frac * 1.02) + 4096 * (n_pos // chunk + 1)
    allk = torch.empty(cap, dtype=torch.int64, device=dev)
    allc = torch.empty(cap, dtype=torch.int16, device=dev)
    n_all = 0
    for a in range(0, n_pos, chunk):
        b = min(n_pos, a + chunk)
        idx = torch.arange(a, b, dtype=torch.int64, device=dev)
        gk = _kmers_at(genome, a, b, k)
        r = splitmix64(idx, seed, 10)
        # four 16-bit uniforms: Irwin-Hall(4), mean 2, variance 1/3 -> roughly normal around lam, sd sqrt(lam)
        s4 = ((r & 0xFFFF) + (_lsr(r, 16) & 0xFFFF) + (_lsr(r, 32) & 0xFFFF) + _lsr(r, 48)).to(torch.float32) / 65536.0
        solid_c = torch.clamp((lam + (s4 - 2.0) * (3.0 * lam) ** 0.5).round().to(torch.int64), min=1, max=cs)
        r2 = splitmix64(idx, seed, 20)
        has_err = (r2 & 0xFFFF) < thr
        src = gk[has_err]
        r2 = r2[has_err]
        epos = (_lsr(r2, 16) & 0xFFFF) % k
        ed = (_lsr(r2, 32) & 0xFFFF) % 3 + 1
        err = src ^ (ed << (2 * epos))
        u = (_lsr(r2, 40) & 0xFFFFFF).to(torch.float32) / float(1 << 24)
        err_c = _geometric_tail(1 - u, torch.float32)
        m = gk.numel() + err.numel()
        assert n_all + m <= cap
        allk[n_all:n_all + m] = canonical_packed(torch.cat([gk, err]), k)
        allc[n_all:n_all + m] = torch.cat([solid_c, err_c]).to(torch.int16)
        n_all += m
        del gk, r, s4, solid_c, r2, has_err, src, epos, ed, err, u, err_c, idx
    del genome
    allk, allc = allk[:n_all], allc[:n_all]
    bins = torch.empty(n_all, dtype=torch.int16, device=dev)
    for a in range(0, n_all, chunk):
        b = min(n_all, a + chunk)
        bins[a:b] = (_lsr(allk[a:b] * _i64(0x9E3779B97F4A7C15), 40) % n_bins).to(torch.int16)
    suf_bytes = (k - lut_prefix_length) // 4
    counter_size = 2
    slots = 4 ** lut_prefix_length
    per_slot = torch.zeros(n_bins * slots, dtype=torch.int64, device=dev)
    per_group = n_bins // n_groups
    total = 0
    present = []
    os.makedirs(os.path.dirname(os.path.abspath(base)), exist_ok=True)
    with open(base + ".kmc_suf", "wb") as f:
        f.write(b"KMCS")
        for g in range(n_groups):
            pk, pc, pb = [], [], []
            for a in range(0, n_all, 1 << 28):                 # selection in pieces: index tensors stay below 2^31 elements
                b = min(n_all, a + (1 << 28))
                bsl = bins[a:b]
                msk = (bsl >= g * per_group) & (bsl < (g + 1) * per_group)
                pk.append(allk[a:b][msk])
                pc.append(allc[a:b][msk])
                pb.append(bsl[msk])
                del msk, bsl
            km, ct, bn = torch.cat(pk), torch.cat(pc).to(torch.int64), torch.cat(pb).to(torch.int64)
            del pk, pc, pb
            km, order = torch.sort(km, stable=True)
            ct, bn = ct[order], bn[order]
            del order
            km, inv = torch.unique_consecutive(km, return_inverse=True)
            if km.numel() != ct.numel():
                cc = torch.zeros(km.numel(), dtype=torch.int64, device=dev)
                cc.index_add_(0, inv, ct)
                first = torch.ones_like(inv, dtype=torch.bool)
                first[1:] = inv[1:] != inv[:-1]
                ct, bn = cc, bn[first]
                del cc, first
            del inv
            ct = torch.clamp(ct, max=cs)
            keep = ct >= ci
            km, ct, bn = km[keep], ct[keep], bn[keep]
            del keep
            bn, order = torch.sort(bn, stable=True)            # records sorted by k-mer within each bin
            km, ct = km[order], ct[order]
            del order
            n = km.numel()
            prefix = _lsr(km, 8 * suf_bytes)
            per_slot += torch.bincount(bn * slots + prefix, minlength=n_bins * slots)
            del prefix, bn
            if n_present:
                want = (n_present + n_groups - 1) // n_groups
                pick = _uniform_int(torch.arange(want, dtype=torch.int64, device=dev), seed, 40 + g, max(n, 1))
                present.append(km[pick])
            step = 1 << 25
            for a in range(0, n, step):
                b = min(n, a + step)
                rec = torch.empty((b - a, suf_bytes + counter_size), dtype=torch.uint8, device=dev)
                for j in range(suf_bytes):
                    rec[:, j] = (_lsr(km[a:b], 8 * (suf_bytes - 1 - j)) & 0xFF).to(torch.uint8)
                for j in range(counter_size):
                    rec[:, suf_bytes + j] = ((ct[a:b] >> (8 * j)) & 0xFF).to(torch.uint8)
                f.write(rec.cpu().numpy().tobytes())
                del rec
            total += n
            del km, ct
        f.write(b"KMCS")
    starts = torch.zeros(n_bins * slots + 1, dtype=torch.int64, device=dev)
    starts[1:] = torch.cumsum(per_slot, 0)
    lut = starts.cpu().numpy().astype(np.uint64)
    header = struct.pack("<7IQB7x5I I", k, 0, counter_size, lut_prefix_length, 7, ci, cs, total, 0, 0, 0, 0, 0, 0, 0x200)
    with open(base + ".kmc_pre", "wb") as f:
        f.write(b"KMCP")
        f.write(lut.tobytes())
        f.write(np.zeros(4 ** 7 + 1, dtype=np.uint32).tobytes())
        f.write(header)
        f.write(struct.pack("<I", len(header)))
        f.write(b"KMCP")
    out = {"n_kmers": int(total), "k": k}
    if n_present:
        out["present"] = torch.cat(present)[:n_present].cpu().numpy().astype(np.uint64)
    return out


def mixed_queries(present: np.ndarray, n_total: int, k: int = 31, seed: int = 7) -> np.ndarray:
    """BASELINE.json configs[4] mix: 50 % present k-mers (random strand), 37.5 % uniform random k-mers
    (absent w.p. ~1), 12.5 % single-base neighbours of present k-mers (the disambiguation path)."""
    rng = np.random.default_rng(seed)
    mask = np.uint64((1 << (2 * k)) - 1)
    n_p = n_total // 2
    n_nb = n_total // 8
    n_a = n_total - n_p - n_nb
    pres = present[rng.integers(0, present.size, n_p)]
    rc = revcomp_packed(torch.from_numpy(pres.astype(np.int64)), k).numpy().astype(np.uint64)
    pres = np.where(rng.integers(0, 2, n_p).astype(bool), rc, pres)
    absent = rng.integers(0, 1 << 62, n_a, dtype=np.uint64) & mask
    nb_src = present[rng.integers(0, present.size, n_nb)]
    nb = ((nb_src << np.uint64(2)) & mask) | rng.integers(0, 4, n_nb).astype(np.uint64)
    q = np.concatenate([pres, absent, nb])
    rng.shuffle(q)
    return q.astype(np.uint64)


def neighbour_rich_queries(sp: Spectrum, n_present: int, n_absent: int, seed: int = 7) -> np.ndarray:
    """Query set: present k-mers (random strand), uniform random k-mers (absent w.p. ~1) and
    single-base neighbours of present k-mers (drives the disambiguation slow path)."""
    rng = np.random.default_rng(seed)
    k = sp.k
    mask = np.uint64((1 << (2 * k)) - 1)
    pres = sp.kmers[rng.integers(0, sp.kmers.size, n_present)]
    flip = rng.integers(0, 2, n_present).astype(bool)
    t = torch.from_numpy(pres.astype(np.int64))
    rc = revcomp_packed(t, k).numpy().astype(np.uint64)
    pres = np.where(flip, rc, pres)
    absent = rng.integers(0, 1 << 62, n_absent, dtype=np.uint64) & mask
    nb_src = sp.kmers[rng.integers(0, sp.kmers.size, max(1, n_absent // 4))]
    nb = ((nb_src << np.uint64(2)) & mask) | rng.integers(0, 4, nb_src.size).astype(np.uint64)
    q = np.concatenate([pres, absent, nb])
    rng.shuffle(q)
    return q.astype(np.uint64)


def to_ascii(kmers: np.ndarray, k: int) -> np.ndarray:
    """(n, k) uint8 matrix of 'ACGT' characters"""
    kmers = np.asarray(kmers, dtype=np.uint64)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = np.empty((kmers.size, k), dtype=np.uint8)
    for i in range(k):
        out[:, i] = lut[((kmers >> np.uint64(2 * (k - 1 - i))) & np.uint64(3)).astype(np.int64)]
    return out


def make_db(base: str, shape: str, seed: int = 1, ci: int = 1, cs: int = 1023, lut_prefix_length: int = 3,
            n_bins: int = 1, device: str | None = None) -> Spectrum:
    """named shapes (BASELINE.json configs, plus small ones for tests)"""
    shapes = {
        # name: (genome_bp, coverage, read_len, mode)
        "tiny": (20_000, 30, 100, "reads"),
        "small": (200_000, 40, 100, "reads"),
        "cfg1": (2_000_000, 50, 100, "reads"),          # 1 M reads x 100 bp
        "rs": (4_600_000, 100, 101, "reads"),          # GAGE R. sphaeroides shaped
        "hc14": (88_000_000, 40, 101, "direct"),
        "wgs350": (350_000_000, 30, 101, "direct"),     # a quarter-gigabase genome: > 2^28 array k-mers, HBM-resident everything
        "na12878": (3_100_000_000, 30, 101, "direct"),
    }
    g, c, l, mode = shapes[shape]
    if mode == "reads":
        sp = synth_reads_spectrum(g, c, l, seed=seed, ci=ci, cs=cs, device=device)
    else:
        sp = synth_direct_spectrum(g, c, l, seed=seed, ci=ci, cs=cs, device=device)
    os.makedirs(os.path.dirname(os.path.abspath(base)), exist_ok=True)
    write_kmc_db(base, sp.kmers, sp.counts, k=sp.k, lut_prefix_length=lut_prefix_length, n_bins=n_bins,
                 min_count=ci, max_count=cs)
    return sp


def synth_fastq(path: str, genome_bp: int = 20_000, coverage: float = 20.0, read_len: int = 100, k: int = 31, seed: int = 1,
                err_rate: float = 0.01, n_rate: float = 0.002, crlf: bool = False):
    """Write a small 4-line FASTQ file of simulated reads (substitution errors, a few N, random case is NOT
    used) and return the canonical k-mer spectrum a counter must find in it: (kmers uint64 sorted, counts int64).
    Pure numpy: meant for tests of the k-mer counting stage."""
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, genome_bp, dtype=np.uint8)
    n_reads = int(genome_bp * coverage / read_len)
    starts = rng.integers(0, genome_bp - read_len + 1, n_reads)
    lens = np.where(rng.random(n_reads) < 0.05, rng.integers(k - 5, read_len, n_reads), read_len)      # a few short reads
    reads = []
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    eol = "\r\n" if crlf else "\n"
    all_k, mask = [], np.uint64((1 << (2 * k)) - 1)
    with open(path, "w", newline="") as f:
        for r in range(n_reads):
            codes = genome[starts[r]:starts[r] + lens[r]].copy()
            if rng.random() < 0.5:                       # reverse strand
                codes = (3 - codes)[::-1]
            err = rng.random(codes.size) < err_rate
            codes[err] = (codes[err] + rng.integers(1, 4, int(err.sum()))) % 4
            seq = alphabet[codes].copy()
            isn = rng.random(codes.size) < n_rate
            seq[isn] = ord("N")
            f.write(f"@read{r}{eol}{seq.tobytes().decode()}{eol}+{eol}{'I' * codes.size}{eol}")
            # expected k-mers of this read
            if codes.size >= k:
                valid = ~isn
                v = np.zeros(codes.size - k + 1, dtype=np.uint64)
                ok = np.ones(codes.size - k + 1, dtype=bool)
                for j in range(k):
                    v = (v << np.uint64(2)) | codes[j:j + v.size].astype(np.uint64)
                    ok &= valid[j:j + v.size]
                v = v[ok] & mask
                if v.size:
                    rc = revcomp_packed(torch.from_numpy(v.astype(np.int64)), k).numpy().astype(np.uint64)
                    all_k.append(np.minimum(v, rc))
    allk = np.concatenate(all_k) if all_k else np.zeros(0, dtype=np.uint64)
    u, c = np.unique(allk, return_counts=True)
    return u, c.astype(np.int64), n_reads
