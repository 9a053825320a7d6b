"""Generate tests/golden/* from the UNMODIFIED reference (oracle/_ref/ref_driver, compiled from
/root/reference by oracle/Makefile).  Run in the authoring container only:

    python tests/golden/make_golden.py

Outputs: kat.txt (hash / canonical / OccuBin / signature known answers printed by the reference's own
functions), ra.json (CheckKmer / GetCountersForRead answers; `--only-ra` / `--only-kat` regenerate just those) and models.json (digests of header / km.bin / rest.bin and of the kmer_to_occ output
vector for every seeded case in cases.py, plus the digests of the generated database files)."""
from __future__ import annotations

import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import cases  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def make_ra(tmp: str) -> None:
    """ra.json: CKMCFile::CheckKmer / GetCountersForRead answers of the reference on the databases of cases.RA_CASES"""
    out = {}
    for name, p in cases.RA_CASES.items():
        base, sp = cases.make_ra_db(name, tmp)
        q = cases.ra_queries(sp, p["seed"] + 100)
        qf, cf = os.path.join(tmp, name + "_raq.bin"), os.path.join(tmp, name + "_rac.bin")
        q.tofile(qf)
        r1 = subprocess.run([REF, "check", base, qf, cf], capture_output=True, text=True, check=True)
        counts = np.fromfile(cf, dtype=np.uint32)
        reads = cases.ra_reads(sp, p["seed"] + 200)
        rf, of = os.path.join(tmp, name + "_reads.txt"), os.path.join(tmp, name + "_readc.bin")
        with open(rf, "wb") as f:
            f.write(b"\n".join(reads) + b"\n")
        r2 = subprocess.run([REF, "reads", base, rf, of], capture_output=True, text=True, check=True)
        rc = np.fromfile(of, dtype=np.uint32)
        out[name] = {
            "params": p, "n_kmers": int(sp.kmers.size),
            "db_md5": {"kmc_pre": cases.md5_file(base + ".kmc_pre"), "kmc_suf": cases.md5_file(base + ".kmc_suf")},
            "query_md5": hashlib.md5(q.tobytes()).hexdigest(), "check_md5": hashlib.md5(counts.tobytes()).hexdigest(),
            "check_hits": int((counts != 0).sum()), "check_sum": int(counts.astype(np.int64).sum()), "check_head": counts[:32].tolist(),
            "reads_md5": hashlib.md5(b"\n".join(reads)).hexdigest(), "read_counters_md5": hashlib.md5(rc.tobytes()).hexdigest(),
            "read_counters": int(rc.size), "read_counters_nonzero": int((rc != 0).sum()), "read_counters_sum": int(rc.astype(np.int64).sum()),
            "reference_stdout": [r1.stdout.strip(), r2.stdout.strip()],
        }
        print(name, out[name]["check_hits"], out[name]["read_counters"], out[name]["read_counters_nonzero"])
    with open(os.path.join(HERE, "ra.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def main() -> None:
    if not os.path.exists(REF):
        raise SystemExit("oracle/_ref/ref_driver missing: run `make -C oracle ref` (needs /root/reference)")
    tmp = tempfile.mkdtemp(prefix="kmx_golden_")
    if "--only-ra" in sys.argv:
        make_ra(tmp)
        return
    # ---- known answers from the reference's own Tools / OccuBin ----
    kat = os.path.join(tmp, "kat.txt")
    subprocess.run([REF, "kat", kat], check=True)
    keep = [ln for ln in open(kat) if not ln.startswith("occubin 65536 ")]
    ob = [ln for ln in open(kat) if ln.startswith("occubin 65536 ")]
    keep += ob[::97] + ob[-300:]                  # the 16-bit case is sampled to keep the fixture small
    with open(os.path.join(HERE, "kat.txt"), "w") as f:
        f.writelines(keep)
    if "--only-kat" in sys.argv:
        return
    # ---- model builds + queries ----
    out = {}
    for name, p in cases.CASES.items():
        base, sp = cases.make_case_db(name, tmp)
        mdir = os.path.join(tmp, name + "_model")
        os.makedirs(mdir, exist_ok=True)
        r = subprocess.run([REF, "build", base, mdir, str(p["ci"]), str(cases.MODEL["cs"]), str(cases.MODEL["n_hash"]), str(cases.MODEL["n_bits"])],
                           capture_output=True, text=True, check=True)
        q = cases.case_queries(sp)
        qf, of = os.path.join(tmp, name + "_q.bin"), os.path.join(tmp, name + "_occ.bin")
        q.tofile(qf)
        subprocess.run([REF, "query", mdir, qf, "31", of, "4"], capture_output=True, text=True, check=True)
        occ = np.fromfile(of, dtype=np.int32)
        # strings with N / lower case, answered by the reference's own kmer_to_occ(vector<string>)
        qa = cases.case_ascii_queries(sp)
        qaf, oaf = os.path.join(tmp, name + "_qa.bin"), os.path.join(tmp, name + "_occa.bin")
        qa.tofile(qaf)
        subprocess.run([REF, "query_ascii", mdir, qaf, "31", oaf, "4"], capture_output=True, text=True, check=True)
        occ_a = np.fromfile(oaf, dtype=np.int32)
        lst = os.path.join(tmp, name + "_list.bin")
        subprocess.run([REF, "list", base, lst], capture_output=True, text=True, check=True)
        out[name] = {
            "params": p, "model": cases.MODEL, "n_kmers": int(sp.kmers.size),
            "db_md5": {"kmc_pre": cases.md5_file(base + ".kmc_pre"), "kmc_suf": cases.md5_file(base + ".kmc_suf")},
            "listing_md5": cases.md5_file(lst),     # (u64 k-mer, u32 count) records in ReadNextKmer order
            "model_md5": {f: cases.md5_file(os.path.join(mdir, f)) for f in ("header", "km.bin", "rest.bin")},
            "model_bytes": {f: os.path.getsize(os.path.join(mdir, f)) for f in ("header", "km.bin", "rest.bin")},
            "query_md5": hashlib.md5(q.tobytes()).hexdigest(),
            "occ_md5": hashlib.md5(occ.tobytes()).hexdigest(),
            "occ_nonzero": int((occ != 0).sum()), "occ_sum": int(occ.astype(np.int64).sum()),
            "occ_head": occ[:64].tolist(),
            "ascii_query_md5": hashlib.md5(qa.tobytes()).hexdigest(), "ascii_occ_md5": hashlib.md5(occ_a.tobytes()).hexdigest(),
            "ascii_occ_nonzero": int((occ_a != 0).sum()), "ascii_differs_from_clean": int((occ_a != occ[:occ_a.size]).sum()),
            "reference_stdout_tail": r.stdout.strip().splitlines()[-1],
        }
        print(name, out[name]["model_bytes"], out[name]["occ_nonzero"])
    with open(os.path.join(HERE, "models.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    make_ra(tmp)


if __name__ == "__main__":
    main()
