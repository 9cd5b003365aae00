#!/usr/bin/env python
"""Generate tests/golden/*.dump.gz: full-precision results of the UNMODIFIED reference
(oracle/_ref/eqtlbma_bf_ref_dump, built by oracle/Makefile from /root/reference) on the seeded
synthetic scenarios of tests/scenarios.py.  Run in the build container only (needs /root/reference
to have been compiled); the fixtures it writes are committed and travel to the GPU box."""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from scenarios import CLI_SCENARIOS, SCENARIOS, build_dataset, cli_extra, dataset_digest, ref_flags  # noqa: E402


def main():
    exe = os.path.join(HERE, "_ref", "eqtlbma_bf_ref_dump")
    if not os.path.exists(exe):
        sys.exit("build oracle/_ref first: make -C oracle ref")
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    only = set(sys.argv[1:])
    meta_path = os.path.join(outdir, "manifest.json")
    meta = json.load(open(meta_path)) if os.path.exists(meta_path) else {}
    for name, sc in list(SCENARIOS.items()) + list(CLI_SCENARIOS.items()):
        if only and name not in only:
            continue
        ds = build_dataset(sc)
        tmp = tempfile.mkdtemp(prefix="golden_")
        ds.write_files(tmp)
        dump = os.path.join(tmp, "dump.txt")
        cmd = [exe] + ds.ref_args(tmp, os.path.join(tmp, "obs")) + ref_flags(sc) + cli_extra(sc, ds, tmp) + ["-v", "0"]
        env = dict(os.environ, EQTLBMA_DUMP=dump)
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout[-2000:], r.stderr[-2000:])
            sys.exit(f"reference failed on scenario {name}")
        if name in SCENARIOS:  # CLI-only scenarios keep the text outputs only
            with open(dump, "rb") as fi, gzip.GzipFile(os.path.join(outdir, name + ".dump.gz"), "wb", mtime=0) as fo:
                fo.write(fi.read())
        # the reference's own text outputs (7 significant digits), kept for the host writer tests
        texts = {}
        for fn in sorted(os.listdir(tmp)):
            if fn.startswith("obs_") and fn.endswith(".txt.gz"):
                texts[fn[4:]] = gzip.open(os.path.join(tmp, fn), "rt").read()
        with gzip.GzipFile(os.path.join(outdir, name + ".text.json.gz"), "wb", mtime=0) as fo:
            fo.write(json.dumps(texts, sort_keys=True).encode())
        meta[name] = {"digest": dataset_digest(ds), "flags": ref_flags(sc)}
        print(name, "ok", len(texts), "text outputs")
        shutil.rmtree(tmp)
    json.dump(meta, open(meta_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
