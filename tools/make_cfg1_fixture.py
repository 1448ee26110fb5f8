#!/usr/bin/env python
"""Make tests/golden/example_set_sketches.npz — stand-in sketches of the reference's own smoke-test genomes.

    python tools/make_cfg1_fixture.py [/root/reference]

Run HERE (the reference tree is not on the GPU box).  Reads test/example_set.tar.bz2 + test/references.txt, sketches every
assembly with tools/standin_sketcher.c (see its header: reference schema, BinDash construction, NOT pp-sketchlib's hash
values), names processed and sorted as PopPUNK/utils.py:453-472 does, and stores sketchsize64 = 16 (sketch size 1024,
BASELINE config 1) for k = 13..29 step 4 (PopPUNK's defaults, __main__.py:77-79) and step 3 up to 28 (the k range of the
reference's smoke test, test/run_test.py:21).  Only numeric sketches are saved — no sequence, no reference source.
"""
import os
import subprocess
import sys
import tarfile
import tempfile

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SS64 = 16
KMERS = sorted(set(range(13, 30, 4)) | set(range(13, 29, 3)))


def main():
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "standin_sketcher")
        subprocess.run(["/usr/bin/gcc", "-O2", "-o", exe, os.path.join(ROOT, "tools", "standin_sketcher.c")], check=True)
        with tarfile.open(os.path.join(REF, "test", "example_set.tar.bz2")) as tar:
            tar.extractall(tmp, filter="data")
        names, files = [], []
        for line in open(os.path.join(REF, "test", "references.txt")):
            fields = line.rstrip().split("\t")
            if len(fields) >= 2:
                names.append(fields[0])
                files.append(fields[1])
        # PopPUNK/utils.py:453, 473-488 isolateNameToLabel; :465-470 sorted on return
        names = [n.split("/")[-1].replace(".", "_").replace(":", "").replace("(", "_").replace(")", "_") for n in names]
        order = sorted(range(len(names)), key=lambda i: names[i])
        names, files = [names[i] for i in order], [files[i] for i in order]
        W = SS64 * 14
        sk = np.empty((len(names), len(KMERS), W), dtype=np.uint64)
        lengths = np.zeros(len(names), dtype=np.int64)
        for i, fa in enumerate(files):
            path = os.path.join(tmp, fa)
            raw = subprocess.run([exe, path, str(SS64), ",".join(map(str, KMERS))], check=True, capture_output=True).stdout
            sk[i] = np.frombuffer(raw, dtype=np.uint64).reshape(len(KMERS), W)
            lengths[i] = sum(len(l.strip()) for l in open(path) if not l.startswith(">"))
        out = os.path.join(ROOT, "tests", "golden", "example_set_sketches.npz")
        np.savez_compressed(out, names=np.array(names), kmers=np.array(KMERS, dtype=np.int32), sketchsize64=np.int32(SS64),
                            bbits=np.int32(14), sketches=sk, length=lengths)
        print(f"{out}: {len(names)} genomes, k = {KMERS}, {os.path.getsize(out) / 1e3:.0f} kB; "
              f"genome lengths {lengths.min()}..{lengths.max()}")


if __name__ == "__main__":
    main()
