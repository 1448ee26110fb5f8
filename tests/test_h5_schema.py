"""The HDF5 branch of the sketch-database reader, against the reference's OWN writer.

h5py / libhdf5 are not in the image, so the file format itself cannot be exercised; what can be pinned is the object
schema: PopPUNK/web.py:14-61 ``sketch_to_hdf5`` (the reference's JSON -> HDF5 converter) is extracted from the reference
tree and run against an in-memory stand-in for the h5py API (groups, datasets, attrs — test infrastructure, below), on
the reference's only real pp-sketchlib sketch (test/json_sketch.txt); ``poppunk_b200.sketchlib.read_db`` then reads the
objects that writer created through the same API.  Skipped where /root/reference is absent (the GPU box)."""
import ast
import json
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"


class FakeDataset:
    def __init__(self, data, dtype=None):
        self.data, self.attrs = np.array(data, dtype=dtype), {}

    def __getitem__(self, key):
        return self.data[key]


class FakeGroup:
    def __init__(self):
        self.attrs, self.children = {}, {}

    def create_group(self, name):
        self.children[name] = FakeGroup()
        return self.children[name]

    def create_dataset(self, name, data=None, dtype=None):
        self.children[name] = FakeDataset(data, dtype)
        return self.children[name]

    def keys(self):
        return sorted(self.children)          # h5py iterates names alphabetically (PopPUNK/sketchlib.py:211)

    def __getitem__(self, name):
        return self.children[name]

    def __contains__(self, name):
        return name in self.children


class FakeH5py:
    """``h5py.File(path, mode)``: 'w' creates (and touches the path so os.path.exists sees it), 'r' reopens."""
    def __init__(self):
        self.files = {}
        outer = self

        class File(FakeGroup):
            def __new__(cls, path, mode="r"):
                if mode == "w":
                    obj = FakeGroup.__new__(cls)
                    FakeGroup.__init__(obj)
                    outer.files[os.path.abspath(path)] = obj
                    open(path, "wb").close()
                    return obj
                return outer.files[os.path.abspath(path)]

            def __init__(self, *a, **k):
                pass

            def close(self):
                pass

            def __enter__(self):
                return self

            def __exit__(self, *a):
                return False

        self.File = File


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (not on the GPU box)")
def test_read_db_reads_what_the_reference_writer_writes(tmp_path, monkeypatch, golden_dir):
    from poppunk_b200 import sketchlib
    fake = FakeH5py()
    # the reference's writer, extracted (not copied) and executed against the stand-in API
    tree = ast.parse(open(os.path.join(REF, "PopPUNK", "web.py")).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "sketch_to_hdf5")
    env = {"h5py": fake, "os": os, "sys": sys, "np": np, "json": json}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "web.py", "exec"), env)
    sketch_json = open(os.path.join(REF, "test", "json_sketch.txt")).read()
    prefix = str(tmp_path / "query_db")
    os.makedirs(prefix)
    names = env["sketch_to_hdf5"]({"sampleB": sketch_json, "sampleA": json.loads(sketch_json)}, prefix)
    assert names == ["sampleB", "sampleA"]
    # the engine's reader, through the same API
    monkeypatch.setattr(sketchlib, "h5py", fake)
    db = sketchlib.read_db(prefix)
    g = np.load(os.path.join(golden_dir, "json_sketch.npz"))
    assert db.names == ["sampleA", "sampleB"]                                     # alphabetical, like h5py keys()
    assert (db.kmers == g["kmers"]).all() and db.sketchsize64 == 156 and db.bbits == 14
    assert db.sketches.dtype == np.uint64 and db.sketches.shape == (2, len(g["kmers"]), 156 * 14)
    assert (db.sketches[0] == g["sketch"]).all() and (db.sketches[1] == g["sketch"]).all()
    assert db.random_table is None                                                # query databases carry no /random
    assert sketchlib.getSketchSize(prefix) == (156, False)
    assert (sketchlib.getKmersFromReferenceDatabase(prefix) == g["kmers"]).all()
    assert sketchlib.getSeqsInDb(prefix) == ["sampleA", "sampleB"]
    sub = sketchlib.read_db(prefix, ["sampleB"])
    assert sub.names == ["sampleB"] and sub.sketches.shape[0] == 1
