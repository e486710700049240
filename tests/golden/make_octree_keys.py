"""Convert the reference's hash-table fixtures into a compact golden file.

Source (read-only, not copied): /root/reference/leven/src/testdata/octree_keys_*.cpp -- twelve
`~0`-terminated uint32 arrays of real octree node codes used by
leven/src/test_cuckoo.cpp:107-178 ("every key inserts, every key is found, value preserved") and
leven/src/test_compute.cpp:46-68 (dedupe of OCTREE_KEYS_3 with each key repeated 1..5 times,
as testdata/gen_duplicate_data.py generates; that generated file is not in the reference tree,
so the duplicated set is regenerated here with a fixed numpy seed).

Run in the build container only:  python tests/golden/make_octree_keys.py
"""
import glob
import os
import re

import numpy as np

REF = "/root/reference/leven/src/testdata"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "octree_keys.npz")


def main():
    arrays = {}
    for path in sorted(glob.glob(os.path.join(REF, "octree_keys_*.cpp"))):
        name = re.search(r"octree_keys_(\d+)\.cpp", path).group(1)
        txt = open(path).read()
        vals = [int(v, 16) for v in re.findall(r"0x[0-9a-fA-F]+", txt)]
        body = txt[txt.index("{"):]
        assert "~0" in body, path
        arrays["keys_" + name] = np.array(vals, dtype=np.uint32)
    rng = np.random.RandomState(3)
    k3 = arrays["keys_3"]
    reps = rng.randint(1, 6, size=len(k3))          # randint(1, 5) inclusive in the python 2 script
    arrays["keys_3_duplicated"] = np.repeat(k3, reps)
    np.savez_compressed(OUT, **arrays)
    print({k: len(v) for k, v in arrays.items()})


if __name__ == "__main__":
    main()
