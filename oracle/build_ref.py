#!/usr/bin/env python
"""Build `oracle/_ref/`: the UNMODIFIED reference implementation of the path, compiled where its sources lie.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is pure Python (no native code), so "compiling" it
means byte-compiling the model files straight from /root/reference into sourceless code objects under `oracle/_ref/`
(`*.refbin` = the bytes of a `.pyc`; git-ignored like a built `.so`; NOT gpurun-ignored, so they travel to the GPU box, where
/root/reference does not exist - plain `.pyc` files are dropped by the snapshot).  `oracle/reference_shim.py` imports them
through a finder that maps `model.*`, `utils.testloss`, `model_dict` onto these files.
No reference source text is copied into the repository.

    python oracle/build_ref.py            # needs /root/reference (build container); a no-op message on the GPU box

Consumers: `oracle/reference_shim.py` / `oracle/baselines.py` -> `bench.py --impl reference`, `bench.py`'s cpu_baseline / gpu_eager_baseline legs,
and tests that check the oracle against the live reference.  The product package never imports it.
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("TBNS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# the files SURVEY.md §8(a) cites for the path (+ what they import)
FILES = [
    "model/Physics_Attention.py",
    "model/Embedding.py",
    "model/Transolver_Structured_Mesh_2D.py",
    "model/Transolver_Irregular_Mesh.py",
    "model/Transolver_Structured_Mesh_3D.py",
    "model/Transolver_Structured_Mesh2D_Encoder.py",
    "model/SOL_Transolver_Structured_Mesh_2D.py",
    "model_dict.py",
    "utils/testloss.py",
]


def build(verbose: bool = True) -> bool:
    if not os.path.isfile(os.path.join(REF_ROOT, FILES[0])):
        if verbose:
            print(f"oracle/build_ref: {REF_ROOT} not mounted - keeping whatever is in {OUT}")
        return os.path.isfile(os.path.join(OUT, "model", "Physics_Attention.refbin"))
    for rel in FILES:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(OUT, os.path.splitext(rel)[0] + ".refbin")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # unchecked-hash pyc: valid without the source file next to it; dfile keeps reference-relative paths in tracebacks
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True, optimize=0,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(OUT, "BUILD_INFO"), "w") as f:
        f.write(f"byte-compiled from {REF_ROOT} by oracle/build_ref.py with python {sys.version.split()[0]}\n" + "\n".join(FILES) + "\n")
    if verbose:
        print(f"oracle/build_ref: {len(FILES)} modules -> {OUT}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
