#!/usr/bin/env python
"""Build recipes for the parity oracle.  TEST INFRASTRUCTURE ONLY.

Nothing in here is product code: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may use what this
script builds.  The product (``umpa_b200``) never imports ``oracle``.

Two artefacts:

``oracle/libumpa_oracle.so``
    gcc build of ``oracle/umpa_oracle.c`` -- our plain-C restatement of the
    reference algorithm (UMPA/lib/Model.cpp, Optim.cpp, Utils.cpp and the
    pixel loop of UMPA/model.pyx).  Buildable everywhere (also on the GPU box).

``oracle/_ref/model.<EXT_SUFFIX>``
    the UNMODIFIED reference extension, compiled from the sources where they
    lie under /root/reference (never copied into the repo): ``cython --cplus``
    on UMPA/model.pyx, then one g++ call with the reference's own flags
    (setup.py:26) except ``-march=native`` -> ``-march=x86-64-v3`` so that the
    binary built in this container also runs on the GPU box's host CPU.
    ``python setup.py build_ext`` itself fails under Cython 3 (implicit
    relative cimports in model.pyx:22-24), hence the direct recipe.
    Only possible where /root/reference exists; the GPU box uses the prebuilt
    file (``oracle/_ref/`` is git-ignored but travels with gpurun).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("UMPA_REFERENCE_DIR", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")
PORT_SO = os.path.join(HERE, "libumpa_oracle.so")
REF_SO = os.path.join(REF_OUT, "model" + sysconfig.get_config_var("EXT_SUFFIX"))


def _newer(target, *sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_port(force=False, quiet=True):
    """gcc -> oracle/libumpa_oracle.so (the C restatement)."""
    src = os.path.join(HERE, "umpa_oracle.c")
    if not force and _newer(PORT_SO, src):
        return PORT_SO
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp",
           "-fno-fast-math", "-Wall", "-Wno-unknown-pragmas",
           src, "-o", PORT_SO, "-lm"]
    if not quiet:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return PORT_SO


def ref_available():
    return os.path.exists(REF_SO)


def build_ref(force=False, quiet=True):
    """cython + g++ -> oracle/_ref/model.*.so (the unmodified reference).

    Returns the path, or None when /root/reference is absent (GPU box) and no
    prebuilt file exists."""
    pyx = os.path.join(REF_SRC, "UMPA", "model.pyx")
    if not os.path.exists(pyx):
        return REF_SO if os.path.exists(REF_SO) else None
    if not force and os.path.exists(REF_SO):
        return REF_SO
    import numpy as np
    os.makedirs(REF_OUT, exist_ok=True)
    cpp = os.path.join(REF_OUT, "umpa_ref_model.cpp")
    inc = os.path.join(REF_SRC, "UMPA")
    cy = [sys.executable, "-m", "cython", "--cplus", "-3", "-I", inc, pyx, "-o", cpp]
    cxx = ["g++", "-shared", "-fPIC", "-std=c++17", "-O3", "-ffast-math",
           "-march=x86-64-v3", "-fopenmp", "-w",
           "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include(),
           "-I" + inc, cpp, "-o", REF_SO, "-lm"]
    for cmd in (cy, cxx):
        if not quiet:
            print(" ".join(cmd))
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL if quiet else None,
                              stderr=subprocess.DEVNULL if quiet else None)
    # the generated C++ is derived from reference source: keep it out of the tree
    try:
        os.remove(cpp)
    except OSError:
        pass
    return REF_SO


if __name__ == "__main__":
    print("port:", build_port(force="--force" in sys.argv, quiet=False))
    print("ref :", build_ref(force="--force" in sys.argv, quiet=False))
