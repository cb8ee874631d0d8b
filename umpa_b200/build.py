"""Build libumpa_b200.so in-tree with nvcc for sm_100a (no GPU needed to compile)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libumpa_b200.so")
SOURCES = ["capi.cu", "lazy_path.cu", "table_path.cu", "kernel_path.cu", "hoststage.cu", "post.cu"]
# shift_table_inst.cu is compiled once per window half-width (-1 = unfiltered table) so that the
# template instantiations build in parallel: (source, extra flags, object tag)
INSTANCES = [("shift_table_inst.cu", ["-DUMPA_INST_NW=%d" % nw], "nw%s" % ("m1" if nw < 0 else nw)) for nw in range(-1, 7)]
HEADERS = ["common.cuh", "walk.cuh", "shift_table.cuh", os.path.join("..", "..", "include", "umpa_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden", "-cudart", "static"]


def nvcc_path():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + [i[0] for i in INSTANCES] + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False, extra=(), dest=None):
    """dest: alternative output path (tuning builds with extra -D flags); default builds LIB in-tree."""
    if dest is None and not force and up_to_date():
        return LIB
    nvcc = nvcc_path()
    objs = []
    procs = []
    units = [(src, [], "") for src in SOURCES] + [(src, fl, "." + tag) for src, fl, tag in INSTANCES]
    for src, flags, tag in units:
        obj = os.path.join(CSRC, src[:-3] + tag + (".o" if dest is None else "." + os.path.basename(dest) + ".o"))
        cmd = [nvcc] + NVCC_FLAGS + list(extra) + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), out))
    link = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC", "-o", dest or LIB] + objs
    if verbose:
        print(" ".join(link))
    subprocess.check_call(link)
    return dest or LIB


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a not in ("--force",)]
    print(build(force=True, verbose=True, extra=extra))
