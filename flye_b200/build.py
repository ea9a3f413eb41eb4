"""Build everything native, in-tree.

  flye_b200/libflye_b200.so   the product: CUDA kernels + C ABI, sm_100a only
  tools/_bin/simreads         synthetic read generator
  oracle/_ref/*               the checkers (reference harness when /root/reference exists, CPU restatement)
  tests/cpu_models/_bin/*     host builds of device headers (introsort emulation check)
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "flye_b200", "csrc")
BUILD = os.path.join(ROOT, "build")
LIB = os.path.join(ROOT, "flye_b200", "libflye_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Wno-deprecated-declarations"]
SOURCES = ["api.cu", "count_index.cu", "overlap.cu", "editdist.cu", "comm.cu", "intpeak.cu", "closure.cu", "ksw.cu"]


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("command failed: " + " ".join(cmd))
    return r.stdout


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "flye_b200.h"))
    objs, jobs = [], []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(BUILD, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj])
    with ThreadPoolExecutor(max_workers=int(os.environ.get("FLYE_B200_BUILD_JOBS", "4"))) as ex:
        list(ex.map(_run, jobs))
    if jobs or _newer(LIB, objs):   # (an object compiled by hand, e.g. with -Xptxas -v, is newer than the library too)
        _run([nvcc, "-shared", "-o", LIB] + objs + ["-ldl"])
    return LIB


def build_tools():
    out = os.path.join(ROOT, "tools", "_bin")
    os.makedirs(out, exist_ok=True)
    src, exe = os.path.join(ROOT, "tools", "simreads.cpp"), os.path.join(out, "simreads")
    if _newer(exe, [src]):
        _run(["g++", "-O2", "-std=c++17", src, "-o", exe])
    # digest of ordered overlap dumps (full-scale parity against committed reference digests; bench.py's digest check)
    src = os.path.join(ROOT, "tools", "ovlpdigest.cpp")
    exe, so = os.path.join(out, "ovlpdigest"), os.path.join(out, "libovlpdigest.so")
    if _newer(exe, [src]):
        _run(["g++", "-O2", "-std=c++17", src, "-o", exe])
    if _newer(so, [src]):
        _run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DOVLPDIGEST_LIB", src, "-o", so])
    out = os.path.join(ROOT, "tests", "cpu_models", "_bin")
    os.makedirs(out, exist_ok=True)
    src, exe = os.path.join(ROOT, "tests", "cpu_models", "introsort_check.cpp"), os.path.join(out, "introsort_check")
    if _newer(exe, [src, os.path.join(CSRC, "introsort_warp.cuh")]):
        _run(["g++", "-O2", "-std=c++17", src, "-o", exe])


    # CPU models: run-compressed DP / chain walk, bounded wavefronts, segmented radix sort, dense k-mer class index, glibc's logf
    for name in ("rundp_check", "wfa_check", "segsort_check", "dense_check", "logf_check"):
        src, exe = os.path.join(ROOT, "tests", "cpu_models", name + ".cpp"), os.path.join(out, name)
        if _newer(exe, [src, os.path.join(CSRC, "kmer_math.cuh"), os.path.join(CSRC, "glibc_logf.cuh")]):
            _run(["g++", "-O2", "-std=c++17", src, "-o", exe])
    src, so = os.path.join(ROOT, "tests", "cpu_models", "stdsort_lib.cpp"), os.path.join(out, "libstdsort.so")
    if _newer(so, [src]):
        _run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", so])


def build_oracle():
    odir = os.path.join(ROOT, "oracle")
    _run(["make", "-C", odir, "restate"])
    if os.path.isdir("/root/reference/src/sequence"):
        _run(["make", "-C", odir, "ref", "-j4"])


def build_host_harness():
    """oracle/harness.cpp (the driver of the reference's public API) compiled UNCHANGED against the C++ host mirror in
    flye_b200/host and linked to libflye_b200.so: the drop-in demonstration (tests/test_gpu_host_mirror.py)."""
    src = os.path.join(ROOT, "oracle", "harness.cpp")
    exe = os.path.join(BUILD, "flye_b200_harness")
    host = os.path.join(ROOT, "flye_b200", "host")
    deps = [src, LIB] + [os.path.join(dp, f) for dp, _, fs in os.walk(host) for f in fs]
    if _newer(exe, deps):
        _run(["g++", "-std=c++17", "-O2", "-DFLYE_B200", "-I" + host, "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
              "-L" + os.path.join(ROOT, "flye_b200"), "-lflye_b200", "-lz", "-pthread", "-Wl,-rpath,$ORIGIN/../flye_b200"])
    # oracle/ingest_check.cpp and oracle/host_api_check.cpp on the mirror (the reference builds of the same sources are
    # oracle/_ref/ingest_ref and oracle/_ref/host_api_ref)
    src2, exe2 = os.path.join(ROOT, "oracle", "ingest_check.cpp"), os.path.join(BUILD, "flye_b200_ingest")
    if _newer(exe2, [src2] + deps[1:]):
        _run(["g++", "-std=c++17", "-O2", "-I" + host, "-I" + os.path.join(ROOT, "include"), src2, "-o", exe2,
              "-L" + os.path.join(ROOT, "flye_b200"), "-lflye_b200", "-lz", "-pthread", "-Wl,-rpath,$ORIGIN/../flye_b200"])
    src3, exe3 = os.path.join(ROOT, "oracle", "host_api_check.cpp"), os.path.join(BUILD, "flye_b200_host_api")
    if _newer(exe3, [src3] + deps[1:]):
        _run(["g++", "-std=c++17", "-O2", "-I" + host, "-I" + os.path.join(ROOT, "include"), src3, "-o", exe3,
              "-L" + os.path.join(ROOT, "flye_b200"), "-lflye_b200", "-lz", "-pthread", "-Wl,-rpath,$ORIGIN/../flye_b200"])
    # oracle/trim_check.cpp on the device build of the KSW2 routine (fg_align_cigar_batch); the reference / restatement / host builds of
    # the same driver are oracle/_ref/trim_{ref,restate,core}
    src4, exe4 = os.path.join(ROOT, "oracle", "trim_check.cpp"), os.path.join(BUILD, "flye_b200_trim_device")
    if _newer(exe4, [src4, LIB, os.path.join(ROOT, "include", "flye_b200.h")]):
        _run(["g++", "-std=c++17", "-O2", "-DTRIM_DEVICE", "-I" + os.path.join(ROOT, "include"), src4, "-o", exe4,
              "-L" + os.path.join(ROOT, "flye_b200"), "-lflye_b200", "-pthread", "-Wl,-rpath,$ORIGIN/../flye_b200"])
    # ... and on the mirror's own trimming tail (hpcRange + trimByCigar) around the routine's host build
    exe5 = os.path.join(BUILD, "flye_b200_trim_mirror")
    if _newer(exe5, [src4, os.path.join(CSRC, "ksw_core.cuh")] + deps[1:]):
        _run(["g++", "-std=c++17", "-O2", "-DTRIM_MIRROR", "-I" + host, "-I" + os.path.join(ROOT, "include"), src4, "-o", exe5,
              "-L" + os.path.join(ROOT, "flye_b200"), "-lflye_b200", "-lz", "-pthread", "-Wl,-rpath,$ORIGIN/../flye_b200"])
    return exe


def build_flye_modules():
    """the reference's whole `flye-modules` twice — unmodified, and its callers on the host mirror (oracle/build_flye_modules.sh);
    build container only (needs the reference sources); tests/test_gpu_flye_modules.py runs both on the GPU box"""
    if not os.path.isdir("/root/reference/src/sequence"):
        return
    host = os.path.join(ROOT, "flye_b200", "host")
    deps = [LIB, os.path.join(ROOT, "include", "flye_b200.h")] + [os.path.join(dp, f) for dp, _, fs in os.walk(host) for f in fs]
    if _newer(os.path.join(ROOT, "oracle", "_ref", "flye-modules-b200"), deps) or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "flye-modules-ref")):
        _run([os.path.join(ROOT, "oracle", "build_flye_modules.sh")])


def build_all(force=False):
    build_lib(force)
    build_tools()
    build_oracle()
    build_host_harness()
    build_flye_modules()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built", LIB)
