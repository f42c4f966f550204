"""Builds libwfsp.so in-tree with nvcc for sm_100a (the only target).  `python -m waveformml_b200.build`.

The library carries a digest of the sources it was compiled from (`wfsp_source_hash`); `_lib.load()` compares it
with `source_hash()` of the tree it runs in, so a binary left over from before a csrc/ or wfsp.h change is rebuilt
(or refused) instead of being called through new ctypes signatures."""
import fcntl
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "rulebook.cu", "conv_simt.cu", "conv_umma.cu", "dense_pack.cu", "bn.cu", "bn_stream.cu", "p2p_sgd.cu", "head.cu", "edges.cu"]
LIB = os.path.join(HERE, "libwfsp.so")
HEADER = os.path.join(ROOT, "include", "wfsp.h")


def source_hash():
    """sha256 over every file of csrc/ (name + bytes, sorted) and include/wfsp.h; 16 hex digits."""
    h = hashlib.sha256()
    files = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) 
    for path in [os.path.join(CSRC, f) for f in files] + [HEADER]:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def built_hash():
    """Digest embedded in the existing libwfsp.so, or None."""
    if not os.path.exists(LIB):
        return None
    import ctypes
    try:
        lib = ctypes.CDLL(LIB)
        fn = lib.wfsp_source_hash
        fn.restype = ctypes.c_char_p
        return fn().decode()
    except (OSError, AttributeError):
        return None


def _stale():
    return built_hash() != source_hash()


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # one builder at a time (several ranks / xdist workers may find the library stale together); the result is
    # moved into place atomically
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not _stale():
            return LIB
        tmp = LIB + ".tmp.%d" % os.getpid()
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
               '-DWFSP_SOURCE_HASH="%s"' % source_hash(), "-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        subprocess.check_call(cmd)
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
