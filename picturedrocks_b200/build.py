"""Build libprb200.so (hand-written sm_100a CUDA + C-ABI) in-tree with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["stream.cu", "vstream.cu", "vell.cu", "vell2.cu", "state.cu", "finalize.cu", "finalize_wide.cu", "selector.cu", "discretize.cu", "synth.cu",
           "step.cu", "peer.cu", "ingest.cu", "host.cu"]
LIB = os.path.join(HERE, "libprb200.so")
OBJ = os.path.join(HERE, "_obj")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [
        os.path.join(CSRC, "common.cuh"),
        os.path.join(CSRC, "select.cuh"),
        os.path.join(CSRC, "selfuse.cuh"),
        os.path.join(CSRC, "vcommon.cuh"),
        os.path.join(HERE, "..", "include", "prb200.h"),
    ]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """nvcc -> libprb200.so.  Safe with one process per GPU: the build runs under an exclusive
    file lock into a temporary file that is renamed into place, so no rank can ever load a
    half-written library and only the first rank to arrive compiles."""
    if not force and not needs_build():
        return LIB
    import fcntl

    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():  # another rank built it while we waited
                return LIB
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            tmp = "%s.tmp.%d" % (LIB, os.getpid())
            flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
                     "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off", "--fmad=true",
                     "-ccbin", ccbin]
            if verbose:
                flags += ["-Xptxas", "-v"]
            os.makedirs(OBJ, exist_ok=True)
            headers = [os.path.join(CSRC, h) for h in ("common.cuh", "select.cuh", "selfuse.cuh", "vcommon.cuh")] + [
                os.path.join(HERE, "..", "include", "prb200.h")]
            hdr_t = max(os.path.getmtime(h) for h in headers)

            def compile_one(src):
                """one translation unit -> build/obj/<name>.o (skipped when up to date)"""
                obj = os.path.join(OBJ, src[:-3] + ".o")
                path = os.path.join(CSRC, src)
                if (not force and os.path.exists(obj)
                        and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_t)):
                    return obj, None
                r = subprocess.run([nvcc] + flags + ["-c", "-o", obj, path], capture_output=True,
                                   text=True)
                return obj, r

            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
                results = list(ex.map(compile_one, SOURCES))
            for obj, r in results:
                if r is not None and r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                    raise RuntimeError("nvcc failed building libprb200.so")
                if r is not None and verbose:
                    sys.stderr.write(r.stderr)
            r = subprocess.run([nvcc, "-shared", "-ccbin", ccbin, "-o", tmp]
                               + [obj for obj, _ in results], capture_output=True, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed linking libprb200.so")
            os.replace(tmp, LIB)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
