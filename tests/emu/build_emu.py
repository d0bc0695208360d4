"""Build tests/emu/_build/libiaf_emu.so: the SIMT sources of iaf_b200/csrc compiled with g++ against the host
emulation header (tests/emu/cuda_emu.h).  TEST INFRASTRUCTURE ONLY -- see cuda_emu.h for what this is and is not."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "iaf_b200", "csrc")
OUT = os.path.join(HERE, "_build")
# IAF_EMU_FLAGS: extra -D switches, e.g. "-DBW_FASTDIV", to run the same tests on a development variant of the kernels
EXTRA = os.environ.get("IAF_EMU_FLAGS", "").split()
LIB = os.path.join(OUT, "libiaf_emu%s.so" % "".join(f.replace("-D", "_") for f in EXTRA))
SOURCES = [os.path.join(CSRC, f) for f in ("iaf_capi.cu", "iaf_pack.cu", "iaf_simt.cu", "iaf_bwd.cu")] + \
          [os.path.join(HERE, "tc_stub.cc")]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in os.listdir(HERE) if
                                                               f.endswith((".h", ".cc"))]
    deps.append(os.path.join(ROOT, "include", "iaf_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False):
    if not force and not _stale():
        return LIB
    os.makedirs(OUT, exist_ok=True)
    cmd = ["g++", "-std=c++20", "-O2", "-g", "-fPIC", "-shared", "-pthread", "-DIAF_EMU", "-Wno-unknown-pragmas",
           "-I", HERE, "-I", CSRC, "-o", LIB] + EXTRA
    for s in SOURCES:
        cmd += ["-x", "c++", s]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
