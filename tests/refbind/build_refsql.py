"""TEST TARGET.  Compile the reference binding (eventql_b200/host/refbind/gpu_binding.{h,cc}) against the reference's REAL
headers and link the test runner
eventql_b200/evqgpu_refsql = reference Runtime (parser, planner, ResultCursor: oracle/_ref/build/libevqlref.a) +
GpuScheduler + GpuCSTableScanProvider + libevqgpu.so.

Only possible where the reference tree and its compiled objects exist (this container: /root/reference and
oracle/_ref/build from oracle/build_ref.py).  The linked binary is git-ignored but travels to the GPU box with the
snapshot, like libevqgpu.so.  No reference source is copied; include paths point into /root/reference.

Usage: python -m tests.refbind.build_refsql [--ref /root/reference]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = os.path.join(ROOT, "eventql_b200")
HERE = os.path.join(PKG, "host", "refbind")   # the binding sources (product); this script and the runner's main are test code
REFBUILD = os.path.join(ROOT, "oracle", "_ref", "build")
OUT = os.path.join(PKG, "evqgpu_refsql")


def available(ref="/root/reference"):
    return os.path.isdir(os.path.join(ref, "src", "eventql")) and os.path.exists(os.path.join(REFBUILD, "libevqlref.a"))


def build(ref="/root/reference", verbose=False):
    if not available(ref):
        return None
    P = os.path.join(ref, "deps/3rdparty/protobuf")
    D = os.path.join(ref, "deps/3rdparty")
    gen = os.path.join(REFBUILD, "gen")
    # the flags of oracle/build_ref.py: the reference's headers need them
    cxx = ["g++", "-std=c++11", "-O2", "-DNDEBUG", "-w", "-fpermissive",
           "-include", "functional", "-include", "memory", "-include", "cstdint",
           "-include", "string", "-include", "limits", "-include", "cstring",
           "-I" + gen, "-I" + os.path.join(ref, "src"), "-I" + D, "-I" + P,
           "-I" + os.path.join(D, "zookeeper/source/include"), "-I" + os.path.join(D, "zookeeper/source/generated"),
           "-DHAVE_PTHREAD=1", "-DHAVE_ZLIB=1", "-DHAVE_SYSLOG_H=1", "-DHAVE_GETHOSTBYNAME_R=1",
           '-DEVQL_VERSION="v0.5.0"', '-DEVQL_BUILDID="evqgpu"']
    srcs = [os.path.join(HERE, "gpu_binding.cc"), os.path.join(ROOT, "tests", "refbind", "evqgpu_refsql_main.cc")]
    deps = srcs + [os.path.join(HERE, "gpu_binding.h"), os.path.join(ROOT, "include", "evqgpu.h"), os.path.join(PKG, "libevqgpu.so"),
                   os.path.join(REFBUILD, "libevqlref.a"), __file__]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps if os.path.exists(d)):
        return OUT
    objdir = os.path.join(PKG, "_build", "refbind")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        cmd = cxx + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout[-6000:])
            raise SystemExit("refbind: compile failed: " + s)
        objs.append(o)
    cobjs = [os.path.join(REFBUILD, "c", p) for p in ("libsimdcomp/simdbitpacking.o", "libsimdcomp/simdcomputil.o", "inih/ini.o",
                                                    "murmurhash/murmur3.o", "liblmdb/mdb.o", "liblmdb/midl.o")]
    tools = [os.path.join(REFBUILD, "tools", p) for p in ("chart_stub.o", "ext_aggregates.o")]
    cmd = (["g++", "-O2"] + objs + tools + ["-Wl,--start-group", os.path.join(REFBUILD, "libevqlref.a"), os.path.join(REFBUILD, "libpb.a"),
                                            "-Wl,--end-group"] + cobjs +
           ["-L" + PKG, "-levqgpu", "-Wl,-rpath,$ORIGIN", "-lpthread", "-lz", "-ldl", "-o", OUT])
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout[-6000:])
        raise SystemExit("refbind: link failed")
    return OUT


if __name__ == "__main__":
    ref = sys.argv[sys.argv.index("--ref") + 1] if "--ref" in sys.argv else "/root/reference"
    print(build(ref, verbose="--verbose" in sys.argv) or "reference tree / oracle/_ref/build not present: nothing built")
