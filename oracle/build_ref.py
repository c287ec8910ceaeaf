#!/usr/bin/env python3
"""Build the UNMODIFIED reference (EventQL v0.5.0 csql + cstable) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path; the
product (eventql_b200/, include/) never imports, links or executes it.

What this does (recipe: SURVEY.md Appendix E; the reference's autotools build
cannot run here - autoconf/automake/libtool/protoc are absent):

  1. compiles the vendored protoc 2.5.0   (deps/3rdparty/protobuf/Makefile.am 'protoc_srcs')
  2. generates the 9 *.pb.{h,cc}          (src/Makefile.am:72-81) into oracle/_ref/build/gen
  3. compiles every .cc named in EVQL_CORE_SOURCES_ (src/Makefile.am:111-960) from where it
     lies under /root/reference, except the files that do not build in this snapshot
     (chartsql, db/database.cc, mapreduce prelude) - SURVEY.md H14
  4. links  oracle/_ref/evqlref  = reference lib + oracle/ref_tools/*.cc (our own runner)

No reference source is copied: the compiler reads the files in place and writes
objects to oracle/_ref/build only.  oracle/_ref/ is git-ignored but NOT
gpurun-ignored, so the binary travels to the GPU box.

Usage: python oracle/build_ref.py [-j N] [--ref /root/reference]
"""
import argparse
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
BUILD = os.path.join(OUT, "build")

EXCLUDE_SUBSTR = (
    ".pb.cc",
    "sql/extensions/chartsql/",
    "eventql/db/database.cc",
    "mapreduce_preludejs.cc",
)
EXCLUDE_MAINS = re.compile(r"eventql/(evql[a-z]*|evqld)\.cc$|cstable_tool\.cc$")


def sh(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    return r.returncode, r.stdout


def newer(src, obj):
    return (not os.path.exists(obj)) or os.path.getmtime(src) > os.path.getmtime(obj)


def compile_many(jobs, nproc):
    """jobs: list of (cmd, src, obj). Returns list of failed sources."""
    failed = []

    def run(job):
        cmd, src, obj = job
        if not newer(src, obj):
            return None
        os.makedirs(os.path.dirname(obj), exist_ok=True)
        rc, out = sh(cmd)
        if rc != 0:
            return (src, out[-2000:])
        return None

    with ThreadPoolExecutor(nproc) as ex:
        for res in ex.map(run, jobs):
            if res:
                failed.append(res)
    return failed


def makefile_list(path, var):
    txt = open(path).read().replace("\\\n", " ")
    m = re.search(r"^%s\s*=\s*(.*)$" % re.escape(var), txt, re.M)
    if not m:
        raise SystemExit("cannot find %s in %s" % (var, path))
    return m.group(1).split()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-j", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    R = args.ref
    if not os.path.isdir(R):
        print("reference tree %s not present - keeping prebuilt oracle/_ref as is" % R)
        return 0
    P = os.path.join(R, "deps/3rdparty/protobuf")
    os.makedirs(BUILD, exist_ok=True)

    # ---- 1. protoc + protobuf runtime -------------------------------------------------
    protoc_srcs = [s.replace("$(abs_srcdir)/", "") for s in makefile_list(os.path.join(P, "Makefile.am"), "protoc_srcs")]
    protoc_srcs = [s for s in protoc_srcs if "msvc" not in s]
    pb_flags = ["-std=c++11", "-O1", "-w", "-fpermissive", "-DHAVE_PTHREAD=1", "-DHAVE_PTHREAD_H=1", "-I" + P]
    jobs = []
    pb_objs, pb_rt_objs = [], []
    for s in protoc_srcs:
        obj = os.path.join(BUILD, "pb", s.replace("/", "_") + ".o")
        jobs.append((["g++"] + pb_flags + ["-c", os.path.join(P, s), "-o", obj], os.path.join(P, s), obj))
        pb_objs.append(obj)
        if "/compiler/" not in s:
            pb_rt_objs.append(obj)
    failed = compile_many(jobs, args.j)
    if failed:
        print(failed[0][1])
        raise SystemExit("protobuf compile failed: %s" % failed[0][0])
    protoc = os.path.join(BUILD, "protoc")
    if not os.path.exists(protoc):
        rc, out = sh(["g++"] + pb_objs + ["-lpthread", "-o", protoc])
        if rc:
            raise SystemExit(out)
    print("[1/4] protoc ok")

    # ---- 2. generated protobuf sources -------------------------------------------------
    gen = os.path.join(BUILD, "gen")
    os.makedirs(gen, exist_ok=True)
    protos = makefile_list(os.path.join(R, "src/Makefile.am"), "evql_protos")
    for p in protos:
        target = os.path.join(gen, p.replace(".proto", ".pb.cc"))
        if not os.path.exists(target):
            rc, out = sh([protoc, "--proto_path=" + os.path.join(R, "src"), "--cpp_out=" + gen, os.path.join(R, "src", p)])
            if rc:
                raise SystemExit(out)
    print("[2/4] %d protos generated" % len(protos))

    # ---- 3. reference core -------------------------------------------------------------
    D = os.path.join(R, "deps/3rdparty")
    cxx = [
        "g++", "-std=c++11", "-O2", "-DNDEBUG", "-w", "-fpermissive",
        "-include", "functional", "-include", "memory", "-include", "cstdint",
        "-include", "string", "-include", "limits", "-include", "cstring",
        "-I" + gen, "-I" + os.path.join(R, "src"), "-I" + D, "-I" + P,
        "-I" + os.path.join(D, "zookeeper/source/include"), "-I" + os.path.join(D, "zookeeper/source/generated"),
        "-DHAVE_PTHREAD=1", "-DHAVE_ZLIB=1", "-DHAVE_SYSLOG_H=1", "-DHAVE_GETHOSTBYNAME_R=1",
        '-DEVQL_VERSION="v0.5.0"', '-DEVQL_BUILDID="oracle"',
    ]
    core = [s for s in makefile_list(os.path.join(R, "src/Makefile.am"), "EVQL_CORE_SOURCES_") if s.endswith(".cc")]
    core = [s for s in core if not any(x in s for x in EXCLUDE_SUBSTR) and not EXCLUDE_MAINS.search(s)]
    jobs, objs = [], []
    for s in core:
        src = os.path.join(R, "src", s)
        obj = os.path.join(BUILD, "core", s[:-3] + ".o")
        jobs.append((cxx + ["-c", src, "-o", obj], src, obj))
        objs.append(obj)
    for p in protos:
        src = os.path.join(gen, p.replace(".proto", ".pb.cc"))
        obj = os.path.join(BUILD, "core", p.replace(".proto", ".pb.o"))
        jobs.append((cxx + ["-c", src, "-o", obj], src, obj))
        objs.append(obj)
    cobjs = []
    for s, extra in (("libsimdcomp/simdbitpacking.c", ["-msse2"]), ("libsimdcomp/simdcomputil.c", ["-msse2"]),
                     ("inih/ini.c", []), ("murmurhash/murmur3.c", []), ("liblmdb/mdb.c", []), ("liblmdb/midl.c", [])):
        src = os.path.join(D, s)
        obj = os.path.join(BUILD, "c", s[:-2] + ".o")
        jobs.append((["gcc", "-O2", "-w"] + extra + ["-I" + D, "-c", src, "-o", obj], src, obj))
        cobjs.append(obj)
    failed = compile_many(jobs, args.j)
    bad = set(f[0] for f in failed)
    for f in failed:
        print("FAILED:", f[0])
        print(f[1][-600:])
    if failed:
        raise SystemExit("%d reference files failed to compile" % len(failed))
    print("[3/4] %d reference sources compiled" % len(jobs))

    lib = os.path.join(BUILD, "libevqlref.a")
    if os.path.exists(lib):
        os.unlink(lib)
    rc, out = sh(["ar", "rcs", lib] + objs)
    if rc:
        raise SystemExit(out)
    libpb = os.path.join(BUILD, "libpb.a")
    if os.path.exists(libpb):
        os.unlink(libpb)
    rc, out = sh(["ar", "rcs", libpb] + pb_rt_objs)
    if rc:
        raise SystemExit(out)

    # ---- 4. our own runner against the reference lib ------------------------------------
    tool_srcs = sorted(f for f in os.listdir(os.path.join(HERE, "ref_tools")) if f.endswith(".cc"))
    tobjs = []
    jobs = []
    for s in tool_srcs:
        src = os.path.join(HERE, "ref_tools", s)
        obj = os.path.join(BUILD, "tools", s[:-3] + ".o")
        jobs.append((cxx + ["-c", src, "-o", obj], src, obj))
        tobjs.append(obj)
    failed = compile_many(jobs, args.j)
    if failed:
        print(failed[0][1])
        raise SystemExit("ref_tools compile failed: %s" % failed[0][0])
    exe = os.path.join(OUT, "evqlref")
    rc, out = sh(["g++", "-O2"] + tobjs + ["-Wl,--start-group", lib, libpb, "-Wl,--end-group"] + cobjs +
                 ["-lpthread", "-lz", "-ldl", "-o", exe])
    if rc:
        print(out[-4000:])
        raise SystemExit("link failed")
    print("[4/4] linked", exe)
    return 0


if __name__ == "__main__":
    sys.exit(main())
