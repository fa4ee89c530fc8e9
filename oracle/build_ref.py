#!/usr/bin/env python3
"""Build the CPU checkers.  TEST INFRASTRUCTURE ONLY.

  oracle/_build/liboracle.so          plain-C restatement (ldpc_oracle.c)      -- always
  oracle/_ref/libref_<variant>.so     the UNMODIFIED reference objects + probe -- only when
  oracle/_ref/wrapper_<variant>       the reference's own console program         /root/reference
                                                                                   is present

The reference selects its code at compile time (ArrayLDPCMacro.h:18-24), so each variant is
compiled against a header whose enum (and the G_mlist bounds, :192-196) was switched by the
regexes below.  That header is generated into a temporary directory and deleted after the
compile: no reference source is ever stored in the repo, and only binaries land in
oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LDPC_REFERENCE_DIR", "/root/reference")
OUT_REF = os.path.join(HERE, "_ref")
OUT_ORACLE = os.path.join(HERE, "_build")

# enum bodies per variant; "wifi" is the header as shipped
VARIANTS = {
    "wifi": None,
    "a5": dict(enum="NUM_VAR = 2209, NUM_CHK = 235, NUM_CGRP = 5, VAR_DEG = 5, NUM_VGRP = 47, "
                    "CHK_DEG = 47, P = 47, CIR_SIZE = 47, INFO_LENGTH = 1978, CWD_LENGTH = 2209",
               g_rows=231, g_cols=1078),
    # cdeg/clist/addr_count are dimensioned by INFO_LENGTH (ArrayLDPCMacro.h:172) -> keep it >= NUM_CHK
    "a24": dict(enum="NUM_VAR = 2209, NUM_CHK = 1128, NUM_CGRP = 24, VAR_DEG = 24, NUM_VGRP = 47, "
                     "CHK_DEG = 47, P = 47, CIR_SIZE = 47, INFO_LENGTH = 1128, CWD_LENGTH = 2209",
                g_rows=8, g_cols=8),
    "c79": dict(enum="NUM_VAR = 2212, NUM_CHK = 316, NUM_CGRP = 4, VAR_DEG = 4, NUM_VGRP = 28, "
                     "CHK_DEG = 28, P = 79, CIR_SIZE = 79, INFO_LENGTH = 1899, CWD_LENGTH = 2212",
                g_rows=8, g_cols=8),
}
DRIVER_VARIANTS = ("wifi", "a5")


def run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        raise RuntimeError("build failed: " + cmd[0])


def build_oracle():
    os.makedirs(OUT_ORACLE, exist_ok=True)
    out = os.path.join(OUT_ORACLE, "liboracle.so")
    src = os.path.join(HERE, "ldpc_oracle.c")
    if os.path.exists(out) and os.path.getmtime(out) >= max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(HERE, "ldpc_oracle.h"))):
        return out
    # no -march=native / -ffast-math: the channel's double arithmetic must not be contracted
    run(["gcc", "-O2", "-std=c99", "-shared", "-fPIC", "-o", out, src, "-lm"])
    return out


def variant_header(name):
    with open(os.path.join(REF, "ArrayLDPCMacro.h")) as fh:
        text = fh.read()
    spec = VARIANTS[name]
    if spec is None:
        return text
    text, n = re.subn(r"enum CodeWifi \{[^}]*\};", "enum CodeVariant {\n\t\t" + spec["enum"] + "};",
                      text, count=1)
    assert n == 1, "enum block not found"
    text, n1 = re.subn(r"int ChkDeg\[972\];", "int ChkDeg[%d];" % max(spec["g_rows"], 8), text)
    text, n2 = re.subn(r"int G_mlist\[972\]\[540\];",
                       "int G_mlist[%d][%d];" % (spec["g_rows"], spec["g_cols"]), text)
    assert n1 == 1 and n2 == 1, "G_mlist bounds not found"
    return text


def build_ref(force=False):
    if not os.path.isdir(REF):
        return False
    os.makedirs(OUT_REF, exist_ok=True)
    srcs = [os.path.join(HERE, f) for f in ("ref_harness.cpp", "ref_driver_main.cpp", "build_ref.py")]
    newest = max(os.path.getmtime(s) for s in srcs)
    for name in VARIANTS:
        lib = os.path.join(OUT_REF, "libref_%s.so" % name)
        exe = os.path.join(OUT_REF, "wrapper_%s" % name)
        need_lib = force or not os.path.exists(lib) or os.path.getmtime(lib) < newest
        need_exe = name in DRIVER_VARIANTS and (force or not os.path.exists(exe)
                                                or os.path.getmtime(exe) < newest)
        if not (need_lib or need_exe):
            continue
        tmp = tempfile.mkdtemp(prefix="ldpc_ref_")
        try:
            with open(os.path.join(tmp, "ArrayLDPCMacro.h"), "w") as fh:
                fh.write(variant_header(name))
            common = ["g++", "-O2", "-w", "-fpermissive", "-I", tmp, "-I", os.path.join(HERE, "shim"),
                      "-DREF_DIR=" + REF]
            if need_lib:
                run(common + ["-shared", "-fPIC", "-o", lib, os.path.join(HERE, "ref_harness.cpp")])
            if need_exe:
                run(common + ["-o", exe, os.path.join(HERE, "ref_driver_main.cpp")])
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    return True


def main():
    build_oracle()
    have_ref = build_ref(force="--force" in sys.argv)
    print("liboracle.so built; reference builds:", "yes" if have_ref else "skipped (no %s)" % REF)


if __name__ == "__main__":
    main()
