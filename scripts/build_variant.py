"""Build an experiment variant of the library next to the shipped one: libldpc_b200_<name>.so, compiled with the
given -D switches; select it at run time with LDPC_B200_LIB=<path> (same-box A/B runs).

usage: python scripts/build_variant.py NAME [-DMACRO ...]"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fixedpointldpc_b200 import build as B


def main():
    name, defs = sys.argv[1], sys.argv[2:]
    out = os.path.join(B.HERE, "libldpc_b200_%s.so" % name)
    cmd = [B._nvcc()] + B.NVCC_FLAGS + defs + ["-o", out] + [os.path.join(B.CSRC, f) for f in B.SOURCES] + ["-ldl", "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    sys.stderr.write(res.stderr[-3000:])
    if res.returncode != 0:
        raise SystemExit("nvcc failed")
    print(out)


if __name__ == "__main__":
    main()
