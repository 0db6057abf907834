"""TEST INFRASTRUCTURE ONLY -- compiles oracle/ciclope_oracle.c into oracle/liboracle.so (gcc)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False):
    src = os.path.join(_HERE, "ciclope_oracle.c")
    out = os.path.join(_HERE, "liboracle.so")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-std=c99", "-Wall", "-o", out, src])
    return out


if __name__ == "__main__":
    print(build(force=True))
