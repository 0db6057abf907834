"""TEST INFRASTRUCTURE ONLY -- compiles oracle/ciclope_oracle.c into oracle/liboracle.so (gcc)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("ciclope_oracle.c", "stream_oracle.c")]
    out = os.path.join(_HERE, "liboracle.so")
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-std=c99", "-Wall", "-o", out] + srcs)
    return out


if __name__ == "__main__":
    print(build(force=True))
