"""Build oracle/_build/liboracle.so from oracle/b2d_oracle.c with gcc. TEST INFRASTRUCTURE ONLY."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")
SRC = os.path.join(HERE, "b2d_oracle.c")


def build_oracle(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = LIB + ".tmp.%d" % os.getpid()
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", tmp, SRC, "-lm"], check=True)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build_oracle(force=True))
