"""Compile the oracle's C restatement (oracle/c/chain_fb.c) -> oracle/c/libpk2_oracle.so.
TEST / BASELINE INFRASTRUCTURE ONLY."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "chain_fb.c")
LIB = os.path.join(HERE, "c", "libpk2_oracle.so")


def build(verbose=True):
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        if verbose:
            print("up to date:", LIB)
        return LIB
    cmd = ["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed:\n" + r.stderr)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build()
