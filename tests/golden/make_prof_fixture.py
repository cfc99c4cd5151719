#!/usr/bin/env python3
"""Regenerates tests/golden/sample_5p.prof and sample_3p.prof: the `-profile` inputs of deamSim (deamSim.cpp:956-1094), written by the
reference's OWN converter src/mapDamage2prof.cpp (built by oracle/Makefile from its source + our libgab shim: oracle/_ref/src/mapDamage2prof)
from tests/golden/mapdamage_sample.txt.  Run in the build container (needs oracle/_ref); the GPU box only sees the two outputs."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
TOOL = os.path.join(HERE, "..", "..", "oracle", "_ref", "src", "mapDamage2prof")

if __name__ == "__main__":
    subprocess.check_call([TOOL, "-5p", os.path.join(HERE, "sample_5p.prof"), "-3p", os.path.join(HERE, "sample_3p.prof"), os.path.join(HERE, "mapdamage_sample.txt")])
    for f in ("sample_5p.prof", "sample_3p.prof"):
        print(f, sum(1 for _ in open(os.path.join(HERE, f))), "lines")
