#!/usr/bin/env python
"""Wall clock of the whole driver on the C2 job: `perl gargammel.pl` (the unmodified script) over the reference filters (oracle/_ref) and
over bin/, against bin/gargammel-b200 --keep on the same directory, options and art stub (bench.py's cpu_baseline.via_gargammel_pl legs,
run alone).  usage: python scripts/driver_walltime.py [--big]   -> one JSON object on stdout"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true", help="also 10 M fragments through the fused driver")
    ap.add_argument("--fused-only", action="store_true", help="skip the gargammel.pl legs")
    a = ap.parse_args()
    args = argparse.Namespace(scale=1.0, frags=0, c5_sizes="l40", c5_emit="artp", gpus=1)
    J = bench.make_job("C2", args, 1)
    bins, ref = os.path.join(ROOT, "bin"), os.path.join(ROOT, "oracle", "_ref", "src")
    out = {"job": J["workload"]}
    out["fused_driver"] = bench.via_gargammel_pl(J, 200000, bins, "bin/gargammel-b200 --keep", fused=True)
    out["fused_driver_10x"] = bench.via_gargammel_pl(J, 2000000, bins, "bin/gargammel-b200 --keep", fused=True)
    if a.big:
        out["fused_driver_50x"] = bench.via_gargammel_pl(J, 10000000, bins, "bin/gargammel-b200 --keep", fused=True)
    if a.fused_only:
        print(json.dumps(out), flush=True)
        return
    out["reference_binaries"] = bench.via_gargammel_pl(J, 200000, ref, "gargammel.pl over oracle/_ref/src")
    out["dropin_binaries"] = bench.via_gargammel_pl(J, 200000, bins, "gargammel.pl over bin/")
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
