#!/usr/bin/env python3
"""Stand-in for art_illumina (absent from this image; gargammel.pl:1856-1869 calls `art_illumina -ss HS25 -amp -na [-p] -i P_a.fa -l RL
-c 1 -qs Q -qs2 Q -o P_s`).  Test infrastructure: cuts each amplicon of the -artp/-arts FASTA into error-free reads with a flat
quality string, so that the driver's final gzip steps find P_s1.fq / P_s2.fq (or P_s.fq)."""
import sys


def main(argv):
    paired = "-p" in argv
    def val(flag):
        return argv[argv.index(flag) + 1]
    src, rl, out = val("-i"), int(val("-l")), val("-o")
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    outs = [open(out + ("1.fq" if paired else ".fq"), "wb")] + ([open(out + "2.fq", "wb")] if paired else [])
    name = None
    for line in open(src, "rb"):
        line = line.rstrip(b"\n")
        if line.startswith(b">"):
            name = line[1:]
            continue
        q = b"I" * rl
        outs[0].write(b"@" + name + b"-1/1\n" + line[:rl] + b"\n+\n" + q + b"\n")
        if paired:
            outs[1].write(b"@" + name + b"-1/2\n" + line[-rl:].translate(comp)[::-1] + b"\n+\n" + q + b"\n")
    for f in outs:
        f.close()


if __name__ == "__main__":
    main(sys.argv[1:])
