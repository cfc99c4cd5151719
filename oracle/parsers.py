"""Pure-Python restatement of the reference's table loaders.  TEST INFRASTRUCTURE ONLY (oracle side of the
parser parity tests; the product's loaders are C++ in gargammel_b200/csrc/gg_host.cpp)."""
import gzip


def _open(path):
    with open(path, "rb") as f:
        magic = f.read(2)
    return gzip.open(path, "rt") if magic == b"\x1f\x8b" else open(path, "rt")


def _lines(path):
    with _open(path) as f:
        data = f.read()
    lines = data.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    return lines


def load_size_freq(path):
    """fragSim.cpp:751-784: 'length<TAB>freq', running sum, 0.99 <= sum <= 1.01."""
    lens, cum, c = [], [], 0.0
    for line in _lines(path):
        f = line.split("\t")
        if len(f) != 2:
            raise ValueError("does not have 2 fields: " + line)
        c += float(f[1])
        lens.append(int(f[0]))
        cum.append(c)
    if c < 0.99 or c > 1.01:
        raise ValueError("the sum of frequencies does not sum to 1, sum=%g" % c)
    return lens, cum


def load_size_list(path):
    """fragSim.cpp:720-745: one length per line, keep t >= 1."""
    out = []
    for line in _lines(path):
        if len(line.split("\t")) != 1:
            raise ValueError("does not have 1 field: " + line)
        t = int(line)
        if t >= 1:
            out.append(t)
    return out


def _rate_file(path, skip):
    rows = []
    for line in _lines(path)[1:]:                                   # header skipped, deamSim.cpp:1121
        first = [fld.split(" ")[0] for fld in line.split("\t")[skip:]]   # deamSim.cpp:1127-1132
        if len(first) != 12:
            raise ValueError("line from deamination does not have 12 fields " + line)
        r = [float(x) for x in first]
        for b in range(4):                                          # deamSim.cpp:1148-1158
            if r[3 * b] + r[3 * b + 1] + r[3 * b + 2] > 1.0:
                raise ValueError("the sum of the substitution probabilities exceeds 1")
        rows.append(r)
    return rows


def load_matrix_dat(prefix):
    """deamSim.cpp:1098-1241: <prefix>5.dat, <prefix>3.dat; first column dropped; 3' reversed (:1236)."""
    s5 = _rate_file(prefix + "5.dat", 1)
    s3 = _rate_file(prefix + "3.dat", 1)
    s3.reverse()
    return s5, s3


def load_profile(prefix):
    """deamSim.cpp:956-1094: <prefix>5p.prof, <prefix>3p.prof; all 12 columns; 3' NOT reversed."""
    return _rate_file(prefix + "5p.prof", 0), _rate_file(prefix + "3p.prof", 0)


def load_mapdamage(path):
    """deamSim.cpp:453-668."""
    lines = _lines(path)
    hdr = lines[3].split("\t")
    idx = []
    for i in range(12):
        h = hdr[9 + i]
        if len(h) != 3 or h[1] != ">":
            raise ValueError("wrong header field " + h)
        fr, to = "ACGT".index(h[0]), "ACGT".index(h[2])
        idx.append(fr * 3 + (to if to < fr else to - 1))
    cnt = {"5p": [], "3p": []}
    first = {"5p": True, "3p": True}
    for line in lines[4:]:
        f = line.split("\t")
        end = f[1]
        if end not in cnt:
            raise ValueError("does not have a 5p or 3p in the 2nd column: " + line)
        if first[end] and f[2] == "-":
            first[end] = False
        if first[end]:
            sub = [0] * 12
            for i in range(12):
                sub[idx[i]] = int(f[9 + i])
            cnt[end].append(([int(x) for x in f[4:8]], sub))
        else:
            base, sub = cnt[end][int(f[3]) - 1]
            for i in range(4):
                base[i] += int(f[4 + i])
            for i in range(12):
                sub[idx[i]] += int(f[9 + i])
    if len(cnt["5p"]) != len(cnt["3p"]):
        raise ValueError("the number of defined bases for the 5' and 3' ends differ")

    def rates(c):
        return [[(sub[b * 3 + a] / base[b]) if base[b] else float("nan") for b in range(4) for a in range(3)] for base, sub in c]
    return rates(cnt["5p"]), rates(cnt["3p"])


def load_dnacomp(path, dist):
    """fragSim.cpp:798-1012: mapDamage dnacomp.txt -> four [2*dist][4] tables of frequency - background + 0.25
    (5p+, 5p-, 3p+, 3p-), positions within `dist` of the fragment end in file order; everything else is background."""
    tabs = {("5p", "+"): [], ("5p", "-"): [], ("3p", "+"): [], ("3p", "-"): []}
    tot = [0, 0, 0, 0, 0]
    for line in _lines(path):
        if line.startswith("#") or line.startswith("Chr"):
            continue
        f = line.split("\t")
        if len(f) <= 7:
            continue
        d = int(f[3]); c = [int(x) for x in f[4:9]]
        if abs(d) <= dist:
            if (f[1], f[2]) in tabs:
                tabs[(f[1], f[2])].append([c[k] / c[4] for k in range(4)])
        else:
            for k in range(5):
                tot[k] += c[k]
    bg = [tot[k] / tot[4] for k in range(4)]
    out = []
    for key in (("5p", "+"), ("5p", "-"), ("3p", "+"), ("3p", "-")):
        rows = tabs[key]
        if len(rows) != 2 * dist:
            raise ValueError("composition file does not hold 2*dist positions for %s %s" % key)
        out.append([[0.25 + (r[k] - bg[k]) for k in range(4)] for r in rows])
    return out
