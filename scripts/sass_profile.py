"""Joins an `ncu --page source --csv` SASS dump with `nvdisasm -gi` line info of the same kernel and prints
instructions / stall samples per source region and per source line.

  ncu -i X.ncu-rep --page source --csv > sass.csv
  cuobjdump -xelf all libsrbd_b200.so ; nvdisasm -gi capi.sm_100a.cubin > dis.txt
  python scripts/sass_profile.py sass.csv dis.txt <kernel-mangled-substring> [file.cuh] [n_units]
"""
import collections
import csv
import re
import sys


def parse_dis(path, kernel):
    """-> list of (opcode, [(file, line), ...innermost first])"""
    out = []
    active = False
    chain = []
    fresh = True
    for ln in open(path, errors="replace"):
        if ln.startswith(".text."):
            active = kernel in ln
            continue
        if not active:
            continue
        if "//## File" in ln:  # consecutive annotation lines spell out one inline chain, innermost first
            if fresh:
                chain = []
                fresh = False
            for m in re.finditer(r'"([^"]+)", line (\d+)', ln):
                c = (m.group(1).split("/")[-1], int(m.group(2)))
                if not chain or chain[-1] != c:
                    chain.append(c)
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            toks = m.group(2).split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            out.append((op, list(chain)))
            fresh = True
    return out


import os
SKIP = [tuple(map(int, x.split("-"))) for x in os.environ.get("SASS_SKIP", "").split(",") if x]
REGIONS = [tuple(map(int, x.split("-"))) for x in os.environ.get("SASS_REGIONS", "").split(",") if x]


def main():
    sass, dis, kernel = sys.argv[1:4]
    fname = sys.argv[4] if len(sys.argv) > 4 else None
    units = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
    rows = list(csv.reader(open(sass)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) >= len(hdr)]
    d = parse_dis(dis, kernel)
    if len(d) != len(data):
        print("warning: instruction count differs", len(d), len(data))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    by_line = collections.defaultdict(lambda: collections.Counter())
    tot = collections.Counter()
    for (op, chain), r in zip(d, data):
        sop = r[ix["Source"]].split()
        sop = sop[1] if sop[0].startswith("@") else sop[0]
        if sop.split(".")[0] != op.split(".")[0]:
            print("mismatch", op, sop)
            break
        n = int(r[ix["Instructions Executed"]])
        s = int(r[ix["# Samples"]])
        # outermost location inside the file of interest
        locs = [c for c in chain if fname is None or c[0] == fname]
        # outermost location that is not inside a "driver" range (SASS_SKIP="lo-hi,lo-hi")
        cand = [c for c in locs if not any(lo <= c[1] <= hi for lo, hi in SKIP)]
        key = cand[-1][1] if cand else (locs[-1][1] if locs else -1)
        if REGIONS:
            key = next((lo for lo, hi in REGIONS if lo <= key <= hi), key)
        inner = locs[0][1] if locs else -1
        c = by_line[(key, inner)]
        c["instr"] += n
        c["samples"] += s
        c["lanes"] += int(r[ix["Predicated-On Thread Instructions Executed"]])
        base = op.split(".")[0]
        c["op_" + base] += n
        for h in stall_cols:
            c[h] += int(r[ix[h]])
        tot["instr"] += n
        tot["samples"] += s
    print("total instr %d  samples %d  per unit %.0f" % (tot["instr"], tot["samples"], tot["instr"] / units))
    # per outer line
    outer = collections.defaultdict(collections.Counter)
    for (key, inner), c in by_line.items():
        outer[key].update(c)
    print("%6s %10s %7s %7s %6s  ops / stalls" % ("line", "instr/unit", "instr%", "samp%", "lanes"))
    for key, c in sorted(outer.items(), key=lambda kv: -kv[1]["samples"])[:int(os.environ.get("SASS_TOP", "70"))]:
        ops = sorted(((k[3:], v) for k, v in c.items() if k.startswith("op_")), key=lambda kv: -kv[1])[:4]
        st = sorted(((k[6:], v) for k, v in c.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:3]
        print("%6d %10.0f %6.2f%% %6.2f%% %6.1f  %s | %s" % (
            key, c["instr"] / units, 100.0 * c["instr"] / tot["instr"], 100.0 * c["samples"] / tot["samples"],
            c["lanes"] / max(c["instr"], 1),
            " ".join("%s:%.0f%%" % (o, 100.0 * v / max(c["instr"], 1)) for o, v in ops),
            " ".join("%s:%.0f%%" % (o, 100.0 * v / max(c["samples"], 1)) for o, v in st)))
    return outer, tot


if __name__ == "__main__":
    main()
