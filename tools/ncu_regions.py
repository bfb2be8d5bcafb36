"""Region profile of one kernel from `ncu -i X.ncu-rep --page source --csv`: consecutive SASS instructions with the same
execution count are merged into regions (loop bodies, phases between barriers), each printed with its share of the executed
warp instructions, its share of the stall samples, its instruction mix and its top stall reasons."""
import csv
import sys


def main(path, min_share=0.005):
    rows = list(csv.reader(open(path)))
    print(rows[0][1] if len(rows[0]) > 1 else rows[0])
    hdr, data = rows[1], rows[2:]
    i_src, i_exec, i_samp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stalls = {n: hdr.index(n) for n in hdr if n.startswith("stall_") and "Not Issued" not in n}
    tot = sum(int(r[i_exec]) for r in data)
    tot_s = max(1, sum(int(r[i_samp]) for r in data))
    print("warp instructions executed %d, stall samples %d, SASS instructions %d" % (tot, tot_s, len(data)))
    groups, cur = [], None
    for k, r in enumerate(data):
        e, s = int(r[i_exec]), int(r[i_samp])
        toks = r[i_src].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        if cur is None or abs(cur["e"] - e) > 0.02 * max(cur["e"], 1):
            cur = {"e": e, "n": 0, "inst": 0, "samp": 0, "start": k, "ops": {}, "st": {}}
            groups.append(cur)
        cur["n"] += 1
        cur["inst"] += e
        cur["samp"] += s
        cur["end"] = k
        cur["ops"][op] = cur["ops"].get(op, 0) + 1
        for n, i in stalls.items():
            cur["st"][n] = cur["st"].get(n, 0) + int(r[i] or 0)
    for g in groups:
        if g["inst"] < min_share * tot:
            continue
        ops = sorted(g["ops"].items(), key=lambda x: -x[1])[:7]
        st = sorted(g["st"].items(), key=lambda x: -x[1])[:4]
        print("sass %4d-%4d (%3d instr) x %8d: inst %5.1f %%  samples %5.1f %%  %s | %s" % (
            g["start"], g["end"], g["n"], g["e"], 100.0 * g["inst"] / tot, 100.0 * g["samp"] / tot_s,
            " ".join("%s:%d" % o for o in ops), " ".join("%s:%d" % (a.replace("stall_", ""), b) for a, b in st)))


if __name__ == "__main__":
    main(sys.argv[1])
