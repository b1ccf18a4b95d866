"""Per-source-line totals (warp instructions executed, stall samples) from an .ncu-rep captured with
--import-source on; development tool, runs here without a GPU.
usage: ncu_lines.py REPORT [top_n] [-s]   (-s: also list the SASS of the top lines)"""
import csv, subprocess, sys, collections

def num(v):
    try:
        return int(v)
    except ValueError:
        return 0

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
show_sass = "-s" in sys.argv
by_samples = "--stalls" in sys.argv
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, key = None, None, None
agg = collections.OrderedDict()
sass = collections.defaultdict(list)
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0].isdigit():
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [r[1].strip(), 0, 0, 0])
        a[1] += num(r[iI]); a[2] += num(r[iS])
    elif key is not None and r[2].startswith("0x"):
        agg[key][3] += 1
        sass[key].append((r[3].strip(), num(r[iI]), num(r[iS])))
tot_i = sum(a[1] for a in agg.values()) or 1
tot_s = sum(a[2] for a in agg.values()) or 1
print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2 if by_samples else 1])[:top]:
    print("%5.1f%% inst %5.1f%% smpl  n=%3d  %s:%d  %s" % (100.0 * a[1] / tot_i, 100.0 * a[2] / tot_s, a[3], key[0], key[1], a[0][:90]))
    if show_sass:
        for ins, ni, ns in sass[key]:
            print("              %10d %6d  %s" % (ni, ns, ins[:100]))
