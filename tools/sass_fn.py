"""Extract one device subroutine from the nvdisasm listing of the in-tree cubin, with source line tags.
usage: sass_fn.py SUBSTRING [--hist]   (development tool, no GPU needed)"""
import os, re, subprocess, sys, tempfile, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(ROOT, "airs-compression_b200", "build", "airs_kernels.o")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
want = sys.argv[1]
starts = [i for i, l in enumerate(dis) if l.lstrip().startswith(".type") and "@function" in l]
for k, i in enumerate(starts):
    if want in dis[i]:
        j = starts[k + 1] if k + 1 < len(starts) else len(dis)
        body = dis[i:j]
        break
else:
    raise SystemExit("not found")
line = None
hist = collections.Counter()
out = []
for l in body:
    m = re.search(r'//## File ".*/([^/"]+)", line (\d+)', l)
    if m:
        line = "%s:%s" % (m.group(1), m.group(2)); continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        op = m.group(2).split()[0] if not m.group(2).startswith("@") else m.group(2).split()[1]
        hist[op.split(".")[0]] += 1
        out.append("%-28s %s" % (line or "", m.group(2)))
    elif l.strip().startswith(".L_"):
        out.append(l.strip())
if "--hist" in sys.argv:
    print(sum(hist.values()), "instructions")
    for op, n in hist.most_common(30):
        print("%6d %s" % (n, op))
else:
    print("\n".join(out))
