"""Attribute ncu stall samples (SASS source page) to CUDA source lines via nvdisasm line info (development aid).
usage: ncu_lines.py <report.ncu-rep> <kernel regex> <object.o> <mangled substring> [top]"""
import csv, re, subprocess, sys, tempfile, os, collections
rep, kre, obj, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
ix = {n: i for i, n in enumerate(rows[h])}
sass = []
for r in rows[h + 1:]:
    if len(r) <= ix["# Samples"] or not r[ix["# Samples"]].isdigit():
        break                                         # next kernel block
    sass.append((r[ix["Source"]].strip(), int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0)))
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l)
lines = []          # source line per instruction, in order
cur = None
for l in dis[start + 1:]:
    if l.startswith("//---------------------") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', l)
        cur = (os.path.basename(m.group(1)), int(m.group(2)), tuple((os.path.basename(a), int(b)) for a, b in inl))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+", l):
        lines.append(cur)
print("sass rows", len(sass), "disasm instrs", len(lines))
n = min(len(sass), len(lines))
agg = collections.Counter(); inst = collections.Counter(); outer = collections.Counter()
tot = sum(s for _, s, _ in sass)
for (txt, s, e), ln in zip(sass[:n], lines[:n]):
    key = (ln[0], ln[1]) if ln else ("?", 0)
    agg[key] += s; inst[key] += e
    ok = ln[2][-1] if ln and ln[2] else key
    outer[ok] += s
src_cache = {}
def src(f, n):
    if f not in src_cache:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    L = src_cache[f]
    return L[n - 1].strip()[:110] if 0 < n <= len(L) else ""
print("total samples", tot)
print("== by innermost line")
for k, s in agg.most_common(top):
    print("%5.1f%% %7d inst %9d  %s:%d  %s" % (100.0 * s / tot, s, inst[k], k[0], k[1], src(*k)))
print("== by outermost (kernel-level) line")
for k, s in outer.most_common(25):
    print("%5.1f%% %7d  %s:%d  %s" % (100.0 * s / tot, s, k[0], k[1], src(*k)))
