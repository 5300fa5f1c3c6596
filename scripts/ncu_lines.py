"""Attributes the executed SASS instructions / stall samples of one kernel in an .ncu-rep to CUDA source lines.
Usage: python scripts/ncu_lines.py report.ncu-rep <kernel regex for ncu> <mangled-name substring> <object file> [top]
ncu's CSV source page carries SASS only; the line table comes from nvdisasm -g on the same object (instruction order is identical)."""
import csv, re, subprocess, sys, tempfile, os
rep, kregex, mangled, obj = sys.argv[1:5]; top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
tmp = tempfile.mkdtemp(); subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.strip().startswith(".section") and ".text." in l and mangled in l][0]
end = [i for i, l in enumerate(dis) if i > start and l.strip().startswith(".section")][0]
cur, seq = None, []
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2)), int(m.group(4)) if m.group(4) else None); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kregex], capture_output=True, text=True).stdout
rows = [r for r in list(csv.reader(out.splitlines()))[2:] if len(r) > 6 and r[5].isdigit()]
rows = rows[:len(seq)]
assert len(rows) == len(seq), (len(rows), len(seq))
tot = sum(int(r[5]) for r in rows); ts = sum(int(r[2]) for r in rows); agg, smp = {}, {}
for cur, r in zip(seq, rows): agg[cur] = agg.get(cur, 0) + int(r[5]); smp[cur] = smp.get(cur, 0) + int(r[2])
print("warp instructions executed: %d, stall samples: %d, SASS instructions: %d" % (tot, ts, len(seq)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]: print("instr %5.1f%%  samples %5.1f%%  %s:%s%s" % (100 * v / tot, 100 * smp[k] / max(ts, 1), k[0], k[1], " (inlined at %d)" % k[2] if k[2] else ""))
