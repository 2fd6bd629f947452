#!/usr/bin/env python3
"""Join an ncu SASS-level source page with nvdisasm line info and aggregate per source line.

usage: ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-substring> [top_n]
Prints instructions executed and stall samples by (file:line) and by top-level stall reason.
"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, so, kname = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
line_of = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn, cur_line, active = None, None, False
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            active = kname in m.group(1)
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            line_of[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
csvp = os.path.join(tmp, "src.csv")
with open(csvp, "w") as fh:
    subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=fh, stderr=subprocess.DEVNULL)
rows = list(csv.reader(open(csvp)))
# pick the block belonging to the kernel
start = None
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name" and (os.environ.get("NCU_KERNEL", "") in r[1]):
        start = i
        break
hdr = rows[start + 1]
col = {h: i for i, h in enumerate(hdr)}
base = None
by_line = collections.defaultdict(lambda: [0, 0, 0])
stalls = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_inst = tot_samp = 0
for r in rows[start + 2:]:
    if not r or r[0] == "Kernel Name":
        break
    addr = int(r[col["Address"]], 16)
    if base is None:
        base = addr
    off = addr - base
    inst = int(r[col["Instructions Executed"]] or 0)
    samp = int(r[col["# Samples"]] or 0)
    key = line_of.get(off, (None, ""))[0]
    e = by_line[key]
    e[0] += inst; e[1] += samp; e[2] += 1
    tot_inst += inst; tot_samp += samp
    for s in stall_cols:
        v = r[col[s]]
        if v and v != "0":
            stalls[s] += int(v)
print("total warp-instructions executed: %d   samples: %d   static SASS: %d" % (tot_inst, tot_samp, sum(e[2] for e in by_line.values())))
print("stall reasons:", ", ".join("%s=%.1f%%" % (k, 100.0 * v / max(tot_samp, 1)) for k, v in stalls.most_common(8)))
byfile = collections.defaultdict(lambda: [0, 0, 0])
for k, e in by_line.items():
    f = k[0] if k else "?"
    for i in range(3):
        byfile[f][i] += e[i]
print("\nby file: inst%  samples%  static")
for f, e in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print("  %-22s %6.2f %6.2f %6d" % (f, 100.0 * e[0] / tot_inst, 100.0 * e[1] / max(tot_samp, 1), e[2]))
# ---- per function (nearest preceding `__device__`/`__global__`/RLM_HD definition line in the same file)
import glob
src_dirs = [os.path.join(os.path.dirname(os.path.abspath(so)), "csrc"), os.path.join(os.path.dirname(os.path.abspath(so)), "..", "include")]
func_of = {}
src_text = {}
for d in src_dirs:
    for path in glob.glob(os.path.join(d, "*")):
        if not os.path.isfile(path):
            continue
        try:
            lines = open(path).read().splitlines()
        except Exception:
            continue
        cur = "?"
        tab = {}
        for i, ln in enumerate(lines, 1):
            m = re.match(r"^(?:template.*>\s*)?(?:static\s+)?(?:__device__|__global__|RLM_HD)\b[^;{]*?(\w+)\s*\(", ln)
            if m and not ln.startswith(" "):
                cur = m.group(1)
            tab[i] = cur
        func_of[os.path.basename(path)] = tab
        src_text[os.path.basename(path)] = lines
byfunc = collections.defaultdict(lambda: [0, 0, 0, 0])
for k, e in by_line.items():
    fn = func_of.get(k[0], {}).get(k[1], k[0]) if k else "?"
    byfunc[fn][0] += e[0]; byfunc[fn][1] += e[1]; byfunc[fn][2] += e[2]
print("\nby function (attributed by line): inst%  samples%  static")
for f, e in sorted(byfunc.items(), key=lambda x: -x[1][0])[:40]:
    print("  %-28s %6.2f %6.2f %6d" % (f, 100.0 * e[0] / tot_inst, 100.0 * e[1] / max(tot_samp, 1), e[2] // 2 if False else e[2]))
print("\ntop lines by instructions executed: inst%  samples%  static  file:line")
for k, e in sorted(by_line.items(), key=lambda x: -x[1][0])[:topn]:
    txt = src_text.get(k[0], [""] * (k[1] + 1))[k[1] - 1].strip()[:90] if k and k[0] in src_text and k[1] <= len(src_text[k[0]]) else ""
    print("  %6.2f %6.2f %5d  %-22s %s" % (100.0 * e[0] / tot_inst, 100.0 * e[1] / max(tot_samp, 1), e[2], "%s:%d" % k if k else "?", txt))
print("\ntop lines by stall samples:")
for k, e in sorted(by_line.items(), key=lambda x: -x[1][1])[:topn]:
    print("  %6.2f %6.2f %5d  %s" % (100.0 * e[0] / tot_inst, 100.0 * e[1] / max(tot_samp, 1), e[2], "%s:%d" % k if k else "?"))
