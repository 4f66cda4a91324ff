"""Warp-stall samples and executed instructions of one kernel of an ncu report, summed per SOURCE LINE (joined through nvdisasm -gi of
the object file built with -lineinfo): python tools/ncu_lines.py REPORT.ncu-rep OBJ NCU_KERNEL_SUBSTR MANGLED_SUBSTR SRC [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj, kern, mangled, srcf = sys.argv[1:6]
top = int(sys.argv[6]) if len(sys.argv) > 6 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"') and kern in l][0]
end = min([i for i, l in enumerate(lines) if l.startswith('"Kernel Name"') and i > start] + [len(lines)])
rd = list(csv.reader(io.StringIO("\n".join(lines[start + 1:end]))))
hdr = rd[0]
ia, isam, iex, isrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
base = int(rd[1][ia], 16)
samp, execd, text = {}, {}, {}
for r in rd[1:]:
    off = int(r[ia], 16) - base
    samp[off] = int(r[isam] or 0); execd[off] = int(r[iex] or 0); text[off] = r[isrc].strip()
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
secs = [i for i, l in enumerate(txt) if l.lstrip().startswith(".section") and ".text." in l]
s = [i for i in secs if mangled in txt[i]][0]
e = min([i for i in secs if i > s] + [len(txt)])
line_of, cur = {}, None
for l in txt[s:e]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
by_s, by_e = collections.Counter(), collections.Counter()
for off in samp:
    k = line_of.get(off)
    by_s[k] += samp[off]; by_e[k] += execd[off]
tot_s, tot_e = sum(by_s.values()), sum(by_e.values())
src = open(srcf).read().splitlines()
print("total samples", tot_s, "instructions executed", tot_e)
for k, c in by_s.most_common(top):
    t = src[k[1] - 1].strip()[:90] if k and srcf.endswith(k[0]) else ""
    print("%6.2f %% samples  %6.2f %% instr   %s:%s  %s" % (100.0 * c / tot_s, 100.0 * by_e[k] / tot_e, k[0] if k else "?", k[1] if k else "?", t))
