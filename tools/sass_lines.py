"""SASS instruction count per source line of one kernel (needs -lineinfo): python tools/sass_lines.py OBJ KERNEL_SUBSTR [SRC] [top]
Code size is what the instruction caches see: a loop body beyond ~32 KB (2000 instructions) streams from L2 on every trip."""
import collections, os, re, subprocess, sys, tempfile
obj, kern = sys.argv[1], sys.argv[2]
srcf = sys.argv[3] if len(sys.argv) > 3 else None
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
secs = [i for i, l in enumerate(txt) if l.lstrip().startswith(".section") and ".text." in l]
s = [i for i in secs if kern in txt[i]][0]
e = min([i for i in secs if i > s] + [len(txt)])
cnt, cur, n = collections.Counter(), None, 0
for l in txt[s:e]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        cnt[cur] += 1
        n += 1
print(txt[s].split(",")[0].split(".text.")[1], ":", n, "instructions =", n * 16 // 1024, "KB")
src = open(srcf).read().splitlines() if srcf else None
for (f, ln), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]:
    print("%5d  %s:%d  %s" % (c, f, ln, src[ln - 1].strip()[:100] if src and srcf.endswith(f) else ""))
