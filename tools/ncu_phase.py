"""Attribute the warp-stall samples of one kernel of an ncu report to source lines / phases.

    python tools/ncu_phase.py REPORT.ncu-rep KERNEL_SUBSTRING [name:lo-hi ...]      (lo-hi: line range of rnmpc_qp.cuh / rnmpc_sls.cuh)

ncu's source page is per SASS instruction; the line table comes from nvdisasm -gi of the cubin inside librnmpc_b200.so
(built with -lineinfo).  Samples are joined on the instruction offset.  With one resident warp per scheduler (small batch)
the sample share of a line is its share of the warp's latency.
"""
import csv, io, os, re, subprocess, sys, collections, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]
phases = []
for a in sys.argv[3:]:
    n, r = a.split(":"); lo, hi = r.split("-"); phases.append((n, int(lo), int(hi)))

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "robust-nonlinear-mpc_b200", "librnmpc_b200.so")], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
# split per kernel
blocks, cur = [], None
for line in src.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]; blocks.append(cur)
    elif cur is not None:
        cur.append(line)
blocks = [b for b in blocks if kern in b[0]]
if not blocks:
    sys.exit("kernel not in report")
mangled = None
samples = collections.Counter(); instr = {}; execd = collections.Counter()
for b in blocks[:1]:
    rd = list(csv.reader(io.StringIO("\n".join(b[1:]))))
    hdr = rd[0]; ia, isrc, isam = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
    base = int(rd[1][ia], 16)
    for r in rd[1:]:
        off = int(r[ia], 16) - base
        samples[off] += int(r[isam] or 0); instr[off] = r[isrc].strip(); execd[off] += int(r[iex] or 0)
# line table
table = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    out = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    inside, chain, fresh = False, [], True
    for line in out.splitlines():
        if line.startswith(".text."):
            name = line[6:].rstrip(":")
            inside = all(t in name for t in re.findall(r"[A-Za-z_]+|\d+", kern.replace("(int)", "")) if len(t) > 2 or t.isdigit())
            continue
        if not inside: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            if fresh: chain, fresh = [], False          # a new group of location lines (innermost first, callers after)
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(\S.*);", line)
        if m:
            fresh = True
            if chain: table.setdefault(int(m.group(1), 16), tuple(chain))
    if table: break
tot = sum(samples.values())
ex_phase = collections.Counter(); ex_stack = collections.Counter(); ex_op = collections.Counter(); by_stack = collections.Counter(); by_line = collections.Counter(); by_phase = collections.Counter(); by_func = collections.Counter()
main = ("rnmpc_qp.cuh", "rnmpc_sls.cuh", "rnmpc_qp_block.cuh")
repo = lambda f: f.startswith("rnmpc_")
for off, n in samples.items():
    chain = table.get(off)
    if chain is None:
        by_line[("?", 0)] += n; by_phase["unmapped"] += n; continue
    own = [c for c in chain if repo(c[0])]
    inner = own[0] if own else chain[0]
    by_line[inner] += n
    by_stack[" < ".join("%s:%d" % (c[0].replace("rnmpc_", "").replace(".cuh", ""), c[1]) for c in own[:4])] += n
    key = next((c[1] for c in own if c[0] in main), None)
    ph = "other"
    if key is not None:
        for nme, lo, hi in phases:
            if lo <= key <= hi: ph = nme; break
        else:
            ph = "unphased"
    by_phase[ph] += n
    ex_phase[ph] += execd[off]
    ex_stack[" < ".join("%s:%d" % (c[0].replace("rnmpc_", "").replace(".cuh", ""), c[1]) for c in own[:4])] += execd[off]
    ex_op[instr[off].replace("@", " ").split()[1 if instr[off].startswith("@") else 0].split(".")[0]] += execd[off]
print("total samples", tot, "instructions", len(instr))
print("--- phases")
for k, v in by_phase.most_common(): print("%-14s %6.2f%%" % (k, 100.0 * v / tot))
print("--- top lines")
for (f, l), v in by_line.most_common(45): print("%-22s %5d %6.2f%%" % (f, l, 100.0 * v / tot))
print("--- top stacks (innermost < callers)")
for k, v in by_stack.most_common(int(os.environ.get("NSTACK", "60"))): print("%6.2f%%  %s" % (100.0 * v / tot, k))
etot = sum(execd.values())
print("--- warp instructions executed: total", etot)
for k, v in ex_phase.most_common(): print("%-14s %6.2f%%  %d" % (k, 100.0 * v / etot, v))
print("--- by opcode")
for k, v in ex_op.most_common(25): print("%-10s %6.2f%%" % (k, 100.0 * v / etot))
print("--- top stacks by instructions")
for k, v in ex_stack.most_common(40): print("%6.2f%%  %s" % (100.0 * v / etot, k))
st_phase = collections.Counter()
for off in instr:
    chain = table.get(off)
    own = [c for c in (chain or ()) if repo(c[0])]
    key = next((c[1] for c in own if c[0] in main), None)
    ph = "other"
    if key is not None:
        ph = next((n for n, lo, hi in phases if lo <= key <= hi), "unphased")
    st_phase[ph] += 1
print("--- static instructions per phase (x16 bytes)")
for k, v in st_phase.most_common(): print("%-14s %6d  %5.1f KB" % (k, v, v * 16 / 1024.0))
