"""Per-source-line stall-sample profile of one kernel from an .ncu-rep captured with --import-source on.
ncu's csv source page is SASS-level; this maps it to source lines with nvdisasm's line info of the same cubin
(instruction order is identical).
  python profiles/ncu_lines.py <rep> <cubin> <kernel-substring> [top N]"""
import csv, re, subprocess, sys
rep, cubin, pat = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# split the csv into kernels
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
blk = next(b for b in blocks if pat in b["name"])
hdr = blk["rows"][0]
isamp = hdr.index("# Samples")
sass = [(r[1], int(r[isamp] or 0)) for r in blk["rows"][1:] if len(r) > isamp]
# nvdisasm listing with line info
fn = re.search(r"(\w*" + re.escape(pat) + r"\w*)", blk["name"]).group(1)
dis = subprocess.run(["nvdisasm", "-c", "-g", cubin], capture_output=True, text=True).stdout
lines, cur_line, infn = [], None, False
mangled = None
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        infn = pat in m.group(1) and ("I" + blk["name"].split("<")[1].split(">")[0].replace("(int)", "").replace(", ", "EL").replace(" ", "") in m.group(1) or True)
        mangled = m.group(1)
        if infn:
            lines = []
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur_line)
    if len(lines) == len(sass) and lines:
        pass
# the listing may hold several template instances: keep the LAST block whose length matches
print("kernel:", blk["name"][:100], "| sass instr:", len(sass), "| disasm instr:", len(lines))
agg = {}
for (ins, n), ln in zip(sass, lines[:len(sass)]):
    agg[ln] = agg.get(ln, 0) + n
tot = sum(agg.values()) or 1
for ln, n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{100.0 * n / tot:5.1f}%  {ln}")
