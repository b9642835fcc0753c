"""Development: per-source-line instruction counts and stall samples of one kernel of an ncu report.
usage: python tools/ncu_lines.py report.ncu-rep build/file.o mangled_kernel_substring [min_share_pct]
Joins `ncu --page source --csv` (SASS rows) with `nvdisasm -g` line info of the same object file."""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj, kern = sys.argv[1:4]
min_share = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
# locate the kernel's text section
line_of = {}
in_fn = False; cur = None
for l in dis:
    if l.startswith("\t.section\t.text."):
        in_fn = kern in l
        continue
    if not in_fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = int(m.group(2)); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m: line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
src_lines = open(re.search(r'File "([^"]+)"', "\n".join(dis)).group(1)).read().splitlines() if dis else []
for b in blocks[1:]:
    rows = list(csv.reader(io.StringIO('"Kernel Name",' + b)))
    name = rows[0][1]
    if kern not in name and not any(part in name for part in re.findall(r"[a-z_]{5,}", kern)): continue
    hdr = rows[1]
    ia, isrc, ii, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = None; per = collections.Counter(); smp = collections.Counter(); tot = 0; tsmp = 0
    for r in rows[2:]:
        if len(r) <= ii: continue
        try: addr = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia]); n = int(r[ii]); s = int(r[ismp] or 0)
        except ValueError: continue
        if base is None: base = addr
        ln = line_of.get(addr - base)
        per[ln] += n; smp[ln] += s; tot += n; tsmp += s
    print(f"== {name}: {tot} warp instructions, {tsmp} samples")
    for ln, n in sorted(per.items(), key=lambda kv: -kv[1]):
        if 100.0 * n / tot < min_share: break
        text = src_lines[ln - 1].strip()[:110] if ln and ln <= len(src_lines) else "?"
        print(f"{100.0 * n / tot:5.1f}% inst {100.0 * smp[ln] / max(1, tsmp):5.1f}% smp  L{ln}: {text}")
    break
