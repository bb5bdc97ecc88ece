#!/usr/bin/env python
"""Attribute executed warp-instructions of an ncu capture to CUDA source lines.

usage: attribute.py <report.ncu-rep> <lib.so> <kernel-substring> [top_n]
The substring must select ONE instantiation when the kernel is a template (e.g. cn_flat_kernelILi0ELi256E: the line
tables of the instantiations are keyed by code offset and would overwrite each other), and <lib.so> must be the binary
the capture was taken from.
Joins `ncu --page source --print-source=sass --csv` (per-SASS-address executed
counts, first captured launch) with `nvdisasm -g` line info of the same cubin.
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, so, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
addr2line = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["nvdisasm", "-gi", "-c", f], cwd=tmp, stdout=subprocess.PIPE, text=True).stdout
    cur_fn, cur, chain, fresh = None, None, [], True
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur_fn = m.group(1); cur = None; chain = []; continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            if fresh:
                chain = []; fresh = False
            chain.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if m:
            fresh = True
            if cur_fn and kname in cur_fn and chain:
                leaf = chain[0]
                # innermost frame that lies in the kernel source (for phase attribution)
                outer = next((c for c in chain if c[0].endswith(".cu")), leaf)
                addr2line[int(m.group(1), 16)] = (leaf, m.group(2).strip(), outer)
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source=sass", "--csv"], stdout=subprocess.PIPE, text=True).stdout
blocks = txt.split('"Kernel Name"')
per_line = collections.Counter(); per_line_samples = collections.Counter(); total = 0; tot_s = 0
per_outer = collections.Counter(); per_outer_s = collections.Counter()
done = False
for b in blocks[1:]:
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + b)))
    if kname not in rows[0][1].replace("(int)", "").replace(" ", "") and kname not in rows[0][1]:
        pass
    hdr = rows[1]
    ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = None
    for r in rows[2:]:
        if len(r) <= ie or not r[ia]:
            continue
        a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        if base is None:
            base = a
        ent = addr2line.get(a - base, ((None, 0), "?", (None, 0)))
        key = ent[0]
        n = int(float(r[ie] or 0)); s = int(float(r[isamp] or 0))
        per_line[key] += n; per_line_samples[key] += s; total += n; tot_s += s
        per_outer[ent[2]] += n; per_outer_s[ent[2]] += s
    break   # first launch only
src_cache = {}
def src(fl):
    f, l = fl
    if f is None: return ""
    for d in ("crowdnav_b200/csrc", "include"):
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][l - 1].strip()[:90] if l - 1 < len(src_cache[p]) else ""
    return ""
print("total warp-instructions executed: %d ; samples %d" % (total, tot_s))
print("%8s %6s %6s  %-18s %s" % ("inst", "%inst", "%samp", "file:line", "source"))
for k, n in per_line.most_common(top_n):
    print("%8d %6.2f %6.2f  %-18s %s" % (n, 100.0 * n / max(total, 1), 100.0 * per_line_samples[k] / max(tot_s, 1),
                                          "%s:%d" % k if k[0] else "?", src(k)))

# ---- by phase: kernel-source lines grouped under the nearest preceding "// ----" marker
csrc_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "crowdnav_b200", "csrc")
labels_by_file = {}
def labels_of(fname):
    if fname not in labels_by_file:
        label, labels = "(prologue)", []
        try:
            lines = open(os.path.join(csrc_dir, fname)).read().splitlines()
        except OSError:
            lines = []
        for l in lines:
            m = re.match(r"\s*// -{3,}\s*(.*?)\s*-*$", l)
            if m and m.group(1):
                label = m.group(1)[:70]
            labels.append(label)
        labels_by_file[fname] = labels
    return labels_by_file[fname]
phase = collections.Counter(); phase_s = collections.Counter()
for (f, l), n in per_outer.items():
    labels = labels_of(f) if f and f.endswith(".cu") else []
    lab = labels[l - 1] if 0 < l <= len(labels) else "(other)"
    phase[lab] += n; phase_s[lab] += per_outer_s[(f, l)]
print("\nby phase (innermost kernel-source frame):")
for lab, n in phase.most_common():
    print("%9d %6.2f%% inst %6.2f%% samples  %s" % (n, 100.0 * n / max(total, 1), 100.0 * phase_s[lab] / max(tot_s, 1), lab))

# optional: per kernel-source line (outer frame) dump, in line order: CN_ATTR_DUMP=<path>
if os.environ.get("CN_ATTR_DUMP"):
    with open(os.environ["CN_ATTR_DUMP"], "w") as fh:
        for (f, l), n in sorted(per_outer.items(), key=lambda kv: (str(kv[0][0]), kv[0][1])):
            fh.write("%-14s %5d %9d %6.2f%% inst %6.2f%% samp  %s\n" % (f, l, n, 100.0 * n / max(total, 1),
                     100.0 * per_outer_s[(f, l)] / max(tot_s, 1), src((f, l))))
