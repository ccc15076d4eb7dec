"""Development aid (no GPU): instructions executed and stall samples of one profiled kernel, summed per source line.

    cuobjdump -xelf all pmhc_diffusion_model_b200/libpmhc_b200.so         # -> egnn_pair_v3.sm_100a.cubin ...
    nvdisasm -gi -c egnn_pair_v3.sm_100a.cubin > v3.sass                    # -gi: with the inlined-at chains
    ncu -i prof.ncu-rep --page source --csv > prof_src.csv                 # SASS rows: address, instructions executed, samples
    python profiles/sass_by_line.py v3.sass prof_src.csv <mangled kernel name> [top N]

(The CUDA-source view of `ncu --page source --csv` carries no metric columns in this ncu version; nvdisasm's line table of
the very same cubin gives the mapping instead — the library must not have been rebuilt since the capture.)"""
import collections
import csv
import re
import sys

sass, prof, kernel = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
line_of = {}       # address -> innermost (file, line)
chain_of = {}      # address -> every frame of the inlined-at chain, innermost first
cur, chain, active, fresh = None, [], False, True
for ln in open(sass):
    if ln.startswith(".text."):
        active = ln.strip().rstrip(":") == ".text." + kernel
        cur, chain = None, []
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:            # first marker after an instruction: a new chain starts
            chain, fresh = [], False
        frame = (m.group(1).split("/")[-1], int(m.group(2)))
        chain.append(frame)
        cur = chain[0]
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m and cur is not None:
        line_of[int(m.group(1), 16)] = cur
        chain_of[int(m.group(1), 16)] = list(chain)
        fresh = True
rows = list(csv.reader(open(prof)))
# several kernels may be concatenated, each after a "Kernel Name" row: keep the section of the requested one (by its demangled
# name's template arguments, given as 6th argument, e.g. "(int)1, (int)2"; default: the first section)
want = sys.argv[6] if len(sys.argv) > 6 else None
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if starts:
    pick = next((i for i in starts if want is None or want in rows[i][1]), starts[0])
    nxt = next((i for i in starts if i > pick), len(rows))
    rows = rows[pick:nxt]
hdr = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
col = {n: i for i, n in enumerate(rows[hdr])}
inst = collections.Counter()
samp = collections.Counter()
addr_stat = []
tot_i = tot_s = 0
base = None
for r in rows[hdr + 1:]:
    try:
        addr = int(r[col["Address"]], 16) if not r[col["Address"]].isdigit() else int(r[col["Address"]])
    except ValueError:
        continue
    if base is None:
        base = addr
    key = line_of.get(addr - base, ("?", 0))
    i = int(float(r[col["Instructions Executed"]] or 0))
    s = int(float(r[col["# Samples"]] or 0))
    inst[key] += i
    samp[key] += s
    addr_stat.append((addr - base, i, s))
    tot_i += i
    tot_s += s
print(f"kernel {kernel}: {tot_i} warp instructions, {tot_s} samples, {len(line_of)} SASS addresses mapped")
print("  line            instr     %   samples     %")
for key, i in inst.most_common(top):
    print(f"  {key[0][:14]:14s}:{key[1]:5d} {i:10d} {100 * i / tot_i:5.1f} {samp[key]:9d} {100 * samp[key] / max(tot_s, 1):5.1f}")
if len(sys.argv) > 5 and sys.argv[5]:      # named line ranges of the main file: name=lo-hi,... ; an instruction goes to the FIRST range
    ranges = []                              # that any frame of its inlined-at chain falls into (list the helpers before their callers)
    for spec in sys.argv[5].split(","):
        name, rng = spec.split("=")
        lo, hi = map(int, rng.split("-"))
        ranges.append((name, lo, hi))
    bi, bs = collections.Counter(), collections.Counter()
    for a, i, s in addr_stat:
        frames = [l for f, l in chain_of.get(a, []) if f.endswith("v3.cu")]
        name = next((n for n, lo, hi in ranges if any(lo <= l <= hi for l in frames)), "other")
        bi[name] += i
        bs[name] += s
    for name, i in bi.most_common():
        print(f"  {name:24s} {i:10d} {100 * i / tot_i:5.1f}% instr {100 * bs[name] / max(tot_s, 1):5.1f}% samples")
