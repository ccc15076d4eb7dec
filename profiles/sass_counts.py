"""Per-kernel SASS mnemonic counts of the in-tree library (cuobjdump -sass): the evidence behind "runs on tcgen05 / TMA".
UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / st (tensor memory), UBLKCP = cp.async.bulk (TMA bulk copy),
UTCBAR = tcgen05.commit, LDGSTS = cp.async, HMMA = warp-level mma.sync, FFMA = fp32 CUDA-core FMA.
Usage: python profiles/sass_counts.py > profiles/sass_counts.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "pmhc_diffusion_model_b200", "libpmhc_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "LDGSTS", "HMMA", "FFMA", "total"]
counts, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = demangle(m.group(1))
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        op = m.group(1)
        counts[name]["total"] += 1
        for k in OPS:
            if op.startswith(k):
                counts[name][k] += 1
print(f"{'kernel':62s} " + " ".join(f"{k:>8s}" for k in OPS))
for k, c in sorted(counts.items(), key=lambda kv: -kv[1]["total"]):
    print(f"{k[:62]:62s} " + " ".join(f"{c[o]:8d}" for o in OPS))
