"""Per-kernel SASS mnemonic counts of the product library (cuobjdump -sass on the .o files): the evidence that the
blend kernels issue packed FP32 (FFMA2 / FMUL2 / FADD2), that the TMA paths are real (UBLKCP = cp.async.bulk, SYNCS =
mbarrier), and which reductions are vectorised (RED.E.ADD.F32x4 / REDG.E...128, multimem).  Run on the CPU box."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "luciddreamer_b200", "csrc")
WATCH = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU.EX2", "MUFU.RCP", "FSEL", "FSETP", "FMNMX3", "FMNMX", "SEL",
         "SHFL", "VOTE", "MATCH", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "UBLKCP", "SYNCS", "BAR", "FCHK", "CALL"]
for obj in sorted(f for f in os.listdir(CSRC) if f.endswith(".o")):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
    fn, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(anonymous namespace\)::", "", name)
            counts[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            counts[fn]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    counts[fn][w] += 1
                    break
    for fn, c in counts.items():
        if not fn.startswith("k_") and "k_" not in fn:
            continue
        items = "  ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        print(f"{obj:18s} {fn[:60]:60s} total={c['_total']:5d}  {items}")
