"""Per-SASS-instruction executed counts of one kernel from an .ncu-rep (source page), hottest regions first."""
import csv, io, subprocess, sys
rep, kname = sys.argv[1], sys.argv[2]
thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
for b in blocks[1:]:
    lines = b.splitlines()
    name = lines[0]
    if kname not in name:
        continue
    rd = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rd[0]
    ie, src, st = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    rows = [(r[src].strip(), int(r[ie] or 0), int(r[st] or 0)) for r in rd[1:] if len(r) > ie]
    tot = sum(r[1] for r in rows); tots = sum(r[2] for r in rows)
    print(f"## {kname}: {len(rows)} SASS instr, {tot/1e6:.1f} M warp-instr executed, {tots} stall samples")
    for i, (s, n, stl) in enumerate(rows):
        if n >= thresh * tot:
            print(f"{i:5d} {n/1e6:8.2f}M {100*n/tot:5.2f}%  st {100*stl/max(tots,1):5.2f}%  {s}")
    break
