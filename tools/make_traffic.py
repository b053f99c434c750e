"""profiles/traffic.json from an `ncu --set full` capture of tools/profile_step.py: per kernel, per launch,
dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum -- stamped with a hash of the kernel sources,
so bench.py can refuse numbers that no longer belong to the kernels it runs.

    python tools/make_traffic.py gpurun_out/<capture>.ncu-rep [scene-name]"""
import csv, hashlib, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sources_sha():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "luciddreamer_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def main():
    rep, scene = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "shell")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[0], rows[2:]
    ki, ri, wi, ii = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                              "smsp__inst_executed.sum"))
    units = rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic, inst = {}, {}
    for r in data:
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "").split("<")[0].strip()
        if not name.startswith("k_"):
            continue
        b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
        traffic.setdefault(name, []).append(b)
        inst.setdefault(name, []).append(float(r[ii]))
    path = os.path.join(ROOT, "profiles", "traffic.json")
    tj = json.load(open(path)) if os.path.exists(path) else {}
    tj[scene] = {k: sum(v) / len(v) for k, v in traffic.items()}
    tj[scene + "_warp_inst"] = {k: sum(v) / len(v) for k, v in inst.items()}
    tj["sources_sha"] = sources_sha()
    tj["source_" + scene] = f"{os.path.basename(rep)} (ncu --set full --clock-control none --import-source on; per launch, mean over the launches captured)"
    json.dump(tj, open(path, "w"), indent=1)
    print(json.dumps(tj, indent=1))


if __name__ == "__main__":
    main()
