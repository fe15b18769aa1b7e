"""DRAM bytes per launch of the dominant kernel from `ncu --set full` captures of bench.py's timed steps, merged into
profiles/traffic.json (read by bench.py for `roofline.traffic`):

    BENCH_CUDA_PROFILER=1 ncu --set full --clock-control none --profile-from-start off -k regex:<kernel> -c <n> -o out \
        python bench.py --config cfgX --steps 1 --no-e2e --no-cpu-baseline --no-also --no-parity
    python tools/traffic_from_ncu.py cfgX[@shards]=out.ncu-rep ...

Per key: the SUM over the captured launches (one step: k_gram launches once per step, k_loo_tiles once per 4096-fold chunk)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
path = os.path.join(ROOT, "profiles", "traffic.json")
tj = json.load(open(path)) if os.path.exists(path) else {}
src = tj.get("sources", {})
for arg in sys.argv[1:]:
    key, rep = arg.split("=", 1)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print("no kernels in", rep)
        continue
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    total, names = 0.0, []
    for r in rows[2:]:
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(r[col[m]]) * scale[units[col[m]]]
        names.append(r[col["Kernel Name"]].split("(")[0])
    tj[key] = total
    src[key] = f"{os.path.basename(rep)}: {len(rows) - 2} launch(es) of {sorted(set(names))}, dram__bytes_read.sum + dram__bytes_write.sum"
    print(key, total)
tj["sources"] = src
json.dump(tj, open(path, "w"), indent=1)
