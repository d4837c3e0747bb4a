"""Turn gpurun_out ncu artefacts into small tracked summaries under profiles/.

    python profiles/summarize.py <tag>      # e.g. r01 -> reads gpurun_out/<tag>_launches.csv, <tag>_prof_*.ncu-rep
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")

RAW_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum",
]


def launches(tag, stem="launches", out="launch_list",
             what="ONE U-ViT-L velocity evaluation at batch 64"):
    p = os.path.join(GO, f"{tag}_{stem}.csv")
    if not os.path.exists(p):
        return
    rows = list(csv.reader(open(p)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > mv:
            agg.setdefault(r[kn].split("(")[0].replace("void ", "").replace("usp::<unnamed>::", "")[-48:], []).append(
                float(r[mv].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(OUT, f"{tag}_{out}.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list of {what} "
                "(`--metrics gpu__time_duration.sum --clock-control none`, cold-cache, serialised: compare shares)\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in agg.items():
            f.write(f"| `{k}` | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f} % |\n")
        f.write(f"| **total** | {sum(len(v) for v in agg.values())} | {tot / 1e3:.1f} | | |\n")


def rep(tag, name):
    p = os.path.join(GO, f"{tag}_prof_{name}.ncu-rep")
    if not os.path.exists(p):
        return
    raw = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    with open(os.path.join(OUT, f"{tag}_ncu_{name}.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none` capture, {name} kernel(s), U-ViT-L batch 64\n\n")
        for r in rows[2:]:
            f.write(f"## {r[h.index('Kernel Name')][:100]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in RAW_METRICS:
                if m in h:
                    i = h.index(m)
                    f.write(f"| {m} | {r[i]} | {rows[1][i]} |\n")
            f.write("\n")
    if name == "gemm":
        # per-launch DRAM traffic of each GEMM class, read by bench.py for `roofline.traffic` (U-ViT-L, batch 64)
        import json
        import re
        cls = {"0": "gemm_qkv", "1": "gemm_fc1", "3": "gemm_skip"}
        kern = {}
        for r in rows[2:]:
            m = re.search(r"gemm2_kernel<\(?(?:int\))?(\d), \(?(?:bool\))?(\d)", r[h.index("Kernel Name")])
            if not m:
                continue
            c = cls.get(m.group(1)) or ("gemm_fc2" if m.group(2) == "1" else "gemm_proj")

            def val(metric):
                i = h.index(metric)
                v = float(r[i].replace(",", ""))
                u = rows[1][i].lower()
                return v * {"mbyte": 1e6, "gbyte": 1e9, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            kern[c] = {"dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                       "time_us": float(r[h.index("gpu__time_duration.sum")].replace(",", "")),
                       "tensor_active_pct": float(r[h.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")])}
        with open(os.path.join(OUT, "ncu_traffic.json"), "w") as f:
            json.dump({"source": f"profiles/{tag}_ncu_gemm.md (ncu --set full --clock-control none, U-ViT-L batch 64, one launch each)",
                       "kernels": kern}, f, indent=1)
    src = subprocess.run(["ncu", "-i", p, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    ks, cur = [], None
    for r in csv.reader(src.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1], rows=[])
            ks.append(cur)
        elif r and r[0] == "Address" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and len(r) > 10:
            cur["rows"].append(r)
    with open(os.path.join(OUT, f"{tag}_ncu_{name}.md"), "a") as f:
        for k in ks:
            if "hdr" not in k:
                continue
            h2 = k["hdr"]
            si = h2.index("# Samples")
            tot = sum(int(r[si]) for r in k["rows"]) or 1
            sc = [i for i, c in enumerate(h2) if c.startswith("stall_") and "Not Issued" not in c]
            f.write(f"### top stall sites: {k['name'][:90]}\n\n| % samples | SASS | top stall reasons |\n|---:|---|---|\n")
            for r in sorted(k["rows"], key=lambda r: -int(r[si]))[:10]:
                st = sorted(((h2[i][6:], int(r[i])) for i in sc if int(r[i]) > 0), key=lambda x: -x[1])[:3]
                f.write(f"| {100 * int(r[si]) / tot:.1f} | `{r[1].strip()[:70]}` | {st} |\n")
            f.write("\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag)
    launches(tag, "vae_launches", "vae_launch_list", "ONE decode of 8 latents through the autoencoder (second call)")
    for n in ("gemm", "attn", "vae"):
        rep(tag, n)
