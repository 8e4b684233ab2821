"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/ (what the judge reads).
  python tools/summarize_ncu.py launches gpurun_out/r1_launches.csv profiles/r1_launches_summary.md
  python tools/summarize_ncu.py report   gpurun_out/r1_gemm.ncu-rep profiles/r1_gemm_ncu.md
"""
import collections, csv, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__inst_executed_pipe_uniform.sum"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", row["Kernel Name"])).replace("void ", "")
        tot[name] += v; cnt[name] += 1
    T = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\nOne bench.py --profile step (Llama-3-8B, seq 4096, 1 GPU); per-launch times are cold-cache and serialised: compare SHARES.\n\n")
        f.write(f"total kernel time {T/1e6:.2f} ms over {sum(cnt.values())} launches\n\n| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"| {k} | {cnt[k]} | {v/1e6:.3f} | {100*v/T:.1f}% | {v/cnt[k]/1e3:.1f} |\n")
    print(open(dst).read())


def report(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full capture ({src})\n\n")
        for row in rows[2:]:
            d = dict(zip(hdr, row)); u = dict(zip(hdr, units))
            f.write(f"## {d.get('Kernel Name','?')[:120]}  grid {d.get('Grid Size','')} block {d.get('Block Size','')}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in hdr:
                if any(k == key or k.endswith(key) for key in KEYS):
                    f.write(f"| {k} | {d[k]} | {u[k]} |\n")
            f.write("\n")
    print(open(dst).read()[:6000])


def traffic(src, dst):
    """ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm  -> average DRAM bytes per launch."""
    import json
    lines = [l for l in open(src) if not l.startswith("==")]
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(u, 1)
        per[row["ID"]][row["Metric Name"]] = v * mult
    n = len(per)
    rd = sum(d["dram__bytes_read.sum"] for d in per.values()); wr = sum(d["dram__bytes_write.sum"] for d in per.values())
    t = sum(d["gpu__time_duration.sum"] for d in per.values())
    out = {"kernel": "gemm_pair_kernel", "launches": n, "dram_bytes_read_per_launch": rd / n, "dram_bytes_write_per_launch": wr / n,
           "traffic_bytes_per_launch": (rd + wr) / n, "avg_duration_us_under_ncu": t / n / 1e3, "source": src}
    json.dump(out, open(dst, "w"), indent=1)
    print(out)


if __name__ == "__main__":
    {"launches": launches, "report": report, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
