"""Per-kernel summary of an ncu csv (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum):
launches, time, DRAM bytes and achieved DRAM GB/s per kernel name, as markdown + json.
Usage: python tools/ncu_summary.py gpurun_out/ncu_block.csv profiles/r02_ncu_per_kernel_16MiB [peak_GBps]"""
import csv
import json
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    peak = float(sys.argv[3]) if len(sys.argv) > 3 else 6572.2
    rows = []
    with open(src, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    per = {}
    for r in rd:
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^(void )?(bz3::)?", "", name)
        d = per.setdefault((r["ID"], name), {})
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(unit, 1.0)
            d["us"] = val
        else:
            val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            d["rd" if "read" in m else "wr"] = val
    agg = {}
    for (_, name), d in per.items():
        a = agg.setdefault(name, {"launches": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["launches"] += 1
        for k in ("us", "rd", "wr"):
            a[k] += d.get(k, 0.0)
    total_us = sum(a["us"] for a in agg.values())
    out = []
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        gbps = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] > 0 else 0.0
        out.append({"kernel": name, "launches": a["launches"], "time_us": round(a["us"], 1), "share": round(a["us"] / total_us, 5),
                    "dram_read_bytes": int(a["rd"]), "dram_write_bytes": int(a["wr"]), "dram_GBps": round(gbps, 1),
                    "frac_of_peak": round(gbps / peak, 4)})
    json.dump({"peak_GBps": peak, "kernels": out}, open(dst + ".json", "w"), indent=1)
    with open(dst + ".md", "w") as f:
        f.write("| kernel | launches | time (us) | share | DRAM read (MB) | DRAM write (MB) | DRAM GB/s | of %.0f GB/s |\n|---|---|---|---|---|---|---|---|\n" % peak)
        for o in out:
            f.write("| `%s` | %d | %.1f | %.2f %% | %.2f | %.2f | %.1f | %.1f %% |\n" % (
                o["kernel"], o["launches"], o["time_us"], 100 * o["share"], o["dram_read_bytes"] / 1e6, o["dram_write_bytes"] / 1e6,
                o["dram_GBps"], 100 * o["frac_of_peak"]))
    print(open(dst + ".md").read())


if __name__ == "__main__":
    main()
