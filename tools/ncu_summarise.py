"""Turn `ncu -i X.ncu-rep --page raw --csv` output into the few numbers DESIGN.md / bench.py quote:
per kernel launch: duration, DRAM bytes read / written, DRAM throughput, PCIe / NVLink bytes where the
counters exist.  usage: ncu_summarise.py <raw.csv> [<raw.csv> ...] -> JSON on stdout.
With --traffic <algorithmic bytes per launch of kernel K>=... writes profiles/r02_ncu_traffic.json."""
import csv
import json
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "pcie__read_bytes.sum", "pcie__write_bytes.sum",
        "nvlrx__bytes.sum", "nvltx__bytes.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__t_bytes.sum",
        "lts__t_sectors_srcunit_ltcfabric.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "msecond": 1e-3,
        "ms": 1e-3, "second": 1, "s": 1, "nsecond": 1e-9}


def parse(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = {"kernel": r[names.index("Kernel Name")], "id": r[names.index("ID")]}
        for k in WANT + [n for n in names if "nvl" in n.lower() or "pcie" in n.lower()]:
            if k in names and k not in d:
                i = names.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[k] = v * UNIT.get(units[i], 1)
        out.append(d)
    return out


if __name__ == "__main__":
    res = {}
    for p in [a for a in sys.argv[1:] if not a.startswith("--")]:
        res[p] = parse(p)
    print(json.dumps(res, indent=1))
